"""Regenerates the golden fixtures of tests/golden/.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

own_code.pt   outputs of the REFERENCE'S OWN functions (executed from /root/reference source by
              oracle/reference_extract.py) on seeded inputs: derive_face_edges_from_faces,
              get_derived_face_features, discretize, scatter_mean.  These pin the oracle's own-code stages.
decode_own_code.pt   the same for the decoder building blocks (groundwork for the decode path): PixelNorm, SqueezeExcite,
              Block, ResnetBlock, undiscretize.
pipeline_*.pt inputs, a small state dict and the ORACLE's stage-by-stage outputs for small models (LFQ tri,
              RVQ tri, LFQ quad + ragged padding).  The third-party stages inside them (SAGEConv,
              ResidualVQ/LFQ) are restatements - parity unpinned, see oracle/meshtok_oracle.py.
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import meshtok_oracle as O  # noqa: E402
from oracle import reference_extract  # noqa: E402
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
import helpers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def own_code():
    R = reference_extract.load()
    g = torch.Generator().manual_seed(7)
    fix = {}
    cases = {}
    cases["kat_degenerate"] = torch.tensor([[5, 5, 7], [5, 8, 9], [-1, -1, -1]])
    cases["kat_quads"] = torch.tensor([[0, 1, 2, 3], [2, 3, 4, 5], [9, 9, 9, 9], [-1] * 4])
    cases["kat_batch"] = torch.tensor([[[0, 1, 2], [1, 2, 3], [2, 3, 4]], [[0, 1, 2], [-1] * 3, [-1] * 3]])
    cases["readme"] = synth.readme_batch(0)[1]
    rnd = torch.randint(0, 40, (3, 60, 3), generator=g)
    rnd[1, 45:] = -1
    rnd[2, 7:] = -1
    cases["random_tri"] = rnd
    rq = torch.randint(0, 50, (2, 48, 4), generator=g)
    rq[1, 30:] = -1
    cases["random_quad"] = rq
    fan = torch.stack([torch.zeros(70, dtype=torch.long), torch.arange(1, 71), torch.arange(2, 72)], -1)
    cases["fan70"] = fan
    cases["grid_tri_6x5"] = synth.grid_mesh("tri", 6, 5, 0)[1]
    cases["grid_quad_4x7"] = synth.grid_mesh("quad", 4, 7, 0)[1]
    edges = {}
    for name, faces in cases.items():
        for flags in (dict(), dict(neighbor_if_share_one_vertex=True), dict(include_self=False)):
            key = name + "|" + ",".join(f"{k}={v}" for k, v in flags.items())
            edges[key] = dict(faces=faces, flags=flags, out=R.derive_face_edges_from_faces(faces, **flags))
    fix["face_edges"] = edges
    fix["grid_counts"] = dict(tri_20x20=int(R.derive_face_edges_from_faces(synth.grid_mesh("tri", 20, 20, 0)[1]).shape[0]),
                              quad_20x40=int(R.derive_face_edges_from_faces(synth.grid_mesh("quad", 20, 40, 0)[1]).shape[0]))
    feats = {}
    for nvf in (3, 4):
        fc = torch.randn(2, 40, nvf, 3, generator=g)
        fc[0, 0] = 0.0  # zero vectors -> normalize eps path
        feats[nvf] = dict(face_coords=fc, **R.get_derived_face_features(fc))
    fix["features"] = feats
    x = torch.cat([torch.randn(4000, generator=g) * 1.3, torch.tensor([-5.0, -1.0, 1.0, 5.0, float("nan")]),
                   torch.arange(6).float().add(1).div(64).sub(1)])
    fix["discretize"] = dict(x=x, out={f"{lo},{hi}": R.discretize(x.clone(), continuous_range=(lo, hi), num_discrete=128)
                                       for lo, hi in ((-1.0, 1.0), (0.0, math.pi), (0.0, 4.0))})
    src = torch.randn(2, 30, 8, generator=g)
    idx = torch.randint(0, 12, (2, 30), generator=g)
    fix["scatter_mean"] = dict(src=src, idx=idx, rows=12,
                               out=R.scatter_mean(torch.zeros(2, 12, 8), idx[:, :, None].expand(2, 30, 8), src, dim=-2))
    torch.save(fix, os.path.join(OUT, "own_code.pt"))


def decode_own_code():
    """Outputs of the reference's OWN decoder building blocks (PixelNorm, SqueezeExcite, Block, ResnetBlock, undiscretize,
    meshgpt_pytorch.py:203-218, 286-379) on seeded weights and inputs: the pins of oracle/decode_oracle.py."""
    R = reference_extract.load()
    g = torch.Generator().manual_seed(11)
    fix = {}
    x = torch.randn(3, 16, 37, generator=g)
    mask = torch.ones(3, 1, 37, dtype=torch.bool)
    mask[1, :, 20:] = False
    mask[2, :, 1:] = False
    fix["x"], fix["mask"] = x, mask
    fix["pixelnorm"] = R.PixelNorm(dim=1)(x)
    blocks = {}
    for name, (din, dout) in dict(same=(16, 16), wider=(16, 24)).items():
        torch.manual_seed(3)
        blk = R.ResnetBlock(din, dout).eval()
        for p_ in blk.parameters():                      # default inits are tiny for the biases: make every term count
            p_.data = torch.randn(p_.shape, generator=g) * 0.3
        with torch.no_grad():
            blocks[name] = dict(dims=(din, dout), state_dict={k: v.clone() for k, v in blk.state_dict().items()},
                                out_masked=blk(x, mask=mask), out_unmasked=blk(x),
                                block1=blk.block1(x, mask=mask), excite=blk.excite(blk.block2(blk.block1(x, mask=mask), mask=mask), mask=mask))
    fix["resnet"] = blocks
    bins = torch.randint(0, 128, (2, 5, 3, 3), generator=g)
    fix["undiscretize"] = dict(bins=bins, out=R.undiscretize(bins.clone(), continuous_range=(-1.0, 1.0), num_discrete=128),
                               out_other=R.undiscretize(bins.clone() % 64, continuous_range=(0.0, 4.0), num_discrete=64))
    torch.save(fix, os.path.join(OUT, "decode_own_code.pt"))


def pipeline(name, cfg, vertices, faces):
    o = helpers.make_oracle(seed=42, **cfg)
    codes, stages = o.forward_codes(vertices=vertices, faces=faces, return_stages=True)
    keep = {k: v for k, v in stages.items()}
    torch.save(dict(cfg=cfg, state_dict={k: v.clone() for k, v in o.state_dict().items()}, vertices=vertices, faces=faces,
                    codes=codes, stages=keep), os.path.join(OUT, f"pipeline_{name}.pt"))


if __name__ == "__main__":
    own_code()
    decode_own_code()
    v, f = synth.readme_batch(0, batch=2, num_vertices=121, num_faces=64)
    pipeline("lfq_tri", dict(helpers.SMALL_CFG), v, f)
    v, f = synth.grid_batch("tri", 6, 5, [0, 1, 2])
    pipeline("rvq_tri", dict(helpers.SMALL_CFG, use_residual_lfq=False), v, f)
    v, f = synth.ragged_batch("quad", 5, 6, [30, 17, 9], [3, 4, 5])
    pipeline("lfq_quad_ragged", dict(helpers.SMALL_CFG, quads=True), v, f)
    print({n: os.path.getsize(os.path.join(OUT, n)) for n in sorted(os.listdir(OUT)) if n.endswith(".pt")})
