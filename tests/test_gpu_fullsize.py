"""GPU (-m gpu): the FULL-SIZE models (dim 192, 16 384 codes, encoder 64-128-256-256-576) against the CPU oracle on the BASELINE.json
workloads that round 1 only covered with small models or with the CUDA path against itself:

  * C4 (configs[3]): a -1 padded batch with faces uniform in 100..800, face_edges derived AND precomputed, padded AND ragged layout,
    RVQ and LFQ - every variant must give the same tokens, and those tokens are checked against o.tokenize token by token;
  * the C5 launch regime (configs[4]): >= 512 meshes in ONE launch, where the double-buffered group aggregation kernel is the default
    path; the oracle runs on a sample of the meshes (meshes are independent end to end, so the sample's tokens are the batch's).

Bar (helpers.assert_tokens_match_or_attributed): token agreement >= 0.999 (measured: 1.0) AND every mismatching token attributed to a
near tie - RVQ: the two candidates' distances at the oracle's quantizer input within RVQ_EPS (relative) plus twice the measured input
difference; LFQ: the flipped bit's value within LFQ_EPS plus the projected input difference of its threshold - AND the quantizer inputs
themselves (averaged vertex rows = the encoder end to end) within 2e-3 relative of the oracle's.
"""
import pytest
import torch

from helpers import assert_tokens_match_or_attributed, make_pair, match_rate
from meshgpt_pytorch_b200 import data as Dt
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)

pytestmark = pytest.mark.gpu


def c4_batch(n_meshes, seed=1234, first_seed=0):
    g = torch.Generator().manual_seed(seed)
    nf = torch.randint(100, 801, (n_meshes,), generator=g).tolist()
    return synth.ragged_batch("tri", 20, 20, nf, range(first_seed, first_seed + n_meshes))


@pytest.mark.parametrize("quantizer", ["rvq", "lfq"])
def test_c4_padded_batch_against_oracle_all_input_variants(quantizer):
    import meshgpt_pytorch_b200 as M
    o, m = make_pair(use_residual_lfq=(quantizer == "lfq"))
    v, f = c4_batch(64)
    assert f.shape[1] > 700 and int((f[:, :, 0] != -1).sum(1).min()) < 200          # really ragged: short and long meshes in one batch
    vd, fd = v.cuda(), f.cuda()
    edges = M.derive_face_edges_from_faces(fd)                                        # what a face-edge cache holds: (b, emax, 2), -1 padded
    r = Dt.padded_to_ragged(v, f)
    rv, rf, fo, vo = r["vertices"].cuda(), r["faces"].cuda(), r["face_offsets"], r["vertex_offsets"]
    keep = (edges != -1).all(-1)
    eo = torch.zeros(f.shape[0] + 1, dtype=torch.int64)
    eo[1:] = keep.sum(1).cpu().cumsum(0)
    variants = {
        "ragged, derived edges": Dt.ragged_to_padded_codes(m.tokenize_ragged(rv, rf, fo, vo), fo, 3, m.pad_id),
        "ragged, precomputed edges": Dt.ragged_to_padded_codes(m.tokenize_ragged(rv, rf, fo, vo, face_edges=edges[keep], edge_offsets=eo), fo, 3, m.pad_id),
        "padded, precomputed edges": m.tokenize(vd, fd, face_edges=edges),
    }
    got = m.tokenize(vd, fd)                                                          # padded, derived edges: LAST call, its taps are read below
    for name, codes in variants.items():
        assert torch.equal(codes, got), f"{name} differs from the padded layout with derived edges"
    rate, n_bad, in_err = assert_tokens_match_or_attributed(o, m, v, f, got, label=f"C4 {quantizer}")
    print(f"C4 {quantizer}: token agreement {rate:.6f}, {n_bad} vertices attributed to near ties, quantizer input within {in_err:.2e}")
    m.check_last_status()


@pytest.mark.parametrize("shape", ["uniform800", "ragged"])
def test_c5_launch_regime_512_meshes_against_oracle_on_a_sample(shape):
    """>= 512 meshes per launch: launch_gather_mean takes gather_mean_smem_group_kernel by default (sage.cu).  The whole launch runs on
    the GPU; the oracle tokenizes every 32nd mesh (16 of them) and must agree with the launch's tokens for those meshes."""
    o, m = make_pair(use_residual_lfq=False)
    B = 512
    if shape == "uniform800":
        v, f = synth.grid_batch("tri", 20, 20, range(B))
    else:
        v, f = c4_batch(B, seed=99)
    got_all = m.tokenize(v.cuda(), f.cuda()).cpu()
    m.check_last_status()
    assert got_all.shape == (B, f.shape[1] * 3, 2)
    sample = list(range(5, B, 32))
    vs, fs = v[sample], f[sample]
    # the same meshes alone (a 16-mesh launch takes the per-slice aggregation kernel): identical tokens, whichever kernel aggregated
    got_s = m.tokenize(vs.cuda(), fs.cuda())
    assert torch.equal(got_s.cpu(), got_all[sample]), "tokens of a mesh depend on the launch size (group vs per-slice aggregation kernel)"
    rate, n_bad, in_err = assert_tokens_match_or_attributed(o, m, vs, fs, got_s, label=f"C5 {shape}")
    print(f"C5 {shape}: token agreement {rate:.6f} on {len(sample)} sampled meshes, {n_bad} attributed, quantizer input within {in_err:.2e}")


@pytest.mark.parametrize("kind,cfg,nxny", [("tri", dict(use_residual_lfq=False), (20, 20)), ("quad", dict(quads=True), (20, 40)),
                                            ("tri", dict(), (20, 20))])
def test_c2_c3_tokens_attributed(kind, cfg, nxny):
    """BASELINE configs[1] (C2: RVQ, triangles) and configs[2] (C3: quads, LFQ), plus the default LFQ model, 8 x 800 faces each."""
    o, m = make_pair(**cfg)
    v, f = synth.grid_batch(kind, *nxny, list(range(8)))
    got = m.tokenize(v.cuda(), f.cuda())
    rate, n_bad, in_err = assert_tokens_match_or_attributed(o, m, v, f, got, label=f"{kind} {cfg}")
    print(f"{kind} {cfg}: token agreement {rate:.6f}, {n_bad} attributed, quantizer input within {in_err:.2e}")


def test_readme_c1_tokens_attributed():
    """BASELINE configs[0]: the README call on random tensors (degenerate faces, coordinates outside the range)."""
    o, m = make_pair()
    v, f = synth.readme_batch(0)
    got = m.tokenize(v.cuda(), f.cuda())
    assert_tokens_match_or_attributed(o, m, v, f, got, min_rate=0.99, max_input_err=5e-3, label="C1")


def test_fused_aggregation_is_bit_identical_to_the_separate_launches():
    """MESHTOK_GATHER_FUSED=1 (gather warps inside the SAGE output linears, linear_tc_pair_gather_kernel; off by default because it
    measured slower) against the separate aggregation launches on the full-size RVQ model, 67 meshes of 1..800 faces: every layer output
    and every code bit-identical.  scripts/gpu_gather_fused_check.py runs the two settings in subprocesses (the switch is read once)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gpu_gather_fused_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "bit-identical" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nxny", [(25, 25), (30, 30)])
def test_meshes_above_800_faces(nxny):
    """1 250 and 1 800 faces per mesh: a 32-column slice of such a mesh no longer fits twice per SM, so the aggregation takes the one-CTA-per-SM
    instantiation with several TMA boxes per slice (1 250) or, beyond the shared memory of an SM, the kernels that read the neighbour rows
    through L2 (1 800) - the per-mesh CSR blocks, the topology kernel's shared-memory tables and the folded-table sum at sizes no BASELINE
    workload reaches.  Same bar as every other full-size test."""
    o, m = make_pair(use_residual_lfq=False)
    v, f = synth.grid_batch("tri", *nxny, [0, 1, 2])
    got = m.tokenize(v.cuda(), f.cuda())
    m.check_last_status()
    rate, n_bad, in_err = assert_tokens_match_or_attributed(o, m, v, f, got, label=f"{2 * nxny[0] * nxny[1]} faces")
    print(f"{2 * nxny[0] * nxny[1]} faces per mesh: token agreement {rate:.6f}, {n_bad} attributed, quantizer input within {in_err:.2e}")
    # and next to small meshes in one -1 padded batch.  A flipped feature bin (an angle or a normal within an ulp of a rounding boundary,
    # scripts/gpu_flip_diag.py) reaches every vertex within five hops: the flipped faces are counted directly and the share of
    # quantizer inputs beyond 2e-3 is bounded by what that many flips can touch
    v2, f2 = synth.ragged_batch("tri", *nxny, [2 * nxny[0] * nxny[1], 37, 600], [3, 4, 5])
    m.keep_taps = True
    got2 = m.tokenize(v2.cuda(), f2.cuda())
    m.keep_taps = False
    m.check_last_status()
    _, st = o.forward_codes(vertices=v2, faces=f2, return_stages=True)
    flips = int(((m.taps()["x0"].cpu() - st["project_in"]).abs().max(-1).values > 1e-4).sum())
    print(f"   ragged: {flips} face rows with a flipped feature bin")
    assert flips <= 3, flips
    assert_tokens_match_or_attributed(o, m, v2, f2, got2, label=f"{2 * nxny[0] * nxny[1]} faces, ragged", max_bad_input_frac=0.005 + 0.02 * flips)
