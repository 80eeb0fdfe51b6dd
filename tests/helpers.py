"""Shared helpers of the parity tests: paired (oracle, CUDA) models with identical weights, comparisons."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import meshtok_oracle as O  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

SMALL_CFG = dict(dim_codebook=64, encoder_dims_through_depth=(32, 64, 64), codebook_size=256, num_quantizers=2,
                 dim_coor_embed=16, dim_area_embed=8, dim_normal_embed=16, dim_angle_embed=8)


def make_oracle(seed=42, codebook_std=0.05, **cfg):
    torch.manual_seed(seed)
    o = O.OracleMeshAutoencoder(**cfg)
    if not cfg.get("use_residual_lfq", True):
        cb = o.quantizer.layers[0]._codebook
        g = torch.Generator().manual_seed(seed + 1)
        cb.embed.copy_(torch.randn(cb.embed.shape, generator=g) * codebook_std)
        cb.embed_avg.copy_(cb.embed)
    return o.eval()


def make_pair(seed=42, codebook_std=0.05, **cfg):
    """(oracle on CPU, CUDA MeshAutoencoder) sharing one state dict."""
    import meshgpt_pytorch_b200 as M
    o = make_oracle(seed, codebook_std, **cfg)
    m = M.MeshAutoencoder(**cfg)
    m.load_reference_state_dict(o.state_dict())
    return o, m.cuda().eval()


def match_rate(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.cpu(), b.cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a == b).float().mean()) if a.numel() else 1.0


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


# ---------------------------------------------------------------------------------------------------
# end-to-end contract of BASELINE.json's north star, token by token:
#   "RVQ code ids match except where the top-2 distances lie within a stated epsilon",
#   "LFQ codes are bit-exact" given the quantizer input, i.e. a bit may differ only where the value it is the sign of
#   lies within epsilon of its threshold (|x| < eps at level 0, ||x| - 1| < eps at level 1, ...).
# End to end the CUDA path's quantizer input x_g differs from the oracle's x_o by the encoder's rounding (activations within
# rtol 1e-3), so the epsilon at vertex v is the stated one plus what |x_g - x_o| can move a distance (2 delta) or a projection
# (|W (x_g - x_o)|).  Every mismatching token must be explained this way; an unexplained one fails the test.
# ---------------------------------------------------------------------------------------------------
RVQ_EPS = 1e-4
LFQ_EPS = 2e-5


def assert_tokens_match_or_attributed(o, m, vertices, faces, got, face_edges=None, min_rate=0.999, max_input_err=2e-3, label="", max_bad_input_frac=0.005):
    """got: codes of the CUDA path from the LAST tokenize call of `m` on (vertices, faces) in the padded layout (its taps are read).
    Returns (token agreement, number of mismatching vertices, largest relative quantizer-input error)."""
    want, st = o.forward_codes(vertices=vertices, faces=faces, face_edges=face_edges, return_stages=True)
    got = got.cpu()
    assert got.shape == want.shape and got.dtype == torch.int64, label
    assert torch.equal(got == o.pad_id, want == o.pad_id), f"{label}: pad positions differ"
    rate = match_rate(got, want)
    t = m.taps()
    vert_row = t["vert_row"].cpu().long()                 # (B, V): quantizer row of each vertex or -1
    x_g_all = t["averaged"].cpu()                         # (n_vrows, D)
    x_o_all = st["averaged_vertices"]                     # (B, nv + 1, D); nv = faces.amax() + 1 <= V
    nvf = faces.shape[-1]
    B, V = vert_row.shape
    nv = x_o_all.shape[1] - 1
    # quantizer-input agreement over ALL referenced vertices (the encoder's end-to-end error at the quantizer boundary)
    ref = vert_row[:, :nv] >= 0
    xg = x_g_all[vert_row[:, :nv][ref]]
    xo = x_o_all[:, :nv][ref]
    err = (xg - xo).norm(dim=-1) / xo.norm(dim=-1).clamp(min=1e-6)
    frac_bad_input = float((err > max_input_err).float().mean()) if err.numel() else 0.0
    assert frac_bad_input <= max_bad_input_frac, f"{label}: {frac_bad_input:.4f} of the quantizer input rows are off by more than {max_input_err} (a flipped feature bin explains a few)"
    # mismatching tokens -> (mesh, vertex) pairs
    bad_tok = (got != want).any(-1).nonzero()
    fl = faces.reshape(B, -1).long()
    pairs = sorted({(int(b), int(fl[b, tk])) for b, tk in bad_tok.tolist()})
    tok_of = {}
    for b, tk in bad_tok.tolist():
        tok_of.setdefault((b, int(fl[b, tk])), tk)
    for b, vid in pairs:
        tk = tok_of[(b, vid)]
        g_codes, w_codes = got[b, tk], want[b, tk]
        x_o = x_o_all[b, vid]
        x_g = x_g_all[vert_row[b, vid]]
        delta = float((x_g - x_o).norm())
        if o.use_residual_lfq:
            q = o.quantizer
            p_o = q.project_in(x_o)
            dp = (q.project_in.weight @ (x_g - x_o)).abs()
            bits = p_o.shape[0]
            r = p_o.clone()
            for lvl in range(o.num_quantizers):
                scale = 2.0 ** (-lvl)
                gb = [(int(g_codes[lvl]) >> (bits - 1 - d)) & 1 for d in range(bits)]
                wb = [(int(w_codes[lvl]) >> (bits - 1 - d)) & 1 for d in range(bits)]
                diff = [d for d in range(bits) if gb[d] != wb[d]]
                for d in diff:
                    assert float(r[d].abs()) <= float(dp[d]) + LFQ_EPS, \
                        f"{label}: LFQ bit (mesh {b}, vertex {vid}, level {lvl}, dim {d}) differs but the value {float(r[d]):.3e} is not within {float(dp[d]) + LFQ_EPS:.3e} of its threshold"
                if diff:
                    break                       # later levels start from a different residual: nothing to attribute there
                r = r - torch.where(r > 0, torch.tensor(scale), torch.tensor(-scale))
        else:
            e = o.quantizer.codebook
            x = x_o.clone()
            for lvl in range(o.num_quantizers):
                if int(g_codes[lvl]) != int(w_codes[lvl]):
                    d_got = float((x - e[int(g_codes[lvl])]).norm())
                    d_want = float((x - e[int(w_codes[lvl])]).norm())
                    assert abs(d_got - d_want) <= 2.0 * delta + RVQ_EPS * d_want + 1e-7, \
                        f"{label}: RVQ code (mesh {b}, vertex {vid}, level {lvl}) differs but the two distances {d_got:.6e} / {d_want:.6e} are not within {2 * delta + RVQ_EPS * d_want:.3e}"
                    break
                x = x - e[int(w_codes[lvl])]
    assert rate >= min_rate, f"{label}: token agreement {rate:.5f} < {min_rate}"
    return rate, len(pairs), float(err.max()) if err.numel() else 0.0
