"""GPU (-m gpu): the per-stage C-ABI entry points of SURVEY 8b (meshtok_stage_*), called one by one through ctypes on the workspace layout
of meshtok_tokenize, exchanging fp32 rows.  Every stage's output is compared with what the fused pipeline computed for the same batch
(its workspace taps), the final codes with MeshAutoencoder.tokenize, and the chain as a whole with the oracle."""
import ctypes as C

import pytest
import torch

from helpers import SMALL_CFG, assert_tokens_match_or_attributed, make_pair, match_rate
import synth

pytestmark = pytest.mark.gpu


def _stage_chain(m, v, f, face_edges=None):
    from meshgpt_pytorch_b200 import _lib
    from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr
    prep = m._prepare()
    model = prep["model"]
    B, F, nvf = f.shape
    V = v.shape[1]
    D, Q = m.dim_codebook, m.num_quantizers
    E = 0 if face_edges is None else face_edges.shape[1]
    cap = max(1024, m.edge_slots_per_face * B * F, B * E)
    nbytes = C.c_size_t()
    check(lib.meshtok_tokenize_workspace_bytes(C.byref(model), B, F, V, E, cap, C.byref(nbytes)))
    ws = torch.zeros(nbytes.value, dtype=torch.uint8, device="cuda")
    codes = torch.empty((B, F * nvf, Q), dtype=torch.int64, device="cuda")
    coords = torch.empty((B, F, nvf * 3), dtype=torch.int64, device="cuda")
    a = _lib.TokenizeArgs()
    a.vertices, a.faces, a.face_edges = ptr(v), ptr(f), ptr(face_edges)
    a.B, a.F, a.V, a.E = B, F, V, E
    a.pad_id, a.edge_capacity = int(m.pad_id), cap
    a.codes_out, a.face_coords_out = ptr(codes), ptr(coords)
    args = (C.byref(model), C.byref(a))
    tail = (ptr(ws), ws.numel(), stream_ptr())
    out = {}
    check(lib.meshtok_stage_topology(*args, *tail))
    t = _lib.Taps()
    check(lib.meshtok_tokenize_taps(C.byref(model), B, F, V, E, cap, C.byref(t)))
    torch.cuda.synchronize()
    n_faces, n_vrows, n_edges = ws[t.counts: t.counts + 12].view(torch.int32).tolist()
    out.update(n_faces=n_faces, n_vrows=n_vrows, n_edges=n_edges,
               indptr=ws[t.indptr: t.indptr + 4 * (n_faces + 1)].view(torch.int32).clone(), src=ws[t.src: t.src + 4 * n_edges].view(torch.int32).clone())
    rows = B * F
    x = torch.zeros(rows, D, device="cuda")
    check(lib.meshtok_stage_face_features_embed(*args, ptr(x), *tail))
    out["x0"] = x[:n_faces].clone()
    out["face_coords"] = coords.clone()
    dims = m.encoder_dims
    for layer, d_out in enumerate(dims):
        d_in = x.shape[1]
        xp = torch.zeros(rows, d_in, device="cuda")
        check(lib.meshtok_stage_sage_project(*args, layer, ptr(x), ptr(xp), *tail))
        y = torch.zeros(rows, d_out, device="cuda")
        check(lib.meshtok_stage_sage_aggregate_linear(*args, layer, ptr(x), ptr(xp), ptr(y), *tail))
        out[f"layer{layer}"] = y[:n_faces].clone()
        x = y
    avg = torch.zeros(B * V, D, device="cuda")
    check(lib.meshtok_stage_project_scatter_mean(*args, ptr(x), ptr(avg), *tail))
    out["averaged"] = avg[:n_vrows].clone()
    vcodes = torch.empty(max(n_vrows, 1), Q, dtype=torch.int32, device="cuda")
    if m.use_residual_lfq:
        check(lib.meshtok_lfq_codes(C.byref(model), ptr(avg), n_vrows, ptr(vcodes), stream_ptr()))
    else:
        qb = C.c_size_t()
        check(lib.meshtok_rvq_workspace_bytes(C.byref(model), n_vrows, C.byref(qb)))
        qws = torch.empty(qb.value, dtype=torch.uint8, device="cuda")
        check(lib.meshtok_rvq_codes(C.byref(model), ptr(avg), n_vrows, 0, ptr(vcodes), None, ptr(qws), qb.value, stream_ptr()))
    check(lib.meshtok_stage_gather_codes(*args, ptr(vcodes), *tail))
    torch.cuda.synchronize()
    assert int(ws[:4].view(torch.int32).item()) == 0, "status word"
    out["codes"] = codes
    return out


@pytest.mark.parametrize("cfg,kind,grid,counts", [
    (dict(SMALL_CFG, use_residual_lfq=False), "tri", (9, 8), [144, 50, 1]),
    (dict(), "tri", (20, 20), [800, 377]),                                  # the default model: LFQ, full size
    (dict(use_residual_lfq=False), "tri", (20, 20), [800, 123, 640]),       # full-size RVQ (tcgen05 search on the staged rows)
    (dict(SMALL_CFG, quads=True), "quad", (6, 7), [42, 17]),
])
def test_stages_one_by_one_reproduce_the_pipeline(cfg, kind, grid, counts):
    o, m = make_pair(**cfg)
    v, f = synth.ragged_batch(kind, grid[0], grid[1], counts, list(range(len(counts))))
    vd, fd = v.cuda(), f.cuda()
    st = _stage_chain(m, vd, fd)
    _, dc = m.encode(vertices=vd, faces=fd, face_edges=torch.full((f.shape[0], 1, 2), -1, device="cuda"), face_mask=(fd != -1).all(-1),
                     face_edges_mask=torch.zeros(f.shape[0], 1, dtype=torch.bool, device="cuda"), return_face_coordinates=True)
    m.keep_taps = True
    want = m.tokenize(vd, fd)          # the LAST pipeline call: its taps are read here and by the attribution below
    m.keep_taps = False
    t = m.taps()
    # topology: the same CSR, the same counts
    assert (st["n_faces"], st["n_vrows"], st["n_edges"]) == (t["n_faces"], t["n_vrows"], t["n_edges"])
    assert torch.equal(st["indptr"], t["indptr"]) and torch.equal(st["src"], t["src"])
    # features: the same table rows summed in the same order -> the same bits; discretised coordinates exact
    assert torch.equal(st["x0"], t["x0"])
    assert torch.equal(st["face_coords"], dc)
    # SAGE layers: the stage path converts fp32 rows to operand images where the pipeline's epilogues write them directly - the same
    # values; the last layer is normalised before project_dim_codebook instead of inside its epilogue (one rounding apart)
    for i in range(len(m.encoder_dims)):
        a, b = st[f"layer{i}"], t[f"layer{i}"]
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), (i, float((a - b).abs().max()))
    assert torch.allclose(st["averaged"], t["averaged"], rtol=1e-5, atol=1e-6)
    assert match_rate(st["codes"], want) >= 0.999
    assert torch.equal(st["codes"] == -1, want == -1)
    # and the chain against the oracle (the taps of the tokenize call above are the CUDA path's quantizer inputs)
    assert_tokens_match_or_attributed(o, m, v, f, want, min_rate=0.995, label=str(cfg))


def test_stage_topology_with_user_edges():
    """csr_from_edge_list: user-supplied face_edges through the stage entry points."""
    from oracle import meshtok_oracle as O
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    v, f = synth.ragged_batch("tri", 8, 8, [128, 50, 7], [0, 1, 2])
    fe = O.derive_face_edges_from_faces(f)
    st = _stage_chain(m, v.cuda(), f.cuda(), face_edges=fe.cuda())
    want = m.tokenize(v.cuda(), f.cuda(), face_edges=fe.cuda())
    assert match_rate(st["codes"], want) >= 0.999 and torch.equal(st["codes"] == -1, want == -1)
