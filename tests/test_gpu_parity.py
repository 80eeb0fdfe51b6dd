"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on identical seeded inputs
and weights, stage by stage and end to end, plus the committed golden fixtures.

Bars (BASELINE.json north_star):
  * face_edges, discretised coordinates: bit-exact.
  * LFQ codes: bit-exact at the quantizer boundary (oracle-injected rows); a code bit may only differ where
    the projected value it is the sign of lies within LFQ_EPS of its threshold.
  * encoder activations: rtol 1e-3 (+ atol 1e-4 on unit-norm rows), except faces whose derived-feature BIN
    differs because acosf / the norm round differently on the two machines (counted and bounded).
  * RVQ code ids: equal except where the two best distances lie within RVQ_EPS (relative).
"""
import ctypes as C
import math
import os

import pytest
import torch

from helpers import GOLDEN_DIR, SMALL_CFG, assert_tokens_match_or_attributed, make_oracle, make_pair, match_rate, rel_err
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
from oracle import meshtok_oracle as O

pytestmark = pytest.mark.gpu

from helpers import LFQ_EPS, RVQ_EPS  # noqa: E402  (2e-5 / 1e-4: the stated epsilons of the quantizer contracts)
ACT_RTOL, ACT_ATOL = 1e-3, 1e-4


def cuda(*ts):
    return [None if t is None else t.cuda() for t in ts]


# ---------------------------------------------------------------------------------------------------
# K1 face adjacency
# ---------------------------------------------------------------------------------------------------


def _edge_cases():
    g = torch.Generator().manual_seed(5)
    cases = dict(
        degenerate=torch.tensor([[5, 5, 7], [5, 8, 9], [-1, -1, -1]]),
        quads=torch.tensor([[0, 1, 2, 3], [2, 3, 4, 5], [9, 9, 9, 9], [-1] * 4]),
        batch=torch.tensor([[[0, 1, 2], [1, 2, 3], [2, 3, 4]], [[0, 1, 2], [-1] * 3, [-1] * 3]]),
        unbatched=torch.tensor([[0, 1, 2], [1, 2, 3]]),
        all_pad=torch.full((2, 5, 3), -1),
        readme=synth.readme_batch(0)[1],
        fan=torch.stack([torch.zeros(300, dtype=torch.long), torch.arange(1, 301), torch.arange(2, 302)], -1),
        grid_tri=synth.grid_batch("tri", 20, 20, [0, 1])[1],
        grid_quad=synth.grid_batch("quad", 20, 40, [0])[1],
        sparse_ids=torch.randint(0, 10 ** 9, (2, 40, 3), generator=g),
        int32=torch.randint(0, 30, (2, 50, 3), generator=g).int(),
    )
    r = torch.randint(0, 40, (4, 130, 3), generator=g)
    r[1, 100:] = -1
    r[2, 1:] = -1
    r[3, 40:90] = -1            # padding in the middle: positions, not ranks, are reported
    cases["random_ragged"] = r
    rq = torch.randint(0, 60, (3, 77, 4), generator=g)
    rq[0, 50:] = -1
    cases["random_quads"] = rq
    return cases


@pytest.mark.parametrize("flags", [dict(), dict(neighbor_if_share_one_vertex=True), dict(include_self=False)])
def test_face_edges_bit_exact(flags):
    import meshgpt_pytorch_b200 as M
    for name, faces in _edge_cases().items():
        want = O.derive_face_edges_from_faces(faces, **flags)
        got = M.derive_face_edges_from_faces(faces.cuda(), **flags)
        assert got.dtype == torch.int64 and got.is_cuda, name
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert torch.equal(got.cpu(), want), name


def test_face_edges_other_pad_ids_and_golden():
    import meshgpt_pytorch_b200 as M
    faces = torch.tensor([[[0, 1, 2], [1, 2, 3], [7, 7, 7]], [[0, 1, 2], [7, 0, 1], [4, 5, 6]]])
    for pad in (7, -3, 0):
        want = O.derive_face_edges_from_faces(faces, pad_id=pad)
        got = M.derive_face_edges_from_faces(faces.cuda(), pad_id=pad)
        assert torch.equal(got.cpu(), want), pad
    own = torch.load(os.path.join(GOLDEN_DIR, "own_code.pt"))
    for key, case in own["face_edges"].items():     # outputs of the reference's own function
        got = M.derive_face_edges_from_faces(case["faces"].cuda(), **case["flags"])
        assert torch.equal(got.cpu(), case["out"]), key
    assert M.derive_face_edges_from_faces(synth.grid_mesh("tri", 20, 20, 0)[1].cuda()).shape == (3120, 2)
    assert M.derive_face_edges_from_faces(synth.grid_mesh("quad", 20, 40, 0)[1].cuda()).shape == (3880, 2)


# ---------------------------------------------------------------------------------------------------
# linears: tcgen05 split-bf16 and fp32 SIMT against an fp64 product
# ---------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("M_,N,K", [(1000, 192, 192), (333, 64, 384), (4096, 576, 512), (129, 768, 576), (128, 256, 64)])
def test_linear_kernels(M_, N, K):
    from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(M_ + N + K)
    A = torch.randn(M_, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = (A.double() @ W.double().t() + b.double())
    nbytes = C.c_size_t()
    check(lib.meshtok_linear_image_bytes(N, K, C.byref(nbytes)))
    img = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_linear_build_image(ptr(W), N, K, ptr(img), stream_ptr()))
    check(lib.meshtok_rows_image_bytes(M_, K, C.byref(nbytes)))
    a_img = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_rows_build_image(ptr(A), M_, K, ptr(a_img), stream_ptr()))
    for simt, relu, tol in ((1, 0, 2e-6), (0, 0, 3e-5), (0, 1, 3e-5)):
        out = torch.full((M_, N), float("nan"), device="cuda")
        check(lib.meshtok_linear(ptr(A), ptr(a_img), ptr(W), ptr(img), ptr(b), ptr(out), M_, N, K, relu, simt, stream_ptr()))
        torch.cuda.synchronize()
        want = ref.clamp(min=0) if relu else ref
        err = float((out.double() - want).abs().max() / want.abs().max())
        assert err < tol, (simt, relu, err)


# ---------------------------------------------------------------------------------------------------
# quantizers with oracle-injected rows
# ---------------------------------------------------------------------------------------------------


def test_lfq_codes_bit_exact_at_the_quantizer_boundary():
    from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr
    o, m = make_pair()                                           # default: LFQ, 16384 codes, dim 192
    prep = m._prepare()
    g = torch.Generator().manual_seed(3)
    rows = torch.randn(20000, 192, generator=g) * 0.3
    rows[:7] = 0
    _, want = o.quantizer(rows)
    codes = torch.empty(rows.shape[0], 2, dtype=torch.int32, device="cuda")
    check(lib.meshtok_lfq_codes(C.byref(prep["model"]), ptr(rows.cuda()), rows.shape[0], ptr(codes), stream_ptr()))
    got = codes.cpu().long()
    bad = (got != want).any(-1)
    x = o.quantizer.project_in(rows)
    near = torch.minimum(x.abs(), (x.abs() - 1).abs()).min(-1).values     # distance of the closest sign threshold
    assert int(bad.sum()) <= 2 and bool((near[bad] < LFQ_EPS).all()), (int(bad.sum()), near[bad])
    assert int(got.min()) >= 0 and int(got.max()) < 16384


def _rvq_check(o, m, rows, exact):
    from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr
    prep = m._prepare()
    R = rows.shape[0]
    nbytes = C.c_size_t()
    check(lib.meshtok_rvq_workspace_bytes(C.byref(prep["model"]), R, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    codes = torch.full((R, 2), -7, dtype=torch.int32, device="cuda")
    resid = torch.empty(R, rows.shape[1], device="cuda")
    check(lib.meshtok_rvq_codes(C.byref(prep["model"]), ptr(rows.cuda()), R, int(exact), ptr(codes), ptr(resid), ptr(ws),
                                nbytes.value, stream_ptr()))
    torch.cuda.synchronize()
    qsum, want = o.quantizer(rows)
    got = codes.cpu().long()
    e = o.quantizer.codebook
    # every disagreement must be a near tie of the two candidates' distances (level 0 of the row decides)
    bad = (got != want).any(-1).nonzero().flatten()
    for r in bad.tolist():
        lvl = 0 if got[r, 0] != want[r, 0] else 1
        x = rows[r] if lvl == 0 else rows[r] - e[want[r, 0]]
        d_got = (x - e[got[r, lvl]]).norm()
        d_want = (x - e[want[r, lvl]]).norm()
        assert abs(float(d_got - d_want)) <= RVQ_EPS * float(d_want) + 1e-7, (r, lvl, float(d_got), float(d_want))
    agree = (got == want).all(-1)
    assert torch.allclose(resid.cpu()[agree], (rows - qsum)[agree], rtol=1e-5, atol=1e-6)
    return 1.0 - len(bad) / R


@pytest.mark.parametrize("exact", [True, False])
def test_rvq_codes_full_size_codebook(exact):
    o, m = make_pair(use_residual_lfq=False)                     # 16384 x 192 shared codebook, 2 levels
    g = torch.Generator().manual_seed(11)
    rows = torch.randn(3000, 192, generator=g) * 0.03
    rows[:3] = 0
    rows[3:40] = o.quantizer.codebook[100:137] + 1e-3 * torch.randn(37, 192, generator=g)   # near-exact hits
    rate = _rvq_check(o, m, rows, exact)
    assert rate > 0.999, rate


def test_rvq_tc_small_codebooks_and_ragged_row_counts():
    for cfg, R in ((dict(SMALL_CFG, use_residual_lfq=False), 777), (dict(SMALL_CFG, use_residual_lfq=False, codebook_size=1024), 1),
                   (dict(SMALL_CFG, use_residual_lfq=False, dim_codebook=128, codebook_size=2048), 5000)):
        o, m = make_pair(**cfg)
        g = torch.Generator().manual_seed(R)
        rows = torch.randn(R, cfg["dim_codebook"], generator=g) * 0.05
        assert _rvq_check(o, m, rows, exact=False) > 0.995


def test_rvq_ambiguous_rows_take_the_exact_scan_inside_the_rescore_launch():
    """A codebook of 16 distinct vectors repeated 64 times: the screened candidate lists of EVERY row are full of exact ties, so every row
    is flagged ambiguous and decided by the CTA-cooperative fp32 scan inside rvq_rescore_kernel (round 1: a launch of its own).  The
    winner among identical codes must be the lowest index, like torch.argmax; also checked: NaN rows do not fault (code 0)."""
    from meshgpt_pytorch_b200._lib import check, lib, ptr, stream_ptr
    cfg = dict(SMALL_CFG, use_residual_lfq=False, codebook_size=1024)
    o, m = make_pair(**cfg)
    g = torch.Generator().manual_seed(5)
    base = torch.randn(16, 64, generator=g) * 0.05
    cb = base.repeat(64, 1)
    o.quantizer.layers[0]._codebook.embed.copy_(cb[None])
    m.load_reference_state_dict(o.state_dict())
    rows = torch.randn(1500, 64, generator=g) * 0.05
    assert _rvq_check(o, m, rows, exact=False) > 0.999
    # through the pipeline: the statistic says the scan ran, and the codes still agree with the oracle
    v, f = synth.grid_batch("tri", 8, 8, [0, 1])
    got = m.tokenize(v.cuda(), f.cuda()).cpu()
    assert m.taps()["rvq_exact_rechecks"] > 0
    assert int(got.max()) < 16, "identical codes: the lowest index (one of the first 16) must win"
    assert match_rate(got, o.tokenize(v, f)) > 0.99
    # a NaN row must not fault or produce an out-of-range id
    prep = m._prepare()
    bad = rows[:64].clone()
    bad[3] = float("nan")
    nbytes = C.c_size_t()
    check(lib.meshtok_rvq_workspace_bytes(C.byref(prep["model"]), 64, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    codes = torch.full((64, 2), -7, dtype=torch.int32, device="cuda")
    for exact in (0, 1):
        check(lib.meshtok_rvq_codes(C.byref(prep["model"]), ptr(bad.cuda()), 64, exact, ptr(codes), None, ptr(ws), nbytes.value, stream_ptr()))
        torch.cuda.synchronize()
        assert int(codes.min()) >= 0 and int(codes.max()) < 1024


# ---------------------------------------------------------------------------------------------------
# encoder, stage by stage (taps) and through encode()
# ---------------------------------------------------------------------------------------------------


def _encoder_stage_check(o, m, vertices, faces, label):
    codes, st = o.forward_codes(vertices=vertices, faces=faces, return_stages=True)
    m.keep_taps = True
    got_codes = m.tokenize(*cuda(vertices, faces))
    m.keep_taps = False
    assert torch.equal(m.tokenize(*cuda(vertices, faces)), got_codes), label      # the taps do not change the result
    t = m.taps()
    n = st["project_in"].shape[0]
    assert t["n_faces"] == n, label
    # CSR == the oracle's edge_index (targets = second column), sources ascending per target
    ei = st["edge_index"]
    order = torch.argsort(ei[1] * n + ei[0])
    assert t["n_edges"] == ei.shape[1]
    assert torch.equal(t["src"].cpu().long(), ei[0][order]), label
    assert torch.equal(t["indptr"].cpu().long(), torch.searchsorted(ei[1][order], torch.arange(n + 1))), label
    # rows whose derived-feature bins all agree must match tightly from project_in on
    x0 = t["x0"].cpu()
    row_ok = ((x0 - st["project_in"]).abs().max(-1).values < 1e-4)
    flip_rate = 1.0 - float(row_ok.float().mean())
    assert flip_rate < 0.02, (label, flip_rate)
    assert torch.allclose(x0[row_ok], st["project_in"][row_ok], rtol=ACT_RTOL, atol=ACT_ATOL)
    # a face is "clean" if neither it nor any face within the receptive field of the 5 layers had a flip;
    # with few flips simply restrict to meshes without any
    mesh_of_row = (faces != -1).all(-1).nonzero()[:, 0]
    clean_mesh = torch.ones(faces.shape[0], dtype=torch.bool)
    clean_mesh[mesh_of_row[~row_ok]] = False
    clean = clean_mesh[mesh_of_row]
    assert bool(clean.any()), label
    names = ["sage0_act_norm"] + [f"sage{i}" for i in range(1, len(m.encoder_dims))]
    for i, name in enumerate(names):
        a, b = t[f"layer{i}"].cpu()[clean], st[name][clean]
        assert torch.allclose(a, b, rtol=ACT_RTOL, atol=ACT_ATOL), (label, name, float((a - b).abs().max()))
    return got_codes.cpu(), codes, clean_mesh, st


@pytest.mark.parametrize("simt", [True, False])
def test_encoder_stages_default_model_readme_and_grid(simt):
    o, m = make_pair()                                           # the default MeshAutoencoder (LFQ, tri)
    m.force_gemm_simt = simt
    v, f = synth.readme_batch(0)                                 # BASELINE config 1
    got, want, clean, st = _encoder_stage_check(o, m, v, f, f"readme simt={simt}")
    v, f = synth.grid_batch("tri", 20, 20, [0, 1, 2, 3])
    got, want, clean, st = _encoder_stage_check(o, m, v, f, f"grid simt={simt}")
    rate = match_rate(got[clean], want[clean])
    assert rate > 0.995, rate


def test_encode_and_quantize_methods_match_oracle():
    o, m = make_pair(**dict(SMALL_CFG, quads=True))
    v, f = synth.ragged_batch("quad", 6, 7, [42, 20, 5], [0, 1, 2])
    fe = O.derive_face_edges_from_faces(f)
    fm, fem = (f != -1).all(-1), (fe != -1).all(-1)
    enc_o, dc_o = o.encode(vertices=v, faces=f, face_edges=fe, face_mask=fm, face_edges_mask=fem, return_face_coordinates=True)
    enc, dc = m.encode(vertices=v.cuda(), faces=f.cuda(), face_edges=fe.cuda(), face_mask=fm.cuda(), face_edges_mask=fem.cuda(),
                       return_face_coordinates=True)
    assert dc.dtype == torch.int64 and torch.equal(dc.cpu()[fm], dc_o[fm])              # discretised coordinates: exact
    assert enc.shape == enc_o.shape and bool((enc.cpu()[~fm] == 0).all())
    assert rel_err(enc, enc_o) < 2e-3
    # quantize() on the ORACLE's encoder output: codes must agree exactly up to sign-threshold ties
    fo, co, _ = o.quantize(faces=f, face_mask=fm, face_embed=enc_o)
    fq, cq, loss = m.quantize(faces=f.cuda(), face_mask=fm.cuda(), face_embed=enc_o.cuda())
    assert cq.shape == co.shape and cq.dtype == torch.int64 and float(loss.abs().sum()) == 0
    assert match_rate(cq, co) > 0.999
    assert bool((cq.cpu()[~fm[:, :, None].expand(-1, -1, 4).reshape(3, -1)] == -1).all())
    same = (cq.cpu() == co).all(-1).reshape(3, -1, 4).all(-1) & fm
    assert torch.allclose(fq.cpu()[same], fo[same], rtol=1e-4, atol=1e-5)
    assert torch.allclose(fq.cpu()[~fm], fo[~fm], rtol=1e-4, atol=1e-5)                # pad faces gather the pad-vertex row


# ---------------------------------------------------------------------------------------------------
# end to end
# ---------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["lfq_tri", "rvq_tri", "lfq_quad_ragged"])
def test_golden_pipeline_fixtures(name):
    import meshgpt_pytorch_b200 as M
    fx = torch.load(os.path.join(GOLDEN_DIR, f"pipeline_{name}.pt"))
    m = M.MeshAutoencoder(**fx["cfg"])
    m.load_reference_state_dict(fx["state_dict"])
    m = m.cuda()
    codes = m.tokenize(fx["vertices"].cuda(), fx["faces"].cuda()).cpu()
    assert codes.shape == fx["codes"].shape and codes.dtype == torch.int64
    pad = (fx["codes"] == -1)
    assert torch.equal(codes == -1, pad)
    # token by token: equal to the frozen codes, or a near tie at the oracle's quantizer input (helpers.py); the oracle rebuilt from
    # the fixture's weights reproduces the frozen codes exactly, so its stages are the fixture's
    o = O.OracleMeshAutoencoder(**fx["cfg"]).eval()
    o.load_state_dict(fx["state_dict"])
    assert torch.equal(o.tokenize(fx["vertices"], fx["faces"]), fx["codes"])
    assert_tokens_match_or_attributed(o, m, fx["vertices"], fx["faces"], codes, min_rate=0.995, label=name)
    enc, dc = m.encode(vertices=fx["vertices"].cuda(), faces=fx["faces"].cuda(), face_edges=fx["stages"]["face_edges"].cuda(),
                       face_mask=(fx["faces"] != -1).all(-1).cuda(), face_edges_mask=(fx["stages"]["face_edges"] != -1).all(-1).cuda(),
                       return_face_coordinates=True)
    fm = (fx["faces"] != -1).all(-1)
    assert torch.equal(dc.cpu()[fm], fx["stages"]["d_coords"][fm])
    assert rel_err(enc, fx["stages"]["encoded"]) < 2e-3


def test_tokenize_variants_agree():
    """batch-less input, user-supplied vs derived face_edges, int32 faces, shuffled user edges."""
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    v, f = synth.ragged_batch("tri", 8, 8, [128, 50, 1], [0, 1, 2])
    base = m.tokenize(v.cuda(), f.cuda())
    assert_tokens_match_or_attributed(o, m, v, f, base, min_rate=0.995, label="variants")
    fe = O.derive_face_edges_from_faces(f)
    with_edges = m.tokenize(v.cuda(), f.cuda(), fe.cuda())
    assert torch.equal(with_edges, base)
    perm = torch.randperm(fe.shape[1], generator=torch.Generator().manual_seed(0))
    shuffled = m.tokenize(v.cuda(), f.cuda(), fe[:, perm].cuda())
    assert match_rate(shuffled, base) > 0.995                      # same graph, different summation order
    assert torch.equal(m.tokenize(v.cuda(), f.int().cuda()), base)
    one = m.tokenize(v[1].cuda(), f[1].cuda())
    assert one.shape == (f.shape[1] * 3, 2) and torch.equal(one, base[1])
    host = m.tokenize_host(v, f)
    assert not host.is_cuda and torch.equal(host, base.cpu())
    # CUDA-graph replay: same codes, also after new data has been copied into the static inputs
    g = m.graphed(v.cuda(), f.cuda())
    assert torch.equal(g(), base)
    v2, f2 = synth.ragged_batch("tri", 8, 8, [77, 128, 9], [7, 8, 9])
    assert torch.equal(g(v2.cuda(), f2.cuda()), m.tokenize(v2.cuda(), f2.cuda()))
    assert torch.equal(g.tokenize_host(v, f), base.cpu())
    bad = f.clone()
    bad[0, 0, 0] = v.shape[1] + 5
    with pytest.raises(IndexError):                       # check before returning (one host sync)
        m.tokenize(v.cuda(), bad.cuda(), check_status=True)
    # the default never syncs: the error of call n surfaces when call n + 1 starts, or at check_last_status()
    m.tokenize(v.cuda(), bad.cuda())
    with pytest.raises(IndexError):
        m.check_last_status()
    m.tokenize(v.cuda(), bad.cuda())
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        m.tokenize(v.cuda(), f.cuda())
    assert torch.equal(m.tokenize(v.cuda(), f.cuda()), base)      # and the model still works afterwards
    m.check_last_status()


def test_histogram_accumulates_codebook_usage():
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    v, f = synth.ragged_batch("tri", 8, 8, [100, 37], [5, 6])
    hist = torch.zeros(2, 256, dtype=torch.int64, device="cuda")
    codes = m.forward(vertices=v.cuda(), faces=f.cuda(), return_codes=True, histogram=hist)
    assert torch.equal(hist.cpu(), O.codebook_usage_histogram(codes.cpu(), 2, 256))
    m.forward(vertices=v.cuda(), faces=f.cuda(), return_codes=True, histogram=hist)
    assert int(hist.sum()) == 2 * int((codes != -1).sum())


def test_full_size_properties_c2_and_c3():
    """BASELINE configs 2 and 3 at full size, through size-independent properties: determinism, independence
    of the meshes of a batch (halves == whole, permutation equivariance), code range, padding."""
    import meshgpt_pytorch_b200 as M
    for kind, cfg, nxny in (("tri", dict(use_residual_lfq=False), (20, 20)), ("quad", dict(quads=True), (20, 40))):
        torch.manual_seed(1)
        m = M.MeshAutoencoder(**cfg).cuda()
        v, f = synth.grid_batch(kind, *nxny, list(range(64)))
        v, f = v.cuda(), f.cuda()
        a = m.tokenize(v, f)
        b = m.tokenize(v, f)
        nvf = f.shape[-1]
        assert a.shape == (64, 800 * nvf, 2) and torch.equal(a, b)
        assert int(a.min()) >= 0 and int(a.max()) < 16384
        halves = torch.cat([m.tokenize(v[:32], f[:32]), m.tokenize(v[32:], f[32:])])
        assert torch.equal(halves, a)
        perm = torch.randperm(64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
        assert torch.equal(m.tokenize(v[perm], f[perm]), a[perm])
        # tokens of one vertex carry one code wherever the vertex appears
        flat_v = f[0].reshape(-1)
        first = {}
        ok = True
        for tok, vid in enumerate(flat_v.tolist()[:600]):
            ok &= first.setdefault(vid, a[0, tok].tolist()) == a[0, tok].tolist()
        assert ok
        fpad = f.clone()
        fpad[:, 500:] = -1
        c = m.tokenize(v, fpad)
        assert bool((c[:, 500 * nvf:] == -1).all()) and bool((c[:, :500 * nvf] >= 0).all())


@pytest.mark.parametrize("kind,cfg,nxny", [("tri", dict(use_residual_lfq=False), (20, 20)), ("quad", dict(quads=True), (20, 40))])
def test_full_model_tokens_against_oracle(kind, cfg, nxny):
    """The full-size model (192-d, 16384 codes, encoder 64-128-256-256-576) on meshes of the BASELINE shape (800 faces),
    token by token against the CPU oracle.  Measured agreement: 100 % (scripts/gpu_match_rates.py); the bar leaves room
    for near-ties only."""
    o, m = make_pair(**cfg)
    v, f = synth.grid_batch(kind, *nxny, [0, 1, 2, 3])
    got = m.tokenize(v.cuda(), f.cuda()).cpu()
    assert_tokens_match_or_attributed(o, m, v, f, got, min_rate=0.999, label=f"{kind} {cfg}")


def test_face_edges_cache_bulk_fill(tmp_path):
    """the GPU adjacency kernel fills the reference's on-disk face-edge cache for a whole batch."""
    import numpy as np
    import meshgpt_pytorch_b200 as M
    _, faces = synth.ragged_batch("quad", 7, 9, [63, 40, 2, 63], [3, 4, 5, 6])
    fe_file, ic_file = M.cache_face_edges(faces.cuda(), tmp_path, dataset_len=4, start=0, max_edges_len=400)
    cache = np.load(fe_file, mmap_mode="r")
    assert bool(np.load(ic_file, mmap_mode="r").all()) and cache.dtype == np.float32
    for i in range(4):
        row = torch.from_numpy(np.array(cache[i]))
        want = O.derive_face_edges_from_faces(faces[i])
        assert torch.equal(row[(row != -1).all(-1)].long(), want[(want != -1).all(-1)])


def test_host_pipeline_matches_single_calls():
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    batches = [synth.ragged_batch("tri", 8, 8, [128, 50 + 9 * i, 1 + i], [i, i + 1, i + 2]) for i in range(5)]   # one padded shape
    pipe = m.host_pipeline(batches[0][0].cuda(), batches[0][1].cuda())
    got = {}
    for v, f in batches:
        done = pipe.submit(v.pin_memory(), f.pin_memory())
        if done is not None:
            got[done[0]] = done[1].clone()
    for ticket, codes in pipe.drain():
        got[ticket] = codes.clone()
    assert sorted(got) == list(range(5))
    for i, (v, f) in enumerate(batches):
        assert torch.equal(got[i], m.tokenize(v.cuda(), f.cuda()).cpu()), i
    # three slots in flight, the codebook-usage histogram accumulated by every batch, and a bad batch raising when it is handed back
    hist = torch.zeros(2, 256, dtype=torch.int64, device="cuda")
    pipe3 = m.host_pipeline(batches[0][0].cuda(), batches[0][1].cuda(), depth=3, histogram=hist)
    outs = []
    for v, f in batches:
        done = pipe3.submit(v.pin_memory(), f.pin_memory())
        if done is not None:
            outs.append(done[1].clone())
    outs += [c.clone() for _, c in pipe3.drain()]
    assert all(torch.equal(a, got[i]) for i, a in enumerate(outs))
    assert torch.equal(hist.cpu(), sum(O.codebook_usage_histogram(c, 2, 256) for c in outs))
    bad = batches[0][1].clone()
    bad[0, 0, 0] = batches[0][0].shape[1] + 3
    pipe3.submit(batches[0][0].pin_memory(), bad.pin_memory())
    with pytest.raises(IndexError):
        pipe3.drain()
    pipe3.submit(batches[1][0].pin_memory(), batches[1][1].pin_memory())
    assert torch.equal(pipe3.drain()[0][1], got[1])                # the pipeline keeps working after a bad batch
    pipe3.close()


def test_graphed_with_precomputed_edges():
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    v, f = synth.ragged_batch("tri", 8, 8, [128, 50, 1], [0, 1, 2])
    fe = O.derive_face_edges_from_faces(f)
    g = m.graphed(v.cuda(), f.cuda(), face_edges=fe.cuda())
    assert torch.equal(g(), m.tokenize(v.cuda(), f.cuda()))
    perm = torch.randperm(fe.shape[1], generator=torch.Generator().manual_seed(1))
    assert match_rate(g(face_edges=fe[:, perm].cuda()), m.tokenize(v.cuda(), f.cuda())) > 0.995


def test_dense_user_edges_through_the_aggregation_kernels():
    """User-supplied face_edges with far more edges per face than the derived default (neighbours by ONE shared vertex,
    about 13 per face): the shared-memory aggregation kernel's staging area for source ids overflows and it reads them
    from global memory; duplicates count twice as in the reference."""
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    v, f = synth.ragged_batch("tri", 12, 12, [288, 288, 100, 288], [0, 1, 2, 3])
    fe = O.derive_face_edges_from_faces(f, neighbor_if_share_one_vertex=True)
    fe = torch.cat([fe, fe[:, :50]], dim=1)          # some edges twice
    got = m.tokenize(v.cuda(), f.cuda(), face_edges=fe.cuda()).cpu()
    assert_tokens_match_or_attributed(o, m, v, f, got, face_edges=fe, min_rate=0.995, label="dense user edges")


@pytest.mark.parametrize("cfg,kind,grid,counts", [
    (dict(SMALL_CFG, use_residual_lfq=False, quads=True), "quad", (6, 7), [42, 17, 3]),              # quads + RVQ (CTA-pair search)
    (dict(SMALL_CFG, use_residual_lfq=False, codebook_size=384), "tri", (8, 8), [128, 1]),          # K % 256 != 0 -> single-CTA search
    (dict(SMALL_CFG, use_residual_lfq=False, codebook_size=200), "tri", (8, 8), [90, 60]),          # no tcgen05 image -> exact fp32 scan
    (dict(SMALL_CFG, num_quantizers=3, codebook_size=512), "tri", (8, 8), [128, 128, 5, 77]),       # LFQ, 3 levels, 9 bits
    (dict(dim_codebook=192, encoder_dims_through_depth=(64, 128), codebook_size=1024, use_residual_lfq=False), "tri", (10, 10), [200, 150]),
    (dict(SMALL_CFG, use_residual_lfq=False), "tri", (45, 20), [1800, 900]),      # meshes too large for a shared-memory slice -> L2 aggregation kernel
    (dict(SMALL_CFG, num_discrete_coors=100, num_discrete_angle=64, num_discrete_area=32, num_discrete_normals=96), "tri", (8, 8), [128, 77]),   # unequal table sizes
    (dict(SMALL_CFG, quads=True, num_discrete_coors=72, num_discrete_angle=128, num_discrete_area=16, num_discrete_normals=40), "quad", (6, 7), [42, 30]),
])
def test_config_matrix_end_to_end(cfg, kind, grid, counts):
    """Other constructor configurations end to end against the oracle (token agreement; exact pad positions)."""
    o, m = make_pair(**cfg)
    v, f = synth.ragged_batch(kind, grid[0], grid[1], counts, list(range(len(counts))))
    got = m.tokenize(v.cuda(), f.cuda()).cpu()
    assert_tokens_match_or_attributed(o, m, v, f, got, min_rate=0.995, label=str(cfg))
    # vertices nobody references and a lone single-face mesh do not disturb anything
    v2 = torch.cat([v, torch.rand(v.shape[0], 7, 3) * 2 - 1], dim=1)
    assert torch.equal(m.tokenize(v2.cuda(), f.cuda()).cpu(), got)
    assert torch.equal(m.tokenize(v[-1:].cuda(), f[-1:].cuda()).cpu(), got[-1:])


_FALLBACK_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from helpers import make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
out = {{}}
for name, cfg, kind, nxny, n in (("rvq", dict(use_residual_lfq=False), "tri", (20, 20), 3), ("lfq_quads", dict(quads=True), "quad", (10, 12), 5)):
    o, m = make_pair(**cfg)
    v, f = synth.grid_batch(kind, *nxny, list(range(n)))
    out[name] = m.tokenize(v.cuda(), f.cuda()).cpu()
torch.save(out, sys.argv[1])
"""


@pytest.mark.parametrize("knob", ["MESHTOK_FACE_SUM_L2=1", "MESHTOK_GATHER_L2=1", "MESHTOK_LINEAR_TC=1", "MESHTOK_LINEAR_B2B=0", "MESHTOK_RVQ_TC=1",
                                  "MESHTOK_GATHER_WIDE=0", "MESHTOK_GATHER_WIDE_GROUP=3", "MESHTOK_GATHER_WIDE_GROUP=1",
                                  "MESHTOK_RESCORE4=0", "MESHTOK_FACE_SUM_PAIR=0", "MESHTOK_PDL=0", "MESHTOK_GATHER_FUSED=1",
                                  "MESHTOK_GATHER_LCSR=0", "MESHTOK_GATHER_SPREAD=0", "MESHTOK_LFQ_ROWS=0", "MESHTOK_LFQ_VARIANT=21",
                                  "MESHTOK_TOPO_ONE=0", "MESHTOK_EDGES_EARLY=1", "MESHTOK_TOPO_CACHE=0"])
def test_fallback_kernels_agree_with_the_default_path(knob, tmp_path):
    """Every kernel that has a fallback (tables / neighbour rows from L2, single-CTA linear and RVQ search, one linear per
    launch) computes the same tokens as the default path.  The switches are read once per process, hence subprocesses."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "run.py"
    script.write_text(_FALLBACK_SCRIPT.format(root=root, tests=os.path.join(root, "tests")))
    outs = {}
    for tag, extra in (("default", {}), ("knob", dict([knob.split("=")]))):
        env = dict(os.environ, **extra)
        path = tmp_path / f"{tag}.pt"
        r = subprocess.run([sys.executable, str(script), str(path)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = torch.load(path)
    for name in outs["default"]:
        a, b = outs["default"][name], outs["knob"][name]
        assert a.shape == b.shape
        # same arithmetic in a different association order at most: tokens may differ only at near-ties
        assert match_rate(a, b) >= 0.999, (knob, name, match_rate(a, b))
