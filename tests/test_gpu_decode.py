"""GPU: the decode path (meshgpt_pytorch_b200/decoder.py + csrc/decode.cu) against the CPU decode oracle
(oracle/decode_oracle.py, pinned to the reference's own decoder blocks by tests/test_decode_oracle.py), through the C ABI."""
import pytest
import torch

from helpers import SMALL_CFG, make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
from oracle import decode_oracle as D

pytestmark = pytest.mark.gpu


def make_decoders(cfg, dims, kernel=7, seed=5):
    import meshgpt_pytorch_b200 as M
    o, m = make_pair(**cfg)
    full = dict(SMALL_CFG, **cfg) if "dim_codebook" not in cfg else cfg
    torch.manual_seed(seed)
    od = D.OracleDecoder(o.quantizer, dim_codebook=o.dim_codebook, num_quantizers=o.num_quantizers, codebook_size=o.codebook_size,
                         use_residual_lfq=o.use_residual_lfq, decoder_dims_through_depth=dims, init_decoder_conv_kernel=kernel,
                         num_discrete_coors=o.num_discrete_coors, quads=o.nvf == 4).eval()
    g = torch.Generator().manual_seed(seed)
    for name, p_ in od.named_parameters():             # every term counts (default inits give tiny biases)
        if name.startswith("quantizer."):
            continue                                   # shared with the tokenizer pair
        if p_.ndim > 1:
            p_.data = torch.randn(p_.shape, generator=g) / (p_[0].numel() ** 0.5)
        else:
            p_.data = torch.randn(p_.shape, generator=g) * 0.2
    md = M.MeshDecoder(m, init_decoder_conv_kernel=kernel, decoder_dims_through_depth=dims)
    md.load_reference_state_dict({k: v for k, v in od.state_dict().items() if not k.startswith("quantizer.")})
    return o, m, od, md.cuda()


@pytest.mark.parametrize("cfg,dims,kind,grid,counts", [
    (dict(SMALL_CFG, use_residual_lfq=False), (32, 32, 64, 64), "tri", (8, 8), [128, 60, 1]),
    (dict(SMALL_CFG), (32, 64), "tri", (6, 5), [60, 17]),
    (dict(SMALL_CFG, quads=True, use_residual_lfq=False), (64, 96, 96), "quad", (6, 7), [42, 30]),
])
def test_decode_stages_and_coordinates(cfg, dims, kind, grid, counts):
    o, m, od, md = make_decoders(cfg, dims)
    v, f = synth.ragged_batch(kind, grid[0], grid[1], counts, list(range(len(counts))))
    codes = o.tokenize(v, f)                                   # (b, nf * nvf, Q), -1 on padded faces
    b, nvf = codes.shape[0], o.nvf
    mask = (codes != -1).reshape(b, -1, nvf * o.num_quantizers).all(-1)
    # de-quantisation
    want_q = od.dequantize(codes.reshape(b, -1, o.num_quantizers))
    got_q = md.dequantize(codes.cuda().reshape(b, -1, o.num_quantizers)).cpu()
    assert torch.allclose(got_q, want_q, rtol=1e-5, atol=1e-6)
    # decoder stack on the oracle's input
    qf = want_q.reshape(b, -1, nvf * o.dim_codebook)
    want_d = od.decode(qf, mask)
    got_d = md.decode(qf.cuda(), mask.cuda()).cpu()
    assert got_d.shape == want_d.shape
    err = (got_d - want_d)[mask].abs().max() / want_d[mask].abs().max()
    assert float(err) < 2e-3, float(err)                        # rtol 1e-3 class: split-bf16 linears through 2 * len(dims) convolutions
    assert float(got_d[~mask].abs().max() if (~mask).any() else 0.0) == 0.0
    # end to end: coordinates
    want_c, want_bins, want_mask = od.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    got_c, got_bins, got_mask = md.decode_from_codes_to_faces(codes.cuda(), return_discrete_codes=True)
    assert torch.equal(got_mask.cpu(), want_mask)
    assert torch.isnan(got_c.cpu()[~want_mask]).all() and not torch.isnan(got_c.cpu()[want_mask]).any()
    same = (got_bins.cpu()[want_mask] == want_bins[want_mask]).float().mean()
    assert float(same) >= 0.99, float(same)                     # arg-max flips only at near-ties of two logits
    eq = got_bins.cpu()[want_mask] == want_bins[want_mask]
    assert torch.equal(got_c.cpu()[want_mask][eq], want_c[want_mask][eq])      # same bin -> the same fp32 coordinate, bit for bit


def test_decode_full_size_decoder_round_trip():
    """The reference's default decoder (17 widths, kernel 7) on 800-face meshes: tokenize on the GPU, decode on the GPU, and
    compare the decoded bins with the CPU oracle's on the same codes."""
    dims = (128, 128, 128, 128, 192, 192, 192, 192, 256, 256, 256, 256, 256, 256, 384, 384, 384)
    o, m, od, md = make_decoders(dict(use_residual_lfq=False), dims)
    v, f = synth.grid_batch("tri", 20, 20, [0, 1])
    codes = m.tokenize(v.cuda(), f.cuda())
    want_c, want_bins, want_mask = od.decode_from_codes_to_faces(codes.cpu(), return_discrete_codes=True)
    got_c, got_bins, got_mask = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    assert got_c.shape == (2, 800, 3, 3) and bool(got_mask.all())
    same = (got_bins.cpu() == want_bins).float().mean()
    assert float(same) >= 0.98, float(same)


def _halo_bytes(R, c, halo):
    import ctypes as C
    from meshgpt_pytorch_b200._lib import lib, check
    n = C.c_size_t()
    check(lib.meshtok_halo_image_bytes(R, c, halo, C.byref(n)))
    return n.value


def _padded(mask, lead):
    """face mask (b, nf) -> (valid bytes of the padded row space, P, R): rows b * P + lead + n, `lead` separators after the last mesh."""
    b, nf = mask.shape
    P = nf + lead
    R = b * P + lead
    valid = torch.zeros(R, dtype=torch.uint8, device=mask.device)
    valid[:b * P].view(b, P)[:, lead:] = mask
    return valid, P, R


def _halo_rows(img, R, c, halo):
    """Read a halo image back: (hi + lo as fp32 of every row < R from the row's own tile, the same rows from the halo copies in the
    neighbouring tiles with NaN where a row has no copy)."""
    rows = 128 + 2 * halo
    tiles = (R + 127) // 128
    blocks = img.view(torch.bfloat16).reshape(tiles, c // 32, 2, 4, rows, 8).float()      # tile, chunk, hi/lo, column group, row, 8 columns
    val = blocks.permute(0, 2, 4, 1, 3, 5).reshape(tiles, 2, rows, c)                       # tile, hi/lo, row, column
    own = val[:, :, halo:halo + 128].permute(1, 0, 2, 3).reshape(2, tiles * 128, c)[:, :R]
    copies = torch.full((2, 2, tiles * 128, c), float("nan"), device=img.device)           # [before / after copy][hi / lo]
    if halo and tiles > 1:
        v = val.permute(1, 0, 2, 3)                                                         # hi/lo, tile, row, column
        idx = torch.arange(tiles * 128, device=img.device).view(tiles, 128)
        copies[0][:, idx[:-1, 128 - halo:].reshape(-1)] = v[:, 1:, :halo].reshape(2, -1, c)       # last rows of tile t in front of tile t + 1
        copies[1][:, idx[1:, :halo].reshape(-1)] = v[:, :-1, halo + 128:].reshape(2, -1, c)       # first rows of tile t + 1 behind tile t
    return own, copies[:, :, :R]


def _split(x):
    hi = x.to(torch.bfloat16).float()
    return torch.stack([hi, (x - hi).to(torch.bfloat16).float()])


@pytest.mark.parametrize("halo", [0, 1, 3])
def test_halo_image_writers_are_bit_exact(halo):
    """rows_halo and dequantize_rvq_halo against a torch restatement of the layout (tc_common.cuh): hi = bf16(x), lo = bf16(x - hi), rows
    of the padded row space, zero rows on separators and masked faces, the first / last `halo` rows of a tile stored again in the
    neighbouring tile (ragged meshes, masked faces inside a mesh, a row count that is not a multiple of 8 or 128)."""
    from meshgpt_pytorch_b200._lib import lib, check, ptr, stream_ptr
    torch.manual_seed(halo)
    b, nf, c, lead = 3, 125, 96, 3
    mask = torch.rand(b, nf, device="cuda") > 0.2
    mask[0, nf - 5:] = False
    valid, P, R = _padded(mask, lead)
    x = torch.randn(b * nf, c, device="cuda")
    want_rows = torch.zeros(R, c, device="cuda")
    want_rows[:b * P].view(b, P, c)[:, lead:] = x.view(b, nf, c) * mask[..., None]
    want = _split(want_rows)

    def check_image(got):
        own, copies = _halo_rows(got, R, c, halo)
        assert torch.equal(own, want)
        r = torch.arange(R, device="cuda")
        tiles = (R + 127) // 128
        expect = ((r % 128 >= 128 - halo) & (r // 128 + 1 < tiles), (r % 128 < halo) & (r // 128 > 0))
        for side in range(2):
            assert torch.equal(~torch.isnan(copies[side][0, :, 0]), expect[side])
            assert torch.equal(copies[side][:, expect[side]], want[:, expect[side]])

    n = _halo_bytes(R, c, halo)
    got = torch.full((n,), 0x7F, dtype=torch.uint8, device="cuda")          # 0x7F7F: a bf16 NaN wherever nothing is written
    check(lib.meshtok_dec_rows_halo(ptr(x), c, ptr(valid), R, P, lead, c, halo, ptr(got), stream_ptr()))
    check_image(got)
    # a strided source: the second half of a (rows, 2 c) buffer
    wide = torch.randn(b * nf, 2 * c, device="cuda")
    wide[:, c:] = x
    got.fill_(0x7F)
    check(lib.meshtok_dec_rows_halo(wide.data_ptr() + 4 * c, 2 * c, ptr(valid), R, P, lead, c, halo, ptr(got), stream_ptr()))
    check_image(got)
    # codes -> image
    nvf, q, d, ksz = 3, 2, 32, 50
    book = torch.randn(ksz, d, device="cuda")
    codes = torch.randint(0, ksz, (b * nf * nvf, q), device="cuda")
    codes[(~mask).reshape(-1).repeat_interleave(nvf)] = -1
    rows = torch.empty(b * nf * nvf, d, device="cuda")
    check(lib.meshtok_dec_dequantize_rvq(ptr(codes), ptr(book), b * nf * nvf, q, d, ptr(rows), stream_ptr()))
    want_rows.zero_()
    want_rows[:b * P].view(b, P, c)[:, lead:] = rows.view(b, nf, c) * mask[..., None]
    want = _split(want_rows)
    got.fill_(0x7F)
    check(lib.meshtok_dec_dequantize_rvq_halo(ptr(codes), ptr(book), ptr(valid), R, P, lead, nvf, q, d, halo, ptr(got), stream_ptr()))
    check_image(got)


@pytest.mark.parametrize("k,halo,c_in,n", [(1, 0, 64, 48), (3, 1, 96, 160), (7, 3, 32, 64), (3, 3, 64, 384), (1, 1, 32, 1152)])
def test_conv1d_on_a_halo_image_matches_torch(k, halo, c_in, n):
    """meshtok_conv1d (tcgen05, tap t = the same shared-memory block read t rows further) against torch's Conv1d in fp64 on every mesh
    separately: zero padding at the ends of a mesh, meshes that straddle 128-row tiles, an image whose halo is wider than the kernel."""
    import ctypes as C
    from meshgpt_pytorch_b200._lib import lib, check, ptr, stream_ptr
    torch.manual_seed(k * 100 + n)
    b, nf, lead = 5, 171, 3
    mask = torch.ones(b, nf, dtype=torch.bool, device="cuda")
    mask[1, 150:] = False
    mask[4, 1:] = False
    valid, P, R = _padded(mask, lead)
    x = torch.randn(b * nf, c_in, device="cuda")
    w = torch.randn(n, c_in, k, device="cuda") / (c_in * k) ** 0.5
    bias = torch.randn(n, device="cuda")
    img = torch.full((_halo_bytes(R, c_in, halo),), 0x7F, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_dec_rows_halo(ptr(x), c_in, ptr(valid), R, P, lead, c_in, halo, ptr(img), stream_ptr()))
    mat = w.permute(0, 2, 1).reshape(n, k * c_in).contiguous()
    nb = C.c_size_t()
    check(lib.meshtok_linear_image_bytes(n, k * c_in, C.byref(nb)))
    wimg = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_linear_build_image(ptr(mat), n, k * c_in, ptr(wimg), stream_ptr()))
    out = torch.empty(R, n, device="cuda")
    check(lib.meshtok_conv1d(ptr(img), ptr(wimg), ptr(bias), ptr(out), R, n, c_in, k, halo, stream_ptr()))
    got = out[:b * P].view(b, P, n)[:, lead:]
    xm = (x.view(b, nf, c_in) * mask[..., None]).double().cpu()
    want = torch.nn.functional.conv1d(xm.transpose(1, 2), w.double().cpu(), bias.double().cpu(), padding=k // 2).transpose(1, 2)
    m = mask.cpu()
    err = (got.cpu().double() - want)[m].abs().max() / want[m].abs().max()
    assert float(err) < 2e-5, float(err)                        # split bf16: 3 products, ~2^-16 relative


@pytest.mark.parametrize("cfg,dims,kind", [
    (dict(SMALL_CFG, use_residual_lfq=False), (32, 32, 64, 64, 96), "tri"),
    (dict(SMALL_CFG, quads=True), (64, 32), "quad"),
])
def test_fused_schedule_matches_the_staged_one(cfg, dims, kind, monkeypatch):
    """Same decoder, same codes: the fused schedule (producers write the next image, residual_conv inside block1's linear, PixelNorm
    folded into the SqueezeExcite mean and the block tail) against the stage-by-stage schedule.  Only summation orders differ."""
    o, m, od, md = make_decoders(cfg, dims)
    v, f = synth.ragged_batch(kind, 9, 8, [130, 61, 8, 1] if kind == "tri" else [72, 30, 5, 1], [0, 1, 2, 3])
    codes = m.tokenize(v.cuda(), f.cuda())
    assert md.fused()
    c1, b1, m1 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    q = md.dequantize(codes.reshape(codes.shape[0], -1, o.num_quantizers)).reshape(codes.shape[0], -1, o.nvf * o.dim_codebook)
    d1 = md.decode(q, m1)
    monkeypatch.setenv("MESHTOK_DEC_FUSED", "0")
    assert not md.fused()
    c0, b0, m0 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    d0 = md.decode(q, m0)
    assert torch.equal(m0, m1)
    assert float((d1 - d0).abs().max() / d0.abs().max()) < 1e-4
    assert float((d1[~m1]).abs().max()) == 0.0
    assert float((b1[m1] == b0[m1]).float().mean()) >= 0.995
    assert torch.isnan(c1[~m1]).all()


@pytest.mark.parametrize("dims", [(32, 32, 64), (64, 160, 160), (128, 192, 256, 256)])
def test_block_tail_in_the_convolution_epilogue_matches_the_separate_launch(dims, monkeypatch):
    """meshtok_conv1d_pixelnorm_silu (PixelNorm + SiLU + mask in the epilogue of block1.proj's convolution; with a residual_conv the
    block columns are pass 0 of two) against meshtok_conv1d + meshtok_dec_pixelnorm_silu_halo, and both against the oracle.
    dims: one pass without residual_conv (32 -> 32), two passes (64 -> 160: N = 320 = 2 x 160; 128 -> 192, 192 -> 256), and blocks the
    fused epilogue does not cover (32 -> 64: N = 128 is one pass of both halves) in the same decoder."""
    cfg = dict(SMALL_CFG, use_residual_lfq=False)
    o, m, od, md = make_decoders(cfg, dims)
    widths = list(zip((dims[0],) + tuple(dims[:-1]), dims))                # (c_in, c_out) of every ResnetBlock
    assert sum(md._tail_in_epilogue(c if ci == c else 2 * c, c) for ci, c in widths) >= 2
    v, f = synth.ragged_batch("tri", 12, 12, [288, 131, 127, 1], [0, 1, 2, 3])
    codes = m.tokenize(v.cuda(), f.cuda())
    monkeypatch.setenv("MESHTOK_DEC_EPI", "1")
    c1, b1, m1 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    monkeypatch.setenv("MESHTOK_DEC_EPI", "0")
    c0, b0, m0 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    assert torch.equal(m0, m1)
    assert float((b1[m1] == b0[m1]).float().mean()) >= 0.995
    assert torch.isnan(c1[~m1]).all()
    want_c, want_bins, want_mask = od.decode_from_codes_to_faces(codes.cpu(), return_discrete_codes=True)
    assert torch.equal(m1.cpu(), want_mask)
    assert float((b1.cpu()[want_mask] == want_bins[want_mask]).float().mean()) >= 0.99


@pytest.mark.parametrize("cfg,dims,kind,counts", [
    (dict(SMALL_CFG, use_residual_lfq=False), (32, 64, 96), "tri", [288, 131, 127, 1]),      # 9 coordinates: 5 passes, the last one half padding
    (dict(SMALL_CFG, quads=True), (64, 32), "quad", [72, 30, 5, 1]),                          # 12 coordinates: 6 full passes
])
def test_argmax_in_the_logits_epilogue_matches_the_separate_launch(cfg, dims, kind, counts, monkeypatch):
    """meshtok_conv1d_argmax_coords (to_coor_logits with the arg-max + undiscretize in the tcgen05 epilogue, logits never stored) against
    meshtok_conv1d + meshtok_dec_argmax_coords_padded: the same accumulators, so bins and coordinates must be identical; and the oracle."""
    o, m, od, md = make_decoders(cfg, dims)
    assert md._prepare()["logits_am"] is not None
    v, f = synth.ragged_batch(kind, 12, 12, counts, [0, 1, 2, 3])
    codes = m.tokenize(v.cuda(), f.cuda())
    monkeypatch.setenv("MESHTOK_DEC_ARGMAX", "1")
    c1, b1, m1 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    monkeypatch.setenv("MESHTOK_DEC_ARGMAX", "0")
    c0, b0, m0 = md.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    assert torch.equal(m0, m1) and torch.equal(b1, b0)
    assert torch.equal(torch.nan_to_num(c1, nan=-7.0), torch.nan_to_num(c0, nan=-7.0)) and torch.isnan(c1[~m1]).all()
    want_c, want_bins, want_mask = od.decode_from_codes_to_faces(codes.cpu(), return_discrete_codes=True)
    assert float((b1.cpu()[want_mask] == want_bins[want_mask]).float().mean()) >= 0.99
