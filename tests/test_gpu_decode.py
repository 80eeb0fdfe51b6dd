"""GPU: the decode path (meshgpt_pytorch_b200/decoder.py + csrc/decode.cu) against the CPU decode oracle
(oracle/decode_oracle.py, pinned to the reference's own decoder blocks by tests/test_decode_oracle.py), through the C ABI."""
import pytest
import torch

from helpers import SMALL_CFG, make_pair
from meshgpt_pytorch_b200 import synth
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


def _image_bytes(R, K):
    import ctypes as C
    from meshgpt_pytorch_b200._lib import lib, check
    n = C.c_size_t()
    check(lib.meshtok_rows_image_bytes(R, K, C.byref(n)))
    return n.value


def _live_image(img, R, K):
    """The bytes of an activation image that belong to rows < R (pad rows of the last 128-row tile are never written)."""
    full = (R // 128) * (K // 32) * 16384
    head = img[:full]
    if R % 128 == 0:
        return head
    tail = img[full:].reshape(K // 32, 2, 16, 4, 8, 16)[:, :, :, :, :, :]       # chunk, hi/lo, 8-row group, k group, row, 16 B
    rows = (torch.arange(16, device=img.device)[:, None] * 8 + torch.arange(8, device=img.device)[None, :]) < (R % 128)
    keep = rows[None, None, :, None, :, None].expand_as(tail)
    return torch.cat([head, tail[keep]])


@pytest.mark.parametrize("k", [1, 3, 7])
def test_fused_image_writers_are_bit_exact_with_the_im2col_pass(k):
    """rows_im2col and dequantize_rvq_im2col write, byte for byte, the image the stand-alone im2col pass builds from the same rows
    (ragged meshes, masked faces in the middle of a mesh, a row count that is not a multiple of 8 or 128)."""
    from meshgpt_pytorch_b200._lib import lib, check, ptr, stream_ptr
    torch.manual_seed(k)
    b, nf, c = 3, 77, 96
    R = b * nf
    x = torch.randn(R, c, device="cuda")
    valid = (torch.rand(R, device="cuda") > 0.2).to(torch.uint8)
    valid[nf - 5:nf] = 0
    n = _image_bytes(R, k * c)
    want = torch.zeros(n, dtype=torch.uint8, device="cuda")
    got = torch.full((n,), 0xAB, dtype=torch.uint8, device="cuda")
    check(lib.meshtok_dec_im2col_image(ptr(x), ptr(valid), R, nf, c, k, ptr(want), stream_ptr()))
    check(lib.meshtok_dec_rows_im2col(ptr(x), c, ptr(valid), R, nf, c, k, ptr(got), stream_ptr()))
    assert torch.equal(_live_image(got, R, k * c), _live_image(want, R, k * c))
    # a strided source: the second half of a (R, 2 c) buffer
    wide = torch.randn(R, 2 * c, device="cuda")
    wide[:, c:] = x
    got.fill_(0xCD)
    check(lib.meshtok_dec_rows_im2col(wide.data_ptr() + 4 * c, 2 * c, ptr(valid), R, nf, c, k, ptr(got), stream_ptr()))
    assert torch.equal(_live_image(got, R, k * c), _live_image(want, R, k * c))
    # codes -> image
    nvf, q, d, ksz = 3, 2, 32, 50
    book = torch.randn(ksz, d, device="cuda")
    codes = torch.randint(0, ksz, (R * nvf, q), device="cuda")
    codes[(valid == 0).repeat_interleave(nvf)] = -1
    rows = torch.empty(R * nvf, d, device="cuda")
    check(lib.meshtok_dec_dequantize_rvq(ptr(codes), ptr(book), R * nvf, q, d, ptr(rows), stream_ptr()))
    check(lib.meshtok_dec_im2col_image(ptr(rows), ptr(valid), R, nf, nvf * d, k, ptr(want), stream_ptr()))
    got.fill_(0xEF)
    check(lib.meshtok_dec_dequantize_rvq_im2col(ptr(codes), ptr(book), ptr(valid), R, nf, nvf, q, d, k, ptr(got), stream_ptr()))
    assert torch.equal(_live_image(got, R, k * c), _live_image(want, R, k * c))


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
