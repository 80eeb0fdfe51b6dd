"""Decode half of MeshAutoencoder on the B200: codes -> face coordinates.

Mirrors the reference's `MeshAutoencoder.decode_from_codes_to_faces` (meshgpt_pytorch.py:898-947) and `decode` (:872-896,
attn_decoder_depth == 0) with the decoder modules under their reference names (init_decoder_conv.{0,3}, decoders.{i}.
block{1,2}.proj, .excite.net.{0,2}, .residual_conv, to_coor_logits.0), so a reference state dict loads key for key.
Every Conv1d / Linear runs as a split-bf16 tcgen05 linear (csrc/linear_tc.cu) on an im2col activation image; the stages in
between are the kernels of csrc/decode.cu, all called through the C ABI (include/meshtok.h, meshtok_dec_*).  PyTorch holds the
device buffers and nothing else; there is no CPU path.

Default ("fused") schedule, when every width is a multiple of 32: NO im2col.  Activations live in a padded row space (separator
rows between the meshes, zero in every image) and each stage writes its output once, as the "halo image" of the convolution that
follows it (de-quantise -> image of the k=7 convolution; SiLU+LayerNorm -> x + image; PixelNorm+SiLU -> image only; PixelNorm +
SqueezeExcite mean without storing the activation; gate * h + res -> x + image); the tcgen05 kernel reads tap t of a kernel from
the same shared-memory block, one row further (meshtok_conv1d).  A block's residual_conv (k=1) rides in the same launch as
block1.proj (its weight sits in the centre tap of a (2 C_out, 3 C_in) matrix).  6 launches per ResnetBlock instead of 10-12.  MESHTOK_DEC_FUSED=0, or a width that is not a multiple of 32, takes the stage-by-stage schedule (same kernels as the
first version; kept for A/B runs and odd widths).

    dec = MeshDecoder(autoencoder)                    # same decoder keywords as the reference constructor
    dec.load_reference_state_dict(reference.state_dict())
    coords, face_mask = dec.cuda().decode_from_codes_to_faces(codes)      # (b, nf, nvf, 3) fp32, NaN on padded faces
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class _Block(nn.Module):                      # meshgpt_pytorch.py:326-353 (parameters only)
    def __init__(self, dim, dim_out):
        super().__init__()
        self.proj = nn.Conv1d(dim, dim_out, 3, padding=1)


class _Excite(nn.Module):                     # :296-324
    def __init__(self, dim, reduction_factor=4, min_dim=16):
        super().__init__()
        inner = max(dim // reduction_factor, min_dim)
        self.net = nn.Sequential(nn.Linear(dim, inner), nn.SiLU(), nn.Linear(inner, dim), nn.Sigmoid())


class _ResnetBlock(nn.Module):                # :355-379
    def __init__(self, dim, dim_out):
        super().__init__()
        self.block1 = _Block(dim, dim_out)
        self.block2 = _Block(dim_out, dim_out)
        self.excite = _Excite(dim_out)
        self.residual_conv = nn.Conv1d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


conv_trace = None      # set to a list by scripts/gpu_decode_bench.py: (rows, N, K, start event, end event) of every convolution launch


class _conv_timer:
    """CUDA events around one convolution launch while `conv_trace` is a list (measurement only; a no-op otherwise)."""

    def __init__(self, rows: int, n: int, k: int):
        self.item = (rows, n, k)

    def __enter__(self):
        if conv_trace is not None:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if conv_trace is not None:
            self.b.record()
            conv_trace.append(self.item + (self.a, self.b))
        return False


class MeshDecoder(nn.Module):
    def __init__(self, autoencoder, init_decoder_conv_kernel: int = 7,
                 decoder_dims_through_depth: Tuple[int, ...] = (128, 128, 128, 128, 192, 192, 192, 192, 256, 256, 256, 256, 256, 256, 384, 384, 384),
                 attn_decoder_depth: int = 0):
        super().__init__()
        if attn_decoder_depth != 0:
            raise NotImplementedError("attn_decoder_depth > 0 is out of scope for the B200 decode path")
        if init_decoder_conv_kernel % 2 != 1:
            raise ValueError("init_decoder_conv_kernel must be odd")
        object.__setattr__(self, "autoencoder", autoencoder)      # not a sub-module: its keys are not this module's
        a = autoencoder
        self.nvf, self.dim_codebook = a.num_vertices_per_face, a.dim_codebook
        self.num_quantizers, self.codebook_size, self.pad_id = a.num_quantizers, a.codebook_size, a.pad_id
        self.num_discrete_coors, self.coor_range = a.num_discrete_coors, a.coor_continuous_range
        init_dim, *rest = decoder_dims_through_depth
        self.init_decoder_conv = nn.ModuleDict({
            "0": nn.Conv1d(self.dim_codebook * self.nvf, init_dim, init_decoder_conv_kernel, padding=init_decoder_conv_kernel // 2),
            "3": nn.LayerNorm(init_dim)})
        self.decoders = nn.ModuleList()
        cur = init_dim
        for d in rest:
            self.decoders.append(_ResnetBlock(cur, d))
            cur = d
        self.to_coor_logits = nn.ModuleDict({"0": nn.Linear(cur, self.num_discrete_coors * self.nvf * 3)})
        for p in self.parameters():
            p.requires_grad_(False)
        self._prep = None
        self._prep_key = None

    def load_reference_state_dict(self, state_dict: Dict[str, Tensor]):
        own = self.state_dict()
        missing = [k for k in own if k not in state_dict]
        if missing:
            raise KeyError(f"state dict lacks decoder keys: {missing[:8]}{'...' if len(missing) > 8 else ''}")
        self.load_state_dict({k: state_dict[k] for k in own}, strict=True)
        self._prep = None

    # ---- weights -> images ------------------------------------------------------------------------
    def _prepare(self):
        key = tuple((t.data_ptr(), t._version) for t in self.parameters())
        if self._prep is not None and self._prep_key == key:
            return self._prep
        dev = self.to_coor_logits["0"].weight.device
        if dev.type != "cuda":
            raise RuntimeError("MeshDecoder must live on a CUDA device (call .cuda()); there is no CPU path")
        f32 = dict(dtype=torch.float32, device=dev)

        def conv(cv: nn.Conv1d):
            w = cv.weight.detach().to(**f32)                       # (co, ci, k)
            co, ci, k = w.shape
            mat = w.permute(0, 2, 1).reshape(co, k * ci).contiguous()   # tap-major columns: t * ci + c, like the im2col image
            nbytes = C.c_size_t()
            check(lib.meshtok_linear_image_bytes(co, k * ci, C.byref(nbytes)))
            img = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            check(lib.meshtok_linear_build_image(ptr(mat), co, k * ci, ptr(img), stream_ptr()))
            return dict(img=img, bias=cv.bias.detach().to(**f32).contiguous(), n=co, c=ci, k=k, mat=mat)

        def image_of(mat, bias, ci, k):
            co = mat.shape[0]
            nbytes = C.c_size_t()
            check(lib.meshtok_linear_image_bytes(co, mat.shape[1], C.byref(nbytes)))
            img = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            check(lib.meshtok_linear_build_image(ptr(mat), co, mat.shape[1], ptr(img), stream_ptr()))
            return dict(img=img, bias=bias.contiguous(), n=co, c=ci, k=k, mat=mat)

        def conv_with_residual(c1: dict, res: dict):
            """block1.proj (C_out, 3 C_in) and residual_conv (C_out, C_in) as one (2 C_out, 3 C_in) matrix on the k=3 image."""
            ci = c1["c"]
            pad = torch.zeros(res["n"], 3 * ci, **f32)
            pad[:, ci:2 * ci] = res["mat"]
            return image_of(torch.cat([c1["mat"], pad]).contiguous(), torch.cat([c1["bias"], res["bias"]]), ci, 3)

        def _as_conv(l: nn.Linear):
            cv = nn.Conv1d(l.in_features, l.out_features, 1)
            cv.weight.data = l.weight.detach()[:, :, None]
            cv.bias.data = l.bias.detach()
            return cv.to(dev)

        with torch.cuda.device(dev):
            prep = dict(init=conv(self.init_decoder_conv["0"]),
                        ln=(self.init_decoder_conv["3"].weight.detach().to(**f32).contiguous(), self.init_decoder_conv["3"].bias.detach().to(**f32).contiguous()),
                        blocks=[], logits=conv(_as_conv(self.to_coor_logits["0"])))
            # to_coor_logits for the arg-max epilogue (meshtok_conv1d_argmax_coords): rows padded to whole passes of two coordinates
            lg, nb, ncoord = prep["logits"], self.num_discrete_coors, self.nvf * 3
            prep["logits_am"] = None
            if nb == 128 and lg["c"] % 32 == 0:
                n_am = -(-ncoord // 2) * 2 * nb
                if n_am <= 1536:
                    mat = torch.zeros(n_am, lg["mat"].shape[1], **f32)
                    mat[: lg["n"]] = lg["mat"]
                    bias = torch.zeros(n_am, **f32)
                    bias[: lg["n"]] = lg["bias"]
                    prep["logits_am"] = image_of(mat, bias, lg["c"], 1)
            for blk in self.decoders:
                e = blk.excite.net
                prep["blocks"].append(dict(
                    c1=conv(blk.block1.proj), c2=conv(blk.block2.proj),
                    res=None if isinstance(blk.residual_conv, nn.Identity) else conv(blk.residual_conv),
                    se=tuple(t.detach().to(**f32).contiguous() for t in (e[0].weight, e[0].bias, e[2].weight, e[2].bias)),
                    se_t=tuple(t.detach().to(**f32).contiguous() for t in (e[0].weight.t(), e[0].bias, e[2].weight.t(), e[2].bias))))
                last = prep["blocks"][-1]
                last["c1r"] = conv_with_residual(last["c1"], last["res"]) if last["res"] is not None and last["c1"]["k"] == 3 else None
        self._prep, self._prep_key = prep, key
        return prep

    # ---- stages -----------------------------------------------------------------------------------
    @staticmethod
    def _conv(x: Tensor, valid: Tensor, nf: int, w: dict) -> Tensor:
        """Conv1d over the face sequence as im2col image + tcgen05 linear.  x (R, C) fp32 -> (R, n) fp32."""
        R, Cc = x.shape
        assert Cc == w["c"]
        K = w["k"] * Cc
        nbytes = C.c_size_t()
        check(lib.meshtok_rows_image_bytes(R, K, C.byref(nbytes)))
        img = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        check(lib.meshtok_dec_im2col_image(ptr(x), ptr(valid), R, nf, Cc, w["k"], ptr(img), stream_ptr()))
        out = torch.empty(R, w["n"], dtype=torch.float32, device=x.device)
        check(lib.meshtok_linear(None, ptr(img), None, ptr(w["img"]), ptr(w["bias"]), ptr(out), R, w["n"], K, 0, 0, stream_ptr()))
        return out

    @torch.no_grad()
    def dequantize(self, codes: Tensor) -> Tensor:
        """get_output_from_indices of the quantizer (:917): codes (b, n, Q) int64 -> (b, n, dim_codebook) fp32."""
        a = self.autoencoder
        b, n, q = codes.shape
        codes = codes.contiguous().long()
        out = torch.empty(b, n, self.dim_codebook, dtype=torch.float32, device=codes.device)
        if a.use_residual_lfq:
            bits = int(round(math.log2(self.codebook_size)))
            w = a.quantizer.project_out.weight.detach().float().contiguous()
            bias = a.quantizer.project_out.bias.detach().float().contiguous()
            check(lib.meshtok_dec_dequantize_lfq(ptr(codes), ptr(w), ptr(bias), b * n, q, bits, self.dim_codebook, ptr(out), stream_ptr()))
        else:
            book = a.quantizer.layers[0]._codebook.embed[0].detach().float().contiguous()
            check(lib.meshtok_dec_dequantize_rvq(ptr(codes), ptr(book), b * n, q, self.dim_codebook, ptr(out), stream_ptr()))
        return out

    # ---- fused schedule ----------------------------------------------------------------------------
    def fused(self) -> bool:
        if os.environ.get("MESHTOK_DEC_FUSED", "1") == "0":
            return False
        widths = [self.init_decoder_conv["0"].out_channels] + [blk.block1.proj.out_channels for blk in self.decoders]
        return (all(w % 32 == 0 for w in widths) and (self.nvf * self.dim_codebook) % 32 == 0 and self.dim_codebook % 8 == 0
                and self.init_decoder_conv["0"].kernel_size[0] // 2 <= 8)

    @staticmethod
    def _tail_in_epilogue(n: int, cout: int) -> bool:
        """True when the tcgen05 kernel's pass 0 is exactly the `cout` block columns (linear_tc_plan: fewest equal passes of a multiple of
        16, at most 256 columns), i.e. when meshtok_conv1d_pixelnorm_silu covers this convolution."""
        p = -(-n // 256)
        while p <= n // 16 and (n % p != 0 or (n // p) % 16 != 0):
            p += 1
        return cout % 32 == 0 and n % p == 0 and n // p == cout and p <= 2

    @property
    def lead(self) -> int:
        """Separator rows in front of every mesh in the padded row space = half width of the widest kernel."""
        return max(1, self.init_decoder_conv["0"].kernel_size[0] // 2)

    def _padded_valid(self, face_mask: Tensor) -> Tuple[Tensor, int, int]:
        """face_mask (b, nf) -> (valid bytes of the padded row space (b * P + lead), P, R)."""
        b, nf = face_mask.shape
        lead = self.lead
        P = nf + lead
        R = b * P + lead
        valid = torch.zeros(R, dtype=torch.uint8, device=face_mask.device)
        valid[:b * P].view(b, P)[:, lead:] = face_mask
        return valid, P, R

    @staticmethod
    def _halo_image(R: int, Cc: int, halo: int, dev) -> Tensor:
        nbytes = C.c_size_t()
        check(lib.meshtok_halo_image_bytes(R, Cc, halo, C.byref(nbytes)))
        return torch.empty(nbytes.value, dtype=torch.uint8, device=dev)

    @staticmethod
    def _conv1d(img: Tensor, R: int, w: dict) -> Tensor:
        out = torch.empty(R, w["n"], dtype=torch.float32, device=img.device)
        with _conv_timer(R, w["n"], w["k"] * w["c"]):
            check(lib.meshtok_conv1d(ptr(img), ptr(w["img"]), ptr(w["bias"]), ptr(out), R, w["n"], w["c"], w["k"], w["k"] // 2, stream_ptr()))
        return out

    def _decode_fused(self, init_img: Tensor, valid: Tensor, b: int, P: int, R: int, logits_image: bool):
        """init_img: halo image of the (masked) quantized rows for init_decoder_conv.0, padded row space.  Returns (decoded (R, C) fp32
        with zero rows on separators / padded faces, the halo-0 image of it for to_coor_logits or None)."""
        p = self._prepare()
        dev = init_img.device
        lead = self.lead
        blocks = p["blocks"]
        f32 = dict(dtype=torch.float32, device=dev)
        x0 = self._conv1d(init_img, R, p["init"])
        c = p["init"]["n"]
        halo = 1 if blocks else 0
        want = bool(blocks) or logits_image
        x = torch.empty(R, c, **f32)
        img = self._halo_image(R, c, halo, dev) if want else None
        check(lib.meshtok_dec_silu_layernorm_halo(ptr(x0), ptr(valid), R, P, c, ptr(p["ln"][0]), ptr(p["ln"][1]), 1e-5, ptr(x), halo, ptr(img),
                                                  stream_ptr()))
        S = lib.meshtok_dec_se_slices(b, P - lead)
        epi_tail = os.environ.get("MESHTOK_DEC_EPI", "1") != "0"
        tickets = None if os.environ.get("MESHTOK_DEC_SE_TAIL", "1") == "0" else torch.zeros(b, dtype=torch.int32, device=dev)   # self-resetting
        for i, blk in enumerate(blocks):
            cout = blk["c1"]["n"]
            w1 = blk["c1r"] if blk["c1r"] is not None else blk["c1"]
            img2 = self._halo_image(R, cout, 1, dev)
            if epi_tail and self._tail_in_epilogue(w1["n"], cout):
                # block1.proj leaves the convolution only as the image of silu(PixelNorm(.)); residual_conv (second pass) as fp32
                h = torch.empty(R, 2 * cout, **f32) if blk["c1r"] is not None else None
                with _conv_timer(R, w1["n"], w1["k"] * w1["c"]):
                    check(lib.meshtok_conv1d_pixelnorm_silu(ptr(img), ptr(w1["img"]), ptr(w1["bias"]), ptr(valid), ptr(h), R, w1["n"], w1["c"], w1["k"],
                                                            w1["k"] // 2, cout, 1e-4, 1, ptr(img2), stream_ptr()))
                res_ptr, ldres = (h.data_ptr() + 4 * cout, 2 * cout) if h is not None else (x.data_ptr(), x.shape[1])
            else:
                h = self._conv1d(img, R, w1)                         # c1r: (R, 2 cout) = [block1.proj | residual_conv]
                if blk["c1r"] is not None:
                    ldh, res_ptr, ldres = 2 * cout, h.data_ptr() + 4 * cout, 2 * cout
                else:
                    ldh, res_ptr, ldres = cout, x.data_ptr(), x.shape[1]
                check(lib.meshtok_dec_pixelnorm_silu_halo(ptr(h), ldh, ptr(valid), R, P, cout, 1e-4, 1, ptr(img2), stream_ptr()))
            h2 = self._conv1d(img2, R, blk["c2"])
            w1, b1, w2, b2 = blk["se_t"]                              # transposed: (C, inner), (inner, C)
            scale = torch.empty(R, **f32)
            ws = torch.empty(b * S * (cout + 1), **f32)
            gate = torch.empty(b, cout, **f32)
            check(lib.meshtok_dec_pixelnorm_squeeze_excite_gate(ptr(h2), cout, ptr(valid), b, P, lead, cout, 1e-4, w1.shape[1], ptr(w1), ptr(b1),
                                                                ptr(w2), ptr(b2), ptr(scale), ptr(ws), ptr(tickets), ptr(gate), stream_ptr()))
            last = i + 1 == len(blocks)
            halo = 0 if last else 1
            out = torch.empty(R, cout, **f32)
            img = self._halo_image(R, cout, halo, dev) if (not last or logits_image) else None
            check(lib.meshtok_dec_gate_residual_halo(ptr(h2), cout, ptr(scale), ptr(gate), res_ptr, ldres, ptr(valid), R, P, cout, ptr(out), halo,
                                                     ptr(img), stream_ptr()))
            x = out
        return x, img

    def _decode_staged(self, x: Tensor, valid: Tensor, b: int, nf: int) -> Tensor:
        """The first version's schedule: one kernel per stage, a separate im2col pass in front of every convolution."""
        p = self._prepare()
        R, dev = b * nf, x.device
        x = self._conv(x, valid, nf, p["init"])
        check(lib.meshtok_dec_silu_layernorm(ptr(x), R, x.shape[1], ptr(p["ln"][0]), ptr(p["ln"][1]), 1e-5, stream_ptr()))
        for blk in p["blocks"]:
            cout = blk["c1"]["n"]
            res = x if blk["res"] is None else self._conv(x, valid, nf, blk["res"])
            h = self._conv(x, valid, nf, blk["c1"])
            check(lib.meshtok_dec_pixelnorm_silu(ptr(h), ptr(valid), R, cout, 1e-4, stream_ptr()))
            h = self._conv(h, valid, nf, blk["c2"])
            check(lib.meshtok_dec_pixelnorm_silu(ptr(h), ptr(valid), R, cout, 1e-4, stream_ptr()))
            w1, b1, w2, b2 = blk["se"]
            mean = torch.empty(b, cout, dtype=torch.float32, device=dev)
            gate = torch.empty(b, cout, dtype=torch.float32, device=dev)
            check(lib.meshtok_dec_squeeze_excite_gate(ptr(h), ptr(valid), b, nf, cout, w1.shape[0], ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                                      ptr(mean), ptr(gate), stream_ptr()))
            out = torch.empty(R, cout, dtype=torch.float32, device=dev)
            check(lib.meshtok_dec_gate_residual(ptr(h), ptr(gate), ptr(res), ptr(valid), R, nf, cout, ptr(out), stream_ptr()))
            x = out
        return x

    @torch.no_grad()
    def decode(self, quantized: Tensor, face_mask: Tensor) -> Tensor:
        """quantized (b, nf, nvf * dim_codebook) fp32, face_mask (b, nf) bool -> (b, nf, last decoder dim); rows of padded faces are 0."""
        p = self._prepare()
        b, nf, d = quantized.shape
        R = b * nf
        dev = quantized.device
        valid = face_mask.reshape(R).to(torch.uint8).contiguous()
        x = quantized.reshape(R, d).float().contiguous()
        with torch.cuda.device(dev):
            if self.fused():
                k = p["init"]["k"]
                lead = self.lead
                vp, P, Rp = self._padded_valid(face_mask)
                img = self._halo_image(Rp, d, k // 2, dev)
                check(lib.meshtok_dec_rows_halo(ptr(x), d, ptr(vp), Rp, P, lead, d, k // 2, ptr(img), stream_ptr()))
                xp, _ = self._decode_fused(img, vp, b, P, Rp, logits_image=False)
                x = xp[:b * P].view(b, P, -1)[:, lead:].contiguous()
            else:
                x = self._decode_staged(x, valid, b, nf)
        return x.reshape(b, nf, -1)

    @torch.no_grad()
    def decode_from_codes_to_faces(self, codes: Tensor, face_mask: Optional[Tensor] = None, return_discrete_codes: bool = False):
        """Same contract as the reference (:898-947): codes (b, nf * nvf, Q) or any (b, ...) layout of them, int64, pad_id on padded
        faces -> (continuous face coordinates (b, nf, nvf, 3) with NaN on padded faces[, discrete bins], face_mask)."""
        if not codes.is_cuda:
            raise RuntimeError("codes must be a CUDA tensor; there is no CPU path")
        p = self._prepare()
        a = self.autoencoder
        b = codes.shape[0]
        codes = codes.reshape(b, -1)
        nvf, q = self.nvf, self.num_quantizers
        if face_mask is None:
            face_mask = (codes != self.pad_id).reshape(b, -1, nvf * q).all(dim=-1)
        codes = codes.reshape(b, -1, q).contiguous().long()
        nf = codes.shape[1] // nvf
        R = b * nf
        dev = codes.device
        valid = face_mask.reshape(R).to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            ncoord = nvf * 3
            coords = torch.empty(b, nf, nvf, 3, dtype=torch.float32, device=dev)
            bins = torch.empty(b, nf, nvf, 3, dtype=torch.int64, device=dev)
            lo, hi = self.coor_range
            if self.fused():
                k, d = p["init"]["k"], nvf * self.dim_codebook
                lead = self.lead
                vp, P, Rp = self._padded_valid(face_mask)
                img = self._halo_image(Rp, d, k // 2, dev)
                if a.use_residual_lfq:
                    quantized = self.dequantize(codes)
                    check(lib.meshtok_dec_rows_halo(ptr(quantized), d, ptr(vp), Rp, P, lead, d, k // 2, ptr(img), stream_ptr()))
                else:
                    book = a.quantizer.layers[0]._codebook.embed[0].detach().float().contiguous()
                    check(lib.meshtok_dec_dequantize_rvq_halo(ptr(codes), ptr(book), ptr(vp), Rp, P, lead, nvf, q, self.dim_codebook, k // 2,
                                                              ptr(img), stream_ptr()))
                decoded, limg = self._decode_fused(img, vp, b, P, Rp, logits_image=True)
                am = p["logits_am"] if os.environ.get("MESHTOK_DEC_ARGMAX", "1") != "0" else None
                if am is not None:      # to_coor_logits + arg-max + undiscretize in one launch: the logits stay in TMEM
                    with _conv_timer(Rp, p["logits"]["n"], am["c"]):      # (the algorithmic width, without the padding rows)
                        check(lib.meshtok_conv1d_argmax_coords(ptr(limg), ptr(am["img"]), ptr(am["bias"]), ptr(vp), Rp, am["n"], am["c"], 1, 0, b, P, lead,
                                                               ncoord, self.num_discrete_coors, float(lo), float(hi), ptr(coords), ptr(bins), stream_ptr()))
                else:
                    logits = self._conv1d(limg, Rp, p["logits"])
                    check(lib.meshtok_dec_argmax_coords_padded(ptr(logits), ptr(vp), b, P, lead, ncoord, self.num_discrete_coors, float(lo), float(hi),
                                                               ptr(coords), ptr(bins), stream_ptr()))
            else:
                quantized = self.dequantize(codes).reshape(R, nvf * self.dim_codebook)
                decoded = self._decode_staged(quantized, valid, b, nf)      # rows of padded faces are already zero (:924)
                ones = torch.ones(R, dtype=torch.uint8, device=dev)        # to_coor_logits is a plain Linear: no neighbours, no mask
                logits = self._conv(decoded, ones, nf, p["logits"])
                check(lib.meshtok_dec_argmax_coords(ptr(logits), ptr(valid), R, ncoord, self.num_discrete_coors, float(lo), float(hi),
                                                    ptr(coords), ptr(bins), stream_ptr()))
        if return_discrete_codes:
            return coords, bins, face_mask
        return coords, face_mask

    def graphed(self, codes: Tensor) -> "GraphedDecoder":
        """decode_from_codes_to_faces for one codes shape as a CUDA graph (the ~160 launches of the decoder stack replayed at once)."""
        return GraphedDecoder(self, codes)


class GraphedDecoder:
    """Static input / output buffers + the captured graph.  Call it with new codes of the captured shape."""

    def __init__(self, decoder: MeshDecoder, codes: Tensor):
        self.codes = codes.clone()
        decoder.decode_from_codes_to_faces(self.codes)                 # eager: weight images, shared-memory attributes
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            decoder.decode_from_codes_to_faces(self.codes)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.coords, self.bins, self.face_mask = decoder.decode_from_codes_to_faces(self.codes, return_discrete_codes=True)

    def __call__(self, codes: Optional[Tensor] = None):
        if codes is not None:
            self.codes.copy_(codes, non_blocking=True)
        self.graph.replay()
        return self.coords, self.face_mask
