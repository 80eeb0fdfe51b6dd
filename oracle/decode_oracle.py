"""
CPU oracle of the DECODE half of MeshAutoencoder: codes -> face coordinates.

*** TEST INFRASTRUCTURE, groundwork for the next round.  No CUDA path for decoding exists yet; nothing in the product
imports this file.  It restates, in plain fp32 torch on the CPU,

    MeshAutoencoder.decode_from_codes_to_faces      /root/reference/meshgpt_pytorch/meshgpt_pytorch.py:898-947
    MeshAutoencoder.decode                           :872-896   (attn_decoder_depth == 0)
    init_decoder_conv / decoders / to_coor_logits    :616-641   (construction)
    ResnetBlock, Block, SqueezeExcite, PixelNorm     :286-379
    undiscretize                                     :203-218

with the reference's state-dict key names (init_decoder_conv.0.*, init_decoder_conv.3.*, decoders.{i}.block{1,2}.proj.*,
decoders.{i}.excite.net.{0,2}.*, decoders.{i}.residual_conv.*, to_coor_logits.0.*), so that a reference checkpoint loads.

Parity status: the blocks above are the reference's OWN code; tests/test_oracle_golden.py pins this restatement against
fixtures produced by executing those classes from the reference source (oracle/reference_extract.py, tests/golden/
make_golden.py -> tests/golden/decode_own_code.pt).  `get_output_from_indices` (codes -> summed codebook vectors) belongs to
vector_quantize_pytorch, which is not installed: restated from its published behaviour, PARITY UNPINNED -
  ResidualVQ: sum over the levels of codebook[code]; a level whose index is -1 contributes zeros (masked positions);
  ResidualLFQ: per level bits -> +-scale_l (scale_l = 2^-l), summed over the levels, then project_out.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn


def undiscretize(t: Tensor, lo: float, hi: float, num_discrete: int = 128) -> Tensor:
    """meshgpt_pytorch.py:203-218: bin centre of an integer bin."""
    t = t.float()
    t = t + 0.5
    t = t / num_discrete
    return t * (hi - lo) + lo


class PixelNormP(nn.Module):
    """:286-294  F.normalize over the channel dim (eps 1e-4) * sqrt(channels)."""

    def __init__(self, dim: int, eps: float = 1e-4):
        super().__init__()
        self.dim, self.eps = dim, eps

    def forward(self, x: Tensor) -> Tensor:
        return F.normalize(x, dim=self.dim, eps=self.eps) * math.sqrt(x.shape[self.dim])


class SqueezeExciteP(nn.Module):
    """:296-324  masked mean over the sequence -> Linear -> SiLU -> Linear -> Sigmoid -> channel gate."""

    def __init__(self, dim: int, reduction_factor: int = 4, min_dim: int = 16):
        super().__init__()
        inner = max(dim // reduction_factor, min_dim)
        # same module indices as the reference's nn.Sequential(Linear, SiLU, Linear, Sigmoid, Rearrange)
        self.net = nn.Sequential(nn.Linear(dim, inner), nn.SiLU(), nn.Linear(inner, dim), nn.Sigmoid())

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:   # x (b, c, n), mask (b, 1, n)
        if mask is not None:
            x = x.masked_fill(~mask, 0.0)
            num = x.sum(dim=-1)
            den = mask.float().sum(dim=-1)
            avg = num / den.clamp(min=1e-5)
        else:
            avg = x.mean(dim=-1)
        return x * self.net(avg)[:, :, None]


class BlockP(nn.Module):
    """:326-353  mask -> Conv1d(k=3, pad=1) -> mask -> PixelNorm -> SiLU (dropout is identity in eval)."""

    def __init__(self, dim: int, dim_out: Optional[int] = None):
        super().__init__()
        dim_out = dim if dim_out is None else dim_out
        self.proj = nn.Conv1d(dim, dim_out, 3, padding=1)
        self.norm = PixelNormP(dim=1)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        if mask is not None:
            x = x.masked_fill(~mask, 0.0)
        x = self.proj(x)
        if mask is not None:
            x = x.masked_fill(~mask, 0.0)
        return F.silu(self.norm(x))


class ResnetBlockP(nn.Module):
    """:355-379"""

    def __init__(self, dim: int, dim_out: Optional[int] = None):
        super().__init__()
        dim_out = dim if dim_out is None else dim_out
        self.block1 = BlockP(dim, dim_out)
        self.block2 = BlockP(dim_out, dim_out)
        self.excite = SqueezeExciteP(dim_out)
        self.residual_conv = nn.Conv1d(dim, dim_out, 1) if dim != dim_out else nn.Identity()

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        res = self.residual_conv(x)
        h = self.block1(x, mask=mask)
        h = self.block2(h, mask=mask)
        h = self.excite(h, mask=mask)
        return h + res


class OracleDecoder(nn.Module):
    """The decoder modules of MeshAutoencoder under their reference names, plus decode / decode_from_codes_to_faces.

    `quantizer` is the tokenization oracle's quantizer module (ResidualVQParams / ResidualLFQParams of
    oracle/meshtok_oracle.py): its codebook / projections turn codes back into vectors."""

    def __init__(self, quantizer: nn.Module, dim_codebook: int = 192, num_quantizers: int = 2, codebook_size: int = 16384,
                 use_residual_lfq: bool = True, init_decoder_conv_kernel: int = 7,
                 decoder_dims_through_depth: Tuple[int, ...] = (128, 128, 128, 128, 192, 192, 192, 192, 256, 256, 256, 256, 256, 256, 384, 384, 384),
                 num_discrete_coors: int = 128, coor_continuous_range: Tuple[float, float] = (-1.0, 1.0), pad_id: int = -1,
                 quads: bool = False, attn_decoder_depth: int = 0):
        super().__init__()
        assert attn_decoder_depth == 0, "decoder attention blocks are out of scope"
        assert init_decoder_conv_kernel % 2 == 1
        self.nvf = 4 if quads else 3
        self.pad_id, self.num_quantizers, self.codebook_size = pad_id, num_quantizers, codebook_size
        self.dim_codebook, self.use_residual_lfq = dim_codebook, use_residual_lfq
        self.num_discrete_coors = num_discrete_coors
        self.coor_range = tuple(float(x) for x in coor_continuous_range)
        self.quantizer = quantizer
        init_dim, *rest = decoder_dims_through_depth
        # nn.Sequential(Conv1d, SiLU, Rearrange, LayerNorm, Rearrange): parameters live at indices 0 and 3 (:621-627)
        self.init_decoder_conv = nn.ModuleDict({
            "0": nn.Conv1d(dim_codebook * self.nvf, init_dim, init_decoder_conv_kernel, padding=init_decoder_conv_kernel // 2),
            "3": nn.LayerNorm(init_dim),
        })
        self.decoders = nn.ModuleList()
        cur = init_dim
        for d in rest:
            self.decoders.append(ResnetBlockP(cur, d))
            cur = d
        self.to_coor_logits = nn.ModuleDict({"0": nn.Linear(cur, num_discrete_coors * self.nvf * 3)})

    # ---- vector_quantize_pytorch get_output_from_indices, restated (see the header) ----------------
    @torch.no_grad()
    def dequantize(self, codes: Tensor) -> Tensor:
        """codes (b, n, Q) int64 -> (b, n, dim_codebook) fp32."""
        b, n, q = codes.shape
        assert q == self.num_quantizers
        if self.use_residual_lfq:
            bits = int(round(math.log2(self.codebook_size)))
            shifts = torch.arange(bits - 1, -1, -1)
            out = torch.zeros(b, n, bits)
            for level in range(q):
                idx = codes[..., level]
                valid = (idx >= 0)[..., None]
                on = ((idx.clamp(min=0)[..., None] >> shifts) & 1).bool()
                scale = 2.0 ** -level
                out = out + torch.where(valid, torch.where(on, scale, -scale), torch.zeros(()))
            return self.quantizer.project_out(out)
        book = self.quantizer.layers[0]._codebook.embed[0]                     # shared codebook (1, K, D)
        out = torch.zeros(b, n, self.dim_codebook)
        for level in range(q):
            idx = codes[..., level]
            vec = book[idx.clamp(min=0)]
            out = out + vec.masked_fill((idx < 0)[..., None], 0.0)
        return out

    # ---- :872-896 ------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, quantized: Tensor, face_mask: Tensor) -> Tensor:
        """quantized (b, nf, nvf * dim_codebook), face_mask (b, nf) bool -> (b, nf, last decoder dim)."""
        conv_mask = face_mask[:, None, :]
        x = quantized.transpose(1, 2)
        x = x.masked_fill(~conv_mask, 0.0)
        x = self.init_decoder_conv["0"](x)
        x = F.silu(x)
        x = self.init_decoder_conv["3"](x.transpose(1, 2)).transpose(1, 2)
        for block in self.decoders:
            x = block(x, mask=conv_mask)
        return x.transpose(1, 2)

    # ---- :898-947 ------------------------------------------------------------------------------
    @torch.no_grad()
    def decode_from_codes_to_faces(self, codes: Tensor, face_mask: Optional[Tensor] = None, return_discrete_codes: bool = False):
        b = codes.shape[0]
        codes = codes.reshape(b, -1)
        if face_mask is None:
            face_mask = (codes != self.pad_id).reshape(b, -1, self.nvf * self.num_quantizers).all(dim=-1)
        codes = codes.reshape(b, -1, self.num_quantizers)
        quantized = self.dequantize(codes)
        quantized = quantized.reshape(b, -1, self.nvf * self.dim_codebook)
        decoded = self.decode(quantized, face_mask)
        decoded = decoded.masked_fill(~face_mask[..., None], 0.0)
        logits = self.to_coor_logits["0"](decoded).reshape(b, -1, self.nvf * 3, self.num_discrete_coors)
        pred = logits.argmax(dim=-1).reshape(b, -1, self.nvf, 3)
        lo, hi = self.coor_range
        coords = undiscretize(pred, lo, hi, self.num_discrete_coors)
        coords = coords.masked_fill(~face_mask[:, :, None, None], float("nan"))
        if return_discrete_codes:
            return coords, pred, face_mask
        return coords, face_mask
