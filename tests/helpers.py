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
