"""CPU: a cross-check of the oracle's ResidualVQ restatement against an INDEPENDENT third-party implementation that is present in this
image: Hugging Face transformers' EnCodec quantizer (transformers.models.encodec.modeling_encodec.EncodecEuclideanCodebook - itself a
port of the Euclidean-codebook residual VQ family vector-quantize-pytorch belongs to): nearest code by squared Euclidean distance
(arg-max of the negated distance, first maximum on ties), residual = residual - code, level after level.

This is NOT the reference's pinned dependency (vector-quantize-pytorch >= 1.18.1 is not installed and not in the wheelhouse, see
DESIGN.md section 2) - parity for SURVEY rows a8 stays "unpinned" in the strict sense - but it is executable evidence from code nobody
here wrote that the restated arithmetic (eval mode: no sampling noise, shared codebook, sqrt and clamp monotone, lowest index on ties)
selects the same codes."""
import pytest
import torch

from helpers import make_oracle

encodec = pytest.importorskip("transformers.models.encodec.modeling_encodec")


@pytest.mark.parametrize("K,D,R", [(1024, 64, 4000), (16384, 192, 600)])
def test_oracle_rvq_agrees_with_the_encodec_residual_quantizer(K, D, R):
    from transformers import EncodecConfig
    cfg = dict(use_residual_lfq=False, codebook_size=K, dim_codebook=D) if D != 192 else dict(use_residual_lfq=False)
    if D != 192:
        cfg.update(encoder_dims_through_depth=(32, 64), dim_coor_embed=16, dim_area_embed=8, dim_normal_embed=16, dim_angle_embed=8)
    o = make_oracle(**cfg)
    codebook = o.quantizer.codebook                                  # (K, D), shared by both levels
    assert codebook.shape == (K, D)
    ecfg = EncodecConfig(codebook_size=K, codebook_dim=D)
    layers = []
    for _ in range(o.num_quantizers):
        cb = encodec.EncodecEuclideanCodebook(ecfg)
        with torch.no_grad():
            cb.embed.copy_(codebook)
        layers.append(cb)
    g = torch.Generator().manual_seed(K + R)
    rows = torch.randn(R, D, generator=g) * 0.05
    rows[:5] = codebook[7:12] + 1e-4 * torch.randn(5, D, generator=g)          # near-exact hits
    residual, want = rows.clone(), []
    with torch.no_grad():
        for cb in layers:                                                        # EncodecResidualVectorQuantizer.encode's loop
            idx = cb.encode(residual)
            residual = residual - cb.decode(idx)
            want.append(idx)
    want = torch.stack(want, dim=-1)
    qsum, got = o.quantizer(rows)
    bad = (got != want).any(-1).nonzero().flatten().tolist()
    for r in bad:                                                                # the two formulas round differently: only near ties may differ
        lvl = 0 if got[r, 0] != want[r, 0] else 1
        x = rows[r] if lvl == 0 else rows[r] - codebook[want[r, 0]]
        d_got, d_want = (x - codebook[got[r, lvl]]).norm(), (x - codebook[want[r, lvl]]).norm()
        assert abs(float(d_got - d_want)) <= 1e-5 * float(d_want) + 1e-8, (r, lvl, float(d_got), float(d_want))
    assert len(bad) <= R // 500
    ok = [r for r in range(R) if r not in set(bad)]
    assert torch.allclose((rows - qsum)[ok], residual[ok], rtol=1e-5, atol=1e-7)  # the same final residual
