"""CPU: the decode oracle (oracle/decode_oracle.py, groundwork for SURVEY 8f rank 1) against fixtures frozen from the
REFERENCE'S OWN decoder blocks (tests/golden/decode_own_code.pt, made by tests/golden/make_golden.py from
meshgpt_pytorch.py:203-218, 286-379), plus shape / masking properties of decode_from_codes_to_faces."""
import os

import pytest
import torch

from helpers import GOLDEN_DIR, SMALL_CFG, make_oracle
from oracle import decode_oracle as D


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN_DIR, "decode_own_code.pt"))


def test_pixelnorm_and_undiscretize_golden(fx):
    assert torch.allclose(D.PixelNormP(dim=1)(fx["x"]), fx["pixelnorm"], rtol=0, atol=1e-6)
    u = fx["undiscretize"]
    assert torch.equal(D.undiscretize(u["bins"], -1.0, 1.0, 128), u["out"])
    assert torch.equal(D.undiscretize(u["bins"] % 64, 0.0, 4.0, 64), u["out_other"])
    assert D.undiscretize(torch.tensor([0, 127]), -1.0, 1.0, 128).tolist() == pytest.approx([-0.9921875, 0.9921875])


@pytest.mark.parametrize("name", ["same", "wider"])
def test_resnet_block_golden(fx, name):
    case = fx["resnet"][name]
    blk = D.ResnetBlockP(*case["dims"]).eval()
    blk.load_state_dict(case["state_dict"])          # the reference's key names, strict
    x, mask = fx["x"], fx["mask"]
    with torch.no_grad():
        assert torch.allclose(blk.block1(x, mask=mask), case["block1"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(blk.excite(blk.block2(blk.block1(x, mask=mask), mask=mask), mask=mask), case["excite"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(blk(x, mask=mask), case["out_masked"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(blk(x), case["out_unmasked"], rtol=1e-5, atol=1e-6)
    # a mesh whose only valid face is the first one: everything behind it is exactly the (masked) residual path
    assert torch.equal(blk(x, mask=mask)[2, :, 5], blk.residual_conv(x)[2, :, 5])


@pytest.mark.reference
def test_blocks_match_reference_source_live(fx):
    from oracle import reference_extract
    R = reference_extract.load()
    torch.manual_seed(0)
    ref = R.ResnetBlock(16, 24).eval()
    mine = D.ResnetBlockP(16, 24).eval()
    mine.load_state_dict(ref.state_dict())
    with torch.no_grad():
        assert torch.allclose(mine(fx["x"], mask=fx["mask"]), ref(fx["x"], mask=fx["mask"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("cfg", [dict(use_residual_lfq=False), dict()])
def test_decode_from_codes_shapes_and_masks(cfg):
    cfg = dict(SMALL_CFG, **cfg)
    o = make_oracle(**cfg)
    dec = D.OracleDecoder(o.quantizer, dim_codebook=cfg["dim_codebook"], num_quantizers=cfg["num_quantizers"], codebook_size=cfg["codebook_size"],
                          use_residual_lfq=cfg.get("use_residual_lfq", True), decoder_dims_through_depth=(32, 32, 48), init_decoder_conv_kernel=7).eval()
    keys = set(dec.state_dict())
    assert {"init_decoder_conv.0.weight", "init_decoder_conv.3.weight", "decoders.0.block1.proj.weight", "decoders.1.residual_conv.weight",
            "decoders.0.excite.net.2.bias", "to_coor_logits.0.weight"} <= keys
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, cfg["codebook_size"], (2, 10 * 3, cfg["num_quantizers"]), generator=g)
    codes[1, 6 * 3:] = -1                                  # four padded faces
    coords, pred, mask = dec.decode_from_codes_to_faces(codes, return_discrete_codes=True)
    assert coords.shape == (2, 10, 3, 3) and pred.shape == (2, 10, 3, 3) and mask.shape == (2, 10)
    assert mask[0].all() and mask[1].tolist() == [True] * 6 + [False] * 4
    assert torch.isnan(coords[1, 6:]).all() and not torch.isnan(coords[1, :6]).any() and not torch.isnan(coords[0]).any()
    assert float(coords[0].min()) >= -1.0 and float(coords[0].max()) <= 1.0
    # padded faces do not influence the valid ones of the same mesh (masked to zero before every conv)
    codes2 = codes.clone()
    codes2[1, 6 * 3:] = -1
    shorter = dec.decode_from_codes_to_faces(codes2[1:, :6 * 3])[0]
    same = torch.allclose(shorter[0], coords[1, :6], equal_nan=True)
    # (zero padding of the convolution = masked faces: the last valid faces see zeros either way)
    assert same
