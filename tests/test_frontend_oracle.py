"""CPU: the token-embedding front-end oracle (oracle/frontend_oracle.py) against fixtures frozen from the
REFERENCE'S OWN statements meshgpt_pytorch.py:1526-1570 (tests/golden/frontend_own_code.pt, made by
tests/golden/make_frontend_golden.py), and - where /root/reference exists - against those statements executed live."""
import os

import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import frontend_oracle as FO


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLDEN_DIR, "frontend_own_code.pt"))


@pytest.mark.parametrize("name", ["tri_whole", "tri_partial", "quad_q3", "eos_token", "append_eos", "append_eos_partial"])
def test_front_end_golden(fx, name):
    case = fx[name]
    m = FO.OracleTokenFrontEnd(**case["kwargs"]).eval()
    m.load_state_dict(case["state_dict"])                       # the reference's key names, strict
    eos = bool(case.get("append_eos", False))                   # (those fixtures ran the reference's own `if append_eos:` block too)
    grouped, face = m(case["codes"], append_eos=eos)
    assert torch.equal(grouped, case["grouped_codes"])           # same torch CPU ops in the same order: bit for bit
    assert torch.equal(face, case["face_codes"])
    n = m.num_quantizers * m.num_vertices_per_face
    L = case["codes"].shape[1] + int(eos)
    assert grouped.shape[1] == -(-L // n) and face.shape[1] == L // n
    if L % n:                                                    # the zero padding of the incomplete last face (:1560)
        assert float(grouped[:, -1, L % n:].abs().max()) == 0.0


def test_pad_codes_read_row_zero(fx):
    case = fx["tri_whole"]
    m = FO.OracleTokenFrontEnd(**case["kwargs"]).eval()
    m.load_state_dict(case["state_dict"])
    codes = case["codes"]
    g0, f0 = m(codes)
    g1, f1 = m(codes.masked_fill(codes == -1, 0))                # :1528
    assert torch.equal(g0, g1) and torch.equal(f0, f1)


@pytest.mark.reference
def test_restatement_equals_reference_lines_live():
    torch.manual_seed(7)
    m = FO.OracleTokenFrontEnd(codebook_size=128, num_quantizers=2, quads=False, dim=64, max_seq_len=600).eval()
    for L in (6, 7, 599, 600):
        codes = torch.randint(0, 129, (2, L))
        codes[1, L // 2:] = -1
        want_g, want_f = FO.reference_lines(m, codes)
        got_g, got_f = m(codes)
        assert torch.equal(got_g, want_g) and torch.equal(got_f, want_f)
        if L < 600:                                                  # append_eos (:1506-1518) executed from the reference source as well
            want_g, want_f = FO.reference_lines(m, codes, append_eos=True)
            got_g, got_f = m(codes, append_eos=True)
            assert torch.equal(got_g, want_g) and torch.equal(got_f, want_f)
