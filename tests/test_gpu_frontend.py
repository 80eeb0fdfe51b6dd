"""GPU: the token-embedding front end (meshgpt_pytorch_b200/frontend.py + csrc/frontend.cu) against the CPU oracle
(oracle/frontend_oracle.py, pinned bit for bit to the reference's own statements meshgpt_pytorch.py:1526-1570 by
tests/test_frontend_oracle.py) and against the committed fixtures of those statements, through the C ABI."""
import os

import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import frontend_oracle as FO

pytestmark = pytest.mark.gpu


def make_pair(seed=0, **kw):
    import meshgpt_pytorch_b200 as M
    torch.manual_seed(seed)
    o = FO.OracleTokenFrontEnd(**kw).eval()
    o.to_face_tokens[1].weight.data = torch.randn(kw["dim"]) * 0.3 + 1.0
    o.to_face_tokens[1].bias.data = torch.randn(kw["dim"]) * 0.3
    m = M.MeshTransformerFrontEnd(codebook_size=kw["codebook_size"], num_quantizers=kw["num_quantizers"], quads=kw["quads"], dim=kw["dim"],
                                  max_seq_len=kw["max_seq_len"], pad_id=kw.get("pad_id", -1), num_sos_tokens=kw.get("num_sos_tokens", 1))
    m.load_reference_state_dict(o.state_dict())
    return o, m.cuda()


def face_close(got, want):
    err = (got - want).abs().max() / want.abs().max()
    assert float(err) < 2e-5, float(err)        # fp32; the folded tables sum n + 1 rounded rows instead of one 3072-term dot product


@pytest.mark.parametrize("kw,b,L", [
    (dict(codebook_size=300, num_quantizers=2, quads=False, dim=128, max_seq_len=600), 3, 600),     # whole faces, full length
    (dict(codebook_size=300, num_quantizers=2, quads=False, dim=256, max_seq_len=600), 2, 305),     # incomplete last face
    (dict(codebook_size=64, num_quantizers=3, quads=True, dim=384, max_seq_len=240), 4, 120),       # quads, 3 levels: 12 codes per face
    (dict(codebook_size=64, num_quantizers=2, quads=False, dim=512, max_seq_len=60, pad_id=-7), 2, 5),   # shorter than one face
    (dict(codebook_size=16, num_quantizers=1, quads=False, dim=1024, max_seq_len=30), 1, 30),
])
def test_front_end_against_the_oracle(kw, b, L):
    o, m = make_pair(**kw)
    pad = kw.get("pad_id", -1)
    codes = torch.randint(0, kw["codebook_size"] + 1, (b, L))                 # includes the eos id (= codebook_size, :1143)
    if b > 1:
        codes[1, L // 3:] = pad                                              # a shorter mesh
    want_g, want_f = o(codes)
    got_g, got_f = m.embed_codes(codes.cuda())
    assert got_g.shape == want_g.shape and got_f.shape == want_f.shape
    assert torch.equal(got_g.cpu(), want_g)                                   # four adds in the reference's order: bit-exact
    if want_f.numel():
        face_close(got_f.cpu(), want_f)


@pytest.mark.parametrize("name", ["tri_whole", "tri_partial", "quad_q3", "eos_token", "append_eos", "append_eos_partial"])
def test_front_end_against_the_reference_fixtures(name):
    """The committed outputs of the reference's own lines (dim 32 there; the face-token kernel needs dim % 128 == 0, so the fixture's
    parameters are embedded into the first 32 of 128 channels - the extra channels are zero and LayerNorm's statistics are rescaled
    by hand below)."""
    import meshgpt_pytorch_b200 as M
    case = torch.load(os.path.join(GOLDEN_DIR, "frontend_own_code.pt"))[name]
    kw, sd = case["kwargs"], case["state_dict"]
    d = kw["dim"]
    n = kw["num_quantizers"] * (4 if kw["quads"] else 3)
    m = M.MeshTransformerFrontEnd(codebook_size=kw["codebook_size"], num_quantizers=kw["num_quantizers"], quads=kw["quads"], dim=128,
                                  max_seq_len=kw["max_seq_len"])
    wide = {k: torch.zeros_like(v) for k, v in m.state_dict().items()}
    for k in ("token_embed.weight", "abs_pos_emb.weight", "quantize_level_embed", "vertex_embed", "sos_token"):
        wide[k][:, :d] = sd[k]
    wide["to_face_tokens.0.weight"].view(128, n, 128)[:d, :, :d] = sd["to_face_tokens.0.weight"].view(d, n, d)
    wide["to_face_tokens.0.bias"][:d] = sd["to_face_tokens.0.bias"]
    wide["to_face_tokens.1.weight"][:] = 1.0
    m.load_reference_state_dict(wide)
    m = m.cuda()
    eos = bool(case.get("append_eos", False))            # fixtures of the reference's own `if append_eos:` block (:1506-1518) + :1526-1570
    got_g, _ = m.embed_codes(case["codes"].cuda(), append_eos=eos)
    assert torch.equal(got_g.cpu()[..., :d], case["grouped_codes"]) and float(got_g[..., d:].abs().max()) == 0.0
    # face tokens: the kernel's LayerNorm ran over 128 channels of which 96 are zero; undo it and redo it over the 32 real ones
    m.to_face_tokens[1].eps = 0.0
    m._prep = None
    pre = m.face_tokens(case["codes"].cuda(), append_eos=eos).cpu()[..., :d]   # = (x - mean128) / std128, an affine map of x per row
    ln = torch.nn.functional.layer_norm(pre, (d,), sd["to_face_tokens.1.weight"], sd["to_face_tokens.1.bias"], 1e-5)
    want = case["face_codes"]
    if want.numel():
        assert float((ln - want).abs().max() / want.abs().max()) < 1e-3      # eps enters at a different scale after the affine map


def test_front_end_on_tokenizer_output():
    """tokenize -> front end, full-size shapes: 800-face meshes, 16384 codes + eos, dim 512, codes straight from the GPU tokenizer."""
    from helpers import make_pair as make_tokenizers
    import synth
    _, tok = make_tokenizers(use_residual_lfq=False)
    v, f = synth.ragged_batch("tri", 20, 20, [800, 423, 1], [0, 1, 2])
    codes = tok.tokenize(v.cuda(), f.cuda())                                 # (3, 2400, 2), -1 on padded faces
    flat = codes.reshape(codes.shape[0], -1)                                 # 'b n q -> b (n q)'
    o, m = make_pair(codebook_size=16384, num_quantizers=2, quads=False, dim=512, max_seq_len=4800)
    want_g, want_f = o(flat.cpu())
    got_g, got_f = m.embed_codes(flat)
    assert got_f.shape == (3, 800, 512) and got_g.shape == (3, 800, 6, 512)
    assert torch.equal(got_g.cpu(), want_g)
    face_close(got_f.cpu(), want_f)


def test_front_end_refuses_what_it_does_not_cover():
    import meshgpt_pytorch_b200 as M
    with pytest.raises(NotImplementedError):
        M.MeshTransformerFrontEnd(codebook_size=16, num_quantizers=2, quads=False, dim=96, max_seq_len=60)
    _, m = make_pair(codebook_size=16, num_quantizers=2, quads=False, dim=128, max_seq_len=60)
    with pytest.raises(RuntimeError):
        m.embed_codes(torch.zeros(1, 6, dtype=torch.int64))                  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        m.embed_codes(torch.zeros(1, 66, dtype=torch.int64, device="cuda"))  # longer than max_seq_len (:1502)


@pytest.mark.parametrize("kw,b,L", [
    (dict(codebook_size=300, num_quantizers=2, quads=False, dim=128, max_seq_len=606), 3, 600),
    (dict(codebook_size=64, num_quantizers=3, quads=True, dim=256, max_seq_len=240), 4, 119),
    (dict(codebook_size=64, num_quantizers=2, quads=False, dim=128, max_seq_len=60, num_sos_tokens=4), 2, 36),
])
def test_append_eos_and_sos_on_the_device(kw, b, L):
    """append_eos (:1506-1518) and the sos rows (:1591-1594), which round 1 left to torch: the eos-extended code tensor bit for bit, the
    embeddings of the extended sequence without materialising it, the sos parameter rows in front of the face tokens."""
    o, m = make_pair(**kw)
    codes = torch.randint(0, kw["codebook_size"], (b, L))
    codes[1, L // 3:] = -1                      # a shorter row: its eos lands in the middle
    if b > 2:
        codes[2, :] = -1                        # an empty row: eos at position 0
    assert torch.equal(m.append_eos(codes.cuda()).cpu(), o.append_eos(codes))
    want_g, want_f = o(codes, append_eos=True, prepend_sos=True)
    got_g, got_f = m.embed_codes(codes.cuda(), append_eos=True, prepend_sos=True)
    assert got_g.shape == want_g.shape and got_f.shape == want_f.shape
    assert torch.equal(got_g.cpu(), want_g)
    k = o.sos_token.shape[0]
    assert torch.equal(got_f.cpu()[:, :k], want_f[:, :k])                     # the parameter itself
    face_close(got_f.cpu()[:, k:], want_f[:, k:])


def test_front_end_fed_from_the_tokenizers_vertex_codes():
    """"codes -> first transformer input" WITHOUT the (b, nf * nvf, q) int64 code tensor: the front-end kernels read the
    tokenizer's per-vertex codes and its face / vertex maps.  Same bits as embedding the code tensor tokenize() returns."""
    from helpers import make_pair as make_tokenizers
    import synth
    _, tok = make_tokenizers(use_residual_lfq=False)
    v, f = synth.ragged_batch("tri", 20, 20, [800, 423, 1], [0, 1, 2])
    fd = f.cuda()
    codes = tok.tokenize(v.cuda(), fd)
    flat = codes.reshape(codes.shape[0], -1)
    _, m = make_pair(codebook_size=16384, num_quantizers=2, quads=False, dim=512, max_seq_len=4806)
    for eos, sos in ((False, False), (True, True)):
        want_g, want_f = m.embed_codes(flat, append_eos=eos, prepend_sos=sos)
        got_g, got_f = m.embed_tokenizer_output(tok, fd, append_eos=eos, prepend_sos=sos)
        assert torch.equal(got_g, want_g) and torch.equal(got_f, want_f), (eos, sos)
