"""Freezes tests/golden/frontend_own_code.pt: outputs of the REFERENCE'S OWN statements meshgpt_pytorch.py:1526-1570
(MeshTransformer.forward_on_codes: token / position / level / vertex embeddings, grouping per face, to_face_tokens), executed from
/root/reference source by oracle/frontend_oracle.reference_lines with small random-init parameters.  Container only.

    python tests/golden/make_frontend_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frontend_oracle as FO  # noqa: E402

CASES = {   # name -> (constructor keywords, batch, code length, trailing pads in mesh 1)
    "tri_whole": (dict(codebook_size=64, num_quantizers=2, quads=False, dim=32, max_seq_len=96), 3, 36, 8),
    "tri_partial": (dict(codebook_size=64, num_quantizers=2, quads=False, dim=32, max_seq_len=96), 2, 41, 5),
    "quad_q3": (dict(codebook_size=32, num_quantizers=3, quads=True, dim=32, max_seq_len=120), 2, 60, 13),
    "eos_token": (dict(codebook_size=16, num_quantizers=2, quads=False, dim=32, max_seq_len=24), 2, 19, 0),
    # append_eos (:1506-1518, executed from the reference source too): a full row (eos lands in the extra column) and a padded one
    "append_eos": (dict(codebook_size=64, num_quantizers=2, quads=False, dim=32, max_seq_len=96), 3, 36, 8),
    "append_eos_partial": (dict(codebook_size=32, num_quantizers=3, quads=True, dim=32, max_seq_len=120), 2, 47, 20),
}


def main():
    out = {}
    for i, (name, (kw, b, L, pads)) in enumerate(CASES.items()):
        torch.manual_seed(100 + i)
        m = FO.OracleTokenFrontEnd(**kw).eval()
        m.to_face_tokens[1].weight.data = torch.randn(kw["dim"]) * 0.3 + 1.0      # LayerNorm defaults (1, 0) would hide a swapped gamma / beta
        m.to_face_tokens[1].bias.data = torch.randn(kw["dim"]) * 0.3
        codes = torch.randint(0, kw["codebook_size"], (b, L))
        if pads:
            codes[1, L - pads:] = -1
        if name == "eos_token":
            codes[0, L - 1] = kw["codebook_size"]                                  # eos id = codebook_size (:1143) is a valid row of token_embed
        eos = name.startswith("append_eos")
        grouped, face = FO.reference_lines(m, codes, append_eos=eos)
        out[name] = dict(kwargs=kw, state_dict=m.state_dict(), codes=codes, grouped_codes=grouped, face_codes=face, append_eos=eos)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontend_own_code.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
