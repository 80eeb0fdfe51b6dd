"""
CPU restatement of the token-embedding front end of MeshTransformer.forward_on_codes.

*** TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and the cpu_baseline leg of the bench scripts may import this.
The product path (meshgpt_pytorch_b200/frontend.py + csrc/frontend.cu) never does. ***

Follows /root/reference/meshgpt_pytorch/meshgpt_pytorch.py:
    MeshTransformer.__init__      :1140-1165 (codebook_size, eos_token_id, token_embed, quantize_level_embed, vertex_embed,
                                   abs_pos_emb, max_seq_len), :1191-1194 (to_face_tokens), :1254 (pad_id)
    forward_on_codes              :1506-1518  append_eos: the eos id at the first pad position of a one-longer sequence
                                  :1526-1570  pad -> 0, token_embed, + abs_pos_emb, + level embed, + vertex embed, zero-pad to a whole
                                   number of faces, group per face, to_face_tokens on the complete faces
                                  :1591-1594  the sos token(s) in front of the face tokens (no kv cache)

PINNED: `reference_lines()` below executes the reference's OWN statements :1526-1570 (located with `ast` in the reference source,
compiled in memory, nothing copied) with this module standing in for `self`; tests/golden/frontend_own_code.pt holds their
outputs (tests/golden/make_frontend_golden.py) and tests/test_frontend_oracle.py compares the restatement with them - bit for bit,
both are the same torch CPU ops in the same order.  Everything on these lines is the reference's own code on torch / einops
primitives; no third-party arithmetic is involved.
"""
from __future__ import annotations

import ast
import os
from math import ceil
from typing import Tuple

import torch
from torch import Tensor, nn

REFERENCE_ROOT = os.environ.get("MESHTOK_REFERENCE_ROOT", "/root/reference")
FIRST_LINE, LAST_LINE = 1526, 1570
EOS_FIRST_LINE = 1506          # `if append_eos:` (:1506-1518); `if return_loss:` (:1522-1524) lies between and runs with return_loss = False


class OracleTokenFrontEnd(nn.Module):
    """The parameters of MeshTransformer that :1526-1570 touch, under the reference's names."""

    def __init__(self, *, codebook_size: int = 16384, num_quantizers: int = 2, quads: bool = False, dim: int = 512,
                 max_seq_len: int = 8192, pad_id: int = -1, num_sos_tokens: int = 1):
        super().__init__()
        self.sos_token = nn.Parameter(torch.randn(num_sos_tokens, dim))                 # :1148-1152
        self.num_vertices_per_face = 4 if quads else 3                                  # :1131
        self.codebook_size = codebook_size                                              # :1140
        self.num_quantizers = num_quantizers                                            # :1141
        self.eos_token_id = codebook_size                                               # :1143
        assert max_seq_len % (self.num_vertices_per_face * num_quantizers) == 0         # :1156
        self.token_embed = nn.Embedding(codebook_size + 1, dim)                         # :1158
        self.quantize_level_embed = nn.Parameter(torch.randn(num_quantizers, dim))      # :1160
        self.vertex_embed = nn.Parameter(torch.randn(self.num_vertices_per_face, dim))  # :1161
        self.abs_pos_emb = nn.Embedding(max_seq_len, dim)                               # :1163
        self.max_seq_len = max_seq_len                                                  # :1165
        self.to_face_tokens = nn.Sequential(                                            # :1191-1194
            nn.Linear(num_quantizers * self.num_vertices_per_face * dim, dim), nn.LayerNorm(dim))
        self.pad_id = pad_id                                                            # :1254
        self.dim = dim

    @torch.no_grad()
    def append_eos(self, codes: Tensor) -> Tensor:
        """:1506-1518: (b, L) -> (b, L + 1), the eos id at the first pad position of every row."""
        code_lens = ((codes == self.pad_id).cumsum(dim=-1) == 0).sum(dim=-1)             # :1509
        out = torch.nn.functional.pad(codes, (0, 1), value=0)                            # :1511
        out[torch.arange(codes.shape[0]), code_lens] = self.eos_token_id                 # :1513-1518
        return out

    @torch.no_grad()
    def forward(self, codes: Tensor, append_eos: bool = False, prepend_sos: bool = False) -> Tuple[Tensor, Tensor]:
        """codes (b, L) int64 -> (grouped_codes (b, ceil(L / n), n, dim), face_codes (b, L // n, dim)), n = nvf * Q;
        with append_eos L counts the eos; with prepend_sos face_codes carries the sos rows first (:1591-1594)."""
        assert codes.shape[-1] <= self.max_seq_len                                       # :1502
        if append_eos:
            codes = self.append_eos(codes)
        g, f = self._embed(codes)
        if prepend_sos:
            f = torch.cat([self.sos_token[None].expand(f.shape[0], -1, -1), f], dim=1)   # :1593-1594
        return g, f

    @torch.no_grad()
    def _embed(self, codes: Tensor) -> Tuple[Tensor, Tensor]:
        x = codes.masked_fill(codes == self.pad_id, 0)                                   # :1528
        x = self.token_embed(x)                                                          # :1529
        code_len = x.shape[1]
        x = x + self.abs_pos_emb(torch.arange(code_len))                                 # :1533-1535
        q, nvf = self.num_quantizers, self.num_vertices_per_face
        level = self.quantize_level_embed.repeat(ceil(code_len / q), 1)                  # :1541  '(r q) d'
        x = x + level[:code_len]                                                         # :1542
        vert = self.vertex_embed.repeat_interleave(q, dim=0).repeat(ceil(code_len / (nvf * q)), 1)   # :1546  '(r nv q) d'
        x = x + vert[:code_len]                                                          # :1547
        n = q * nvf                                                                      # :1552
        whole = code_len % n == 0                                                        # :1556
        padded = ceil(code_len / n) * n                                                  # :1558
        x = torch.nn.functional.pad(x, (0, 0, 0, padded - code_len))                     # :1560
        grouped = x.reshape(x.shape[0], padded // n, n, self.dim)                        # :1564
        face = grouped if whole else grouped[:, :-1]                                     # :1568
        face = self.to_face_tokens(face.reshape(face.shape[0], face.shape[1], n * self.dim))   # :1569-1570
        return grouped, face


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "meshgpt_pytorch", "meshgpt_pytorch.py"))


@torch.no_grad()
def reference_lines(module: OracleTokenFrontEnd, codes: Tensor, append_eos: bool = False) -> Tuple[Tensor, Tensor]:
    """Runs the reference's statements :1526-1570 of MeshTransformer.forward_on_codes (with append_eos: from :1506 on, i.e. including
    its `if append_eos:` block), as they stand in its source, with `module` as `self`.  Container only (needs /root/reference)."""
    from einops import rearrange, repeat
    path = os.path.join(REFERENCE_ROOT, "meshgpt_pytorch", "meshgpt_pytorch.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MeshTransformer")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward_on_codes")
    first = EOS_FIRST_LINE if append_eos else FIRST_LINE
    body = [st for st in fn.body if first <= st.lineno and st.end_lineno <= LAST_LINE]
    assert body and body[0].lineno == (1506 if append_eos else 1528) and body[-1].end_lineno == LAST_LINE, "the reference source moved: re-derive the line range"
    helpers = {}
    for name in ("divisible_by", "pad_at_dim", "pad_to_length"):
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
        exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), helpers)
    env = dict(helpers, self=module, codes=codes.clone(), device=codes.device, rearrange=rearrange, repeat=repeat, ceil=ceil, torch=torch,
               F=torch.nn.functional, append_eos=append_eos, return_loss=False, batch=codes.shape[0], exists=lambda v: v is not None)
    helpers["F"] = torch.nn.functional
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), env)
    return env["grouped_codes"], env["face_codes"]
