"""Token-embedding front end of the reference's MeshTransformer.forward_on_codes (meshgpt_pytorch.py:1526-1570): the
step that follows tokenization.  codes (b, L) -> the per-code "fine" tokens (token + absolute position + quantizer-level + vertex
embeddings, zero-padded to whole faces and grouped per face) and the per-face "coarse" tokens (to_face_tokens = Linear + LayerNorm
over the n = nvf * Q fine tokens of a face).

    fe = MeshTransformerFrontEnd(autoencoder, dim=512, max_seq_len=8192)      # the transformer's keywords that matter here
    fe.load_reference_state_dict(transformer.state_dict())                     # token_embed.*, abs_pos_emb.*, quantize_level_embed,
    grouped_codes, face_codes = fe.cuda().embed_codes(codes)                   #   vertex_embed, to_face_tokens.{0,1}.*

Both outputs come from csrc/frontend.cu through the C ABI (meshtok_frontend_*); there is no CPU path.  The fine tokens are bit-exact
(four adds in the reference's order).  The face tokens never run the (n * dim -> dim) Linear: it is folded at load time into n
tables token_embed @ W_j^T plus one position row per face index (fp64 products rounded once to fp32), so a face token is n + 1 row
gathers and a LayerNorm - HBM traffic instead of 3.1 MFLOP per face.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor, nn

import ctypes as C

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class MeshTransformerFrontEnd(nn.Module):
    def __init__(self, autoencoder=None, *, dim: int = 512, max_seq_len: int = 8192, pad_id: int = -1, codebook_size: int | None = None,
                 num_quantizers: int | None = None, quads: bool | None = None, num_sos_tokens: int = 1):
        super().__init__()
        if autoencoder is not None:
            codebook_size = autoencoder.codebook_size if codebook_size is None else codebook_size
            num_quantizers = autoencoder.num_quantizers if num_quantizers is None else num_quantizers
            quads = (autoencoder.num_vertices_per_face == 4) if quads is None else quads
        if codebook_size is None or num_quantizers is None:
            raise ValueError("give an autoencoder or codebook_size / num_quantizers / quads")
        self.num_vertices_per_face = 4 if quads else 3                                    # :1131
        self.codebook_size, self.num_quantizers = codebook_size, num_quantizers           # :1140-1141
        self.eos_token_id = codebook_size                                                 # :1143
        n = self.num_vertices_per_face * num_quantizers
        if max_seq_len % n != 0:                                                          # :1156
            raise ValueError(f"max_seq_len ({max_seq_len}) must be divisible by {n}")
        if dim % 128 != 0 or dim > 1024:
            raise NotImplementedError(f"dim {dim}: the face-token kernel needs a multiple of 128, at most 1024")
        self.dim, self.max_seq_len, self.pad_id = dim, max_seq_len, pad_id
        self.token_embed = nn.Embedding(codebook_size + 1, dim)                           # :1158
        self.quantize_level_embed = nn.Parameter(torch.randn(num_quantizers, dim))        # :1160
        self.vertex_embed = nn.Parameter(torch.randn(self.num_vertices_per_face, dim))    # :1161
        self.abs_pos_emb = nn.Embedding(max_seq_len, dim)                                 # :1163
        self.to_face_tokens = nn.Sequential(nn.Linear(n * dim, dim), nn.LayerNorm(dim))   # :1191-1194
        self.sos_token = nn.Parameter(torch.randn(num_sos_tokens, dim))                   # :1148-1152
        for p in self.parameters():
            p.requires_grad_(False)
        self._prep = None
        self._prep_key = None

    def load_reference_state_dict(self, state_dict: Dict[str, Tensor]):
        """Takes the front end's keys out of a MeshTransformer state dict (every other key is ignored)."""
        own = self.state_dict()
        missing = [k for k in own if k not in state_dict]
        if missing:
            raise KeyError(f"state dict lacks front-end keys: {missing}")
        self.load_state_dict({k: state_dict[k] for k in own}, strict=True)
        self._prep = None

    # ---- load time: fold to_face_tokens.0 into per-slot tables ---------------------------------------
    def _prepare(self):
        key = tuple((t.data_ptr(), t._version) for t in self.parameters())
        if self._prep is not None and self._prep_key == key:
            return self._prep
        dev = self.token_embed.weight.device
        if dev.type != "cuda":
            raise RuntimeError("MeshTransformerFrontEnd must live on a CUDA device (call .cuda()); there is no CPU path")
        q, nvf, d = self.num_quantizers, self.num_vertices_per_face, self.dim
        n = q * nvf
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            W = self.to_face_tokens[0].weight.detach().double().view(d, n, d)             # [out, slot j, in]
            tok = self.token_embed.weight.detach().double()
            tables = torch.empty(n, tok.shape[0], d, **f32)
            for j in range(n):
                tables[j] = (tok @ W[:, j].t()).float()                                   # T_j = token_embed @ W_j^T
            # position / level / vertex part of slot j at face f: W_j (abs_pos_emb[n f + j] + level[j % q] + vertex[j // q]), summed over j
            pos = self.abs_pos_emb.weight.detach().double().view(self.max_seq_len // n, n, d)
            slot = (self.quantize_level_embed.detach().double().repeat(nvf, 1)
                    + self.vertex_embed.detach().double().repeat_interleave(q, dim=0))    # (n, d): t = (vertex, level), level fastest
            face_pos = torch.einsum("fjd,ojd->fo", pos + slot[None], W) + self.to_face_tokens[0].bias.detach().double()
            prep = dict(tables=tables.contiguous(), face_pos=face_pos.float().contiguous(),
                        tok=self.token_embed.weight.detach().to(**f32).contiguous(), pos=self.abs_pos_emb.weight.detach().to(**f32).contiguous(),
                        level=self.quantize_level_embed.detach().to(**f32).contiguous(), vert=self.vertex_embed.detach().to(**f32).contiguous(),
                        sos=self.sos_token.detach().to(**f32).contiguous(),
                        gamma=self.to_face_tokens[1].weight.detach().to(**f32).contiguous(),
                        beta=self.to_face_tokens[1].bias.detach().to(**f32).contiguous(), eps=float(self.to_face_tokens[1].eps))
        self._prep, self._prep_key = prep, key
        return prep

    # ---- :1526-1570 -------------------------------------------------------------------------------
    def _check(self, codes: Tensor) -> Tensor:
        if codes.device.type != "cuda":
            raise RuntimeError("codes must be a CUDA tensor; there is no CPU path")
        if codes.ndim != 2:
            raise ValueError("codes must be (batch, seq_len); flatten (b, n, q) codes with 'b n q -> b (n q)' like the reference (:1494)")
        if codes.shape[1] > self.max_seq_len:                                             # :1502
            raise ValueError(f"received codes of length {codes.shape[1]} but needs to be less than or equal to set max_seq_len {self.max_seq_len}")
        return codes.to(torch.int64).contiguous()

    def _source(self, codes, source):
        """(codes pointer, vertex-source struct or None, batch, L, device, keep-alive)."""
        if source is None:
            codes = self._check(codes)
            return ptr(codes), None, codes.shape[0], codes.shape[1], codes.device, codes
        autoencoder, faces = source
        t = autoencoder.taps()                                    # views into the workspace of the tokenizer's last call
        faces = faces.to(torch.int64).contiguous()
        B, F, nvf = faces.shape
        assert t["face_row"].shape == (B, F) and nvf == self.num_vertices_per_face, "faces must be the tensor the autoencoder's last call tokenized"
        vs = _lib.VertexCodeSource()
        vs.faces, vs.face_row, vs.vert_row, vs.vertex_codes = ptr(faces), ptr(t["face_row"]), ptr(t["vert_row"]), ptr(t["vertex_codes"])
        vs.F, vs.nvf, vs.V, vs.Q = F, nvf, t["vert_row"].shape[1], self.num_quantizers
        L = F * nvf * self.num_quantizers
        if L > self.max_seq_len:
            raise ValueError(f"received codes of length {L} but needs to be less than or equal to set max_seq_len {self.max_seq_len}")
        return 0, vs, B, L, faces.device, (faces, t)

    def _eos(self, append_eos, B, dev):
        if not append_eos:
            return -1, None
        return self.eos_token_id, torch.empty(max(B, 1), dtype=torch.int32, device=dev)

    @torch.no_grad()
    def append_eos(self, codes: Tensor) -> Tensor:
        """:1506-1518 on the device: (b, L) -> (b, L + 1) with the eos id at the first pad position of every row."""
        cp, vs, b, L, dev, keep = self._source(codes, None)
        out = torch.empty(b, L + 1, dtype=torch.int64, device=dev)
        lens = torch.empty(max(b, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.meshtok_frontend_append_eos(cp, None, b, L, self.pad_id, self.eos_token_id, ptr(lens), ptr(out), stream_ptr()))
        return out

    @torch.no_grad()
    def fine_tokens(self, codes: Tensor = None, append_eos: bool = False, source=None) -> Tensor:
        """grouped_codes (b, ceil(L / n), n, dim) of :1564 - every code's embedding, zero rows behind an incomplete last face.
        append_eos: of the eos-extended sequence (:1506-1518), without materialising it.  source=(autoencoder, faces): read the codes
        from the tokenizer's per-vertex result instead of a codes tensor (see embed_tokenizer_output)."""
        cp, vs, b, L, dev, keep = self._source(codes, source)
        p = self._prepare()
        n = self.num_quantizers * self.num_vertices_per_face
        eos, lens = self._eos(append_eos, b, dev)
        Le = L + int(append_eos)
        if Le > self.max_seq_len:
            raise ValueError(f"received codes of length {Le} but needs to be less than or equal to set max_seq_len {self.max_seq_len}")
        nfp = -(-Le // n)
        out = torch.empty(b, nfp, n, self.dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.meshtok_frontend_fine_tokens(cp, None if vs is None else C.byref(vs), b, L, self.dim, self.pad_id, eos, ptr(lens),
                                                   self.num_quantizers, self.num_vertices_per_face, self.codebook_size + 1, ptr(p["tok"]),
                                                   ptr(p["pos"]), ptr(p["level"]), ptr(p["vert"]), ptr(out), stream_ptr()))
        return out

    @torch.no_grad()
    def face_tokens(self, codes: Tensor = None, append_eos: bool = False, prepend_sos: bool = False, source=None) -> Tensor:
        """face_codes (b, L // n, dim) of :1570 - one token per complete face; prepend_sos: the sos row(s) first (:1591-1594)."""
        cp, vs, b, L, dev, keep = self._source(codes, source)
        p = self._prepare()
        n = self.num_quantizers * self.num_vertices_per_face
        eos, lens = self._eos(append_eos, b, dev)
        Le = L + int(append_eos)
        n_sos = self.sos_token.shape[0] if prepend_sos else 0
        out = torch.empty(b, n_sos + Le // n, self.dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.meshtok_frontend_face_tokens(cp, None if vs is None else C.byref(vs), b, L, n, self.dim, self.pad_id, eos, ptr(lens),
                                                   self.codebook_size + 1, ptr(p["tables"]), ptr(p["face_pos"]), ptr(p["gamma"]), ptr(p["beta"]),
                                                   p["eps"], ptr(p["sos"]) if n_sos else None, n_sos, ptr(out), stream_ptr()))
        return out

    def embed_codes(self, codes: Tensor, append_eos: bool = False, prepend_sos: bool = False) -> Tuple[Tensor, Tensor]:
        """(grouped_codes, face_codes) as forward_on_codes holds them after :1570 (after :1594 with prepend_sos)."""
        return self.fine_tokens(codes, append_eos), self.face_tokens(codes, append_eos, prepend_sos)

    def embed_tokenizer_output(self, autoencoder, faces: Tensor, append_eos: bool = False, prepend_sos: bool = False) -> Tuple[Tensor, Tensor]:
        """embed_codes for the batch `autoencoder` has just tokenized (its last tokenize / forward call on `faces`, padded layout),
        WITHOUT the (b, nf * nvf, q) int64 code tensor in between: the kernels read the per-vertex codes and the face / vertex maps
        the tokenizer left in its workspace ("codes -> first transformer input" without writing the codes to HBM)."""
        src = (autoencoder, faces)
        return self.fine_tokens(None, append_eos, src), self.face_tokens(None, append_eos, prepend_sos, src)

    forward = embed_codes
