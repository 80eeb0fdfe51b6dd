"""
CPU oracle for the MeshGPT mesh-tokenization hot path.

*** TEST INFRASTRUCTURE - NOT PRODUCT CODE. ***
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker / CPU baseline.  The product
(meshgpt_pytorch_b200) never imports it and has no CPU path.

What it is: a plain fp32 torch-CPU restatement of
    MeshAutoencoder.tokenize -> forward(return_codes=True) -> encode + quantize
of lucidrains/meshgpt-pytorch v1.7.2 (reference paths relative to /root/reference):
    derive_face_edges_from_faces   meshgpt_pytorch/data.py:297-340
    get_derived_face_features      meshgpt_pytorch/meshgpt_pytorch.py:151-183
    discretize                     meshgpt_pytorch/meshgpt_pytorch.py:187-201
    scatter_mean                   meshgpt_pytorch/meshgpt_pytorch.py:243-257
    MeshAutoencoder.encode         meshgpt_pytorch/meshgpt_pytorch.py:686-785
    MeshAutoencoder.quantize       meshgpt_pytorch/meshgpt_pytorch.py:787-870
    tokenize / forward             meshgpt_pytorch/meshgpt_pytorch.py:949-1019

PARITY PIN STATUS
  * own-code stages (the five helper functions above): PINNED - checked against the
    reference's own functions executed from /root/reference source by
    oracle/reference_extract.py (tests/golden/make_golden.py freezes the outputs as
    fixtures; tests/test_oracle_golden.py re-checks them without /root/reference).
  * third-party stages: PARITY UNPINNED.  torch_geometric.SAGEConv,
    vector_quantize_pytorch.ResidualVQ / ResidualLFQ and einx.get_at are un-vendored pip
    dependencies (setup.py:22-44, lower bounds only: vector-quantize-pytorch>=1.18.1,
    einx>=0.3.0, torch_geometric unpinned) that are neither under /root/reference nor
    installable here (no network).  Their eval-mode arithmetic is restated below from the
    published algorithms; the reference's own tests (tests/test_meshgpt.py) pin no numbers
    on this path, so there is no golden vector to anchor them.  Each restated function says
    exactly what it assumes.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

# --------------------------------------------------------------------------------------
# (a1) face adjacency                                   reference: data.py:297-340
# --------------------------------------------------------------------------------------


def derive_face_edges_from_faces(
    faces: Tensor,
    pad_id: int = -1,
    neighbor_if_share_one_vertex: bool = False,
    include_self: bool = True,
) -> Tensor:
    """Directed face-adjacency list, (b, e, 2) int64 padded with pad_id.

    Semantics (data.py:322-333): for faces i, j of one mesh,
        n_shared[i, j] = #{ slots c of face i : face_i[c] occurs anywhere in face j }
    and (i, j) is an edge iff n_shared >= (1 if neighbor_if_share_one_vertex else 2) and
    both faces are un-padded (a face is padding if ANY slot equals pad_id, data.py:317).
    include_self=False removes every pair with n_shared == 3 (data.py:329-330).  Pairs are
    listed row-major (i ascending, then j ascending; data.py:332) and meshes are padded to
    the batch max with pad_id (data.py:335).  Output is int64 whatever the input dtype.

    Restated with a face-by-vertex membership table instead of the reference's
    (F,F,nvf,nvf) broadcast compare.
    """
    one_mesh = faces.ndim == 2
    if one_mesh:
        faces = faces[None]
    faces = faces.long()
    batch, nf, nvf = faces.shape
    need = 1 if neighbor_if_share_one_vertex else 2

    per_mesh = []
    for b in range(batch):
        fb = faces[b]
        ok = (fb != pad_id).all(dim=-1)                      # (F,)
        # remap vertex ids (any integers) to 0..U-1 so the table stays small
        uniq, inv = torch.unique(fb, return_inverse=True)    # inv: (F, nvf)
        member = torch.zeros(nf, uniq.numel(), dtype=torch.bool)
        member[torch.arange(nf)[:, None].expand_as(inv), inv] = True     # member[j, u]
        # n_shared[i, j] = sum_c member[j, inv[i, c]]
        n_shared = torch.zeros(nf, nf, dtype=torch.int64)
        for c in range(nvf):
            n_shared += member[:, inv[:, c]].t().long()
        keep = (n_shared >= need) & ok[:, None] & ok[None, :]
        if not include_self:
            keep &= n_shared != 3
        per_mesh.append(keep.nonzero())                       # row-major (i, j), int64

    emax = max((e.shape[0] for e in per_mesh), default=0)
    out = torch.full((batch, emax, 2), pad_id, dtype=torch.int64)
    for b, e in enumerate(per_mesh):
        out[b, : e.shape[0]] = e
    return out[0] if one_mesh else out


# --------------------------------------------------------------------------------------
# (a2) derived face features              reference: meshgpt_pytorch.py:94-95, 151-183
# --------------------------------------------------------------------------------------


def _unit(t: Tensor) -> Tensor:
    # F.normalize(p=2, dim=-1, eps=1e-12)  (meshgpt_pytorch.py:94-95)
    return F.normalize(t, p=2, dim=-1)


def derived_face_features(face_coords: Tensor) -> Dict[str, Tensor]:
    """face_coords (b, nf, nvf, 3) fp32 -> angles (b,nf,nvf), area (b,nf,1), normals (b,nf,3).

    angles[k] = acos(clip(<unit(p_k), unit(p_{k-1})>, -1+1e-5, 1-1e-5)) - the angle between
    the two POSITION vectors seen from the origin (meshgpt_pytorch.py:151-153, 164-166).
    prev = one cyclic step back for triangles, two for quads (:168-170); the first two rows of
    (p - prev) are the edge vectors (:172); cross -> normal (unit) and area = |cross|/2 (:174-177).
    """
    nvf = face_coords.shape[-2]
    back1 = torch.roll(face_coords, shifts=1, dims=2)
    cosang = (_unit(face_coords) * _unit(back1)).sum(dim=-1)
    angles = torch.arccos(cosang.clamp(-1 + 1e-5, 1 - 1e-5))
    prev = back1 if nvf == 3 else torch.roll(back1, shifts=1, dims=2)
    diff = face_coords - prev
    cr = torch.cross(diff[:, :, 0], diff[:, :, 1], dim=-1)
    return dict(angles=angles, area=cr.norm(dim=-1, keepdim=True) * 0.5, normals=_unit(cr))


# --------------------------------------------------------------------------------------
# (a3) discretize                                 reference: meshgpt_pytorch.py:187-201
# --------------------------------------------------------------------------------------


def discretize(t: Tensor, lo: float, hi: float, num_discrete: int = 128) -> Tensor:
    """fp32, in exactly this order (meshgpt_pytorch.py:197-201): (t-lo)/(hi-lo); *n; -0.5;
    round (half-to-even); .long(); clamp(0, n-1)."""
    assert hi > lo
    u = (t - lo) / (hi - lo)
    u = u * num_discrete
    u = u - 0.5
    return u.round().long().clamp(min=0, max=num_discrete - 1)


# --------------------------------------------------------------------------------------
# (a6) scatter_mean                               reference: meshgpt_pytorch.py:243-257
# --------------------------------------------------------------------------------------


def scatter_mean_rows(num_rows: int, index: Tensor, src: Tensor, eps: float = 1e-5) -> Tensor:
    """index (b, n) int64 rows, src (b, n, d) -> (b, num_rows, d): sum / clamp(count, eps).
    Rows nobody writes stay 0 (meshgpt_pytorch.py:255-257)."""
    b, n, d = src.shape
    idx = index[:, :, None].expand(b, n, d)
    total = torch.zeros(b, num_rows, d, dtype=src.dtype).scatter_add(1, idx, src)
    count = torch.zeros(b, num_rows, d, dtype=src.dtype).scatter_add(1, idx, torch.ones_like(src))
    return total / count.clamp(min=eps)


# --------------------------------------------------------------------------------------
# (a5) SAGEConv(normalize=True, project=True)   THIRD PARTY - PARITY UNPINNED
#      call sites meshgpt_pytorch.py:543, 553-557, 764, 769; kwargs :471-474
# --------------------------------------------------------------------------------------


class SAGEConvParams(nn.Module):
    """Parameter container with torch_geometric.nn.SAGEConv's state-dict key names
    (lin, lin_l, lin_r).  Assumed upstream eval arithmetic (torch_geometric, unpinned):
        xp   = relu(lin(x))                              (project=True)
        agg  = mean over edges (s -> t) of xp[s],  s = edge_index[0], t = edge_index[1]
               (sum / clamp(count, min=1); isolated targets get 0)
        out  = lin_l(agg) + lin_r(x)                     (root weight uses the UN-projected x;
                                                          lin_l has the bias, lin_r has none)
        out  = F.normalize(out, p=2, dim=-1)             (normalize=True, eps 1e-12)
    """

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.lin = nn.Linear(dim_in, dim_in, bias=True)
        self.lin_l = nn.Linear(dim_in, dim_out, bias=True)
        self.lin_r = nn.Linear(dim_in, dim_out, bias=False)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        src, dst = edge_index[0], edge_index[1]
        xp = F.relu(self.lin(x))
        n, d = x.shape
        total = torch.zeros(n, d, dtype=x.dtype).index_add_(0, dst, xp[src])
        count = torch.zeros(n, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.shape[0], dtype=x.dtype))
        agg = total / count.clamp(min=1)[:, None]
        out = self.lin_l(agg) + self.lin_r(x)
        return F.normalize(out, p=2, dim=-1)


# --------------------------------------------------------------------------------------
# (a9) ResidualLFQ                               THIRD PARTY - PARITY UNPINNED
#      ctor meshgpt_pytorch.py:579-586, call :842
# --------------------------------------------------------------------------------------


class ResidualLFQParams(nn.Module):
    """vector_quantize_pytorch.ResidualLFQ(dim, num_quantizers, codebook_size, soft_clamp_input_value=10.)
    eval-mode restatement (assumed, >=1.18.1):
        bits = log2(codebook_size); x = project_in(v)  (Linear dim->bits)
        level l: scale = 2^-l ; clamp value c_l = 10 * 2^-l
                 y = tanh(r / c_l) * c_l            (sign preserving; only feeds the sign test)
                 bit_d = y_d > 0 (strict)          code = sum_d bit_d * 2^(bits-1-d)   (dim 0 = MSB)
                 q_d = +scale if bit_d else -scale
                 r <- r - q                         (residual uses the un-clamped r)
        quantized_out = project_out(sum_l q_l);  the mask argument only affects training losses.
    """

    def __init__(self, dim: int, num_quantizers: int, codebook_size: int, soft_clamp: float = 10.0):
        super().__init__()
        bits = int(round(math.log2(codebook_size)))
        assert 2 ** bits == codebook_size
        self.bits, self.num_quantizers, self.soft_clamp = bits, num_quantizers, soft_clamp
        self.project_in = nn.Linear(dim, bits)
        self.project_out = nn.Linear(bits, dim)

    def forward(self, v: Tensor) -> Tuple[Tensor, Tensor]:
        r = self.project_in(v)
        weights = 2 ** torch.arange(self.bits - 1, -1, -1, dtype=torch.int64)
        qsum = torch.zeros_like(r)
        codes = []
        for level in range(self.num_quantizers):
            scale = 2.0 ** -level
            c = self.soft_clamp * scale
            y = torch.tanh(r / c) * c
            bit = y > 0
            codes.append((bit.long() * weights).sum(dim=-1))
            q = torch.where(bit, torch.full_like(r, scale), torch.full_like(r, -scale))
            r = r - q
            qsum = qsum + q
        return self.project_out(qsum), torch.stack(codes, dim=-1)


# --------------------------------------------------------------------------------------
# (a8) ResidualVQ (shared codebook)              THIRD PARTY - PARITY UNPINNED
#      ctor meshgpt_pytorch.py:588-598, call :842
# --------------------------------------------------------------------------------------


class _Codebook(nn.Module):
    def __init__(self, codebook_size: int, dim: int):
        super().__init__()
        self.register_buffer("embed", torch.zeros(1, codebook_size, dim))
        self.register_buffer("embed_avg", torch.zeros(1, codebook_size, dim))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("initted", torch.tensor([True]))


class _VQLayer(nn.Module):
    def __init__(self, codebook: _Codebook):
        super().__init__()
        self._codebook = codebook


class ResidualVQParams(nn.Module):
    """vector_quantize_pytorch.ResidualVQ(dim, num_quantizers, codebook_size, shared_codebook=True)
    eval-mode restatement (assumed, >=1.18.1; Euclidean codebook, heads=1, no projections):
        level l, same codebook e (K, D) at every level:
            d2   = |r|^2 + |e|^2 - 2 r.e          (fp32)
            dist = -sqrt(clamp(d2, min=0))
            idx  = argmax(dist)                   (lowest index wins ties; no gumbel noise in eval)
            q    = e[idx]
            masked rows (mask False): q <- r (so the residual becomes 0), idx <- -1
            r   <- r - q
        returns (sum_l q_l, indices (.., Q) int64).
    The codebook must be supplied explicitly (upstream k-means-initialises on the first batch when
    `initted` is False - SURVEY hard part 5); `initted` is stored True.
    State-dict keys: layers.{l}._codebook.{embed (1,K,D), embed_avg, cluster_size, initted}, aliased.
    """

    def __init__(self, dim: int, num_quantizers: int, codebook_size: int):
        super().__init__()
        shared = _Codebook(codebook_size, dim)
        self.layers = nn.ModuleList([_VQLayer(shared) for _ in range(num_quantizers)])
        self.num_quantizers = num_quantizers

    @property
    def codebook(self) -> Tensor:
        return self.layers[0]._codebook.embed[0]

    def forward(self, v: Tensor, mask: Optional[Tensor] = None, chunk: int = 2048) -> Tuple[Tensor, Tensor]:
        e = self.codebook
        e2 = (e * e).sum(dim=-1)
        lead = v.shape[:-1]
        r = v.reshape(-1, v.shape[-1]).clone()
        flat_mask = None if mask is None else mask.reshape(-1)
        qsum = torch.zeros_like(r)
        all_idx = []
        for _ in range(self.num_quantizers):
            idx = torch.empty(r.shape[0], dtype=torch.int64)
            for s in range(0, r.shape[0], chunk):            # chunked only to bound memory
                rc = r[s : s + chunk]
                d2 = (rc * rc).sum(dim=-1, keepdim=True) + e2[None, :] + (rc @ e.t()) * -2
                idx[s : s + chunk] = (-d2.clamp(min=0).sqrt()).argmax(dim=-1)
            q = e[idx]
            if flat_mask is not None:
                q = torch.where(flat_mask[:, None], q, r)
                idx = torch.where(flat_mask, idx, torch.full_like(idx, -1))
            r = r - q
            qsum = qsum + q
            all_idx.append(idx)
        return qsum.reshape(*lead, -1), torch.stack(all_idx, dim=-1).reshape(*lead, -1)


# --------------------------------------------------------------------------------------
# the autoencoder's tokenization half            reference: meshgpt_pytorch.py:431-602
# --------------------------------------------------------------------------------------


class OracleMeshAutoencoder(nn.Module):
    """Encoder + quantizer half of MeshAutoencoder with the reference's state-dict key names.

    Only the constructor arguments that influence tokenize() are accepted; the decoder, the loss
    and the optional attention blocks (attn_encoder_depth must be 0, meshgpt_pytorch.py:477) are
    out of scope.
    """

    def __init__(
        self,
        num_discrete_coors: int = 128,
        coor_continuous_range: Tuple[float, float] = (-1.0, 1.0),
        dim_coor_embed: int = 64,
        num_discrete_area: int = 128,
        dim_area_embed: int = 16,
        num_discrete_normals: int = 128,
        dim_normal_embed: int = 64,
        num_discrete_angle: int = 128,
        dim_angle_embed: int = 16,
        encoder_dims_through_depth: Tuple[int, ...] = (64, 128, 256, 256, 576),
        dim_codebook: int = 192,
        num_quantizers: int = 2,
        codebook_size: int = 16384,
        use_residual_lfq: bool = True,
        pad_id: int = -1,
        quads: bool = False,
        attn_encoder_depth: int = 0,
    ):
        super().__init__()
        assert attn_encoder_depth == 0, "encoder attention blocks are out of scope"
        self.nvf = 4 if quads else 3
        self.pad_id = pad_id
        self.num_discrete_coors = num_discrete_coors
        self.num_discrete_area = num_discrete_area
        self.num_discrete_normals = num_discrete_normals
        self.num_discrete_angle = num_discrete_angle
        self.coor_range = tuple(float(x) for x in coor_continuous_range)
        self.dim_codebook = dim_codebook
        self.num_quantizers = num_quantizers
        self.codebook_size = codebook_size
        self.use_residual_lfq = use_residual_lfq

        self.coor_embed = nn.Embedding(num_discrete_coors, dim_coor_embed)
        self.angle_embed = nn.Embedding(num_discrete_angle, dim_angle_embed)
        self.area_embed = nn.Embedding(num_discrete_area, dim_area_embed)
        self.normal_embed = nn.Embedding(num_discrete_normals, dim_normal_embed)

        init_dim = dim_coor_embed * 3 * self.nvf + dim_angle_embed * self.nvf + dim_normal_embed * 3 + dim_area_embed
        self.project_in = nn.Linear(init_dim, dim_codebook)

        first_dim, *rest = encoder_dims_through_depth
        self.init_sage_conv = SAGEConvParams(dim_codebook, first_dim)
        self.init_encoder_act_and_norm = nn.Sequential(nn.SiLU(), nn.LayerNorm(first_dim))
        self.encoders = nn.ModuleList()
        cur = first_dim
        for d in rest:
            self.encoders.append(SAGEConvParams(cur, d))
            cur = d
        self.project_dim_codebook = nn.Linear(cur, dim_codebook * self.nvf)

        if use_residual_lfq:
            self.quantizer = ResidualLFQParams(dim_codebook, num_quantizers, codebook_size)
        else:
            self.quantizer = ResidualVQParams(dim_codebook, num_quantizers, codebook_size)

    # ---- (a4) encode: meshgpt_pytorch.py:686-785 ------------------------------------
    @torch.no_grad()
    def encode(self, *, vertices, faces, face_edges, face_mask, face_edges_mask,
               return_face_coordinates=False, return_stages=False):
        b, nf, nvf = faces.shape
        assert nvf == self.nvf
        lo, hi = self.coor_range
        safe_faces = faces.masked_fill(~face_mask[:, :, None], 0).long()                  # :711
        bidx = torch.arange(b)[:, None, None]
        face_coords = vertices[bidx, safe_faces]                                          # :715  (b,nf,nvf,3)

        feats = derived_face_features(face_coords)                                        # :719
        d_angle = discretize(feats["angles"], 0.0, math.pi, self.num_discrete_angle)      # :721
        d_area = discretize(feats["area"], 0.0, (hi - lo) ** 2, self.num_discrete_area)   # :724
        d_normal = discretize(feats["normals"], lo, hi, self.num_discrete_normals)        # :727
        d_coords = discretize(face_coords, lo, hi, self.num_discrete_coors).reshape(b, nf, nvf * 3)   # :732-733

        parts = [
            self.coor_embed(d_coords).reshape(b, nf, -1),                                 # :735-736
            self.angle_embed(d_angle).reshape(b, nf, -1),
            self.area_embed(d_area).reshape(b, nf, -1),
            self.normal_embed(d_normal).reshape(b, nf, -1),
        ]
        x = self.project_in(torch.cat(parts, dim=-1))                                     # :740-741

        counts = face_mask.long().sum(dim=1)                                              # :748-750
        offsets = torch.cumsum(counts, 0) - counts
        edges = (face_edges.long() + offsets[:, None, None])[face_edges_mask]             # :752-753
        edge_index = edges.t().contiguous()                                               # :754  row0=i, row1=j

        x = x[face_mask]                                                                  # :760
        stages = {"project_in": x.clone(), "edge_index": edge_index}
        x = self.init_sage_conv(x, edge_index)                                            # :764
        stages["sage0"] = x.clone()
        x = self.init_encoder_act_and_norm(x)                                             # :766
        stages["sage0_act_norm"] = x.clone()
        for li, conv in enumerate(self.encoders):                                         # :768-769
            x = conv(x, edge_index)
            stages[f"sage{li + 1}"] = x.clone()

        out = torch.zeros(b, nf, x.shape[-1], dtype=x.dtype)                              # :771-773
        out[face_mask] = x
        ret = (out, d_coords) if return_face_coordinates else out
        if return_stages:
            stages.update(d_angle=d_angle, d_area=d_area, d_normal=d_normal, d_coords=d_coords)
            return ret, stages
        return ret

    # ---- (a7) quantize: meshgpt_pytorch.py:787-870 ----------------------------------
    @torch.no_grad()
    def quantize(self, *, faces, face_mask, face_embed, pad_id=None, return_stages=False):
        pad_id = self.pad_id if pad_id is None else pad_id
        b, nf, nvf = faces.shape
        num_vertices = int(faces.amax().item()) + 1                                       # :800-801
        y = self.project_dim_codebook(face_embed).reshape(b, nf * nvf, self.dim_codebook) # :803-804, 820
        vidx = faces.long().masked_fill(~face_mask[:, :, None], num_vertices)             # :814 pad vertex row
        averaged = scatter_mean_rows(num_vertices + 1, vidx.reshape(b, nf * nvf), y)      # :818-824
        mask = torch.ones(b, num_vertices + 1, dtype=torch.bool)                          # :828-829
        mask[:, -1] = False
        if self.use_residual_lfq:
            quantized, codes = self.quantizer(averaged)
        else:
            quantized, codes = self.quantizer(averaged, mask=mask)
        bidx = torch.arange(b)[:, None, None]
        face_embed_out = quantized[bidx, vidx].reshape(b, nf, nvf * self.dim_codebook)    # :856
        codes_out = codes[bidx, vidx].reshape(b, nf * nvf, -1)                            # :861
        tok_mask = face_mask[:, :, None].expand(b, nf, nvf).reshape(b, nf * nvf, 1)
        codes_out = codes_out.masked_fill(~tok_mask, self.pad_id)                         # :865-866 (uses self.pad_id)
        commit_loss = torch.zeros(self.num_quantizers)
        if return_stages:
            return (face_embed_out, codes_out, commit_loss), dict(averaged_vertices=averaged, vertex_codes=codes)
        return face_embed_out, codes_out, commit_loss

    # ---- (a10) forward(return_codes=True) / tokenize: meshgpt_pytorch.py:949-1019 ----
    @torch.no_grad()
    def forward_codes(self, *, vertices, faces, face_edges=None, return_stages=False):
        if face_edges is None:
            face_edges = derive_face_edges_from_faces(faces, pad_id=self.pad_id)          # :991-992
        face_mask = (faces != self.pad_id).all(dim=-1)                                    # :996
        face_edges_mask = (face_edges != self.pad_id).all(dim=-1)                         # :997
        enc = self.encode(vertices=vertices, faces=faces, face_edges=face_edges, face_mask=face_mask,
                          face_edges_mask=face_edges_mask, return_face_coordinates=True,
                          return_stages=return_stages)
        stages = {}
        if return_stages:
            (encoded, d_coords), stages = enc
        else:
            encoded, d_coords = enc
        q = self.quantize(face_embed=encoded, faces=faces, face_mask=face_mask, return_stages=return_stages)
        if return_stages:
            (_, codes, _), qstages = q
            stages.update(qstages)
            stages["encoded"] = encoded
            stages["face_edges"] = face_edges
        else:
            _, codes, _ = q
        nvf = self.nvf
        tok_mask = face_mask[:, :, None].expand(-1, -1, nvf).reshape(faces.shape[0], -1, 1)
        codes = codes.masked_fill(~tok_mask, self.pad_id)                                 # :1018
        return (codes, stages) if return_stages else codes

    @torch.no_grad()
    def tokenize(self, vertices, faces, face_edges=None):
        given = [t for t in (vertices, faces, face_edges) if t is not None]
        ndims = {t.ndim for t in given}
        assert len(ndims) == 1                                                            # :957
        batchless = ndims.pop() == 2
        if batchless:
            vertices, faces = vertices[None], faces[None]
            face_edges = None if face_edges is None else face_edges[None]
        self.eval()                                                                       # :965
        codes = self.forward_codes(vertices=vertices, faces=faces, face_edges=face_edges)
        return codes[0] if batchless else codes


def codebook_usage_histogram(codes: Tensor, num_quantizers: int, codebook_size: int, pad_id: int = -1) -> Tensor:
    """hist[q, c] = number of non-pad output tokens whose level-q code is c (SURVEY section 8e;
    new feature, no reference counterpart).  codes (..., Q) int64 -> (Q, K) int64."""
    flat = codes.reshape(-1, num_quantizers)
    out = torch.zeros(num_quantizers, codebook_size, dtype=torch.int64)
    for q in range(num_quantizers):
        col = flat[:, q]
        col = col[col != pad_id]
        out[q] = torch.bincount(col, minlength=codebook_size)
    return out
