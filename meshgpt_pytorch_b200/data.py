"""Drop-in for meshgpt_pytorch.data.derive_face_edges_from_faces (reference data.py:297-340),
running on the K1 CUDA kernel of libmeshtok.  Same signature, same result (values, order, padding,
int64 dtype); CUDA tensors only - there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

_INTERNAL_PAD = -1


def _normalise_vertex_ids(faces64: Tensor, pad_id: int):
    """The kernel wants vertex ids in [0, V) and a pad marker outside that range.  Ids only matter up to
    equality (data.py:324), so anything else (negative non-pad ids, huge sparse ids, a non-negative
    pad_id) is renumbered with torch.unique first.  Returns (faces, kernel_pad_id, V)."""
    is_pad = faces64 == pad_id
    if bool(is_pad.all()):
        return torch.full_like(faces64, _INTERNAL_PAD), _INTERNAL_PAD, 1
    non_pad = faces64[~is_pad]
    lo, hi = int(non_pad.min()), int(non_pad.max())
    budget = faces64.shape[1] * faces64.shape[2] + 64      # at most nf*nvf distinct ids exist
    if pad_id < 0 and lo >= 0 and hi < budget:
        return faces64, pad_id, hi + 1
    uniq, inv = torch.unique(non_pad, return_inverse=True)
    out = torch.full_like(faces64, _INTERNAL_PAD)
    out[~is_pad] = inv
    return out, _INTERNAL_PAD, int(uniq.numel())


def derive_face_edges_from_faces(
    faces: Tensor,
    pad_id: int = -1,
    neighbor_if_share_one_vertex: bool = False,
    include_self: bool = True,
) -> Tensor:
    """faces (b, nf, nvf) or (nf, nvf) integer, padded with pad_id -> (b, e, 2) / (e, 2) int64.

    (i, j) is listed iff both faces are un-padded and at least 2 (1 with neighbor_if_share_one_vertex)
    slots of face i hold a vertex id that occurs in face j; include_self=False drops pairs with exactly 3
    such slots; rows are ordered (i, then j) and padded with pad_id to the longest mesh of the batch.
    """
    if not faces.is_cuda:
        raise RuntimeError("meshgpt_pytorch_b200.derive_face_edges_from_faces needs a CUDA tensor (no CPU path)")
    one_mesh = faces.ndim == 2
    if one_mesh:
        faces = faces[None]
    assert faces.ndim == 3 and faces.shape[-1] in (3, 4), "faces must be (b, nf, 3|4)"
    B, F, nvf = faces.shape
    dev = faces.device
    if B == 0 or F == 0:
        out = torch.full((B, 0, 2), pad_id, dtype=torch.int64, device=dev)
        return out[0] if one_mesh else out
    with torch.cuda.device(dev):
        faces64, kpad, V = _normalise_vertex_ids(faces.long().contiguous(), int(pad_id))
        nbytes = C.c_size_t()
        check(lib.meshtok_topology_workspace_bytes(B, F, nvf, V, 0, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        thr = 1 if neighbor_if_share_one_vertex else 2
        check(lib.meshtok_face_edges_count(ptr(faces64), B, F, nvf, V, kpad, thr, int(bool(include_self)), ptr(ws),
                                           nbytes.value, ptr(counts), stream_ptr()))
        emax = int(counts.max())           # the one unavoidable sync: the output shape is data dependent
        out = torch.empty((B, emax, 2), dtype=torch.int64, device=dev)
        check(lib.meshtok_face_edges_fill(ptr(faces64), B, F, nvf, V, kpad, thr, int(bool(include_self)), ptr(ws),
                                          nbytes.value, emax, ptr(out), stream_ptr()))
        if kpad != pad_id and emax > 0:
            tail = torch.arange(emax, device=dev)[None, :] >= counts[:, None]
            out[tail] = pad_id
    return out[0] if one_mesh else out


# ---------------------------------------------------------------------------------------------------
# bulk producer of the reference's on-disk face-edge cache
# ---------------------------------------------------------------------------------------------------

FACE_EDGES_CACHE_FILE = "cache.face_edges_embed.memmap.npy"
IS_CACHED_FILE = "cache.is_cached.memmap.npy"


def write_face_edges_cache(cache_path, dataset_len: int, start: int, face_edges: Tensor, max_edges_len: int,
                           pad_id: int = -1, assert_edge_len_lt_max: bool = True, memmap_file_mode: str = "r+"):
    """Stores the edge lists of meshes [start, start + b) in the cache the reference's
    `cache_face_edges_for_dataset` decorator reads (meshgpt_pytorch/data.py:158-257): `cache.face_edges_embed.memmap.npy`
    (dataset_len, max_edges_len, 2) **float32** (the reference's dtype, data.py:191) and `cache.is_cached.memmap.npy`
    (dataset_len,) bool.  face_edges: (b, e, 2) integer, rows padded with pad_id (what derive_face_edges_from_faces
    returns); every mesh is padded or cut to max_edges_len rows like the reference's pad_or_slice_to (data.py:205-210)."""
    import os
    import numpy as np
    from numpy.lib.format import open_memmap

    os.makedirs(cache_path, exist_ok=True)
    fe_file = os.path.join(cache_path, FACE_EDGES_CACHE_FILE)
    ic_file = os.path.join(cache_path, IS_CACHED_FILE)
    mode = memmap_file_mode if (os.path.isfile(fe_file) and os.path.isfile(ic_file)) else "w+"
    if mode == "w+":
        cache = open_memmap(fe_file, mode="w+", dtype="float32", shape=(dataset_len, max_edges_len, 2))
        cached = open_memmap(ic_file, mode="w+", dtype="bool", shape=(dataset_len,))
    else:
        cache = open_memmap(fe_file, mode=mode)
        cached = open_memmap(ic_file, mode=mode)
        assert cache.shape == (dataset_len, max_edges_len, 2) and cached.shape == (dataset_len,), "cache shape mismatch"
    fe = face_edges.detach().cpu()
    b, e, _ = fe.shape
    assert 0 <= start and start + b <= dataset_len
    if assert_edge_len_lt_max:
        edge_len = (fe != pad_id).all(-1).sum(-1)
        worst = int(edge_len.max()) if b else 0
        assert worst <= max_edges_len, f"a mesh has {worst} edges which exceeds the cache length of {max_edges_len}"
    out = torch.full((b, max_edges_len, 2), float(pad_id), dtype=torch.float32)
    keep = min(e, max_edges_len)
    out[:, :keep] = fe[:, :keep].to(torch.float32)
    cache[start:start + b] = out.numpy()
    cached[start:start + b] = True
    cache.flush()
    cached.flush()
    return fe_file, ic_file


def cache_face_edges(faces: Tensor, cache_path, dataset_len: int, start: int, max_edges_len: int, pad_id: int = -1,
                     assert_edge_len_lt_max: bool = True):
    """derive_face_edges_from_faces on the GPU for a whole padded batch (b, nf, nvf) + write_face_edges_cache: fills the
    reference's face-edge cache for meshes [start, start + b) in one go instead of one Python call per __getitem__."""
    edges = derive_face_edges_from_faces(faces, pad_id=pad_id)
    return write_face_edges_cache(cache_path, dataset_len, start, edges, max_edges_len, pad_id, assert_edge_len_lt_max)


# ---------------------------------------------------------------------------------------------------
# ragged batches: a CSR of meshes instead of the reference's pad_id padding
# ---------------------------------------------------------------------------------------------------
def ragged_collate(data, keys=("vertices", "faces")):
    """Counterpart of the reference's custom_collate (meshgpt_pytorch/data.py:347-369) that does not pad: a list of meshes (dicts with
    'vertices' (nv_i, 3) and 'faces' (nf_i, nvf), or (vertices, faces) tuples) -> dict(vertices (sum nv, 3), faces (sum nf, nvf),
    face_offsets (B + 1) int32, vertex_offsets (B + 1) int32), the arguments of MeshAutoencoder.tokenize_ragged.  Vertex ids stay
    mesh-local.  Offsets are CPU tensors (the loader knows the sizes; keeping them on the host spares tokenize_ragged a device sync)."""
    vs, fs = [], []
    for item in data:
        v, f = (item[keys[0]], item[keys[1]]) if isinstance(item, dict) else (item[0], item[1])
        assert v.ndim == 2 and v.shape[1] == 3 and f.ndim == 2, "a mesh is vertices (nv, 3) and faces (nf, nvf)"
        vs.append(v)
        fs.append(f)
    nvf = fs[0].shape[1] if fs else 3
    face_offsets = torch.zeros(len(fs) + 1, dtype=torch.int32)
    vertex_offsets = torch.zeros(len(vs) + 1, dtype=torch.int32)
    if fs:
        face_offsets[1:] = torch.tensor([f.shape[0] for f in fs]).cumsum(0)
        vertex_offsets[1:] = torch.tensor([v.shape[0] for v in vs]).cumsum(0)
    return dict(vertices=torch.cat(vs) if vs else torch.zeros(0, 3), faces=torch.cat(fs) if fs else torch.zeros(0, nvf, dtype=torch.long),
                face_offsets=face_offsets, vertex_offsets=vertex_offsets)


def padded_to_ragged(vertices: Tensor, faces: Tensor, pad_id: int = -1, num_vertices=None):
    """A custom_collate batch (vertices (b, nv, 3), faces (b, nf, nvf), trailing pad_id padding, data.py:347-369) -> the ragged_collate
    form of the same meshes.  A mesh's faces are its leading rows without a pad_id entry (the reference's face mask, meshgpt_pytorch.py:
    996, on trailing padding); its vertices are the first num_vertices[i] rows (default: up to the largest vertex id its faces use)."""
    assert vertices.ndim == 3 and faces.ndim == 3 and vertices.shape[0] == faces.shape[0]
    meshes = []
    for i in range(faces.shape[0]):
        valid = (faces[i] != pad_id).all(dim=-1)
        nf = int(valid.sum())
        assert bool(valid[:nf].all()), "padding must be trailing (data.py:358)"
        f = faces[i, :nf]
        nv = int(num_vertices[i]) if num_vertices is not None else (int(f.max()) + 1 if nf else 0)
        meshes.append((vertices[i, :nv], f))
    return ragged_collate(meshes)


def ragged_to_padded_codes(codes: Tensor, face_offsets, nvf: int, pad_id: int = -1) -> Tensor:
    """codes (sum nf * nvf, Q) from tokenize_ragged -> (b, max nf * nvf, Q) padded with pad_id: the tensor the reference's tokenize
    returns for the custom_collate batch of the same meshes (meshgpt_pytorch.py:861-866)."""
    fo = torch.as_tensor(face_offsets).long().cpu()
    n = (fo[1:] - fo[:-1]) * nvf
    B, T = n.numel(), (int(n.max()) if n.numel() else 0)
    out = torch.full((B, T, codes.shape[-1]), pad_id, dtype=codes.dtype, device=codes.device)
    for b in range(B):
        out[b, : int(n[b])] = codes[int(fo[b]) * nvf: int(fo[b + 1]) * nvf]
    return out


def ragged_slice(batch: dict, lo: int, hi: int) -> dict:
    """Meshes [lo, hi) of a ragged_collate batch as a batch of their own (offsets re-based to 0; tensors are views).  With
    sharding.balanced_contiguous_shards over the face counts this is a rank's share of a ragged collection (meshes shard by
    contiguous blocks, no activation crosses a link)."""
    fo, vo = batch["face_offsets"], batch["vertex_offsets"]
    assert 0 <= lo <= hi <= fo.numel() - 1
    f0, f1, v0, v1 = int(fo[lo]), int(fo[hi]), int(vo[lo]), int(vo[hi])
    return dict(vertices=batch["vertices"][v0:v1], faces=batch["faces"][f0:f1],
                face_offsets=(fo[lo: hi + 1] - f0).to(torch.int32), vertex_offsets=(vo[lo: hi + 1] - v0).to(torch.int32))
