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
