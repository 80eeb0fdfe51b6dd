"""Synthetic meshes of the shapes BASELINE.json names.  Pure torch on the CPU; used by
bench.py, the smoke test and the parity tests to build identical inputs for the CUDA path and the oracle."""
from __future__ import annotations

from typing import List, Tuple

import torch


def grid_mesh(kind: str, nx: int, ny: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Regular nx x ny grid.  Vertex (i, j) has id i*(ny+1)+j; a cell (a,b,c,d) with a=(i,j), b=(i+1,j),
    c=(i+1,j+1), d=(i,j+1) gives triangles [a,b,c],[a,c,d] or the quad [a,b,c,d].  x,y span [-0.9,0.9],
    z is a smooth height field plus U(-0.02,0.02) jitter; every coordinate stays inside (-1,1)."""
    g = torch.Generator().manual_seed(int(seed))
    ii, jj = torch.meshgrid(torch.arange(nx + 1), torch.arange(ny + 1), indexing="ij")
    x = -0.9 + 1.8 * ii.float() / nx
    y = -0.9 + 1.8 * jj.float() / ny
    phase = torch.rand(2, generator=g) * 6.2831853
    z = 0.3 * torch.sin(3.0 * x + phase[0]) * torch.cos(2.0 * y + phase[1])
    verts = torch.stack([x, y, z], dim=-1).reshape(-1, 3)
    verts = verts + (torch.rand(verts.shape, generator=g) - 0.5) * 0.04
    ci, cj = torch.meshgrid(torch.arange(nx), torch.arange(ny), indexing="ij")
    a = (ci * (ny + 1) + cj).reshape(-1)
    b = ((ci + 1) * (ny + 1) + cj).reshape(-1)
    c = ((ci + 1) * (ny + 1) + cj + 1).reshape(-1)
    d = (ci * (ny + 1) + cj + 1).reshape(-1)
    if kind == "tri":
        faces = torch.stack([torch.stack([a, b, c], -1), torch.stack([a, c, d], -1)], dim=1).reshape(-1, 3)
    elif kind == "quad":
        faces = torch.stack([a, b, c, d], -1)
    else:
        raise ValueError(kind)
    return verts.contiguous(), faces.long().contiguous()


def grid_batch(kind: str, nx: int, ny: int, seeds) -> Tuple[torch.Tensor, torch.Tensor]:
    vs, fs = zip(*(grid_mesh(kind, nx, ny, s) for s in seeds))
    return torch.stack(vs), torch.stack(fs)


def ragged_batch(kind: str, nx: int, ny: int, num_faces: List[int], seeds, pad_id: int = -1):
    """Meshes made of the first n faces of a grid, padded like the reference's custom_collate
    (data.py:347-369): faces with pad_id, vertices with float(pad_id), trailing."""
    vs, fs = [], []
    for n, s in zip(num_faces, seeds):
        v, f = grid_mesh(kind, nx, ny, s)
        f = f[:n]
        used = int(f.max()) + 1
        vs.append(v[:used])
        fs.append(f)
    V = max(v.shape[0] for v in vs)
    F = max(f.shape[0] for f in fs)
    vertices = torch.full((len(vs), V, 3), float(pad_id))
    faces = torch.full((len(fs), F, fs[0].shape[1]), pad_id, dtype=torch.long)
    for i, (v, f) in enumerate(zip(vs, fs)):
        vertices[i, : v.shape[0]] = v
        faces[i, : f.shape[0]] = f
    return vertices, faces


def readme_batch(seed: int = 0, batch: int = 2, num_vertices: int = 121, num_faces: int = 64):
    """BASELINE config 1: the README tensors (README.md:55-56) - random, includes degenerate faces and
    coordinates outside the quantisation range."""
    torch.manual_seed(seed)
    vertices = torch.randn((batch, num_vertices, 3))
    faces = torch.randint(0, num_vertices, (batch, num_faces, 3))
    return vertices, faces


# ---- generation ON THE DEVICE, keyed by the global mesh index ("seeds = global mesh index, generated on-device per rank") ----
def _hash01(idx: torch.Tensor, salt: int) -> torch.Tensor:
    """Counter-based uniform [0, 1): a 64-bit integer mix of (idx, salt); integer arithmetic only, so CPU and GPU give the same bits."""
    def s64(k):                                                # a Python int as the int64 with the same bits modulo 2^64
        k &= (1 << 64) - 1
        return k - (1 << 64) if k >= (1 << 63) else k
    x = idx.to(torch.int64) * s64(0x9E3779B97F4A7C15) + s64((salt * 2 + 1) * 0x6C62272E07BB0142)   # tensor arithmetic wraps modulo 2^64
    x = (x ^ (x >> 31)) * s64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> 29)) * s64(0x94D049BB133111EB)
    x = x ^ (x >> 32)
    return ((x >> 20) & 0xFFFFFF).to(torch.float32) / 16777216.0     # bits 20..43 (the top bits carry the arithmetic shifts' sign extension)


def grid_batch_device(kind: str, nx: int, ny: int, first_index: int, count: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """`count` grid meshes with the global indices first_index .. first_index + count - 1, built by torch ops on `device` (no host
    data, no H2D copy): the same family as grid_mesh - x, y on the grid in [-0.9, 0.9], z a smooth height field whose two phases and
    the U(-0.02, 0.02) jitter of every coordinate are hashes of (mesh index, vertex, coordinate) - so every rank can generate any
    slice of a collection of any size, and a test can regenerate the same meshes for the oracle (copy them to the CPU: sin / cos
    round differently there)."""
    dev = torch.device(device)
    idx = torch.arange(first_index, first_index + count, device=dev, dtype=torch.int64)
    ii, jj = torch.meshgrid(torch.arange(nx + 1, device=dev), torch.arange(ny + 1, device=dev), indexing="ij")
    x = (-0.9 + 1.8 * ii.float() / nx).reshape(-1)
    y = (-0.9 + 1.8 * jj.float() / ny).reshape(-1)
    nv = x.numel()
    p0 = (_hash01(idx, 1) * 6.2831853)[:, None]
    p1 = (_hash01(idx, 2) * 6.2831853)[:, None]
    z = 0.3 * torch.sin(3.0 * x[None] + p0) * torch.cos(2.0 * y[None] + p1)
    verts = torch.stack([x[None].expand(count, nv), y[None].expand(count, nv), z], dim=-1)
    cell = idx[:, None, None] * (nv * 3) + torch.arange(nv * 3, device=dev, dtype=torch.int64).reshape(1, nv, 3)
    verts = verts + (_hash01(cell, 3) - 0.5) * 0.04
    ci, cj = torch.meshgrid(torch.arange(nx, device=dev), torch.arange(ny, device=dev), indexing="ij")
    a = (ci * (ny + 1) + cj).reshape(-1)
    b = ((ci + 1) * (ny + 1) + cj).reshape(-1)
    c = ((ci + 1) * (ny + 1) + cj + 1).reshape(-1)
    d = (ci * (ny + 1) + cj + 1).reshape(-1)
    if kind == "tri":
        faces = torch.stack([torch.stack([a, b, c], -1), torch.stack([a, c, d], -1)], dim=1).reshape(-1, 3)
    elif kind == "quad":
        faces = torch.stack([a, b, c, d], -1)
    else:
        raise ValueError(kind)
    return verts.contiguous(), faces.long()[None].expand(count, -1, -1).contiguous()


def ragged_batch_device(kind: str, nx: int, ny: int, first_index: int, count: int, device, lo: int = 100, hi: int = 800, pad_id: int = -1):
    """grid_batch_device cut to F_i = lo + hash(mesh index) % (hi - lo + 1) faces per mesh, padded like custom_collate
    (faces with pad_id, vertices with float(pad_id), trailing).  Returns (vertices, faces, face_counts)."""
    v, f = grid_batch_device(kind, nx, ny, first_index, count, device)
    idx = torch.arange(first_index, first_index + count, device=v.device, dtype=torch.int64)
    nf = lo + (_hash01(idx, 4) * (hi - lo + 1)).long().clamp(max=hi - lo)
    F = f.shape[1]
    fmask = torch.arange(F, device=v.device)[None, :] < nf[:, None]
    f = torch.where(fmask[:, :, None], f, torch.full_like(f, pad_id))
    used = torch.where(fmask[:, :, None], f, torch.zeros_like(f)).amax(dim=(1, 2)) + 1          # vertices a mesh's kept faces reference
    vmask = torch.arange(v.shape[1], device=v.device)[None, :] < used[:, None]
    v = torch.where(vmask[:, :, None], v, torch.full_like(v, float(pad_id)))
    fmax, vmax = int(nf.max()), int(used.max())
    return v[:, :vmax].contiguous(), f[:, :fmax].contiguous(), nf
