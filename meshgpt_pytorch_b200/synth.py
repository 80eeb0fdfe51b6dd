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
