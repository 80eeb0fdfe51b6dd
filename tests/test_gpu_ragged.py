"""GPU: the ragged input layout (CSR of meshes; MeshAutoencoder.tokenize_ragged + the three offset-aware kernels)
against the padded layout of the reference's custom_collate (data.py:347-369) on the same meshes - bit for bit, since only the
addressing of the inputs and of the output differs - and against the CPU oracle."""
import pytest
import torch

from helpers import SMALL_CFG, make_pair, match_rate
from meshgpt_pytorch_b200 import data as Dt, synth

pytestmark = pytest.mark.gpu


def ragged_cuda(v, f, num_vertices=None):
    r = Dt.padded_to_ragged(v, f, num_vertices=num_vertices)
    return r["vertices"].cuda(), r["faces"].cuda(), r["face_offsets"], r["vertex_offsets"]


@pytest.mark.parametrize("cfg,kind,grid,counts", [
    (dict(SMALL_CFG), "tri", (9, 8), [130, 61, 8, 1, 144]),
    (dict(SMALL_CFG, use_residual_lfq=False), "tri", (9, 8), [144, 1, 77]),
    (dict(SMALL_CFG, quads=True), "quad", (9, 8), [72, 30, 5, 1]),
    (dict(SMALL_CFG, quads=True, use_residual_lfq=False), "quad", (6, 7), [42, 17]),
])
def test_ragged_layout_equals_the_padded_layout(cfg, kind, grid, counts):
    o, m = make_pair(**cfg)
    v, f = synth.ragged_batch(kind, grid[0], grid[1], counts, list(range(len(counts))))
    hist_p = torch.zeros(m.num_quantizers, m.codebook_size, dtype=torch.int64, device="cuda")
    hist_r = torch.zeros_like(hist_p)
    padded = m.forward(vertices=v.cuda(), faces=f.cuda(), return_codes=True, histogram=hist_p)
    rv, rf, fo, vo = ragged_cuda(v, f)
    codes, coords = m.tokenize_ragged(rv, rf, fo, vo, histogram=hist_r, return_face_coordinates=True)
    nvf = m.num_vertices_per_face
    assert codes.shape == (sum(counts) * nvf, m.num_quantizers) and codes.dtype == torch.int64
    assert bool((codes >= 0).all())                                             # no pad tokens in the flat sequence
    assert torch.equal(Dt.ragged_to_padded_codes(codes, fo, nvf, m.pad_id), padded)
    assert torch.equal(hist_r, hist_p)
    # discretised coordinates (encode()'s second result): the padded run's valid rows, in order
    want_coords = m._run(vertices=v.cuda(), faces=f.cuda(), want_codes=False, want_encoded=True, want_face_coords=True)["face_coords"]
    mask = (f != -1).all(-1)
    assert torch.equal(coords.cpu(), want_coords.cpu()[mask])
    # and the oracle on the padded batch
    want = o.tokenize(v, f)
    assert match_rate(Dt.ragged_to_padded_codes(codes, fo, nvf, m.pad_id).cpu(), want) > 0.97


def test_ragged_empty_meshes_offsets_on_the_device_and_explicit_bounds():
    o, m = make_pair(**SMALL_CFG)
    v, f = synth.ragged_batch("tri", 9, 8, [50, 144, 20], [0, 1, 2])
    rv, rf, fo, vo = ragged_cuda(v, f)
    ref = m.tokenize_ragged(rv, rf, fo, vo)
    # the same three meshes with empty ones in front, between and behind (an offset repeated), offsets as CUDA tensors, loose bounds
    fo2 = torch.tensor([0, 0, int(fo[1]), int(fo[1]), int(fo[1]), int(fo[2]), int(fo[3]), int(fo[3])], dtype=torch.int64).cuda()
    vo2 = torch.tensor([0, 0, int(vo[1]), int(vo[1]), int(vo[1]), int(vo[2]), int(vo[3]), int(vo[3])], dtype=torch.int64).cuda()
    got = m.tokenize_ragged(rv, rf, fo2, vo2, max_faces=200, max_vertices=128)
    assert torch.equal(got, ref)
    got = m.tokenize_ragged(rv, rf, fo2, vo2)                                    # bounds from the device offsets
    assert torch.equal(got, ref)
    # no meshes at all / no faces at all
    e = m.tokenize_ragged(rv[:0], rf[:0], [0], [0])
    assert e.shape == (0, m.num_quantizers)
    e = m.tokenize_ragged(rv[:0], rf[:0], [0, 0, 0], [0, 0, 0])
    assert e.shape == (0, m.num_quantizers)


def test_ragged_refusals():
    o, m = make_pair(**SMALL_CFG)
    v, f = synth.ragged_batch("tri", 9, 8, [50, 144], [0, 1])
    rv, rf, fo, vo = ragged_cuda(v, f)
    with pytest.raises(IndexError):                                              # a mesh larger than the stated bound
        m.tokenize_ragged(rv, rf, fo.cuda(), vo.cuda(), max_faces=100, max_vertices=128)
    with pytest.raises(IndexError):                                              # mesh-local ids: mesh 0 must not reach into mesh 1's vertices
        bad = rf.clone()
        bad[3, 1] = int(vo[1])
        m.tokenize_ragged(rv, bad, fo, vo)
    with pytest.raises(AssertionError):
        m.tokenize_ragged(rv, rf, [0, 10, 5, rf.shape[0]], [0, 1, 2, rv.shape[0]])
    with pytest.raises(RuntimeError):
        m.tokenize_ragged(rv.cpu(), rf.cpu(), fo, vo)
    assert m.tokenize_ragged(rv, rf, fo, vo).shape[0] == rf.shape[0] * 3         # and the model still works afterwards


def test_ragged_full_size_c4_like_batch():
    """BASELINE configs[3]-like: 32 meshes with 100..800 faces, default RVQ model - ragged against padded, whole batch."""
    o, m = make_pair(use_residual_lfq=False)
    g = torch.Generator().manual_seed(1234)
    counts = torch.randint(100, 801, (32,), generator=g).tolist()
    v, f = synth.ragged_batch("tri", 20, 20, counts, list(range(32)))
    padded = m.tokenize(v.cuda(), f.cuda())
    rv, rf, fo, vo = ragged_cuda(v, f)
    codes = m.tokenize_ragged(rv, rf, fo, vo)
    assert torch.equal(Dt.ragged_to_padded_codes(codes, fo, 3, m.pad_id), padded)
    assert rf.shape[0] == sum(counts) and rv.shape[0] < v.shape[0] * v.shape[1]
