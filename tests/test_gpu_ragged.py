"""GPU: the ragged input layout (CSR of meshes; MeshAutoencoder.tokenize_ragged + the three offset-aware kernels)
against the padded layout of the reference's custom_collate (data.py:347-369) on the same meshes - bit for bit, since only the
addressing of the inputs and of the output differs - and against the CPU oracle."""
import pytest
import torch

from helpers import SMALL_CFG, make_pair, match_rate
import synth
from meshgpt_pytorch_b200 import data as Dt

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
    assert match_rate(Dt.ragged_to_padded_codes(codes, fo, nvf, m.pad_id).cpu(), want) >= 0.995


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
        m.tokenize_ragged(rv, rf, fo.cuda(), vo.cuda(), max_faces=100, max_vertices=128, check_status=True)
    with pytest.raises(IndexError):                                              # mesh-local ids: mesh 0 must not reach into mesh 1's vertices
        bad = rf.clone()
        bad[3, 1] = int(vo[1])
        m.tokenize_ragged(rv, bad, fo, vo)
        m.check_last_status()                                                    # (the default defers the check: no host sync in the call)
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


@pytest.mark.parametrize("cfg,kind,counts", [
    (dict(SMALL_CFG), "tri", [130, 61, 8, 144]),
    (dict(SMALL_CFG, quads=True, use_residual_lfq=False), "quad", [72, 30, 5]),
])
def test_ragged_precomputed_face_edges(cfg, kind, counts):
    """Precomputed edges in the ragged layout (flat (sum ne, 2) list with mesh-local face indices + edge_offsets: the face-edge cache's
    content without its padding) against tokenize(face_edges=...) on the padded batch, and against the edges derived on the device."""
    import meshgpt_pytorch_b200 as M
    o, m = make_pair(**cfg)
    v, f = synth.ragged_batch(kind, 9, 8, counts, list(range(len(counts))))
    edges = M.derive_face_edges_from_faces(f.cuda())                          # (b, e, 2), pad_id rows behind each mesh's list
    padded = m.tokenize(v.cuda(), f.cuda(), face_edges=edges)
    keep = (edges != -1).all(-1)
    flat_edges = edges[keep]
    eo = torch.zeros(len(counts) + 1, dtype=torch.int64)
    eo[1:] = keep.sum(1).cpu().cumsum(0)
    rv, rf, fo, vo = ragged_cuda(v, f)
    codes = m.tokenize_ragged(rv, rf, fo, vo, face_edges=flat_edges, edge_offsets=eo)
    nvf = m.num_vertices_per_face
    assert torch.equal(Dt.ragged_to_padded_codes(codes, fo, nvf, m.pad_id), padded)
    assert torch.equal(codes, m.tokenize_ragged(rv, rf, fo, vo))              # derived on the device: same lists in the same order
    assert torch.equal(codes, m.tokenize_ragged(rv, rf, fo, vo, face_edges=flat_edges, edge_offsets=eo.cuda()))
    # a different edge set must change the result (the lists are really used): self-loops only
    loops = torch.cat([torch.arange(c) for c in counts]).cuda()
    lo = torch.zeros(len(counts) + 1, dtype=torch.int64)
    lo[1:] = torch.tensor(counts).cumsum(0)
    alone = m.tokenize_ragged(rv, rf, fo, vo, face_edges=torch.stack([loops, loops], 1), edge_offsets=lo)
    want = m.tokenize(v.cuda(), f.cuda(), face_edges=Dt_self_loops(f).cuda())
    assert torch.equal(Dt.ragged_to_padded_codes(alone, fo, nvf, m.pad_id), want) and not torch.equal(alone, codes)
    # no edges at all: every aggregation is zero, still well defined
    none = m.tokenize_ragged(rv, rf, fo, vo, face_edges=flat_edges[:0], edge_offsets=[0] * (len(counts) + 1))
    assert none.shape == codes.shape and bool((none >= 0).all())
    with pytest.raises(IndexError):                                            # an edge that points outside the batch's faces
        bad = flat_edges.clone()
        bad[-1, 0] = 10 ** 6
        m.tokenize_ragged(rv, rf, fo, vo, face_edges=bad, edge_offsets=eo, check_status=True)
    with pytest.raises(IndexError):                                            # ... or into ANOTHER mesh's faces (inside the batch, outside its mesh)
        bad = flat_edges.clone()
        bad[0, 1] = counts[0]                                                  # first mesh's edge -> first face of the second mesh
        m.tokenize_ragged(rv, rf, fo, vo, face_edges=bad, edge_offsets=eo, check_status=True)


def Dt_self_loops(faces):
    """(b, nf, 2) padded edge list with (i, i) for every un-padded face."""
    b, nf, _ = faces.shape
    e = torch.arange(nf)[None, :, None].expand(b, nf, 2).clone()
    e[~(faces != -1).all(-1)] = -1
    return e


def test_ragged_graph_serves_batches_of_any_composition():
    """GraphedRaggedTokenizer: ONE graph captured on upper bounds, replayed on batches with different mesh counts and sizes, offsets from
    the host and from the device - every replay equal to the eager ragged call (and so to the padded layout)."""
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    hist = torch.zeros(m.num_quantizers, m.codebook_size, dtype=torch.int64, device="cuda")
    g = m.graphed_ragged(max_meshes=6, max_faces=144, max_vertices=90, max_total_faces=500, max_total_vertices=400, histogram=hist)
    want_hist = torch.zeros_like(hist)
    for counts in ([130, 61, 8, 1, 144], [144], [5, 5, 5, 5, 5, 5], [77, 144, 144, 100]):
        v, f = synth.ragged_batch("tri", 9, 8, counts, list(range(len(counts))))
        rv, rf, fo, vo = ragged_cuda(v, f)
        want = m.tokenize_ragged(rv, rf, fo, vo, histogram=want_hist)
        got = g(rv, rf, fo, vo).clone()
        assert torch.equal(got, want), counts
        got = g(rv.cpu().pin_memory(), rf.cpu().pin_memory(), fo.cuda(), vo.cuda()).clone()      # host data, device offsets
        assert torch.equal(got, want), counts
        want_hist += torch.stack([torch.bincount(want[:, q], minlength=m.codebook_size) for q in range(m.num_quantizers)])
        g.check()
    assert torch.equal(hist, want_hist)
    with pytest.raises(AssertionError):
        v, f = synth.ragged_batch("tri", 9, 8, [144] * 4, [0, 1, 2, 3])
        g(*ragged_cuda(v, f))                                                                      # 576 faces > 500


def test_ragged_host_pipeline_matches_single_calls():
    """RaggedHostPipeline (meshtok_pipeline_create_ragged / _submit_ragged): host buffers in, flat host codes out, two and three batches in
    flight on one graph per slot - every batch equal to the eager ragged call on the same meshes; histogram = bincount over all batches;
    a bad batch raises when it is handed back and the ring keeps working afterwards."""
    o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
    hist = torch.zeros(m.num_quantizers, m.codebook_size, dtype=torch.int64, device="cuda")
    compositions = ([130, 61, 8, 1, 144], [144], [5, 5, 5, 5, 5, 5], [77, 144, 144, 100], [1], [144, 144, 144])
    for depth in (2, 3):
        hist.zero_()
        pipe = m.host_pipeline_ragged(6, 144, 90, 500, 400, depth=depth, histogram=hist)
        want, got = [], {}
        for i, counts in enumerate(compositions):
            v, f = synth.ragged_batch("tri", 9, 8, counts, [10 * i + k for k in range(len(counts))])
            rv, rf, fo, vo = ragged_cuda(v, f)
            want.append(m.tokenize_ragged(rv, rf, fo, vo).cpu())
            done = pipe.submit(rv.cpu().pin_memory(), rf.cpu().pin_memory(), fo.cpu(), vo.cpu())
            if done is not None:
                got[done[0]] = done[1].clone()
        for ticket, codes in pipe.drain():
            got[ticket] = codes.clone()
        assert sorted(got) == list(range(len(compositions)))
        for i, w in enumerate(want):
            assert got[i].shape == w.shape and torch.equal(got[i], w), (depth, compositions[i])
        allc = torch.cat(want)
        assert torch.equal(hist.cpu(), torch.stack([torch.bincount(allc[:, q], minlength=m.codebook_size) for q in range(m.num_quantizers)]))
        # an empty batch, then a vertex id that reaches into the next mesh: reported with the batch, not silently tokenized
        assert pipe.submit(torch.zeros(0, 3).pin_memory(), torch.zeros(0, 3, dtype=torch.int64).pin_memory(), torch.zeros(1, dtype=torch.int32),
                           torch.zeros(1, dtype=torch.int32)) is None
        v, f = synth.ragged_batch("tri", 9, 8, [20, 20], [0, 1])
        rv, rf, fo, vo = ragged_cuda(v, f)
        bad = rf.cpu().clone()
        bad[3, 1] = int(vo[1]) + 2
        pipe.submit(rv.cpu().pin_memory(), bad.pin_memory(), fo.cpu(), vo.cpu())
        with pytest.raises(Exception):
            pipe.drain()
        done = pipe.submit(rv.cpu().pin_memory(), rf.cpu().pin_memory(), fo.cpu(), vo.cpu())
        res = pipe.drain()
        assert torch.equal(res[-1][1], m.tokenize_ragged(rv, rf, fo, vo).cpu())
        with pytest.raises(Exception):
            v, f = synth.ragged_batch("tri", 9, 8, [144] * 4, [0, 1, 2, 3])
            rv, rf, fo, vo = ragged_cuda(v, f)
            pipe.submit(rv.cpu(), rf.cpu(), fo.cpu(), vo.cpu())                          # 576 faces > 500
        pipe.close()


_GROUP_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from helpers import SMALL_CFG, make_pair
import synth  # tests/synth.py: pure-torch mesh generator (test + bench infrastructure, not product)
o, m = make_pair(**dict(SMALL_CFG, use_residual_lfq=False))
g = torch.Generator().manual_seed(11)
counts = torch.randint(1, 41, (320,), generator=g).tolist()
v, f = synth.ragged_batch("tri", 5, 4, counts, list(range(320)))
torch.save(m.tokenize(v.cuda(), f.cuda()).cpu(), sys.argv[1])
"""


def test_large_batch_aggregation_kernel_gives_the_same_tokens(tmp_path):
    """The aggregation kernel's work split - one CTA per (mesh, slice) or per (mesh, group of slices) with double-buffered slices (what
    large launches take) - and the round-1 kernels (MESHTOK_GATHER_WIDE=0, four threads per face) on a 320-mesh ragged batch: same
    neighbour order, same rounding, so the tokens must be identical.  The switches are read once per process, hence subprocesses."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "run.py"
    script.write_text(_GROUP_SCRIPT.format(root=root, tests=os.path.join(root, "tests")))
    outs = {}
    for name, env in (("one slice", dict(MESHTOK_GATHER_WIDE_GROUP="1")), ("groups of 2", dict(MESHTOK_GATHER_WIDE_GROUP="2")), ("default", {}),
                      ("round 1 per slice", dict(MESHTOK_GATHER_WIDE="0", MESHTOK_GATHER_GROUP="0")), ("round 1 groups", dict(MESHTOK_GATHER_WIDE="0", MESHTOK_GATHER_GROUP="1")),
                      ("fused into the output linears", dict(MESHTOK_GATHER_FUSED="1"))):
        path = tmp_path / f"{name.replace(' ', '_')}.pt"
        r = subprocess.run([sys.executable, str(script), str(path)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = torch.load(path)
    for name, codes in outs.items():
        assert codes.shape == outs["default"].shape and torch.equal(codes, outs["default"]), name


@pytest.mark.parametrize("cfg,kind,counts", [
    (dict(SMALL_CFG), "tri", [130, 61, 8, 144]),
    (dict(SMALL_CFG, quads=True, use_residual_lfq=False), "quad", [72, 30, 5]),
])
def test_encode_and_quantize_on_the_ragged_layout(cfg, kind, counts):
    """encode() / quantize() as separate calls on a CSR of meshes (round 1 refused them): the same rows as the padded calls, one per
    real face in input order, and quantize_ragged(encode_ragged(...)) == tokenize_ragged(...)."""
    o, m = make_pair(**cfg)
    v, f = synth.ragged_batch(kind, 9, 8, counts, list(range(len(counts))))
    rv, rf, fo, vo = ragged_cuda(v, f)
    nvf = m.num_vertices_per_face
    fm = (f != -1).all(-1)
    vd, fd = v.cuda(), f.cuda()
    import meshgpt_pytorch_b200 as M
    edges = M.derive_face_edges_from_faces(fd)
    enc_p, dc_p = m.encode(vertices=vd, faces=fd, face_edges=edges, face_mask=fm.cuda(), face_edges_mask=(edges != -1).all(-1),
                           return_face_coordinates=True)
    enc_r, dc_r = m.encode_ragged(rv, rf, fo, vo, return_face_coordinates=True)
    assert enc_r.shape == (rf.shape[0], m.encoder_dims[-1])
    assert torch.equal(enc_r, enc_p[fm.cuda()]) and torch.equal(dc_r, dc_p[fm.cuda()])
    q_p, codes_p, _ = m.quantize(faces=fd, face_mask=fm.cuda(), face_embed=enc_p)
    q_r, codes_r, loss = m.quantize_ragged(rf, fo, vo, enc_r)
    assert loss.shape == (m.num_quantizers,)
    assert torch.equal(Dt.ragged_to_padded_codes(codes_r, fo, nvf, m.pad_id), codes_p)
    assert torch.equal(q_r, q_p[fm.cuda()])
    assert torch.equal(codes_r, m.tokenize_ragged(rv, rf, fo, vo))
    # against the oracle's encode on the padded batch (activations within the stated tolerance)
    enc_o = o.encode(vertices=v, faces=f, face_edges=edges.cpu(), face_mask=fm, face_edges_mask=(edges.cpu() != -1).all(-1))
    assert torch.allclose(enc_r.cpu(), enc_o[fm], rtol=1e-3, atol=1e-4)
