"""CPU: host-side mirror of the reference interface - constructor contract, state-dict compatibility,
synthetic inputs, sharding helpers, the no-CPU-fallback rule."""
import pytest
import torch

import meshgpt_pytorch_b200 as M
from helpers import SMALL_CFG, make_oracle
import synth
from meshgpt_pytorch_b200 import sharding
from oracle import meshtok_oracle as O

REFERENCE_KEYS_TRI_LFQ = {
    "coor_embed.weight": (128, 64), "angle_embed.weight": (128, 16), "area_embed.weight": (128, 16),
    "normal_embed.weight": (128, 64), "project_in.weight": (192, 832), "project_in.bias": (192,),
    "init_sage_conv.lin.weight": (192, 192), "init_sage_conv.lin.bias": (192,), "init_sage_conv.lin_l.weight": (64, 192),
    "init_sage_conv.lin_l.bias": (64,), "init_sage_conv.lin_r.weight": (64, 192),
    "init_encoder_act_and_norm.1.weight": (64,), "init_encoder_act_and_norm.1.bias": (64,),
    "encoders.0.lin_l.weight": (128, 64), "encoders.3.lin_l.weight": (576, 256), "encoders.3.lin_r.weight": (576, 256),
    "project_dim_codebook.weight": (576, 576), "project_dim_codebook.bias": (576,),
    "quantizer.project_in.weight": (14, 192), "quantizer.project_in.bias": (14,),
    "quantizer.project_out.weight": (192, 14), "quantizer.project_out.bias": (192,),
}


def test_default_constructor_matches_reference_key_layout():
    m = M.MeshAutoencoder(num_discrete_coors=128)            # README example: default => ResidualLFQ, triangles
    sd = m.state_dict()
    for k, shape in REFERENCE_KEYS_TRI_LFQ.items():
        assert tuple(sd[k].shape) == shape, k
    assert "init_sage_conv.lin_r.bias" not in sd
    q = M.MeshAutoencoder(quads=True)
    assert q.project_in.weight.shape == (192, 1040) and q.project_dim_codebook.weight.shape == (768, 576)
    r = M.MeshAutoencoder(use_residual_lfq=False)
    sd = r.state_dict()
    assert tuple(sd["quantizer.layers.0._codebook.embed"].shape) == (1, 16384, 192)
    assert sd["quantizer.layers.1._codebook.embed"].data_ptr() == sd["quantizer.layers.0._codebook.embed"].data_ptr()
    assert bool(sd["quantizer.layers.0._codebook.initted"])


def test_state_dict_round_trip_with_oracle_and_extra_decoder_keys():
    o = make_oracle(**SMALL_CFG)
    m = M.MeshAutoencoder(**SMALL_CFG)
    sd = dict(o.state_dict())
    sd["decoders.0.block1.proj.weight"] = torch.zeros(3)       # a decoder key of a real checkpoint
    ignored = m.load_reference_state_dict(sd)
    assert ignored == ["decoders.0.block1.proj.weight"]
    for k, v in o.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k
    del sd["project_in.bias"]
    with pytest.raises(KeyError):
        m.load_reference_state_dict(sd)


def test_unsupported_configurations_are_refused():
    with pytest.raises(NotImplementedError):
        M.MeshAutoencoder(attn_encoder_depth=1)
    with pytest.raises(NotImplementedError):
        M.MeshAutoencoder(sageconv_kwargs=dict(normalize=False, project=True))
    m = M.MeshAutoencoder(**SMALL_CFG)
    v, f = synth.readme_batch()
    with pytest.raises(NotImplementedError):
        m.forward(vertices=v, faces=f)                         # training branch is out of scope
    with pytest.raises(RuntimeError, match="CUDA"):
        m.tokenize(v, f)                                       # CPU tensors / CPU module: no fallback
    with pytest.raises(RuntimeError, match="CUDA"):
        M.derive_face_edges_from_faces(f)
    with pytest.raises(AssertionError):
        m.tokenize(v, f, return_codes=True)                    # meshgpt_pytorch.py:951


def test_product_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(M.__file__)
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn


def test_synthetic_meshes():
    v, f = synth.grid_mesh("tri", 20, 20, 3)
    assert v.shape == (441, 3) and f.shape == (800, 3) and float(v.abs().max()) < 1 and int(f.max()) == 440
    v, f = synth.grid_mesh("quad", 20, 40, 3)
    assert v.shape == (861, 3) and f.shape == (800, 4)
    v1, _ = synth.grid_mesh("tri", 4, 4, 11)
    v2, _ = synth.grid_mesh("tri", 4, 4, 11)
    assert torch.equal(v1, v2)
    vs, fs = synth.ragged_batch("tri", 20, 20, [100, 800, 333], [0, 1, 2])
    assert fs.shape == (3, 800, 3) and (fs[0, 100:] == -1).all() and (fs[2, 333:] == -1).all() and (fs[1] >= 0).all()
    assert vs.shape[1] == 441 and (vs[0, int(fs[0].max()) + 1:] == -1).all()


def test_shard_ranges_and_balancing():
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharding.shard_range(3, 3, 4) == (3, 3)
    counts = [800, 100, 100, 100, 700, 200, 300, 500]
    cuts = sharding.balanced_contiguous_shards(counts, 3)
    assert cuts[0][0] == 0 and cuts[-1][1] == len(counts) and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    loads = [sum(counts[a:b]) for a, b in cuts]
    assert max(loads) <= 1.35 * sum(counts) / 3


def test_histogram_definition():
    codes = torch.tensor([[[3, 1], [3, 0], [-1, -1]], [[2, 1], [-1, -1], [-1, -1]]])
    h = O.codebook_usage_histogram(codes, 2, 4)
    assert h.tolist() == [[0, 0, 1, 2], [1, 2, 0, 0]]


def test_face_edges_cache_file_format(tmp_path):
    """write_face_edges_cache produces what the reference's cache_face_edges_for_dataset decorator reads back
    (meshgpt_pytorch/data.py:184-214): float32 (dataset_len, max_edges_len, 2) padded with pad_id + bool flags."""
    import numpy as np
    import meshgpt_pytorch_b200 as M
    import synth
    from oracle import meshtok_oracle as O
    _, faces = synth.ragged_batch("tri", 6, 6, [72, 30, 1], [0, 1, 2])
    edges = O.derive_face_edges_from_faces(faces)                     # (3, e, 2) padded with -1
    max_len = int((edges != -1).all(-1).sum(-1).max()) + 5
    fe_file, ic_file = M.write_face_edges_cache(tmp_path, dataset_len=5, start=1, face_edges=edges, max_edges_len=max_len)
    cache, cached = np.load(fe_file, mmap_mode="r"), np.load(ic_file, mmap_mode="r")
    assert cache.dtype == np.float32 and cache.shape == (5, max_len, 2) and cached.dtype == np.bool_
    assert cached.tolist() == [False, True, True, True, False]
    for i in range(3):
        # the reference's read path: rows with no pad entry are the mesh's edges, in order
        row = torch.from_numpy(np.array(cache[1 + i]))
        got = row[(row != -1).all(-1)].long()
        want = O.derive_face_edges_from_faces(faces[i])
        want = want[(want != -1).all(-1)]
        assert torch.equal(got, want)
    # a second batch goes into the same files; too-long meshes are refused like the reference's assert
    M.write_face_edges_cache(tmp_path, 5, 4, edges[:1], max_len)
    assert np.load(ic_file, mmap_mode="r").tolist() == [False, True, True, True, True]
    with pytest.raises(AssertionError):
        M.write_face_edges_cache(tmp_path / "small", 5, 0, edges, max_edges_len=3)


# ---- the decode path and the token-embedding front end: host-side contract (no kernel runs on CPU) ----

def test_decoder_and_front_end_key_layout_matches_the_reference_names():
    a = M.MeshAutoencoder(use_residual_lfq=False)
    d = M.MeshDecoder(a)
    sd = d.state_dict()
    assert tuple(sd["init_decoder_conv.0.weight"].shape) == (128, 576, 7)            # Conv1d(dim_codebook * 3 -> 128, k = 7), :621-623
    assert tuple(sd["init_decoder_conv.3.weight"].shape) == (128,)                   # LayerNorm sits at index 3 of the Sequential, :626
    assert tuple(sd["decoders.3.residual_conv.weight"].shape) == (192, 128, 1)       # 128 -> 192: a 1x1 residual conv, :367
    assert "decoders.0.residual_conv.weight" not in sd                               # 128 -> 128: nn.Identity
    assert tuple(sd["decoders.15.block2.proj.weight"].shape) == (384, 384, 3)
    assert tuple(sd["decoders.0.excite.net.0.weight"].shape) == (32, 128)            # max(dim // 4, 16), :304
    assert tuple(sd["to_coor_logits.0.weight"].shape) == (128 * 9, 384)              # :639
    fe = M.MeshTransformerFrontEnd(a, dim=512, max_seq_len=8190)
    sd = fe.state_dict()
    assert tuple(sd["token_embed.weight"].shape) == (16385, 512)                     # codebook_size + 1 (eos), :1158
    assert tuple(sd["abs_pos_emb.weight"].shape) == (8190, 512)
    assert tuple(sd["quantize_level_embed"].shape) == (2, 512) and tuple(sd["vertex_embed"].shape) == (3, 512)
    assert tuple(sd["to_face_tokens.0.weight"].shape) == (512, 6 * 512) and tuple(sd["to_face_tokens.1.weight"].shape) == (512,)


def test_decoder_and_front_end_refuse_what_they_do_not_cover():
    a = M.MeshAutoencoder(**SMALL_CFG)
    with pytest.raises(NotImplementedError):
        M.MeshDecoder(a, attn_decoder_depth=2)
    with pytest.raises(ValueError):
        M.MeshDecoder(a, init_decoder_conv_kernel=4)
    with pytest.raises(KeyError):
        M.MeshDecoder(a, decoder_dims_through_depth=(32, 32)).load_reference_state_dict({})
    with pytest.raises(ValueError):
        M.MeshTransformerFrontEnd(a, dim=128, max_seq_len=100)                       # not a multiple of 3 x 2 codes per face (:1156)
    with pytest.raises(NotImplementedError):
        M.MeshTransformerFrontEnd(a, dim=100, max_seq_len=60)
    fe = M.MeshTransformerFrontEnd(a, dim=128, max_seq_len=60)
    with pytest.raises(KeyError):
        fe.load_reference_state_dict({"token_embed.weight": torch.zeros(257, 128)})
    with pytest.raises(RuntimeError):
        fe.embed_codes(torch.zeros(1, 6, dtype=torch.int64))                         # CPU tensor: there is no CPU path
    d = M.MeshDecoder(a, decoder_dims_through_depth=(32, 32))
    with pytest.raises(RuntimeError):
        d.decode_from_codes_to_faces(torch.zeros(1, 6, 2, dtype=torch.int64))


def test_ragged_collate_is_custom_collate_without_the_padding():
    """ragged_collate / padded_to_ragged / ragged_to_padded_codes against the padded form of the same meshes
    (synth.ragged_batch pads like the reference's custom_collate, data.py:347-369)."""
    from meshgpt_pytorch_b200 import data as Dt
    counts = [13, 1, 40]
    v, f = synth.ragged_batch("tri", 5, 4, counts, [0, 1, 2])
    r = Dt.padded_to_ragged(v, f)
    assert r["face_offsets"].tolist() == [0, 13, 14, 54] and r["face_offsets"].dtype == torch.int32
    assert r["faces"].shape == (54, 3) and int(r["faces"].min()) >= 0
    nv = [int(f[i][: counts[i]].max()) + 1 for i in range(3)]
    assert r["vertex_offsets"].tolist() == [0, nv[0], nv[0] + nv[1], sum(nv)]
    for i in range(3):
        fo, vo = r["face_offsets"], r["vertex_offsets"]
        assert torch.equal(r["faces"][fo[i]:fo[i + 1]], f[i, : counts[i]])                        # vertex ids stay mesh-local
        assert torch.equal(r["vertices"][vo[i]:vo[i + 1]], v[i, : nv[i]])
    # dict items, the reference's dataset item form
    r2 = Dt.ragged_collate([dict(vertices=v[i, : nv[i]], faces=f[i, : counts[i]]) for i in range(3)])
    assert all(torch.equal(r[k], r2[k]) for k in r)
    # flat codes -> the padded tensor tokenize returns
    codes = torch.arange(54 * 3 * 2).view(54 * 3, 2)
    padded = Dt.ragged_to_padded_codes(codes, r["face_offsets"], 3)
    assert padded.shape == (3, 40 * 3, 2)
    assert torch.equal(padded[1, :3], codes[13 * 3: 14 * 3]) and bool((padded[1, 3:] == -1).all())
    assert torch.equal(padded[2], codes[14 * 3:])
    # padding that is not trailing is refused (the reference's prefix assumption, meshgpt_pytorch.py:748-750)
    g = f.clone()
    g[0, 2] = -1
    with pytest.raises(AssertionError):
        Dt.padded_to_ragged(v, g)
    empty = Dt.ragged_collate([])
    assert empty["face_offsets"].tolist() == [0] and empty["faces"].shape == (0, 3)
    # a rank's share of a ragged collection: contiguous meshes, cut by the number of faces, offsets re-based
    cuts = sharding.balanced_contiguous_shards(counts, 2)
    parts = [Dt.ragged_slice(r, lo, hi) for lo, hi in cuts]
    assert cuts == [(0, 2), (2, 3)] and parts[0]["face_offsets"].tolist() == [0, 13, 14] and parts[1]["face_offsets"].tolist() == [0, 40]
    assert torch.equal(torch.cat([p["faces"] for p in parts]), r["faces"]) and torch.equal(torch.cat([p["vertices"] for p in parts]), r["vertices"])
    assert parts[1]["vertex_offsets"].tolist() == [0, nv[2]] and Dt.ragged_slice(r, 1, 1)["faces"].shape == (0, 3)


def test_tokenize_ragged_needs_cuda_tensors():
    m = M.MeshAutoencoder(**SMALL_CFG)
    with pytest.raises((RuntimeError, AssertionError)):
        m.tokenize_ragged(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.long), [0, 1], [0, 3])
