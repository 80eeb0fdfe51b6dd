"""CPU: the C-ABI library loads, exports every symbol include/meshtok.h declares, and its host-side argument
checks answer without a GPU (no compute is launched here)."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "meshtok.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(meshtok_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from meshgpt_pytorch_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/meshtok.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.lib.meshtok_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    from meshgpt_pytorch_b200 import _lib
    # meshtok_sage_layer: 2 x int32 + 6 pointers; meshtok_model embeds 8 of them
    assert C.sizeof(_lib.SageLayer) == 8 + 6 * 8
    assert _lib.Model.sage.size == 8 * C.sizeof(_lib.SageLayer)
    assert C.sizeof(_lib.TokenizeArgs) == 4 * 8 + 4 * 4 + 2 * 8 + 5 * 8 + 4 * 4 + 6 * 8
    assert C.sizeof(_lib.PipelineSlot) == 7 * 8 and C.sizeof(_lib.PipelineResult) == 8 + 4 + 4 + 8 + 8
    assert C.sizeof(_lib.PipelineRaggedSlot) == 9 * 8
    assert C.sizeof(_lib.Taps) == (7 + 8 + 3) * 8


def test_host_side_argument_checks():
    from meshgpt_pytorch_b200 import _lib
    lib = _lib.lib
    n = C.c_size_t()
    assert lib.meshtok_topology_workspace_bytes(64, 800, 3, 441, 0, C.byref(n)) == 0 and n.value > 64 * 800 * 4
    assert lib.meshtok_topology_workspace_bytes(2, 10, 5, 10, 0, C.byref(n)) == -1          # nvf must be 3 or 4
    assert b"nvf" in lib.meshtok_last_error_string()
    assert lib.meshtok_topology_workspace_bytes(1, 200000, 3, 100000, 0, C.byref(n)) == -2  # too large for one SM
    assert lib.meshtok_linear_image_bytes(576, 512, C.byref(n)) == 0 and n.value == 576 * 512 * 4
    assert lib.meshtok_linear_image_bytes(100, 30, C.byref(n)) == -2
    assert lib.meshtok_rvq_codebook_image_bytes(16384, 192, C.byref(n)) == 0 and n.value == 16384 * 192 * 2 + (16384 // 128) * 4096
    assert lib.meshtok_profile_sections() >= 10
    assert lib.meshtok_profile_section_name(0) == b"topology"
    m = _lib.Model()
    assert lib.meshtok_tokenize_workspace_bytes(C.byref(m), 1, 1, 1, 0, 16, C.byref(n)) == -1  # abi_version 0
    with pytest.raises(_lib.MeshtokError):
        _lib.check(-1)


def test_ragged_workspace_is_sized_by_the_sum_of_the_mesh_sizes():
    """BASELINE config 4: 256 meshes of 100..800 faces - about 116 k faces in 204 800 padded slots.  The ragged layout's workspace is
    sized by total_faces / total_vertices (pure host arithmetic: no GPU needed)."""
    from meshgpt_pytorch_b200 import _lib
    lib = _lib.lib
    m = _lib.Model()
    m.abi_version = _lib.ABI_VERSION
    m.nvf, m.dim_codebook, m.num_quantizers, m.codebook_size, m.use_lfq, m.num_slots = 3, 192, 2, 16384, 1, 16
    dims = [192, 64, 128, 256, 256, 576]
    m.num_sage_layers = 5
    for i in range(5):
        m.sage[i].dim_in, m.sage[i].dim_out = dims[i], dims[i + 1]
    B, F, V, cap = 256, 800, 441, 8 * 256 * 800
    padded, ragged, same = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _lib.check(lib.meshtok_tokenize_workspace_bytes(C.byref(m), B, F, V, 0, cap, C.byref(padded)))
    _lib.check(lib.meshtok_tokenize_workspace_bytes_ragged(C.byref(m), B, F, V, 0, cap, 116206, 65000, C.byref(ragged)))
    _lib.check(lib.meshtok_tokenize_workspace_bytes_ragged(C.byref(m), B, F, V, 0, cap, 0, 0, C.byref(same)))
    assert same.value == padded.value
    assert ragged.value < 0.62 * padded.value, (ragged.value, padded.value)
    assert lib.meshtok_tokenize_workspace_bytes_ragged(C.byref(m), B, F, V, 0, cap, -1, 0, C.byref(ragged)) == -1
