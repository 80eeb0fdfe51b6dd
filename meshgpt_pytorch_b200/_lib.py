"""ctypes binding of libmeshtok.so (see include/meshtok.h).  There is no CPU implementation: if the
library is missing the import of this module fails loudly, it never falls back to anything."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MESHTOK_LIBRARY") or os.path.join(HERE, "lib", "libmeshtok.so")   # override: A/B of two builds

MAX_SAGE_LAYERS = 8
ABI_VERSION = 3

WS_STATUS_VERTEX_RANGE = 1
WS_STATUS_EDGE_OVERFLOW = 2
WS_STATUS_EDGE_RANGE = 4
WS_STATUS_RAGGED_BOUNDS = 8


class SageLayer(C.Structure):
    _fields_ = [
        ("dim_in", C.c_int32), ("dim_out", C.c_int32),
        ("lin_w", C.c_void_p), ("lin_b", C.c_void_p), ("cat_w", C.c_void_p), ("lin_l_b", C.c_void_p),
        ("lin_w_img", C.c_void_p), ("cat_w_img", C.c_void_p),
    ]


class Model(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("nvf", C.c_int32), ("dim_codebook", C.c_int32), ("num_quantizers", C.c_int32),
        ("codebook_size", C.c_int32), ("use_lfq", C.c_int32), ("num_sage_layers", C.c_int32),
        ("num_discrete_coors", C.c_int32), ("num_discrete_angle", C.c_int32), ("num_discrete_area", C.c_int32),
        ("num_discrete_normals", C.c_int32),
        ("coor_lo", C.c_float), ("coor_hi", C.c_float),
        ("num_slots", C.c_int32),
        ("folded_tables", C.c_void_p), ("project_in_bias", C.c_void_p),
        ("sage", SageLayer * MAX_SAGE_LAYERS),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
        ("pdc_w", C.c_void_p), ("pdc_b", C.c_void_p), ("pdc_w_img", C.c_void_p),
        ("codebook", C.c_void_p), ("codebook_sqnorm", C.c_void_p), ("codebook_f16_tiles", C.c_void_p),
        ("codebook_image_scale", C.c_float), ("codebook_max_norm", C.c_float),
        ("lfq_in_w", C.c_void_p), ("lfq_in_b", C.c_void_p), ("lfq_out_w", C.c_void_p), ("lfq_out_b", C.c_void_p),
    ]


class TokenizeArgs(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("faces", C.c_void_p), ("face_edges", C.c_void_p), ("encoded_in", C.c_void_p),
        ("B", C.c_int32), ("F", C.c_int32), ("V", C.c_int32), ("E", C.c_int32),
        ("pad_id", C.c_int64), ("edge_capacity", C.c_int64),
        ("codes_out", C.c_void_p), ("encoded_out", C.c_void_p), ("face_coords_out", C.c_void_p),
        ("quantized_out", C.c_void_p), ("histogram", C.c_void_p),
        ("stop_after_encode", C.c_int32), ("rvq_exact", C.c_int32), ("gemm_simt", C.c_int32), ("keep_taps", C.c_int32),
        ("face_offsets", C.c_void_p), ("vertex_offsets", C.c_void_p), ("total_faces", C.c_int64),
        ("edge_offsets", C.c_void_p), ("total_edges", C.c_int64), ("total_vertices", C.c_int64),
    ]


class PipelineSlot(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("faces", C.c_void_p), ("codes", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("codes_host", C.c_void_p), ("status_host", C.c_void_p),
    ]


class PipelineResult(C.Structure):
    _fields_ = [("ticket", C.c_int64), ("slot", C.c_int32), ("status", C.c_int32), ("codes_host", C.c_void_p), ("tokens", C.c_int64)]


class PipelineRaggedSlot(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p), ("faces", C.c_void_p), ("offsets", C.c_void_p), ("codes", C.c_void_p), ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t), ("codes_host", C.c_void_p), ("status_host", C.c_void_p), ("offsets_host", C.c_void_p),
    ]


class VertexCodeSource(C.Structure):
    _fields_ = [("faces", C.c_void_p), ("face_row", C.c_void_p), ("vert_row", C.c_void_p), ("vertex_codes", C.c_void_p),
                ("F", C.c_int32), ("nvf", C.c_int32), ("V", C.c_int32), ("Q", C.c_int32)]


class Taps(C.Structure):
    _fields_ = [
        ("status", C.c_int64), ("counts", C.c_int64), ("face_row", C.c_int64), ("vert_row", C.c_int64),
        ("indptr", C.c_int64), ("src", C.c_int64), ("x0", C.c_int64),
        ("layer_out", C.c_int64 * MAX_SAGE_LAYERS), ("averaged", C.c_int64), ("vertex_codes", C.c_int64),
        ("rvq_stats", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/meshtok.h declares
SIGNATURES = {
    "meshtok_version": (C.c_int, []),
    "meshtok_last_error_string": (C.c_char_p, []),
    "meshtok_kernel_launch_count": (C.c_int64, []),
    "meshtok_debug_linear_trace": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "meshtok_debug_gather_trace": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "meshtok_dec_dequantize_rvq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_dequantize_lfq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_im2col_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_pixelnorm_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "meshtok_dec_silu_layernorm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "meshtok_dec_squeeze_excite_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "meshtok_dec_gate_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_argmax_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "meshtok_halo_image_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_conv1d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "meshtok_conv1d_pixelnorm_silu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_conv1d_argmax_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "meshtok_dec_se_slices": (C.c_int, [C.c_int, C.c_int]),
    "meshtok_dec_rows_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_dequantize_rvq_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_silu_layernorm_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_pixelnorm_silu_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_pixelnorm_squeeze_excite_gate": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "meshtok_dec_gate_residual_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_dec_argmax_coords_padded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "meshtok_frontend_append_eos": (C.c_int, [C.c_void_p, C.POINTER(VertexCodeSource), C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "meshtok_frontend_fine_tokens": (C.c_int, [C.c_void_p, C.POINTER(VertexCodeSource), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "meshtok_frontend_face_tokens": (C.c_int, [C.c_void_p, C.POINTER(VertexCodeSource), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_void_p]),
    "meshtok_profile_enable": (C.c_int, [C.c_int]),
    "meshtok_profile_sections": (C.c_int, []),
    "meshtok_profile_section_name": (C.c_char_p, [C.c_int]),
    "meshtok_profile_collect": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]),
    "meshtok_topology_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_face_edges_count": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "meshtok_face_edges_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                          C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_tokenize_workspace_bytes": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                                   C.POINTER(C.c_size_t)]),
    "meshtok_tokenize_workspace_bytes_ragged": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                                          C.POINTER(C.c_size_t)]),
    "meshtok_tokenize_taps_ragged": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.POINTER(Taps)]),
    "meshtok_tokenize": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_pipeline_create": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(PipelineSlot), C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "meshtok_pipeline_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PipelineResult)]),
    "meshtok_pipeline_create_ragged": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                                 C.POINTER(PipelineRaggedSlot), C.c_int, C.POINTER(C.c_void_p)]),
    "meshtok_pipeline_submit_ragged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(PipelineResult)]),
    "meshtok_pipeline_drain": (C.c_int, [C.c_void_p, C.POINTER(PipelineResult), C.POINTER(C.c_int)]),
    "meshtok_pipeline_stream": (C.c_void_p, [C.c_void_p, C.c_int]),
    "meshtok_pipeline_destroy": (C.c_int, [C.c_void_p]),
    "meshtok_stage_topology": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_stage_face_features_embed": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_stage_sage_project": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_stage_sage_aggregate_linear": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_size_t, C.c_void_p]),
    "meshtok_stage_project_scatter_mean": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_stage_gather_codes": (C.c_int, [C.POINTER(Model), C.POINTER(TokenizeArgs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "meshtok_tokenize_taps": (C.c_int, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(Taps)]),
    "meshtok_lfq_codes": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_rvq_workspace_bytes": (C.c_int, [C.POINTER(Model), C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_rvq_codes": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "meshtok_rvq_codebook_image_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_rvq_build_codebook_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "meshtok_linear_image_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_linear_build_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_rows_image_bytes": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "meshtok_rows_build_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "meshtok_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p]),
}


class MeshtokError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python meshgpt_pytorch_b200/build.py` (needs nvcc). "
            "meshgpt_pytorch_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    got = lib.meshtok_version()
    if got != ABI_VERSION:
        raise ImportError(f"libmeshtok ABI {got} != binding ABI {ABI_VERSION}: rebuild the library")
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise MeshtokError(f"libmeshtok error {rc}: {lib.meshtok_last_error_string().decode()}")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
