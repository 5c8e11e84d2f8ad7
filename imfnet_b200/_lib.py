"""ctypes binding of include/imfnet_b200.h (the in-tree libimfnet_b200.so).  No fallbacks: if the library
is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libimfnet_b200.so")


def lib_path() -> str:
    return LIB_PATH

_p, _i32, _i64, _sz, _f32, _f64 = C.c_void_p, C.c_int32, C.c_longlong, C.c_size_t, C.c_float, C.c_double


class KmapJob(C.Structure):
    _fields_ = [("out_coords", _p), ("n_out_dev", _p), ("table_in", _p), ("nbr_t", _p), ("tile_mask", _p), ("perm", _p), ("scale", _i32),
                ("dense_meta", _p), ("dense_cells", _p)]


class AttnWeights(C.Structure):
    _fields_ = [(n, _p) for n in ("ln_q_w", "ln_q_b", "ln_c_w", "ln_c_b", "wq", "wkv", "wo", "bo", "ln_f_w", "ln_f_b",
                                  "w1", "b1", "w2", "b2")] + [("latent", _i32), ("dim", _i32), ("inner", _i32)]


class AttnPacked(C.Structure):
    _fields_ = [(n, _p) for n in ("wq", "wkv", "wo", "w1", "w2")] + [(n, _f32) for n in ("mq", "mkv", "mo", "m1", "m2")]


# name -> (restype, argtypes); mirrors include/imfnet_b200.h one to one (tests/test_abi.py checks the header).
SIGNATURES = {
    "imf_last_error": (C.c_char_p, []),
    "imf_version": (C.c_int, []),
    "imf_launch_count": (_i64, []),
    "imf_device_sm_count": (C.c_int, []),
    "imf_hash_capacity": (_i64, [_i64]),
    "imf_hash_bytes": (_sz, [_i64]),
    "imf_hash_clear": (C.c_int, [_p, _i64, _p]),
    "imf_hash_build": (C.c_int, [_p, _p, _i32, _p, _i64, _p, _p]),
    "imf_stride_map_workspace_bytes": (_sz, [_i32]),
    "imf_stride_map": (C.c_int, [_p, _p, _i32, _i32, _p, _i64, _p, _p, _p, _p, _sz, _p, _p]),
    "imf_kernel_map": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _i32, _p, _p]),
    "imf_kernel_map_t": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _i32, _p, _i32, _p, _p]),
    "imf_kernel_map_t_batch": (C.c_int, [C.POINTER(KmapJob), _i32, _i32, _i64, _i32, _i32, _p]),
    "imf_sparse_conv_g4_workspace_bytes": (_sz, [_i32]),
    "imf_sparse_conv_g4_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i32,
                                        _p, _i32, _i32, _i32, _p, _sz, _p, _p]),
    "imf_sparse_conv_g4_fwd_perm": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i32,
                                             _p, _i32, _i32, _i32, _p, _p, _sz, _p, _p]),
    "imf_parity_perm_workspace_bytes": (_sz, [_i32]),
    "imf_parity_perm": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _sz, _p]),
    "imf_debug_conv_g4_trace": (C.c_int, [_p, _i32, _i32, _i32]),
    "imf_quantize_points": (C.c_int, [_p, _i32, _f64, _i32, _p, _p]),
    "imf_quantize_points_f32": (C.c_int, [_p, _i32, _f32, _i32, _p, _p]),
    "imf_batch_segments": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "imf_batch_segments_n": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "imf_sparse_conv_fwd": (C.c_int, [_p, _i32, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _p, _i32, _p]),
    "imf_h2_pack": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _i32, _p, _p]),
    "imf_h2_unpack": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "imf_h2_pack_n": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p, _p]),
    "imf_h2_unpack_n": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "imf_h2_pack_scaled_n": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _f32, _p, _i32, _p, _p]),
    "imf_h2_unpack_scaled_n": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _f32, _p, _i32, _p]),
    "imf_h2_unpack_l2norm": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "imf_identity_table": (C.c_int, [_p, _i32, _p, _i32, _p, _p]),
    "imf_tc_gemm_m": (C.c_int, [_p, _i32, _p, _i32, _p, _i32, _i32, _p, _i32, _i32, _f32, _p, _p, _i32, _i32, _p, _sz, _p, _p]),
    "imf_attention_fusion_fwd_m": (C.c_int, [C.POINTER(AttnWeights), _p, _i32, _i32, _p, _p, _i32, _p, _i32, _p, _sz, _p]),
    "imf_sparse_conv_h2_packed_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "imf_sparse_conv_h2_pack": (C.c_int, [_p, _i32, _i32, _i32, _i32, _f32, _p, _p]),
    "imf_conv_first_tc_columns": (_i32, [_i32]),
    "imf_conv_first_tc_workspace_bytes": (_sz, [_i32, _i32]),
    "imf_conv_first_tc_h2_fwd": (C.c_int, [_p, _i32, _p, _p, _p, _i32, _i32, _p, _i64, _i32, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _sz, _p, _p]),
    "imf_conv_first_tc_h2_fwd_keep": (C.c_int, [_p, _i32, _p, _p, _p, _i32, _i32, _p, _i64, _i32, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _sz, _p, _p]),
    "imf_conv_first_tc_release": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _sz, _p]),
    "imf_conv_first_tc_grid": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "imf_conv_first_h2_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i32, _p, _i64, _i32, _i32, _i32, _p, _p, _i32, _p, _i32, _i32, _p]),
    "imf_tail_fused_h2_fwd": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _i32, _p, _p]),
    "imf_image_p8_bytes": (_sz, [_i32, _i32, _i32]),
    "imf_image_maxpool_p8": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "imf_image_conv3x3_p8_fwd": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _p, _i32, _i32, _p, _p]),
    "imf_image_stem_workspace_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "imf_image_stem_h2_fwd": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p, _i32, _p, _p]),
    "imf_pointwise_tail_h2_fwd": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _p, _i32, _p]),
    "imf_conv_first_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i32, _p, _i64, _i32, _i32, _i32, _p, _p, _i32, _p, _i32, _p]),
    "imf_pointwise_tail_fwd": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _i32, _p]),
    "imf_image_conv_table": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _p, _i32, _p, _p]),
    "imf_image_im2col_h2": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "imf_image_maxpool_h2": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "imf_image_im2col_h2_batch": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p]),
    "imf_image_maxpool_h2_batch": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p]),
    "imf_transpose_tokens": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "imf_nn_search_workspace_bytes": (_sz, [_i32]),
    "imf_nn_search": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "imf_nn_search_tc_workspace_bytes": (_sz, [_i32, _i32]),
    "imf_nn_search_tc": (C.c_int, [_p, _i32, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "imf_linear_fwd": (C.c_int, [_p, _i32, _p, _p, _i32, _i32, _i32, _p, _i32, _p]),
    "imf_tc_gemm_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "imf_tc_gemm": (C.c_int, [_p, _i32, _p, _i32, _p, _i32, _i32, _i32, _i32, _f32, _p, _p, _i32, _i32, _p, _sz, _p, _p]),
    "imf_attention_kv_bytes": (_sz, [_i32, _i32]),
    "imf_attention_kv_workspace_bytes": (_sz, [_i32, _i32]),
    "imf_attention_kv": (C.c_int, [C.POINTER(AttnWeights), _p, _i32, _i32, _p, _p, _sz, _p]),
    "imf_attention_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "imf_attention_kv_batched_bytes": (_sz, [_i32, _i32]),
    "imf_attention_kv_batched_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "imf_attention_kv_batched": (C.c_int, [C.POINTER(AttnWeights), C.POINTER(AttnPacked), _p, _i32, _i32, _p, _p, _sz, _p, _p]),
    "imf_attention_batched_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "imf_attention_fusion_fwd_batched": (C.c_int, [C.POINTER(AttnWeights), C.POINTER(AttnPacked), _p, _i32, _i32, _p, _p, _p, _i32, _p, _i32, _p, _i32,
                                                   _p, _sz, _p, _p]),
    "imf_h2_gemm": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _i32, _f32, _p, _p, _i32, _i32, _p, _i32, _p, _p]),
    "imf_attention_fusion_fwd": (C.c_int, [C.POINTER(AttnWeights), _p, _i32, _i32, _p, _i32, _p, _i32, _p, _sz, _p]),
}

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built (python -m imfnet_b200.build)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -m imfnet_b200.build` "
                               "(the CUDA extension is required; there is no CPU fallback)")
        l = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def sm_count() -> int:
    """SMs of the current CUDA device (the persistent kernels' grid size)."""
    return int(lib().imf_device_sm_count())


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"imfnet_b200 C-ABI call failed ({rc}): {lib().imf_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device (got {t.device}); imfnet_b200 has no CPU path")
