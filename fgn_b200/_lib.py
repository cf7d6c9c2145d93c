"""ctypes binding of libfgn_b200.so (declared in include/fgn_b200.h).

The library is built in-tree by ``fgn_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is
no CPU or PyTorch fallback anywhere in this package: if the shared object is missing or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfgn_b200.so")

FGN_MAX_LEVELS = 8
ABI_VERSION = 3
LAYOUT_NCHW = 0
LAYOUT_NHWC = 1


class FgnError(RuntimeError):
    """A libfgn_b200 entry point returned a negative status."""


class Pyramid(Structure):
    _fields_ = [
        ("num_levels", c_int),
        ("feat", c_void_p * FGN_MAX_LEVELS),
        ("H", c_int * FGN_MAX_LEVELS),
        ("W", c_int * FGN_MAX_LEVELS),
        ("spatial_scale", c_float * FGN_MAX_LEVELS),
    ]


_P = c_void_p  # device pointers travel as integers

# name -> (restype, argtypes); must list every symbol include/fgn_b200.h declares
SIGNATURES = {
    "fgn_abi_version": (c_int, []),
    "fgn_last_error_string": (c_char_p, []),
    "fgn_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "fgn_launch_count": (c_uint64, []),
    "fgn_rpn_proposals_workspace_bytes": (c_size_t, [POINTER(c_int), POINTER(c_int), c_int, c_int, c_int, c_int]),
    "fgn_rpn_proposals": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                  c_int, c_int, c_int, _P, _P, POINTER(c_float), POINTER(c_float), c_float,
                                  c_int, c_float, c_int, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "fgn_det_postprocess_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_det_postprocess": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                    POINTER(c_float), POINTER(c_float), c_float, c_float, c_float, c_int,
                                    _P, _P, _P, _P, c_size_t, _P]),
    "fgn_fold_attention_weights": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_mask_paste_rle_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_mask_paste_rle": (c_int, [_P, _P, c_int, _P, _P, c_int, c_int, c_float, c_int, c_int, _P, _P, _P, _P, c_int, c_int,
                                   _P, c_size_t, _P]),
    "fgn_mask_rle_encode": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    "fgn_mask_paste": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P]),
    "fgn_debug_roi_window_violations": (ctypes.c_uint, []),
    "fgn_map_roi_levels": (c_int, [_P, c_int, c_int, c_float, _P, _P]),
    "fgn_roi_align_ml_fwd": (c_int, [POINTER(Pyramid), c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int,
                                     c_float, _P, _P, _P, c_int, _P, _P]),
    "fgn_roi_align_ml_fwd_direct": (c_int, [POINTER(Pyramid), c_int, c_int, c_int, _P, c_int, c_int, c_int,
                                            c_int, c_float, _P, _P, _P, c_int, _P, _P]),
    "fgn_roi_align_sample_indices": (c_int, [POINTER(Pyramid), _P, c_int, c_int, c_int, c_int, c_float, c_int,
                                             _P, _P, _P, _P, _P]),
    "fgn_nchw_to_nhwc": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_nhwc_to_nchw": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_support_mask_pool": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_support_prologue_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fgn_support_prologue_fwd": (c_int, [POINTER(Pyramid), c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float,
                                         _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "fgn_support_pool": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P]),
    "fgn_attention_vectors_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "fgn_attention_vectors": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "fgn_channel_attention": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_attention_vectors_ml_workspace_bytes": (c_size_t, [POINTER(Pyramid), c_int, c_int, c_int]),
    "fgn_attention_vectors_ml": (c_int, [POINTER(Pyramid), c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "fgn_channel_attention_ml": (c_int, [POINTER(Pyramid), _P, c_int, c_int, c_int, POINTER(c_void_p), _P]),
    "fgn_best_class_select": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "fgn_relation_fusion_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_relation_split_weights_bytes": (c_size_t, [c_int]),
    "fgn_relation_split_weights": (c_int, [_P, c_int, _P, _P]),
    "fgn_relation_fusion_fwd": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                        _P, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P,
                                        _P, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "fgn_relation_fusion_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_relation_fusion_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                        _P, _P, _P, c_int, c_float, _P, _P,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "fgn_gemm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fgn_gemm_nt": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "fgn_gemm_nt_presplit": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "fgn_conv1x1_nhwc": (c_int, [_P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "fgn_conv_split_weights_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fgn_conv_split_weights": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "fgn_conv3x3_nhwc": (c_int, [_P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "fgn_deconv2x2_logits_nhwc": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          _P, c_size_t, _P]),
    "fgn_gemm_nt_bf16": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "fgn_cls_bbox_reassemble": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "fgn_attention_vectors_ml_bf16": (c_int, [POINTER(Pyramid), c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "fgn_channel_attention_ml_bf16": (c_int, [POINTER(Pyramid), _P, c_int, c_int, c_int, POINTER(c_void_p), _P]),
    "fgn_roi_align_ml_fwd_bf16": (c_int, [POINTER(Pyramid), c_int, c_int, _P, c_int, c_int, c_int, c_int, c_float,
                                          _P, _P, _P, c_int, _P, _P]),
    "fgn_guided_roi_fused_bf16_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_guided_roi_fused_fwd_bf16": (c_int, [POINTER(Pyramid), c_int, c_int, _P, c_int, c_int, c_int, c_int, c_float,
                                              _P, c_int, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P,
                                              _P, _P, _P, _P, c_size_t, _P]),
    "fgn_roi_align_ml_bwd": (c_int, [POINTER(Pyramid), c_int, c_int, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "fgn_roi_align_ml_bwd_det_workspace_bytes": (c_size_t, [POINTER(Pyramid), c_int, c_int]),
    "fgn_roi_align_ml_bwd_det": (c_int, [POINTER(Pyramid), c_int, c_int, _P, c_int, c_int, c_int, c_int, c_float, _P, c_int, _P, _P,
                                         _P, c_size_t, _P]),
    "fgn_channel_attention_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "fgn_channel_attention_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "fgn_attention_vectors_bwd": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_support_pool_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "fgn_guided_roi_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fgn_guided_roi_fused_fwd": (c_int, [POINTER(Pyramid), c_int, c_int, _P, c_int, c_int, c_int, c_int, c_float,
                                         _P, _P, c_int, _P, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P,
                                         _P, _P, _P, c_int, _P, c_size_t, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libfgn_b200.so (once) and attach argument types.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FgnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C fgn_b200/csrc`.  fgn_b200 has no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.fgn_abi_version() != ABI_VERSION:
        raise FgnError(f"libfgn_b200 ABI {lib.fgn_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fgn_last_error_string()
        raise FgnError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
