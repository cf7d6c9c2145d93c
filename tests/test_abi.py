"""The C-ABI library loads and exports every symbol include/fgn_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fgn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fgn_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ["fgn_map_roi_levels", "fgn_roi_align_ml_fwd", "fgn_roi_align_sample_indices",
                 "fgn_support_mask_pool", "fgn_support_pool", "fgn_attention_vectors", "fgn_channel_attention",
                 "fgn_best_class_select", "fgn_relation_fusion_fwd", "fgn_cls_bbox_reassemble",
                 "fgn_guided_roi_fused_fwd", "fgn_abi_version", "fgn_last_error_string"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from fgn_b200 import _lib
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/fgn_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in fgn_b200/_lib.py"
    assert lib.fgn_abi_version() == 3


def test_zero_size_calls_do_not_touch_the_device():
    from fgn_b200 import _lib
    lib = _lib.load()
    # R == 0 returns FGN_OK before any CUDA call (reference early-outs fgn_roi_head.py:558-567)
    assert lib.fgn_map_roi_levels(None, 0, 4, 56.0, None, None) == 0
    assert lib.fgn_support_mask_pool(None, None, 0, 64, 64, 7, None, None) == 0
    assert lib.fgn_cls_bbox_reassemble(None, None, 0, 3, None, None, None) == 0
    pyr = _lib.Pyramid()
    pyr.num_levels = 1
    pyr.H[0], pyr.W[0], pyr.spatial_scale[0] = 8, 8, 1.0 / 16
    assert lib.fgn_roi_align_ml_fwd(ctypes.byref(pyr), 1, 64, 1, None, 0, 7, 0, 1, 56.0, None, None, None, 0, None, None) == 0


def test_bad_arguments_set_the_error_string():
    from fgn_b200 import _lib
    lib = _lib.load()
    assert lib.fgn_map_roi_levels(None, 5, 99, 56.0, None, None) == -1
    assert b"num_levels" in lib.fgn_last_error_string()
    assert lib.fgn_relation_fusion_fwd(None, 0, None, None, None, 4, 1, 1, 30, 7, None, None, None, None, None, 32, 1e-5,
                                       None, None, None, None, None, None, None, None, 0, None, 0, None) == -1
    assert b"GroupNorm" in lib.fgn_last_error_string()
    with pytest.raises(_lib.FgnError):
        _lib.check(-1, "x")


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from fgn_b200 import ops, FgnError
    with pytest.raises(FgnError):
        ops.map_roi_levels(torch.zeros(3, 5), 4)
    with pytest.raises(FgnError):
        ops.roi_align_multilevel([torch.zeros(1, 4, 8, 8)], torch.zeros(2, 5), [1 / 16])
    with pytest.raises(FgnError):
        ops.channel_attention(torch.zeros(1, 4, 8, 8), torch.zeros(1, 2, 4, 1, 1))
    with pytest.raises(FgnError):
        ops.fold_attention_weights(torch.zeros(4, 4, 3, 3), torch.zeros(2, 4))
    with pytest.raises(FgnError):
        ops.mask_paste(torch.zeros(2, 1, 28, 28), torch.zeros(2, 4), 32, 32)
    with pytest.raises(FgnError):
        ops.mask_paste_rle(torch.zeros(2, 1, 28, 28), torch.zeros(2, 4), [(32, 32)])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fgn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                src = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{path} imports oracle"
                assert "liboracle" not in src and "import torchvision" not in src and "ops.torchvision" not in src, \
                    f"{path} reaches for a CPU op"
            elif f.endswith((".cu", ".cuh")):
                src = open(path).read()
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f"{path} includes oracle code"
