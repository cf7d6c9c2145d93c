"""Tensor-level wrappers over the C ABI (include/fgn_b200.h).

PyTorch is used for device memory and streams only; every computation below is a call into
libfgn_b200.so.  There is no CPU path: CPU tensors raise ``FgnError``.

Layouts: tensors keep the reference's *logical* NCHW shapes.  Storage may be contiguous NCHW
(the reference's layout) or ``torch.channels_last`` (NHWC, the fast layout).  Kernels read
either; outputs follow ``out_format``.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, FgnError, Pyramid

__all__ = [
    "map_roi_levels", "roi_align_multilevel", "roi_align_sample_indices", "to_nhwc", "support_mask_pool",
    "support_pool", "attention_vectors", "channel_attention", "attention_multilevel", "best_class_select",
    "relation_fusion", "guided_roi_fused", "cls_bbox_reassemble", "gemm_nt", "launch_count", "mask_rle_encode", "conv1x1",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_CONST_CACHE: "dict" = {}


def _const_tensor(values, dtype: torch.dtype, device) -> torch.Tensor:
    """A small constant (per-image offsets, image sizes, scale factors) as a PERSISTENT device tensor.

    These used to be staged through temporary pinned tensors; a CUDA-graph capture of that copy bakes the pinned
    host pointer into the graph, and the temporary goes back to the allocator right after capture -- a later replay
    would read whatever occupies the block by then.  Here the device tensor itself is kept (keyed by its values), so
    a captured kernel reads memory that stays valid and constant; a value set first seen during capture raises instead
    of capturing a host copy (warm up once before capturing, as torch.cuda.graph requires anyway)."""
    dev = torch.device(device)
    flat = tuple(float(v) if dtype.is_floating_point else int(v) for v in torch.as_tensor(values).reshape(-1).tolist())
    shape = tuple(torch.as_tensor(values).shape)
    key = (dev.type, dev.index, dtype, shape, flat)
    t = _CONST_CACHE.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise FgnError("constant operands first seen during CUDA-graph capture: run the call once before capturing")
        if len(_CONST_CACHE) > 4096:
            _CONST_CACHE.clear()
        t = torch.tensor(flat, dtype=dtype).reshape(shape).to(dev)
        _CONST_CACHE[key] = t
    return t


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise FgnError("fgn_b200 has no CPU fallback: expected CUDA tensors, got device=%s" % t.device)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise FgnError(f"{name}: expected float32, got {t.dtype}")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def storage_layout(t: torch.Tensor) -> Optional[int]:
    """LAYOUT_NHWC / LAYOUT_NCHW for a densely packed 4-D tensor, None if it is neither."""
    if t.dim() != 4:
        return None
    b, c, h, w = t.shape
    st = t.stride()

    def ok(expected):
        # strides of size-1 dims are irrelevant
        return all(s == e or n == 1 for s, e, n in zip(st, expected, t.shape))

    nchw = ok((c * h * w, h * w, w, 1))
    nhwc = ok((h * w * c, 1, w * c, c))
    if nhwc and not nchw:
        return LAYOUT_NHWC
    if nchw:
        return LAYOUT_NCHW
    return None


def _dense(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    lay = storage_layout(t)
    if lay is None:
        t = t.contiguous()
        lay = LAYOUT_NCHW
    return t, lay


def _empty_like_format(shape, device, layout: int) -> torch.Tensor:
    """Logical NCHW-shaped fp32 tensor whose storage is NCHW or NHWC."""
    n, c, h, w = shape
    if layout == LAYOUT_NHWC:
        return torch.empty((n, h, w, c), device=device, dtype=torch.float32).permute(0, 3, 1, 2)
    return torch.empty((n, c, h, w), device=device, dtype=torch.float32)


def _fmt(name: str) -> int:
    if name in ("nchw", "contiguous"):
        return LAYOUT_NCHW
    if name in ("nhwc", "channels_last"):
        return LAYOUT_NHWC
    raise ValueError(f"unknown memory format {name!r}")


def launch_count() -> int:
    return int(_lib.load().fgn_launch_count())


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Repack a contiguous NCHW tensor to channels_last storage with the library's own kernel."""
    _need_cuda(x)
    x, lay = _dense(_f32(x, "x"))
    if lay == LAYOUT_NHWC:
        return x
    b, c, h, w = x.shape
    out = _empty_like_format(x.shape, x.device, LAYOUT_NHWC)
    _lib.check(_lib.load().fgn_nchw_to_nhwc(x.data_ptr(), b, c, h, w, out.data_ptr(), _stream()), "fgn_nchw_to_nhwc")
    return out


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    x, lay = _dense(_f32(x, "x"))
    if lay == LAYOUT_NCHW:
        return x
    b, c, h, w = x.shape
    out = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().fgn_nhwc_to_nchw(x.data_ptr(), b, c, h, w, out.data_ptr(), _stream()), "fgn_nhwc_to_nchw")
    return out


# --------------------------------------------------------------------------------------------
def map_roi_levels(rois: torch.Tensor, num_levels: int, finest_scale: float = 56.0) -> torch.Tensor:
    """mmdet SingleRoIExtractor.map_roi_levels; returns int64 like ``.long()`` in the reference."""
    _need_cuda(rois)
    rois = _f32(rois, "rois").contiguous()
    r = rois.shape[0]
    out = torch.empty((r,), device=rois.device, dtype=torch.int32)
    _lib.check(_lib.load().fgn_map_roi_levels(rois.data_ptr(), r, int(num_levels), float(finest_scale),
                                              out.data_ptr(), _stream()), "fgn_map_roi_levels")
    return out.long()


def _is_bf16(feats: Sequence[torch.Tensor]) -> bool:
    return all(f.dtype == torch.bfloat16 for f in feats)


def _make_pyramid(feats: Sequence[torch.Tensor], scales: Sequence[float], dtype=torch.float32):
    if len(feats) != len(scales) or not (1 <= len(feats) <= _lib.FGN_MAX_LEVELS):
        raise FgnError(f"pyramid needs 1..{_lib.FGN_MAX_LEVELS} levels with one scale each")
    keep = []
    lay0 = None
    b0, c0 = feats[0].shape[:2]
    pyr = Pyramid()
    pyr.num_levels = len(feats)
    for i, (f, s) in enumerate(zip(feats, scales)):
        _need_cuda(f)
        if f.dtype != dtype:
            raise FgnError(f"feats[{i}]: expected {dtype}, got {f.dtype}")
        if dtype == torch.bfloat16 and storage_layout(f) != LAYOUT_NHWC:
            raise FgnError("bf16 feature maps must be channels_last (the bf16 variant has no NCHW kernels)")
        f, lay = _dense(f)
        if f.shape[0] != b0 or f.shape[1] != c0:
            raise FgnError("all pyramid levels must share batch and channel dims")
        if lay0 is None:
            lay0 = lay
        elif lay != lay0:   # mixed storage: bring everything to the first level's layout
            f = to_nhwc(f) if lay0 == LAYOUT_NHWC else to_nchw(f)
        keep.append(f)
        pyr.feat[i] = f.data_ptr()
        pyr.H[i] = f.shape[2]
        pyr.W[i] = f.shape[3]
        pyr.spatial_scale[i] = float(s)
    return pyr, keep, lay0, b0, c0


def roi_align_multilevel(feats: Sequence[torch.Tensor], rois: torch.Tensor, scales: Sequence[float],
                         output_size: int = 7, sampling_ratio: int = 0, aligned: bool = True,
                         finest_scale: float = 56.0, chan_scale: Optional[torch.Tensor] = None,
                         scale_index: Optional[torch.Tensor] = None, out_format: str = "nchw",
                         return_levels: bool = False, nchw_input: str = "repack", force_direct: bool = False,
                         out_dtype: Optional[torch.dtype] = None):
    """Level assignment + RoIAlign (avg) in one kernel.  ``feats``: list of [B,C,H_l,W_l].

    ``nchw_input``: what to do with reference-layout (contiguous NCHW) inputs --
    "repack" converts them to channels_last with fgn_nchw_to_nhwc and runs the fast kernel,
    "direct" runs the NCHW kernel that keeps the reference's exact summation order.
    """
    _need_cuda(rois, chan_scale, scale_index)
    rois = _f32(rois, "rois").contiguous()
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise FgnError("rois must be [R,5] (batch_idx, x1, y1, x2, y2)")
    feats = list(feats)
    c = feats[0].shape[1]
    if _is_bf16(feats):       # bf16 variant: NHWC in, NHWC out (bf16 unless out_dtype=torch.float32)
        pyr, keep, lay, b, c = _make_pyramid(feats, scales, torch.bfloat16)
        r, p = rois.shape[0], int(output_size)
        odt = out_dtype or torch.bfloat16
        out = torch.empty((r, p, p, c), device=rois.device, dtype=odt).permute(0, 3, 1, 2)
        lvl = torch.empty((r,), device=rois.device, dtype=torch.int32) if return_levels else None
        if chan_scale is not None:
            chan_scale = _f32(chan_scale, "chan_scale").reshape(-1, c).contiguous()
            if scale_index is not None:
                scale_index = scale_index.to(torch.int32).contiguous()
        _lib.check(_lib.load().fgn_roi_align_ml_fwd_bf16(
            ctypes.byref(pyr), b, c, rois.data_ptr(), r, p, int(sampling_ratio), int(bool(aligned)), float(finest_scale),
            _ptr(chan_scale), _ptr(scale_index), out.data_ptr(), int(odt == torch.bfloat16), _ptr(lvl), _stream()),
            "fgn_roi_align_ml_fwd_bf16")
        return (out, lvl.long()) if return_levels else out
    if nchw_input == "repack" and not force_direct and c % 4 == 0:
        feats = [to_nhwc(f) if storage_layout(f) != LAYOUT_NHWC else f for f in feats]
    pyr, keep, lay, b, c = _make_pyramid(feats, scales)
    r = rois.shape[0]
    p = int(output_size)
    out_lay = _fmt(out_format)
    out = _empty_like_format((r, c, p, p), rois.device, out_lay)
    lvl = torch.empty((r,), device=rois.device, dtype=torch.int32) if return_levels else None
    if chan_scale is not None:
        chan_scale = _f32(chan_scale, "chan_scale").reshape(-1, c).contiguous()
        if scale_index is not None:
            scale_index = scale_index.to(torch.int32).contiguous()
    lib = _lib.load()
    fn = lib.fgn_roi_align_ml_fwd_direct if force_direct else lib.fgn_roi_align_ml_fwd
    _lib.check(fn(ctypes.byref(pyr), b, c, lay, rois.data_ptr(), r, p, int(sampling_ratio), int(bool(aligned)),
                  float(finest_scale), _ptr(chan_scale), _ptr(scale_index), out.data_ptr(), out_lay,
                  _ptr(lvl), _stream()), "fgn_roi_align_ml_fwd")
    return (out, lvl.long()) if return_levels else out


def roi_align_sample_indices(hw: Sequence[Tuple[int, int]], rois: torch.Tensor, scales: Sequence[float],
                             output_size: int = 7, sampling_ratio: int = 0, aligned: bool = True,
                             finest_scale: float = 56.0, max_grid: int = 32):
    """Integer side of roi_align_multilevel: (levels [R], grid [R,2], ytab [R,P,G,3], xtab [R,P,G,3])."""
    _need_cuda(rois)
    rois = _f32(rois, "rois").contiguous()
    pyr = Pyramid()
    pyr.num_levels = len(hw)
    for i, ((h, w), s) in enumerate(zip(hw, scales)):
        pyr.feat[i] = None
        pyr.H[i], pyr.W[i], pyr.spatial_scale[i] = int(h), int(w), float(s)
    r, p = rois.shape[0], int(output_size)
    dev = rois.device
    lvl = torch.empty((r,), device=dev, dtype=torch.int32)
    grid = torch.empty((r, 2), device=dev, dtype=torch.int32)
    ytab = torch.empty((r, p, max_grid, 3), device=dev, dtype=torch.int32)
    xtab = torch.empty((r, p, max_grid, 3), device=dev, dtype=torch.int32)
    _lib.check(_lib.load().fgn_roi_align_sample_indices(
        ctypes.byref(pyr), rois.data_ptr(), r, p, int(sampling_ratio), int(bool(aligned)), float(finest_scale),
        int(max_grid), lvl.data_ptr(), grid.data_ptr(), ytab.data_ptr(), xtab.data_ptr(), _stream()),
        "fgn_roi_align_sample_indices")
    return lvl, grid, ytab, xtab


# --------------------------------------------------------------------------------------------
def support_mask_pool(masks: torch.Tensor, boxes: torch.Tensor, output_size: int = 7) -> torch.Tensor:
    """roi_align(spp_isegmaps.float(), boxes, 7) of fgn_roi_head.py:429 -> [M,1,P,P]."""
    _need_cuda(masks, boxes)
    if masks.dim() == 4:
        masks = masks[:, 0]
    if masks.dtype == torch.bool:
        m8 = masks.contiguous().view(torch.uint8)
    elif masks.dtype == torch.uint8:
        m8 = masks.contiguous()
    else:
        raise FgnError("support masks must be bool/uint8 (binary), got %s" % masks.dtype)
    m, sh, sw = m8.shape
    boxes = _f32(boxes, "boxes").reshape(m, 4).contiguous()
    p = int(output_size)
    out = torch.empty((m, 1, p, p), device=m8.device, dtype=torch.float32)
    _lib.check(_lib.load().fgn_support_mask_pool(m8.data_ptr(), boxes.data_ptr(), m, sh, sw, p, out.data_ptr(),
                                                 _stream()), "fgn_support_mask_pool")
    return out


def support_pool(f: torch.Tensor, m: torch.Tensor, n_ways: int, k_shots: int, out_format: Optional[str] = None):
    """Class mean and masked GAP of fgn_roi_head.py:439-447.

    f [B*N*K,C,P,P], m [B*N*K,1,P,P] -> cat_mean [B,N,C,P,P], masked_gap [B,N,C,1,1].
    """
    _need_cuda(f, m)
    f, lay = _dense(_f32(f, "f"))
    bnk, c, p, _ = f.shape
    if bnk % (n_ways * k_shots):
        raise FgnError(f"{bnk} support maps is not a multiple of N*K={n_ways * k_shots}")
    bn = bnk // k_shots
    m = _f32(m, "m").reshape(bnk, p * p).contiguous()
    out_lay = lay if out_format is None else _fmt(out_format)
    cat = _empty_like_format((bn, c, p, p), f.device, out_lay)
    gap = torch.empty((bn, c), device=f.device, dtype=torch.float32)
    _lib.check(_lib.load().fgn_support_pool(f.data_ptr(), lay, m.data_ptr(), bn, int(k_shots), c, p,
                                            cat.data_ptr(), out_lay, gap.data_ptr(), _stream()), "fgn_support_pool")
    b = bn // n_ways
    return cat.view(b, n_ways, c, p, p) if out_lay == LAYOUT_NCHW else cat.unflatten(0, (b, n_ways)), \
        gap.view(b, n_ways, c, 1, 1)


def support_prologue(spp_feats: Sequence[torch.Tensor], scales: Sequence[float], spp_bboxes: torch.Tensor,
                     spp_isegmaps: torch.Tensor, n_ways: int, k_shots: int, output_size: int = 7, finest_scale: float = 56.0,
                     conv_w: Optional[torch.Tensor] = None, conv_b: Optional[torch.Tensor] = None):
    """count_spp in one launch (fgn_support_prologue_fwd; fgn_roi_head.py:419-449): mask pooling, support RoIAlign with
    level assignment, class mean, masked GAP and -- when the relation conv's ``conv_w`` [C,2C] / ``conv_b`` are given -- the
    class half of that convolution.  ``spp_feats``: levels [B*N*K,C,H_l,W_l] (repacked to channels_last if they are not),
    ``spp_bboxes`` [B*N*K,4] XYXY px, ``spp_isegmaps`` [B*N*K,(1,)S,S] bool.
    -> cat_mean [B,N,C,P,P] (channels_last storage), masked_gap [B,N,C,1,1], class_term [B*N*P*P, C] or None."""
    _need_cuda(spp_bboxes, spp_isegmaps, conv_w, conv_b, *spp_feats)
    feats = [to_nhwc(_f32(f, "spp_feats")) if storage_layout(f) != LAYOUT_NHWC else _f32(f, "spp_feats") for f in spp_feats]
    pyr, keep, lay, m, c = _make_pyramid(feats, scales, torch.float32)
    if m % (n_ways * k_shots):
        raise FgnError(f"{m} support maps is not a multiple of N*K={n_ways * k_shots}")
    bn, p = m // k_shots, int(output_size)
    boxes = _f32(spp_bboxes, "spp_bboxes").reshape(m, 4).contiguous()
    masks = spp_isegmaps.reshape(m, spp_isegmaps.shape[-2], spp_isegmaps.shape[-1])
    if masks.dtype == torch.bool:
        masks = masks.contiguous().view(torch.uint8)
    elif masks.dtype != torch.uint8:
        raise FgnError(f"spp_isegmaps: expected bool or uint8, got {masks.dtype}")
    masks = masks.contiguous()
    dev = boxes.device
    cat = _empty_like_format((bn, c, p, p), dev, LAYOUT_NHWC)
    gap = torch.empty((bn, c), device=dev, dtype=torch.float32)
    term = None
    if conv_w is not None:
        cw = _f32(conv_w, "conv_w").reshape(conv_w.shape[0], -1).contiguous()
        if tuple(cw.shape) != (c, 2 * c) or conv_b is None:
            raise FgnError(f"support_prologue: conv_w must be [C,2C] = [{c},{2 * c}] with its bias")
        term = torch.empty((bn * p * p, c), device=dev, dtype=torch.float32)
    lib = _lib.load()
    wsb = int(lib.fgn_support_prologue_workspace_bytes(bn, c, p))
    ws = torch.empty((max(wsb, 1),), device=dev, dtype=torch.uint8)
    _lib.check(lib.fgn_support_prologue_fwd(ctypes.byref(pyr), c, boxes.data_ptr(), masks.data_ptr(), masks.shape[1], masks.shape[2],
                                            bn, int(k_shots), p, float(finest_scale), _ptr(cw if conv_w is not None else None),
                                            _ptr(None if conv_b is None else _f32(conv_b, "conv_b").contiguous()),
                                            cat.data_ptr(), gap.data_ptr(), _ptr(term), ws.data_ptr(), wsb, _stream()),
               "fgn_support_prologue_fwd")
    b = bn // n_ways
    return cat.unflatten(0, (b, n_ways)), gap.view(b, n_ways, c, 1, 1), term


def attention_vectors(spp_fmaps: torch.Tensor, n_ways: int, k_shots: int) -> torch.Tensor:
    """fgn_ag_rpn_head.py:37-41 -> [B,N,C,1,1]."""
    _need_cuda(spp_fmaps)
    x, lay = _dense(_f32(spp_fmaps, "spp_fmaps"))
    bnk, c, h, w = x.shape
    if bnk % (n_ways * k_shots):
        raise FgnError(f"{bnk} support maps is not a multiple of N*K={n_ways * k_shots}")
    bn = bnk // k_shots
    if lay == LAYOUT_NHWC and (c % 4 or c > 1024):
        x, lay = to_nchw(x), LAYOUT_NCHW
    lib = _lib.load()
    ws_bytes = lib.fgn_attention_vectors_workspace_bytes(bn, k_shots, c, h, w, lay)
    ws = torch.empty((max(ws_bytes, 1),), device=x.device, dtype=torch.uint8)
    vec = torch.empty((bn, c), device=x.device, dtype=torch.float32)
    _lib.check(lib.fgn_attention_vectors(x.data_ptr(), lay, bn, int(k_shots), c, h, w, vec.data_ptr(),
                                         ws.data_ptr(), ws_bytes, _stream()), "fgn_attention_vectors")
    return vec.view(bn // n_ways, n_ways, c, 1, 1)


def channel_attention(qry: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """fgn_ag_rpn_head.py:44-46: [B,C,H,W] x [B,N,C,1,1] -> [B*N,C,H,W] (same storage format as qry)."""
    _need_cuda(qry, vec)
    q, lay = _dense(_f32(qry, "qry_fmap"))
    b, c, h, w = q.shape
    n = vec.shape[1]
    v = _f32(vec, "vec").reshape(b * n, c).contiguous()
    if lay == LAYOUT_NHWC and c % 4:
        q, lay = to_nchw(q), LAYOUT_NCHW
    out = _empty_like_format((b * n, c, h, w), q.device, lay)
    _lib.check(_lib.load().fgn_channel_attention(q.data_ptr(), v.data_ptr(), b, n, c, h, w, lay, out.data_ptr(),
                                                 _stream()), "fgn_channel_attention")
    return out


def fold_attention_weights(weight: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    """Channel attention folded into the consumer's weights: ``conv(qry * vec[bn]) == conv(qry, w'[bn])`` with
    ``w'[bn,o,c,ky,kx] = weight[o,c,ky,kx] * vec[bn,c]``.  ``weight`` [Co,Ci,kh,kw]; ``vec`` [...,Ci] or
    [B,N,Ci,1,1] (any leading shape).  Returns [BN,Co,Ci,kh,kw]."""
    _need_cuda(weight, vec)
    w = _f32(weight, "weight").contiguous()
    co, ci, kh, kw = w.shape
    v = _f32(vec, "vec").reshape(-1, ci).contiguous()
    out = torch.empty((v.shape[0], co, ci, kh, kw), device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().fgn_fold_attention_weights(w.data_ptr(), v.data_ptr(), v.shape[0], co, ci, kh * kw,
                                                      out.data_ptr(), _stream()), "fgn_fold_attention_weights")
    return out


def attention_vectors_multilevel(spp_feats: Sequence[torch.Tensor], n_ways: int, k_shots: int) -> torch.Tensor:
    """The class vectors of every level in two launches (channels_last fp32 support maps): [L,B*N,C]."""
    ss = list(spp_feats)
    _need_cuda(*ss)
    c = ss[0].shape[1]
    if not (c % 4 == 0 and c <= 1024 and all(storage_layout(t) == LAYOUT_NHWC and t.dtype == torch.float32 for t in ss)):
        return torch.stack([attention_vectors(s, n_ways, k_shots).reshape(-1, c) for s in ss])
    bn = ss[0].shape[0] // k_shots
    spyr = Pyramid()
    spyr.num_levels = len(ss)
    for i, s in enumerate(ss):
        if s.shape[0] != bn * k_shots or s.shape[1] != c:
            raise FgnError("attention_vectors_multilevel: level shapes must be [B*N*K,C,h,w]")
        spyr.feat[i], spyr.H[i], spyr.W[i] = s.data_ptr(), s.shape[2], s.shape[3]
    lib = _lib.load()
    vec = torch.empty((len(ss), bn, c), device=ss[0].device, dtype=torch.float32)
    wsb = lib.fgn_attention_vectors_ml_workspace_bytes(ctypes.byref(spyr), bn, k_shots, c)
    ws = torch.empty((max(wsb, 1),), device=vec.device, dtype=torch.uint8)
    _lib.check(lib.fgn_attention_vectors_ml(ctypes.byref(spyr), bn, int(k_shots), c, vec.data_ptr(), ws.data_ptr(), wsb,
                                            _stream()), "fgn_attention_vectors_ml")
    return vec


def attention_multilevel(qry_feats: Sequence[torch.Tensor], spp_feats: Sequence[torch.Tensor], n_ways: int,
                         k_shots: int):
    """AG-RPN attention for every level of a pyramid in three launches (channels_last inputs).
    Returns (vecs: list of [B,N,C,1,1], mods: list of [B*N,C,H_l,W_l]).  Falls back to the per-level
    entry points for other storage layouts."""
    qs, ss = list(qry_feats), list(spp_feats)
    _need_cuda(*qs, *ss)
    c = qs[0].shape[1]
    bf16 = _is_bf16(qs + ss)
    if bf16:
        if c % 8 or c > 1024 or not all(storage_layout(t) == LAYOUT_NHWC for t in qs + ss):
            raise FgnError("bf16 attention needs channels_last maps with C%8==0 and C<=1024")
        fast = True
    else:
        fast = c % 4 == 0 and c <= 1024 and all(storage_layout(t) == LAYOUT_NHWC and t.dtype == torch.float32 for t in qs + ss)
    if not fast:
        vecs = [attention_vectors(s, n_ways, k_shots) for s in ss]
        return vecs, [channel_attention(q, v) for q, v in zip(qs, vecs)]
    b = qs[0].shape[0]
    bn = b * n_ways
    L = len(qs)
    spyr, qpyr = Pyramid(), Pyramid()
    spyr.num_levels = qpyr.num_levels = L
    outs = []
    for i, (q, s) in enumerate(zip(qs, ss)):
        if s.shape[0] != bn * k_shots or s.shape[1] != c or q.shape[1] != c or q.shape[0] != b:
            raise FgnError("attention_multilevel: level shapes must be [B,C,H,W] / [B*N*K,C,h,w]")
        spyr.feat[i], spyr.H[i], spyr.W[i] = s.data_ptr(), s.shape[2], s.shape[3]
        qpyr.feat[i], qpyr.H[i], qpyr.W[i] = q.data_ptr(), q.shape[2], q.shape[3]
        if bf16:
            outs.append(torch.empty((bn, q.shape[2], q.shape[3], c), device=q.device, dtype=torch.bfloat16).permute(0, 3, 1, 2))
        else:
            outs.append(_empty_like_format((bn, c, q.shape[2], q.shape[3]), q.device, LAYOUT_NHWC))
    lib = _lib.load()
    dev = qs[0].device
    vec = torch.empty((L, bn, c), device=dev, dtype=torch.float32)
    wsb = lib.fgn_attention_vectors_ml_workspace_bytes(ctypes.byref(spyr), bn, k_shots, c)
    ws = torch.empty((max(wsb, 1),), device=dev, dtype=torch.uint8)
    fvec = lib.fgn_attention_vectors_ml_bf16 if bf16 else lib.fgn_attention_vectors_ml
    fmul = lib.fgn_channel_attention_ml_bf16 if bf16 else lib.fgn_channel_attention_ml
    _lib.check(fvec(ctypes.byref(spyr), bn, int(k_shots), c, vec.data_ptr(), ws.data_ptr(), wsb, _stream()),
               "fgn_attention_vectors_ml")
    ptrs = (ctypes.c_void_p * L)(*[o.data_ptr() for o in outs])
    _lib.check(fmul(ctypes.byref(qpyr), vec.data_ptr(), b, n_ways, c, ptrs, _stream()), "fgn_channel_attention_ml")
    return [vec[i].view(b, n_ways, c, 1, 1) for i in range(L)], outs


def best_class_select(cls: torch.Tensor, reg: torch.Tensor, batch: int, n_ways: int):
    """fgn_ag_rpn_head.py:87-108 for sigmoid objectness (one score per anchor)."""
    _need_cuda(cls, reg)
    cls = to_nchw(_f32(cls, "rpn_cls_score"))
    reg = to_nchw(_f32(reg, "rpn_bbox_pred"))
    bn, a, h, w = cls.shape
    if bn != batch * n_ways or reg.shape != (bn, 4 * a, h, w):
        raise FgnError("best_class_select: cls [B*N,A,H,W] / reg [B*N,4A,H,W] expected")
    cls_out = torch.empty((batch, a, h, w), device=cls.device, dtype=torch.float32)
    reg_out = torch.empty((batch, 4 * a, h, w), device=cls.device, dtype=torch.float32)
    _lib.check(_lib.load().fgn_best_class_select(cls.data_ptr(), reg.data_ptr(), batch, n_ways, a, h, w,
                                                 cls_out.data_ptr(), reg_out.data_ptr(), _stream()),
               "fgn_best_class_select")
    return cls_out, reg_out


# --------------------------------------------------------------------------------------------
class RelationParams:
    """Device-resident, contiguous copies of the relation head's parameters."""

    def __init__(self, conv_w, conv_b, gn_w, gn_b, fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b,
                 gn_groups: int = 32, gn_eps: float = 1e-5):
        c = conv_w.shape[0]
        self.C = c
        self.conv_w = conv_w.detach().reshape(c, 2 * c).contiguous().float()
        self.conv_b = (conv_b.detach() if conv_b is not None else torch.zeros(c, device=conv_w.device)).contiguous().float()
        self.gn_w, self.gn_b = gn_w.detach().contiguous().float(), gn_b.detach().contiguous().float()
        self.fc_cls_w, self.fc_cls_b = fc_cls_w.detach().contiguous().float(), fc_cls_b.detach().contiguous().float()
        self.fc_reg_w, self.fc_reg_b = fc_reg_w.detach().contiguous().float(), fc_reg_b.detach().contiguous().float()
        if self.fc_cls_w.shape != (2, c) or self.fc_reg_w.shape != (4, c):
            raise FgnError("relation head expects fc_cls [2,C] (num_classes=1) and fc_reg [4,C]")
        self.gn_groups, self.gn_eps = int(gn_groups), float(gn_eps)
        _need_cuda(self.conv_w, self.gn_w, self.fc_cls_w, self.fc_reg_w)
        # load-time TF32 split of the conv weights (fgn_relation_split_weights): two launches less per call
        lib = _lib.load()
        self.conv_w_split = torch.empty((max(int(lib.fgn_relation_split_weights_bytes(c)) // 4, 1),),
                                        device=self.conv_w.device, dtype=torch.float32)
        _lib.check(lib.fgn_relation_split_weights(self.conv_w.data_ptr(), c, self.conv_w_split.data_ptr(), _stream()),
                   "fgn_relation_split_weights")


def relation_fusion(roi_feat: torch.Tensor, roi_batch: torch.Tensor, spp_cat_mean: torch.Tensor, n_ways: int,
                    params: RelationParams, precision: str = "fp32", return_raw: bool = False,
                    class_term: Optional[torch.Tensor] = None):
    """count_one_roi_by_n_spp + bbox_head.forward + count_modified_cls_bbox (fgn_roi_head.py:336-339).

    roi_feat [R,C,P,P]; roi_batch [R] (rois[:,0]); spp_cat_mean [B,N,C,P,P] -> cls [R,N+1], reg [R,4N].
    """
    _need_cuda(roi_feat, roi_batch, spp_cat_mean)
    x, lay = _dense(_f32(roi_feat, "roi_feat"))
    r, c, p, _ = x.shape
    s = spp_cat_mean.reshape(-1, c, p, p)
    s, slay = _dense(_f32(s, "spp_cat_mean"))
    if slay != lay:
        s = to_nhwc(s) if lay == LAYOUT_NHWC else to_nchw(s)
    bn = s.shape[0]
    b = bn // n_ways
    rb = roi_batch.to(torch.int32).contiguous()
    dev = x.device
    cls = torch.empty((r, n_ways + 1), device=dev, dtype=torch.float32)
    reg = torch.empty((r, 4 * n_ways), device=dev, dtype=torch.float32)
    raw_c = torch.empty((r * n_ways, 2), device=dev, dtype=torch.float32) if return_raw else None
    raw_r = torch.empty((r * n_ways, 4), device=dev, dtype=torch.float32) if return_raw else None
    lib = _lib.load()
    wsb = lib.fgn_relation_fusion_workspace_bytes(r, bn, c, p)
    ws = torch.empty((max(wsb, 1),), device=dev, dtype=torch.uint8)
    pr = params
    _lib.check(lib.fgn_relation_fusion_fwd(
        x.data_ptr(), lay, rb.data_ptr(), s.data_ptr(), _ptr(class_term), r, b, n_ways, c, p,
        pr.conv_w.data_ptr(), pr.conv_w_split.data_ptr(), pr.conv_b.data_ptr(), pr.gn_w.data_ptr(), pr.gn_b.data_ptr(),
        pr.gn_groups, pr.gn_eps,
        pr.fc_cls_w.data_ptr(), pr.fc_cls_b.data_ptr(), pr.fc_reg_w.data_ptr(), pr.fc_reg_b.data_ptr(),
        cls.data_ptr(), reg.data_ptr(), _ptr(raw_c), _ptr(raw_r), {"fp32": 0, "tf32": 1}[precision],
        ws.data_ptr(), wsb, _stream()), "fgn_relation_fusion_fwd")
    return (cls, reg, raw_c, raw_r) if return_raw else (cls, reg)


def guided_roi_fused(feats: Sequence[torch.Tensor], rois: torch.Tensor, scales: Sequence[float],
                     spp_cat_mean: torch.Tensor, n_ways: int, params: RelationParams, output_size: int = 7,
                     sampling_ratio: int = 0, aligned: bool = True, finest_scale: float = 56.0,
                     precision: str = "fp32", return_levels: bool = False, class_term: Optional[torch.Tensor] = None):
    """FPN-mode single call: level assignment + RoIAlign + relation fusion + heads.  ``class_term``: the class half of the
    relation convolution as ``support_prologue`` leaves it ([B*N*P*P, C]); saves the class-term launch."""
    _need_cuda(rois, spp_cat_mean)
    rois = _f32(rois, "rois").contiguous()
    bf16 = _is_bf16(list(feats))
    if not bf16:
        feats = [to_nhwc(f) if storage_layout(f) != LAYOUT_NHWC else f for f in feats]
    pyr, keep, lay, b, c = _make_pyramid(feats, scales, torch.bfloat16 if bf16 else torch.float32)
    p = int(output_size)
    s = to_nhwc(_f32(spp_cat_mean.reshape(-1, c, p, p), "spp_cat_mean"))
    bn = s.shape[0]
    if bn != b * n_ways:
        raise FgnError(f"spp_cat_mean has {bn} class maps, expected B*N={b * n_ways}")
    r = rois.shape[0]
    dev = rois.device
    cls = torch.empty((r, n_ways + 1), device=dev, dtype=torch.float32)
    reg = torch.empty((r, 4 * n_ways), device=dev, dtype=torch.float32)
    lvl = torch.empty((r,), device=dev, dtype=torch.int32) if return_levels else None
    lib = _lib.load()
    pr = params
    if bf16:
        wsb = lib.fgn_guided_roi_fused_bf16_workspace_bytes(r, bn, c, p)
        ws = torch.empty((max(wsb, 1),), device=dev, dtype=torch.uint8)
        _lib.check(lib.fgn_guided_roi_fused_fwd_bf16(
            ctypes.byref(pyr), b, c, rois.data_ptr(), r, p, int(sampling_ratio), int(bool(aligned)), float(finest_scale),
            s.data_ptr(), n_ways, pr.conv_w.data_ptr(), pr.conv_b.data_ptr(), pr.gn_w.data_ptr(), pr.gn_b.data_ptr(),
            pr.gn_groups, pr.gn_eps, pr.fc_cls_w.data_ptr(), pr.fc_cls_b.data_ptr(), pr.fc_reg_w.data_ptr(),
            pr.fc_reg_b.data_ptr(), cls.data_ptr(), reg.data_ptr(), _ptr(lvl), ws.data_ptr(), wsb, _stream()),
            "fgn_guided_roi_fused_fwd_bf16")
        return (cls, reg, lvl.long()) if return_levels else (cls, reg)
    wsb = lib.fgn_guided_roi_fused_workspace_bytes(r, bn, c, p)
    ws = torch.empty((max(wsb, 1),), device=dev, dtype=torch.uint8)
    _lib.check(lib.fgn_guided_roi_fused_fwd(
        ctypes.byref(pyr), b, c, rois.data_ptr(), r, p, int(sampling_ratio), int(bool(aligned)), float(finest_scale),
        s.data_ptr(), _ptr(class_term), n_ways, pr.conv_w.data_ptr(), pr.conv_w_split.data_ptr(), pr.conv_b.data_ptr(),
        pr.gn_w.data_ptr(), pr.gn_b.data_ptr(),
        pr.gn_groups, pr.gn_eps, pr.fc_cls_w.data_ptr(), pr.fc_cls_b.data_ptr(), pr.fc_reg_w.data_ptr(),
        pr.fc_reg_b.data_ptr(), cls.data_ptr(), reg.data_ptr(), _ptr(lvl), {"fp32": 0, "tf32": 1}[precision],
        ws.data_ptr(), wsb, _stream()), "fgn_guided_roi_fused_fwd")
    return (cls, reg, lvl.long()) if return_levels else (cls, reg)


def gemm_nt(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, precision: str = "fp32",
            use_workspace: bool = True, b_split: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] = A[M,K] @ B[N,K]^T (+ bias): the relation head's contraction alone (rows may be strided).
    ``b_split`` = ``conv_split_weights(b.contiguous()[None])`` (fp32 precision only): the weights' TF32 split made once."""
    _need_cuda(a, b, bias, b_split)
    if a.stride(1) != 1 or b.stride(1) != 1:
        raise FgnError("gemm_nt operands must be K-major (unit stride along K)")
    m, k = a.shape
    n = b.shape[0]
    c = torch.empty((m, n), device=a.device, dtype=torch.float32)
    lib = _lib.load()
    if b_split is not None:
        if precision != "fp32" or b_split.numel() != 2 * n * k or b_split.dtype != torch.float32:
            raise FgnError("gemm_nt: b_split is the fp32 route's [2,N,K] split of b")
        _lib.check(lib.fgn_gemm_nt_presplit(_f32(a, "a").data_ptr(), a.stride(0), b_split.contiguous().data_ptr(), _ptr(bias),
                                            c.data_ptr(), n, m, n, k, _stream()), "fgn_gemm_nt_presplit")
        return c
    if a.dtype == torch.bfloat16 or b.dtype == torch.bfloat16:
        if not (a.dtype == b.dtype == torch.bfloat16):
            raise FgnError("bf16 contraction needs both operands in bfloat16")
        _lib.check(lib.fgn_gemm_nt_bf16(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _ptr(bias), c.data_ptr(), n,
                                        m, n, k, _stream()), "fgn_gemm_nt_bf16")
        return c
    wsb = lib.fgn_gemm_workspace_bytes(n, k) if use_workspace else 0
    ws = torch.empty((max(wsb, 1),), device=a.device, dtype=torch.uint8)
    _lib.check(lib.fgn_gemm_nt(_f32(a, "a").data_ptr(), a.stride(0), _f32(b, "b").data_ptr(), b.stride(0), _ptr(bias),
                               c.data_ptr(), n, m, n, k, {"fp32": 0, "tf32": 1}[precision], ws.data_ptr(), wsb,
                               _stream()), "fgn_gemm_nt")
    return c


def cls_bbox_reassemble(raw_cls: torch.Tensor, raw_reg: torch.Tensor, rois_amount: int, n_ways: int):
    """count_modified_cls_bbox (fgn_roi_head.py:302-326), generalised to any N."""
    _need_cuda(raw_cls, raw_reg)
    raw_cls = _f32(raw_cls, "cls_score").contiguous()
    raw_reg = _f32(raw_reg, "bbox_pred").contiguous()
    dev = raw_cls.device
    cls = torch.empty((rois_amount, n_ways + 1), device=dev, dtype=torch.float32)
    reg = torch.empty((rois_amount, 4 * n_ways), device=dev, dtype=torch.float32)
    _lib.check(_lib.load().fgn_cls_bbox_reassemble(raw_cls.data_ptr(), raw_reg.data_ptr(), rois_amount, n_ways,
                                                   cls.data_ptr(), reg.data_ptr(), _stream()),
               "fgn_cls_bbox_reassemble")
    return cls, reg


def det_postprocess(rois: torch.Tensor, cls_score: torch.Tensor, bbox_pred: torch.Tensor, num_per_img: Sequence[int],
                    img_shapes: Optional[Sequence[Sequence[float]]] = None,
                    scale_factors: Optional[Sequence[Sequence[float]]] = None,
                    score_thr: float = 0.05, iou_thr: float = 0.5, max_per_img: int = 100,
                    means: Sequence[float] = (0., 0., 0., 0.), stds: Sequence[float] = (0.1, 0.1, 0.2, 0.2),
                    wh_ratio_clip: float = 16 / 1000):
    """BBoxHead.get_bboxes [3P] per image (fgn_roi_head.py:606-613): softmax, delta decode, clip / rescale,
    multiclass NMS, top ``max_per_img``.  ``rois`` [R,5] grouped by image, ``num_per_img`` the per-image counts
    (host ints, as in the reference's ``rois.split``), ``img_shapes`` [(h, w, ...)] or None, ``scale_factors``
    [4 floats per image] or None (= rescale False).
    Returns ``(det [B,max_per_img,5], labels [B,max_per_img] int32, counts [B] int32)`` on the device; rows past
    ``counts[b]`` are unspecified (see FGNRoIHead.simple_test_bboxes for the list-of-tensors form)."""
    _need_cuda(rois, cls_score, bbox_pred)
    rois = _f32(rois, "rois").contiguous()
    cls_score = _f32(cls_score, "cls_score").contiguous()
    bbox_pred = _f32(bbox_pred, "bbox_pred").contiguous()
    r, b = rois.shape[0], len(num_per_img)
    if b < 1 or sum(int(x) for x in num_per_img) != r:
        raise FgnError("num_per_img must sum to the number of rois")
    n = cls_score.shape[1] - 1
    if cls_score.shape[0] != r or n < 1 or tuple(bbox_pred.shape) != (r, 4 * n):
        raise FgnError(f"cls_score {tuple(cls_score.shape)} / bbox_pred {tuple(bbox_pred.shape)} do not match R={r}")
    dev = rois.device
    det = torch.empty((b, max_per_img, 5), device=dev, dtype=torch.float32)
    lab = torch.empty((b, max_per_img), device=dev, dtype=torch.int32)
    cnt = torch.empty((b,), device=dev, dtype=torch.int32)
    offs, acc = [0], 0
    for x in num_per_img:
        acc += int(x)
        offs.append(acc)
    rmax = max(int(x) for x in num_per_img)
    off_t = _const_tensor(offs, torch.int32, dev)           # persistent device constants (see _const_tensor)
    hw_t = None
    if img_shapes is not None:
        hw_t = _const_tensor([[float(s[0]), float(s[1])] for s in img_shapes], torch.float32, dev)
    sf_t = None
    if scale_factors is not None:
        sf_t = _const_tensor([[float(v) for v in s][:4] for s in scale_factors], torch.float32, dev).reshape(b, 4)
    lib = _lib.load()
    nbytes = int(lib.fgn_det_postprocess_workspace_bytes(r, n, b, rmax))
    ws = torch.empty((max(nbytes, 256),), device=dev, dtype=torch.uint8)
    m4 = (ctypes.c_float * 4)(*[float(v) for v in means])
    s4 = (ctypes.c_float * 4)(*[float(v) for v in stds])
    _lib.check(lib.fgn_det_postprocess(_ptr(rois), _ptr(cls_score), _ptr(bbox_pred), _ptr(off_t), r, n, b, rmax,
                                       _ptr(hw_t), _ptr(sf_t), m4, s4, float(wh_ratio_clip), float(score_thr),
                                       float(iou_thr), int(max_per_img), det.data_ptr(), lab.data_ptr(), cnt.data_ptr(),
                                       ws.data_ptr(), nbytes, _stream()), "fgn_det_postprocess")
    return det, lab, cnt


def base_anchors(base_size: float, scales: Sequence[float], ratios: Sequence[float]) -> torch.Tensor:
    """mmdet AnchorGenerator.gen_single_level_base_anchors [3P] (scale_major=True, center_offset=0), evaluated
    with the same fp32 torch expressions: [len(ratios)*len(scales), 4], ratio-major."""
    sc = torch.tensor(list(scales), dtype=torch.float32)
    ra = torch.tensor(list(ratios), dtype=torch.float32)
    w = h = float(base_size)
    h_ratios = torch.sqrt(ra)
    w_ratios = 1 / h_ratios
    ws = (w * w_ratios[:, None] * sc[None, :]).view(-1)
    hs = (h * h_ratios[:, None] * sc[None, :]).view(-1)
    xc = yc = 0.0
    return torch.stack([xc - 0.5 * ws, yc - 0.5 * hs, xc + 0.5 * ws, yc + 0.5 * hs], dim=-1)


def rpn_proposals(cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor], strides: Sequence[int],
                  anchors: torch.Tensor, img_shapes: Optional[Sequence[Sequence[float]]] = None, nms_pre: int = 6000,
                  iou_thr: float = 0.7, max_per_img: int = 300, min_bbox_size: float = 0.0,
                  means: Sequence[float] = (0., 0., 0., 0.), stds: Sequence[float] = (1., 1., 1., 1.),
                  wh_ratio_clip: float = 16 / 1000):
    """RPNHead.get_bboxes [3P] (fgn.py:229-235): ``cls_scores[l]`` [B,A,H,W] sigmoid logits, ``bbox_preds[l]``
    [B,4A,H,W], ``anchors`` [L,A,4] base anchors (ops.base_anchors per level, stacked).
    Returns ``(proposals [B,max_per_img,5], levels [B,max_per_img] int32, counts [B] int32)`` on the device."""
    _need_cuda(*cls_scores, *bbox_preds)
    nl = len(cls_scores)
    cls_scores = [_f32(t, "cls_score").contiguous() for t in cls_scores]
    bbox_preds = [_f32(t, "bbox_pred").contiguous() for t in bbox_preds]
    b, a = cls_scores[0].shape[0], cls_scores[0].shape[1]
    for c, r in zip(cls_scores, bbox_preds):
        if c.dim() != 4 or r.shape != (b, 4 * a, c.shape[2], c.shape[3]) or c.shape[:2] != (b, a):
            raise FgnError(f"rpn_proposals: cls {tuple(c.shape)} / reg {tuple(r.shape)} do not match B={b} A={a}")
    dev = cls_scores[0].device
    anchors = _f32(anchors, "anchors").reshape(nl, a, 4).contiguous()
    if anchors.device != dev:
        anchors = _const_tensor(anchors, torch.float32, dev)
    hs = (ctypes.c_int * nl)(*[int(c.shape[2]) for c in cls_scores])
    wsz = (ctypes.c_int * nl)(*[int(c.shape[3]) for c in cls_scores])
    st = (ctypes.c_int * nl)(*[int(s) for s in strides])
    cp = (ctypes.c_void_p * nl)(*[c.data_ptr() for c in cls_scores])
    rp = (ctypes.c_void_p * nl)(*[r.data_ptr() for r in bbox_preds])
    hw_t = None
    if img_shapes is not None:
        hw_t = _const_tensor([[float(s[0]), float(s[1])] for s in img_shapes], torch.float32, dev)
    prop = torch.empty((b, max_per_img, 5), device=dev, dtype=torch.float32)
    lvl = torch.empty((b, max_per_img), device=dev, dtype=torch.int32)
    cnt = torch.empty((b,), device=dev, dtype=torch.int32)
    lib = _lib.load()
    nbytes = int(lib.fgn_rpn_proposals_workspace_bytes(hs, wsz, nl, a, b, int(nms_pre)))
    ws = torch.empty((max(nbytes, 256),), device=dev, dtype=torch.uint8)
    m4 = (ctypes.c_float * 4)(*[float(v) for v in means])
    s4 = (ctypes.c_float * 4)(*[float(v) for v in stds])
    _lib.check(lib.fgn_rpn_proposals(cp, rp, hs, wsz, st, nl, a, b, anchors.data_ptr(), _ptr(hw_t), m4, s4,
                                     float(wh_ratio_clip), int(nms_pre), float(iou_thr), int(max_per_img),
                                     float(min_bbox_size), prop.data_ptr(), lvl.data_ptr(), cnt.data_ptr(),
                                     ws.data_ptr(), nbytes, _stream()), "fgn_rpn_proposals")
    return prop, lvl, cnt


def mask_paste(mask_pred: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int,
               mask_thr_binary: float = 0.5) -> torch.Tensor:
    """FCNMaskHead.get_seg_masks' tensor [3P] for one image (fgn_roi_head.py:668-671): sigmoid, paste every
    [M,M] mask over the whole ``img_h x img_w`` image (F.grid_sample semantics of _do_paste_mask), ``>= thr``.
    ``mask_pred`` [D,1,M,M] or [D,M,M] logits; ``boxes`` [D,4+] (x1,y1,x2,y2,...).  Returns bool [D,img_h,img_w]."""
    _need_cuda(mask_pred, boxes)
    mask_pred = _f32(mask_pred, "mask_pred").contiguous()
    boxes = _f32(boxes, "boxes").contiguous()
    d, m = mask_pred.shape[0], mask_pred.shape[-1]
    if mask_pred.numel() != d * m * m or boxes.dim() != 2 or boxes.shape[0] != d or boxes.shape[1] < 4:
        raise FgnError(f"mask_pred {tuple(mask_pred.shape)} / boxes {tuple(boxes.shape)}: expected [D,(1,)M,M] and [D,>=4]")
    out = torch.empty((d, int(img_h), int(img_w)), device=mask_pred.device, dtype=torch.uint8)
    _lib.check(_lib.load().fgn_mask_paste(_ptr(mask_pred), _ptr(boxes), int(boxes.shape[1]), d, m, int(img_h), int(img_w),
                                          float(mask_thr_binary), _ptr(out), _stream()), "fgn_mask_paste")
    return out.view(torch.bool)


def mask_paste_rle(mask_pred: torch.Tensor, boxes: torch.Tensor, img_hw: Sequence[Sequence[int]],
                   det_img: Optional[torch.Tensor] = None, mask_thr_binary: float = 0.5, cap: int = 4096,
                   return_counts: bool = False):
    """get_seg_masks + encode_mask_results in one kernel (fgn_roi_head.py:668-671, fgn.py:281): the COCO RLE of
    every pasted mask without materialising the masks.  ``img_hw``: (h, w) per image; ``det_img`` [D] image index
    of every detection (None: all image 0).  Returns a list of ``dict(size=[h, w], counts=bytes)`` (the
    pycocotools form encode_mask_results returns); with ``return_counts`` also the uncompressed run lengths as a
    list of int lists.  One device->host copy of the encoded bytes; ``cap`` (runs per mask) grows on overflow."""
    _need_cuda(mask_pred, boxes, det_img)
    mask_pred = _f32(mask_pred, "mask_pred").contiguous()
    boxes = _f32(boxes, "boxes").contiguous()
    d, m = mask_pred.shape[0], mask_pred.shape[-1]
    if mask_pred.numel() != d * m * m or boxes.dim() != 2 or boxes.shape[0] != d or boxes.shape[1] < 4:
        raise FgnError(f"mask_pred {tuple(mask_pred.shape)} / boxes {tuple(boxes.shape)}: expected [D,(1,)M,M] and [D,>=4]")
    hw = [[int(s[0]), int(s[1])] for s in img_hw]
    if d == 0:
        return ([], []) if return_counts else []
    dev = mask_pred.device
    hw_t = _const_tensor(hw, torch.int32, dev)
    if det_img is not None:
        det_img = det_img.to(torch.int32).contiguous()
        img_of = det_img.tolist()
    else:
        img_of = [0] * d
    lib = _lib.load()
    hmax, wmax = max(1, max(s[0] for s in hw)), max(1, max(s[1] for s in hw))
    while True:
        cap_bytes = 2 * cap                           # a run costs 1-2 characters in practice, 7 at most
        counts = torch.empty((d, cap), device=dev, dtype=torch.int32)
        ncounts = torch.empty((d,), device=dev, dtype=torch.int32)
        sbuf = torch.empty((d, cap_bytes), device=dev, dtype=torch.uint8)
        slen = torch.empty((d,), device=dev, dtype=torch.int32)
        wsb = int(lib.fgn_mask_paste_rle_workspace_bytes(d, cap, hmax, wmax))
        ws = torch.empty((max(wsb, 4),), device=dev, dtype=torch.uint8)
        _lib.check(lib.fgn_mask_paste_rle(_ptr(mask_pred), _ptr(boxes), int(boxes.shape[1]), _ptr(det_img), _ptr(hw_t),
                                          d, m, float(mask_thr_binary), hmax, wmax, _ptr(counts), _ptr(ncounts),
                                          _ptr(sbuf), _ptr(slen), cap, cap_bytes, ws.data_ptr(), wsb, _stream()),
                   "fgn_mask_paste_rle")
        nc, sl = ncounts.tolist(), slen.tolist()
        need = max([-v for v in nc] + [(-v + 1) // 2 for v in sl] + [0])
        if need == 0:
            break
        cap = max(2 * cap, need + 16)
    smax = max(sl)
    sbytes = sbuf[:, :max(smax, 1)].cpu().numpy()
    rles = [dict(size=list(hw[img_of[i]]), counts=sbytes[i, :sl[i]].tobytes()) for i in range(d)]
    if not return_counts:
        return rles
    cmax = max(nc)
    ch = counts[:, :max(cmax, 1)].cpu().numpy()
    return rles, [ch[i, :nc[i]].tolist() for i in range(d)]


def mask_rle_encode(masks: torch.Tensor, cap: int = 4096, return_counts: bool = False):
    """mmdet.core.encode_mask_results / pycocotools mask.encode [3P] of given masks (fgn.py:296-298, the
    ``qry_isegmaps`` of the result dict): ``masks`` [D,H,W] bool / uint8 on the device -> list of
    ``dict(size=[H, W], counts=bytes)``; with ``return_counts`` also the uncompressed run lengths."""
    _need_cuda(masks)
    if masks.dim() != 3:
        raise FgnError(f"masks {tuple(masks.shape)}: expected [D,H,W]")
    m8 = masks.contiguous().view(torch.uint8) if masks.dtype == torch.bool else masks.to(torch.uint8).contiguous()
    d, h, w = m8.shape
    if d == 0:
        return ([], []) if return_counts else []
    dev = m8.device
    lib = _lib.load()
    while True:
        cap_bytes = 2 * cap
        counts = torch.empty((d, cap), device=dev, dtype=torch.int32)
        ncounts = torch.empty((d,), device=dev, dtype=torch.int32)
        sbuf = torch.empty((d, cap_bytes), device=dev, dtype=torch.uint8)
        slen = torch.empty((d,), device=dev, dtype=torch.int32)
        wsb = int(lib.fgn_mask_paste_rle_workspace_bytes(d, cap, max(h, 1), max(w, 1)))
        ws = torch.empty((max(wsb, 4),), device=dev, dtype=torch.uint8)
        _lib.check(lib.fgn_mask_rle_encode(_ptr(m8), d, h, w, _ptr(counts), _ptr(ncounts), _ptr(sbuf), _ptr(slen), cap, cap_bytes,
                                           ws.data_ptr(), wsb, _stream()), "fgn_mask_rle_encode")
        nc, sl = ncounts.tolist(), slen.tolist()
        need = max([-v for v in nc] + [(-v + 1) // 2 for v in sl] + [0])
        if need == 0:
            break
        cap = max(2 * cap, need + 16)
    sbytes = sbuf[:, :max(max(sl), 1)].cpu().numpy()
    rles = [dict(size=[h, w], counts=sbytes[i, :sl[i]].tobytes()) for i in range(d)]
    if not return_counts:
        return rles
    ch = counts[:, :max(max(nc), 1)].cpu().numpy()
    return rles, [ch[i, :nc[i]].tolist() for i in range(d)]


def conv1x1(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, relu: bool = False, precision: Optional[str] = None) -> torch.Tensor:
    """A 1x1 convolution on RoI tiles through the tcgen05 contraction (fgn_conv1x1_nhwc): ``x`` [R,Cin,H,W] (repacked to
    channels_last if it is not), ``weight`` [Cout,Cin] or [Cout,Cin,1,1] with any BatchNorm already folded in, optional
    ``residual`` [R,Cout,H,W] and ReLU in the epilogue.  Returns [R,Cout,H,W] in channels_last storage.
    ``precision``: "fp32" (3xTF32, fp32 parity), "tf32" (single pass), None = follow ``torch.backends.cudnn.allow_tf32``,
    i.e. the precision cuDNN gives the convolutions around this one."""
    _need_cuda(x, weight, bias, residual)
    x = _f32(x, "x")
    if storage_layout(x) != LAYOUT_NHWC:
        x = to_nhwc(x if storage_layout(x) is not None else x.contiguous())
    r, cin, h, w = x.shape
    wt = _f32(weight, "weight").reshape(weight.shape[0], -1).contiguous()
    cout = wt.shape[0]
    if wt.shape[1] != cin or cin % 4 or cout % 4:
        raise FgnError(f"conv1x1: weight {tuple(weight.shape)} does not match Cin={cin} (channel counts must be multiples of 4)")
    if residual is not None:
        residual = _f32(residual, "residual")
        if tuple(residual.shape) != (r, cout, h, w):
            raise FgnError("conv1x1: residual must be [R,Cout,H,W]")
        if storage_layout(residual) != LAYOUT_NHWC:
            residual = to_nhwc(residual if storage_layout(residual) is not None else residual.contiguous())
    out = _empty_like_format((r, cout, h, w), x.device, LAYOUT_NHWC)
    lib = _lib.load()
    wsb = int(lib.fgn_gemm_workspace_bytes(cout, cin))
    ws = torch.empty((max(wsb, 1),), device=x.device, dtype=torch.uint8)
    _lib.check(lib.fgn_conv1x1_nhwc(x.data_ptr(), wt.data_ptr(), _ptr(None if bias is None else _f32(bias, "bias").contiguous()),
                                    _ptr(residual), int(bool(relu)), out.data_ptr(), r * h * w, cin, cout,
                                    {"fp32": 0, "tf32": 1}[precision or ("tf32" if torch.backends.cudnn.allow_tf32 else "fp32")],
                                    ws.data_ptr(), wsb, _stream()), "fgn_conv1x1_nhwc")
    return out


def _precision_code(precision: Optional[str]) -> int:
    return {"fp32": 0, "tf32": 1}[precision or ("tf32" if torch.backends.cudnn.allow_tf32 else "fp32")]


def conv_taps(weight: torch.Tensor, transposed: bool = False) -> torch.Tensor:
    """Weights in the layout the tcgen05 convolutions read: ``[taps, Cout, Cin]`` (K-major rows per tap).
    Conv2d weight [Cout,Cin,kh,kw] -> tap = ky*kw+kx; ConvTranspose2d weight [Cin,Cout,2,2] (``transposed``) -> tap = i*2+j."""
    w = _f32(weight, "weight")
    w = w.permute(2, 3, 1, 0) if transposed else w.permute(2, 3, 0, 1)
    return w.reshape(-1, w.shape[2], w.shape[3]).contiguous()


def conv_split_weights(w_taps: torch.Tensor) -> torch.Tensor:
    """Load-time TF32 hi/lo split of ``conv_taps`` weights (fgn_conv_split_weights) for the fp32-parity passes."""
    _need_cuda(w_taps)
    t, cout, cin = w_taps.shape
    lib = _lib.load()
    out = torch.empty((2, t, cout, cin), device=w_taps.device, dtype=torch.float32)
    _lib.check(lib.fgn_conv_split_weights(w_taps.data_ptr(), t, cout, cin, out.data_ptr(), _stream()), "fgn_conv_split_weights")
    return out


def conv3x3(x: torch.Tensor, w_taps: torch.Tensor, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            relu: bool = False, precision: Optional[str] = None, w_split: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 convolution on RoI tiles as a tcgen05 implicit GEMM (fgn_conv3x3_nhwc): ``x`` [R,Cin,H,W]
    (repacked to channels_last if it is not), ``w_taps`` = ``conv_taps(conv.weight)`` with any BatchNorm folded in,
    optional ``bias`` / ``residual`` / ReLU in the epilogue.  Returns [R,Cout,H,W] in channels_last storage.
    ``precision`` as in ``conv1x1``."""
    _need_cuda(x, w_taps, bias, residual, w_split)
    x = _f32(x, "x")
    r, cin, h, w = x.shape
    if w_taps.dim() != 3 or w_taps.shape[0] != 9 or w_taps.shape[2] != cin:
        raise FgnError(f"conv3x3: w_taps {tuple(w_taps.shape)} must be [9,Cout,{cin}] (ops.conv_taps)")
    cout = w_taps.shape[1]
    if r == 0:
        return _empty_like_format((0, cout, h, w), x.device, LAYOUT_NHWC)
    if storage_layout(x) != LAYOUT_NHWC:
        x = to_nhwc(x if storage_layout(x) is not None else x.contiguous())
    if residual is not None:
        residual = _f32(residual, "residual")
        if tuple(residual.shape) != (r, cout, h, w):
            raise FgnError("conv3x3: residual must be [R,Cout,H,W]")
        if storage_layout(residual) != LAYOUT_NHWC:
            residual = to_nhwc(residual if storage_layout(residual) is not None else residual.contiguous())
    out = _empty_like_format((r, cout, h, w), x.device, LAYOUT_NHWC)
    lib = _lib.load()
    prec = _precision_code(precision)
    wsb = int(lib.fgn_conv_split_weights_bytes(9, cout, cin)) if (prec == 0 and w_split is None) else 0
    ws = torch.empty((max(wsb, 1),), device=x.device, dtype=torch.uint8)
    _lib.check(lib.fgn_conv3x3_nhwc(x.data_ptr(), w_taps.contiguous().data_ptr(), _ptr(w_split),
                                    _ptr(None if bias is None else _f32(bias, "bias").contiguous()), _ptr(residual),
                                    int(bool(relu)), out.data_ptr(), r, h, w, cin, cout, prec, ws.data_ptr(), wsb, _stream()),
               "fgn_conv3x3_nhwc")
    return out


def deconv2x2_logits(x: torch.Tensor, w_taps: torch.Tensor, b_deconv: Optional[torch.Tensor], w_logits: torch.Tensor,
                     b_logits: Optional[torch.Tensor], precision: Optional[str] = None,
                     w_split: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FCNMaskHead's tail in one launch (fgn_deconv2x2_logits_nhwc): ConvTranspose2d(k=2, s=2) + ReLU + 1x1 logits on
    ``x`` [R,Cin,H,W]; ``w_taps`` = ``conv_taps(upsample.weight, transposed=True)`` [4,Cout,Cin], ``w_logits`` [ncls,Cout]
    (or [ncls,Cout,1,1]).  Returns mask_pred [R,ncls,2H,2W] (contiguous NCHW)."""
    _need_cuda(x, w_taps, b_deconv, w_logits, b_logits, w_split)
    x = _f32(x, "x")
    if storage_layout(x) != LAYOUT_NHWC:
        x = to_nhwc(x if storage_layout(x) is not None else x.contiguous())
    r, cin, h, w = x.shape
    if w_taps.dim() != 3 or w_taps.shape[0] != 4 or w_taps.shape[2] != cin:
        raise FgnError(f"deconv2x2_logits: w_taps {tuple(w_taps.shape)} must be [4,Cout,{cin}] (ops.conv_taps(..., transposed=True))")
    cout = w_taps.shape[1]
    wl = _f32(w_logits, "w_logits").reshape(w_logits.shape[0], -1).contiguous()
    if wl.shape[1] != cout:
        raise FgnError("deconv2x2_logits: w_logits must be [ncls,Cout]")
    ncls = wl.shape[0]
    out = torch.empty((r, ncls, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    prec = _precision_code(precision)
    wsb = int(lib.fgn_conv_split_weights_bytes(4, cout, cin)) if (prec == 0 and w_split is None) else 0
    ws = torch.empty((max(wsb, 1),), device=x.device, dtype=torch.uint8)
    _lib.check(lib.fgn_deconv2x2_logits_nhwc(x.data_ptr(), w_taps.contiguous().data_ptr(), _ptr(w_split),
                                             _ptr(None if b_deconv is None else _f32(b_deconv, "b_deconv").contiguous()),
                                             wl.data_ptr(), _ptr(None if b_logits is None else _f32(b_logits, "b_logits").contiguous()),
                                             out.data_ptr(), r, h, w, cin, cout, ncls, prec, ws.data_ptr(), wsb, _stream()),
               "fgn_deconv2x2_logits_nhwc")
    return out
