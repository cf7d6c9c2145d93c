"""autograd.Function wrappers: forward = the C-ABI forward kernels, backward = the hand-written adjoint
kernels of fgn_b200/csrc/backward.cu and relation_bwd.cu.  The guided heads
(FGNRoIHead, AGRPNHead, SingleRoIExtractor) call the functions at the bottom of this file whenever autograd is
recording and an input or parameter requires grad; otherwise they call the plain ``ops``.

The reference gets these gradients from autograd over mmcv / torchvision / ATen ops
(fgn_roi_head.py:344-358,451-529, fgn_ag_rpn_head.py:58-79).  Feature tensors travel channels_last.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib, ops
from ._lib import LAYOUT_NHWC, FgnError, Pyramid


def _cl(t: torch.Tensor) -> torch.Tensor:
    """channels_last fp32 view/copy of a logical-NCHW tensor."""
    t = t.float() if t.dtype != torch.float32 else t
    return t if ops.storage_layout(t) == LAYOUT_NHWC else ops.to_nhwc(t.contiguous() if ops.storage_layout(t) is None else t)


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in ts)


_DETERMINISTIC = None          # None: follow torch.are_deterministic_algorithms_enabled()


def set_deterministic_backward(flag) -> None:
    """RoIAlign backward with order-independent (64-bit fixed-point) accumulation: True / False, or None to follow
    ``torch.use_deterministic_algorithms``.  The default scatter uses float atomics like mmcv's own backward."""
    global _DETERMINISTIC
    _DETERMINISTIC = flag


def deterministic_backward() -> bool:
    return torch.are_deterministic_algorithms_enabled() if _DETERMINISTIC is None else bool(_DETERMINISTIC)


class _RoIAlignML(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rois, scales, output_size, sampling_ratio, aligned, finest_scale, *feats):
        feats = [_cl(f) for f in feats]
        out = ops.roi_align_multilevel(feats, rois, scales, output_size, sampling_ratio, aligned, finest_scale,
                                       out_format="nhwc")
        ctx.save_for_backward(rois)
        ctx.meta = (tuple(scales), int(output_size), int(sampling_ratio), bool(aligned), float(finest_scale),
                    [tuple(f.shape) for f in feats])
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        scales, p, sr, aligned, finest, shapes = ctx.meta
        g = _cl(g)
        b, c = shapes[0][:2]
        grads = [torch.zeros((s[0], s[2], s[3], s[1]), device=g.device, dtype=torch.float32).permute(0, 3, 1, 2)
                 for s in shapes]
        pyr = Pyramid()
        pyr.num_levels = len(shapes)
        for i, (gr, s) in enumerate(zip(grads, shapes)):
            pyr.feat[i], pyr.H[i], pyr.W[i], pyr.spatial_scale[i] = gr.data_ptr(), s[2], s[3], float(scales[i])
        r = rois.shape[0]
        lib = _lib.load()
        if deterministic_backward():
            # bit-identical gradients from run to run: 64-bit fixed-point accumulation (fgn_roi_align_ml_bwd_det)
            wsb = int(lib.fgn_roi_align_ml_bwd_det_workspace_bytes(ctypes.byref(pyr), b, c))
            ws = torch.empty((max(wsb, 1),), device=g.device, dtype=torch.uint8)
            _lib.check(lib.fgn_roi_align_ml_bwd_det(ctypes.byref(pyr), b, c, rois.contiguous().data_ptr(), r, p, sr, int(aligned),
                                                    finest, None, 0, None, g.data_ptr(), ws.data_ptr(), wsb,
                                                    torch.cuda.current_stream().cuda_stream), "fgn_roi_align_ml_bwd_det")
        else:
            _lib.check(lib.fgn_roi_align_ml_bwd(ctypes.byref(pyr), b, c, rois.contiguous().data_ptr(), r, p, sr,
                                                int(aligned), finest, None, None, g.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream), "fgn_roi_align_ml_bwd")
        return (None,) * 6 + tuple(grads)


class _ChannelAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qry, vec):
        q = _cl(qry)
        ctx.save_for_backward(q, vec)
        return ops.channel_attention(q, vec)

    @staticmethod
    def backward(ctx, g):
        q, vec = ctx.saved_tensors
        g = _cl(g)
        b, c, h, w = q.shape
        n = vec.shape[1]
        v = vec.reshape(b * n, c).contiguous()
        lib = _lib.load()
        need_q, need_v = ctx.needs_input_grad
        gq = torch.empty((b, h, w, c), device=g.device, dtype=torch.float32).permute(0, 3, 1, 2) if need_q else None
        gv = torch.empty((b * n, c), device=g.device, dtype=torch.float32) if need_v else None
        wsb = lib.fgn_channel_attention_bwd_workspace_bytes(b, n, c, h, w) if need_v else 0
        ws = torch.empty((max(wsb, 1),), device=g.device, dtype=torch.uint8)
        _lib.check(lib.fgn_channel_attention_bwd(q.data_ptr(), v.data_ptr(), g.data_ptr(), b, n, c, h, w,
                                                 None if gq is None else gq.data_ptr(),
                                                 None if gv is None else gv.data_ptr(), ws.data_ptr(), wsb,
                                                 torch.cuda.current_stream().cuda_stream), "fgn_channel_attention_bwd")
        return gq, None if gv is None else gv.view_as(vec)


class _AttentionVectors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spp, n_ways, k_shots):
        x = _cl(spp)
        ctx.meta = (tuple(x.shape), int(n_ways), int(k_shots))
        return ops.attention_vectors(x, n_ways, k_shots)

    @staticmethod
    def backward(ctx, g):
        (bnk, c, h, w), n, k = ctx.meta
        bn = bnk // k
        gv = g.reshape(bn, c).contiguous().float()
        out = torch.empty((bnk, h, w, c), device=g.device, dtype=torch.float32).permute(0, 3, 1, 2)
        _lib.check(_lib.load().fgn_attention_vectors_bwd(gv.data_ptr(), bn, k, c, h, w, out.data_ptr(),
                                                         torch.cuda.current_stream().cuda_stream), "fgn_attention_vectors_bwd")
        return out, None, None


class _SupportPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, m, n_ways, k_shots):
        x = _cl(f)
        mm = m.reshape(x.shape[0], -1).contiguous().float()
        ctx.save_for_backward(mm)
        ctx.meta = (tuple(x.shape), int(n_ways), int(k_shots))
        cat, gap = ops.support_pool(x, mm, n_ways, k_shots, out_format="nhwc")
        return cat, gap

    @staticmethod
    def backward(ctx, g_cat, g_gap):
        (mm,) = ctx.saved_tensors
        (bnk, c, p, _), n, k = ctx.meta
        bn = bnk // k
        gc = _cl(g_cat.reshape(bn, c, p, p)) if g_cat is not None else None
        gg = g_gap.reshape(bn, c).contiguous().float() if g_gap is not None else None
        out = torch.empty((bnk, p, p, c), device=mm.device, dtype=torch.float32).permute(0, 3, 1, 2)
        _lib.check(_lib.load().fgn_support_pool_bwd(None if gc is None else gc.data_ptr(), None if gg is None else gg.data_ptr(),
                                                    mm.data_ptr(), bn, k, c, p, out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "fgn_support_pool_bwd")
        return out, None, None, None


class _RelationFusion(torch.autograd.Function):
    """count_one_roi_by_n_spp + BBoxHead.forward + count_modified_cls_bbox (fgn_roi_head.py:336-339) with the
    hand-written adjoint of fgn_b200/csrc/relation_bwd.cu (fgn_relation_fusion_bwd)."""

    @staticmethod
    def forward(ctx, roi_feat, roi_batch, spp_cat_mean, n_ways, gn_groups, gn_eps, conv_w, conv_b, gn_w, gn_b,
                fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b):
        x = _cl(roi_feat)
        c = x.shape[1]
        s = _cl(spp_cat_mean.reshape(-1, c, x.shape[2], x.shape[3]))
        params = ops.RelationParams(conv_w, conv_b, gn_w, gn_b, fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b, gn_groups, gn_eps)
        rb = roi_batch.to(torch.int32).contiguous()
        cls, reg = ops.relation_fusion(x, rb, s, n_ways, params)
        ctx.save_for_backward(x, s, rb, cls, params.conv_w, params.conv_b, params.gn_w, params.gn_b, params.fc_cls_w,
                              params.fc_reg_w)
        ctx.meta = (int(n_ways), int(gn_groups), float(gn_eps), tuple(spp_cat_mean.shape))
        ctx.conv4d = conv_w.dim() == 4
        return cls, reg

    @staticmethod
    def backward(ctx, g_cls, g_reg):
        x, s, rb, cls, conv_w, conv_b, gn_w, gn_b, fc_cls_w, fc_reg_w = ctx.saved_tensors
        n, groups, eps, spp_shape = ctx.meta
        r, c, p, _ = x.shape
        bn = s.shape[0]
        b = bn // n
        dev = x.device
        # the forward's split-conv outputs, re-derived on the forward's own contraction kernels
        xq = x.permute(0, 2, 3, 1).reshape(r * p * p, c)              # NHWC storage: a view
        xs = s.permute(0, 2, 3, 1).reshape(bn * p * p, c)
        yq = ops.gemm_nt(xq, conv_w[:, :c])
        ys = ops.gemm_nt(xs, conv_w[:, c:], conv_b)
        need = ctx.needs_input_grad
        f32 = dict(device=dev, dtype=torch.float32)
        d_x = torch.empty((r, p, p, c), **f32) if need[0] else None
        d_s = torch.empty((bn, p, p, c), **f32) if need[2] else None
        d_cw = torch.empty((c, 2 * c), **f32) if need[6] else None
        d_cb = torch.empty((c,), **f32) if need[7] else None
        d_gw = torch.empty((c,), **f32) if need[8] else None
        d_gb = torch.empty((c,), **f32) if need[9] else None
        d_fcw = torch.empty((2, c), **f32) if need[10] else None
        d_fcb = torch.empty((2,), **f32) if need[11] else None
        d_frw = torch.empty((4, c), **f32) if need[12] else None
        d_frb = torch.empty((4,), **f32) if need[13] else None
        lib = _lib.load()
        wsb = int(lib.fgn_relation_fusion_bwd_workspace_bytes(r, bn, n, c))
        ws = torch.empty((max(wsb, 256),), device=dev, dtype=torch.uint8)
        ptr = lambda t: None if t is None else t.data_ptr()
        gc = g_cls.contiguous().float() if g_cls is not None else torch.zeros_like(cls)
        gr = g_reg.contiguous().float() if g_reg is not None else torch.zeros((r, 4 * n), **f32)
        _lib.check(lib.fgn_relation_fusion_bwd(
            xq.data_ptr(), xs.data_ptr(), rb.data_ptr(), yq.data_ptr(), ys.data_ptr(), cls.data_ptr(), gc.data_ptr(),
            gr.data_ptr(), r, b, n, c, p, conv_w.data_ptr(), gn_w.data_ptr(), gn_b.data_ptr(), groups, eps,
            fc_cls_w.data_ptr(), fc_reg_w.data_ptr(), ptr(d_x), ptr(d_s), ptr(d_cw), ptr(d_cb), ptr(d_gw), ptr(d_gb),
            ptr(d_fcw), ptr(d_fcb), ptr(d_frw), ptr(d_frb), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream),
            "fgn_relation_fusion_bwd")
        g_x = d_x.permute(0, 3, 1, 2) if d_x is not None else None
        g_s = d_s.permute(0, 3, 1, 2).reshape(spp_shape) if d_s is not None else None
        return (g_x, None, g_s, None, None, None, d_cw.view(c, 2 * c, 1, 1) if d_cw is not None and ctx.conv4d else d_cw,
                d_cb, d_gw, d_gb, d_fcw, d_fcb, d_frw, d_frb)


# ---- what the heads call ---------------------------------------------------------------------------
def roi_align_multilevel(feats: Sequence[torch.Tensor], rois, scales, output_size=7, sampling_ratio=0, aligned=True,
                         finest_scale=56.0, **kw):
    """ops.roi_align_multilevel, differentiable w.r.t. the feature maps when they require grad."""
    if _needs_grad(*feats):
        if kw.get("chan_scale") is not None:
            raise FgnError("fused chan_scale has no adjoint: use roi_align_multilevel + channel_attention under autograd")
        if any(f.dtype == torch.bfloat16 for f in feats):
            raise FgnError("the bf16 variant has no adjoint kernels: train in fp32")
        out = _RoIAlignML.apply(rois, list(scales), output_size, sampling_ratio, aligned, finest_scale, *feats)
        if kw.get("out_format", "nchw") in ("nchw", "contiguous"):
            out = out.contiguous()                        # (logical shape is [R,C,P,P] either way; this only fixes the storage)
        return (out, ops.map_roi_levels(rois, len(feats), finest_scale)) if kw.get("return_levels") else out
    return ops.roi_align_multilevel(feats, rois, scales, output_size, sampling_ratio, aligned, finest_scale, **kw)


def channel_attention(qry: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    return _ChannelAttention.apply(qry, vec) if _needs_grad(qry, vec) else ops.channel_attention(qry, vec)


def attention_vectors(spp: torch.Tensor, n_ways: int, k_shots: int) -> torch.Tensor:
    return _AttentionVectors.apply(spp, n_ways, k_shots) if _needs_grad(spp) else ops.attention_vectors(spp, n_ways, k_shots)


def support_pool(f: torch.Tensor, m: torch.Tensor, n_ways: int, k_shots: int, out_format: Optional[str] = None):
    if _needs_grad(f):
        return _SupportPool.apply(f, m, n_ways, k_shots)
    return ops.support_pool(f, m, n_ways, k_shots, out_format)


def relation_fusion(roi_feat, roi_batch, spp_cat_mean, n_ways: int, conv_w, conv_b, gn_w, gn_b, fc_cls_w, fc_cls_b,
                    fc_reg_w, fc_reg_b, gn_groups: int = 32, gn_eps: float = 1e-5):
    """ops.relation_fusion, differentiable w.r.t. the RoI features, the class maps and all eight parameter tensors
    (``conv_w`` may be the Conv2d weight [C,2C,1,1] or [C,2C]).  -> (cls_score [R,N+1], bbox_pred [R,4N])."""
    return _RelationFusion.apply(roi_feat, roi_batch, spp_cat_mean, n_ways, gn_groups, gn_eps, conv_w, conv_b, gn_w, gn_b,
                                 fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b)
