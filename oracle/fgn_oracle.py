"""oracle/fgn_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch CPU, fp32) of FGN's guided RoIAlign + support-guided fusion hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; nothing under fgn_b200/ does.

Every function cites the reference lines it follows (paths relative to
/root/reference/subprojects/sp02_omniiseg_fgn_mmdet/).  Third-party arithmetic ([3P]: mmcv-full
1.3.16 RoIAlign, mmdet 2.18.0 SingleRoIExtractor/bbox2roi/BBoxHead, torchvision 0.10 roi_align --
requirements.txt:45-46,100-101, none of them under /root/reference) is restated from SURVEY.md
appendix A; RoIAlign itself is served either by oracle/roi_align_ref.c (plain C, also emits the
sample indices) or by torch.ops.torchvision.roi_align on CPU (the very op fgn_roi_head.py:429 calls).

Parity pin (see tests/test_oracle.py, tests/golden/make_golden.py):
  * the FGN-owned functions here are checked against fixtures produced by executing the
    reference's own, unmodified methods (imported from /root/reference with the mmdet/mmcv
    imports stubbed) -- tests/golden/fgn_reference_*.npz;
  * roi_align_ref.c is checked bit-for-bit against torchvision's CPU op;
  * map_roi_levels and the mmcv RoIAlign (aligned=True) have no reference-side test or golden
    vector and mmcv/mmdet cannot be installed here: that part is "parity unpinned" beyond
    torchvision's own aligned=True mode.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(HERE, "roi_align_ref.c")
_C_LIB = os.path.join(HERE, "_build", "liboracle.so")
_clib = None


def build_c_oracle(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no FMA, like the CPU wheels) -> oracle/_build/liboracle.so."""
    if force or not os.path.exists(_C_LIB) or os.path.getmtime(_C_LIB) < os.path.getmtime(_C_SRC):
        os.makedirs(os.path.dirname(_C_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _C_LIB, _C_SRC, "-lm"])
    return _C_LIB


def _c():
    global _clib
    if _clib is None:
        _clib = ctypes.CDLL(build_c_oracle())
    return _clib


_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_i32p)


# ------------------------------------------------------------------------------------------------
# [3P] pieces
# ------------------------------------------------------------------------------------------------
def bbox2roi(bbox_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """mmdet.core.bbox2roi [3P] (called at fgn_roi_head.py:348,390,556,654); SURVEY A.2."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def map_roi_levels(rois: torch.Tensor, num_levels: int, finest_scale: float = 56.0) -> torch.Tensor:
    """mmdet SingleRoIExtractor.map_roi_levels [3P], SURVEY A.2, as the torch expression."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    target_lvls = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    target_lvls = target_lvls.clamp(min=0, max=num_levels - 1).long()
    return target_lvls


def map_roi_levels_c(rois: torch.Tensor, num_levels: int, finest_scale: float = 56.0) -> torch.Tensor:
    """Same, through oracle/roi_align_ref.c (the exact contract the CUDA kernel implements)."""
    r = np.ascontiguousarray(rois.detach().cpu().numpy(), dtype=np.float32)
    out = np.zeros((r.shape[0],), np.int32)
    _c().fgn_oracle_map_roi_levels(_fp(r), r.shape[0], int(num_levels), ctypes.c_float(finest_scale), _ip(out))
    return torch.from_numpy(out).long()


def roi_align_tv(feat: torch.Tensor, rois: torch.Tensor, spatial_scale: float, output_size: int,
                 sampling_ratio: int, aligned: bool) -> torch.Tensor:
    """torch.ops.torchvision.roi_align on CPU (Detectron ROIAlign; same algorithm as mmcv's)."""
    import torchvision  # noqa: F401  (registers the op)
    return torch.ops.torchvision.roi_align(feat.float().contiguous(), rois.float().contiguous(), float(spatial_scale),
                                           int(output_size), int(output_size), int(sampling_ratio), bool(aligned))


def roi_align_c(feat: torch.Tensor, rois: torch.Tensor, spatial_scale: float, output_size: int,
                sampling_ratio: int, aligned: bool) -> torch.Tensor:
    """oracle/roi_align_ref.c: SURVEY A.1 restated in plain C, no FMA."""
    f = np.ascontiguousarray(feat.detach().cpu().numpy(), dtype=np.float32)
    r = np.ascontiguousarray(rois.detach().cpu().numpy(), dtype=np.float32)
    b, c, h, w = f.shape
    n, p = r.shape[0], int(output_size)
    out = np.zeros((n, c, p, p), np.float32)
    if n:
        _c().fgn_oracle_roi_align(_fp(f), b, c, h, w, _fp(r), n, ctypes.c_float(spatial_scale), p, p,
                                  int(sampling_ratio), int(bool(aligned)), _fp(out))
    return torch.from_numpy(out)


def roi_align(feat, rois, spatial_scale, output_size, sampling_ratio, aligned, impl: str = "tv"):
    return (roi_align_tv if impl == "tv" else roi_align_c)(feat, rois, spatial_scale, output_size, sampling_ratio, aligned)


def roi_align_indices_c(rois: torch.Tensor, lvls: torch.Tensor, hw: Sequence[Tuple[int, int]],
                        scales: Sequence[float], output_size: int, sampling_ratio: int, aligned: bool,
                        max_grid: int):
    """Per-axis sample tables (valid, low, high) of RoIAlign for each RoI at its level."""
    r = np.ascontiguousarray(rois.detach().cpu().numpy(), dtype=np.float32)
    n, p = r.shape[0], int(output_size)
    lv = lvls.cpu().numpy().astype(np.int64)
    sc = np.asarray([scales[l] for l in lv], np.float32)
    hws = np.asarray([hw[l] for l in lv], np.int32).reshape(n, 2)
    grid = np.zeros((n, 2), np.int32)
    ytab = np.zeros((n, p, max_grid, 3), np.int32)
    xtab = np.zeros((n, p, max_grid, 3), np.int32)
    if n:
        _c().fgn_oracle_roi_align_indices(_fp(r), n, _fp(sc), _ip(hws), p, p, int(sampling_ratio),
                                          int(bool(aligned)), int(max_grid), _ip(grid), _ip(ytab), _ip(xtab))
    return torch.from_numpy(grid), torch.from_numpy(ytab), torch.from_numpy(xtab)


def single_roi_extractor(feats: Sequence[torch.Tensor], rois: torch.Tensor, strides: Sequence[int],
                         output_size: int = 7, sampling_ratio: int = 0, aligned: bool = True,
                         finest_scale: float = 56.0, impl: str = "tv"):
    """mmdet SingleRoIExtractor.forward [3P], reached from fgn_roi_head.py:331-332,366-367.
    Returns (roi_feats [R,C,P,P], levels [R])."""
    num_levels = len(feats)
    c = feats[0].shape[1]
    out = feats[0].new_zeros((rois.shape[0], c, output_size, output_size))
    if num_levels == 1:
        lv = torch.zeros((rois.shape[0],), dtype=torch.long)
        if rois.shape[0] == 0:
            return out, lv
        return roi_align(feats[0], rois, 1.0 / strides[0], output_size, sampling_ratio, aligned, impl), lv
    lv = map_roi_levels(rois, num_levels, finest_scale)
    for i in range(num_levels):
        inds = (lv == i).nonzero(as_tuple=False).squeeze(1)
        if inds.numel() > 0:
            out[inds] = roi_align(feats[i], rois[inds], 1.0 / strides[i], output_size, sampling_ratio, aligned, impl)
    return out, lv


def bbox_head_forward(x: torch.Tensor, fc_cls_w, fc_cls_b, fc_reg_w, fc_reg_b):
    """mmdet BBoxHead.forward [3P] with with_avg_pool=True (fgn_r50_c4_densecl.py:76-93), SURVEY A.6."""
    x = F.avg_pool2d(x, x.shape[-1])
    x = x.view(x.size(0), -1)
    return F.linear(x, fc_cls_w, fc_cls_b), F.linear(x, fc_reg_w, fc_reg_b)


# ------------------------------------------------------------------------------------------------
# FGN-owned pieces (restated from the reference's own lines)
# ------------------------------------------------------------------------------------------------
def agrpn_attention(qry_fmap: torch.Tensor, spp_fmaps: torch.Tensor, n_ways: int, k_shots: int):
    """fgn_ag_rpn_head.py:33-46 -> (spp_fvecs_cat_mean [B,N,C,1,1], qry_fmap_mod [B*N,C,H,W])."""
    batch, c, x_h, x_w = qry_fmap.shape
    qry = qry_fmap[:, None, :, :, :]
    c, h, w = spp_fmaps.shape[-3:]
    vec = spp_fmaps.view(batch, n_ways, k_shots, c, h, w).mean(axis=(2, 4, 5)).view(batch, n_ways, c, 1, 1)
    mod = (qry * vec).view(batch * n_ways, c, x_h, x_w)
    return vec, mod


def best_class_selection(rpn_cls_score: torch.Tensor, rpn_bbox_pred: torch.Tensor, batch: int, n_ways: int):
    """fgn_ag_rpn_head.py:82-113."""
    _, c, x_h, x_w = rpn_cls_score.shape
    cls = rpn_cls_score.view(batch, n_ways, c, x_h, x_w)
    _, c4, x_h, x_w = rpn_bbox_pred.shape
    reg = rpn_bbox_pred.view(batch, n_ways, c4, x_h, x_w)
    if n_ways > 1:
        cls_all, reg_all = [], []
        for i in range(batch):
            a_scores = cls[i].permute(0, 2, 3, 1).reshape(n_ways, -1, 1)
            a_deltas = reg[i].permute(0, 2, 3, 1).reshape(n_ways, -1, 4)
            index = 1 if a_scores.shape[-1] == 2 else 0
            argmax = torch.argmax(a_scores[:, :, index], dim=0)
            arranged = torch.arange(len(argmax))
            cls_new = a_scores[argmax, arranged, :]
            reg_new = a_deltas[argmax, arranged, :]
            _, na, h, w = cls[i].shape
            cls_all.append(cls_new.view(1, h, w, na).permute(0, 3, 1, 2))
            _, nb, h, w = reg[i].shape
            reg_all.append(reg_new.view(1, h, w, nb).permute(0, 3, 1, 2))
        return torch.cat(cls_all, 0), torch.cat(reg_all, 0)
    return cls.view(batch, c, x_h, x_w), reg.view(batch, c4, x_h, x_w)


def count_spp(spp_fmaps: torch.Tensor, spp_bboxes: torch.Tensor, spp_isegmaps: torch.Tensor, n_ways: int,
              k_shots: int, subsampling_ratio: float = 16, shared_head=None, impl: str = "tv"):
    """fgn_roi_head.py:419-449.  Mutates spp_bboxes in place (/= subsampling_ratio) like the reference.
    Returns (cat_mean [B,N,C,7,7], masked_gap [B,N,C,1,1], mask_roi [BNK,1,7,7], feat_roi [BNK,C,7,7])."""
    m = spp_bboxes.shape[0]
    idx = torch.arange(m, dtype=torch.float32).view(m, 1)
    rois = torch.cat([idx, spp_bboxes.reshape(m, 4)], 1)      # torchvision list-of-[1,4] form (:429)
    mask_ra = roi_align(spp_isegmaps.float(), rois, 1.0, 7, -1, False, impl)
    spp_bboxes /= subsampling_ratio                             # :430, in place
    rois = torch.cat([idx, spp_bboxes.reshape(m, 4)], 1)
    feat_ra = roi_align(spp_fmaps, rois, 1.0, 7, -1, False, impl)
    if shared_head is not None:
        feat_ra = shared_head(feat_ra)                          # :435-436 (C4 only)
    c, h, w = feat_ra.shape[-3:]
    cat_mean = feat_ra.view(-1, n_ways, k_shots, c, h, w).mean(dim=2).view(-1, n_ways, c, h, w)
    mp = (feat_ra * mask_ra).view(-1, n_ways, k_shots, c, h, w).mean(dim=(2, 4, 5)).view(-1, n_ways, c, 1, 1)
    return cat_mean, mp, mask_ra, feat_ra


def count_spp_fpn(spp_feats: Sequence[torch.Tensor], strides: Sequence[int], spp_bboxes: torch.Tensor,
                  spp_isegmaps: torch.Tensor, n_ways: int, k_shots: int, finest_scale: float = 56.0,
                  impl: str = "tv"):
    """FPN generalisation of count_spp (SURVEY A.9 assumption A-FPN): the support box is pooled from
    the level map_roi_levels assigns it, with spatial_scale = 1/stride_l, otherwise identical
    (sampling_ratio=-1, aligned=False).  With one level of stride 16 this is count_spp."""
    m = spp_bboxes.shape[0]
    idx = torch.arange(m, dtype=torch.float32).view(m, 1)
    rois = torch.cat([idx, spp_bboxes.reshape(m, 4)], 1)
    mask_ra = roi_align(spp_isegmaps.float(), rois, 1.0, 7, -1, False, impl)
    feat_ra, _ = single_roi_extractor(spp_feats, rois, strides, 7, -1, False, finest_scale, impl)
    c, h, w = feat_ra.shape[-3:]
    cat_mean = feat_ra.view(-1, n_ways, k_shots, c, h, w).mean(dim=2).view(-1, n_ways, c, h, w)
    mp = (feat_ra * mask_ra).view(-1, n_ways, k_shots, c, h, w).mean(dim=(2, 4, 5)).view(-1, n_ways, c, 1, 1)
    return cat_mean, mp, mask_ra, feat_ra


def count_one_roi_by_n_spp(bbox_feats: torch.Tensor, rois: torch.Tensor, spp_cat_mean: torch.Tensor, n_ways: int,
                           conv_w, conv_b, gn_w, gn_b, gn_groups: int = 32, gn_eps: float = 1e-5):
    """fgn_roi_head.py:253-279, materialised concat form, line for line."""
    rois_amount = bbox_feats.shape[0]
    batch = spp_cat_mean.shape[0]
    indexes = torch.repeat_interleave(torch.arange(len(rois)), n_ways)
    c, h, w = bbox_feats.shape[-3:]
    rois_repeated = bbox_feats[indexes].view(rois_amount * n_ways, c, h, w)
    c, h, w = spp_cat_mean.shape[-3:]
    spps_repeated = spp_cat_mean.view(batch, n_ways, c, h, w)
    indexes = rois[:, 0].long()
    spps_repeated = spps_repeated[indexes].view(rois_amount * n_ways, c, h, w)
    x = torch.cat((rois_repeated, spps_repeated), dim=1)
    y = F.conv2d(x, conv_w.view(conv_w.shape[0], -1, 1, 1), conv_b)
    y = F.group_norm(y, gn_groups, gn_w, gn_b, gn_eps)
    return rois_amount, F.relu(y)


def count_modified_cls_bbox(rois_amount: int, cls_score: torch.Tensor, bbox_pred: torch.Tensor, n_ways: int):
    """fgn_roi_head.py:302-326; the hard-coded [1,3,5] generalised to 1::2."""
    if n_ways == 1:
        return cls_score[:, [1, 0]], bbox_pred
    reshaped = cls_score.view(rois_amount, n_ways * 2)
    top = reshaped[:, 1::2].argmax(dim=-1) * 2
    indexes = torch.arange(rois_amount)
    bg_class = reshaped[indexes, top].view(rois_amount, 1)
    pr_class = reshaped[:, 1::2]
    return torch.cat((pr_class, bg_class), dim=1), bbox_pred.view(rois_amount, n_ways * 4)


def bbox_forward(feats: Sequence[torch.Tensor], strides: Sequence[int], rois: torch.Tensor,
                 spp_cat_mean: torch.Tensor, n_ways: int, w: dict, shared_head=None, impl: str = "tv",
                 chunk: Optional[int] = None):
    """fgn_roi_head.py:328-342 (_bbox_forward).  `chunk` bounds the [R*N,2C,7,7] concat by looping over
    RoI chunks (results are per-RoI independent); used for the N=20 stress config."""
    bbox_feats, lv = single_roi_extractor(feats, rois, strides, 7, 0, True, 56.0, impl)
    if shared_head is not None:
        bbox_feats = shared_head(bbox_feats)
    r = rois.shape[0]
    step = r if not chunk else chunk
    cls_l, reg_l = [], []
    for s in range(0, r, max(step, 1)):
        e = min(r, s + step)
        n_r, fused = count_one_roi_by_n_spp(bbox_feats[s:e], rois[s:e], spp_cat_mean, n_ways, w["conv_w"], w["conv_b"],
                                            w["gn_w"], w["gn_b"], w.get("gn_groups", 32), w.get("gn_eps", 1e-5))
        cls_raw, reg_raw = bbox_head_forward(fused, w["fc_cls_w"], w["fc_cls_b"], w["fc_reg_w"], w["fc_reg_b"])
        c_, r_ = count_modified_cls_bbox(n_r, cls_raw, reg_raw, n_ways)
        cls_l.append(c_)
        reg_l.append(r_)
    if r == 0:
        return dict(cls_score=rois.new_zeros((0, n_ways + 1)), bbox_pred=rois.new_zeros((0, 4 * n_ways)),
                    bbox_feats=bbox_feats, levels=lv)
    return dict(cls_score=torch.cat(cls_l), bbox_pred=torch.cat(reg_l), bbox_feats=bbox_feats, levels=lv)


def mask_attention(feats: Sequence[torch.Tensor], strides: Sequence[int], rois: torch.Tensor,
                   masked_gap: torch.Tensor, det_labels: List[torch.Tensor], n_ways: int,
                   output_size: int = 7, shared_head=None, impl: str = "tv"):
    """fgn_roi_head.py:360-382 with the vector gather of :707-714 (test) / :516-522 (train)."""
    gather = torch.cat([det_labels[i] + n_ways * i for i in range(len(det_labels))])
    batch, n, c = masked_gap.shape[:3]
    vecs = masked_gap.view(batch * n_ways, c, 1, 1)[gather]
    mask_feats, _ = single_roi_extractor(feats, rois, strides, output_size, 0, True, 56.0, impl)
    if shared_head is not None:
        mask_feats = shared_head(mask_feats)
    return mask_feats * vecs


# ---- test-time box post-processing: BBoxHead.get_bboxes [3P] (called at fgn_roi_head.py:606-613) -------------
def delta2bbox(rois4: torch.Tensor, deltas: torch.Tensor, means, stds, max_shape=None, wh_ratio_clip: float = 16 / 1000):
    """mmdet 2.18 DeltaXYWHBBoxCoder.decode / delta2bbox [3P; mmdet absent from /root/reference, parity unpinned]:
    restated from its published source, operation by operation.  rois4 [R,4], deltas [R,4N] -> [R,4N]."""
    num_bboxes, num_classes = deltas.size(0), deltas.size(1) // 4
    if num_bboxes == 0:
        return deltas
    deltas = deltas.reshape(-1, 4)
    means_t = deltas.new_tensor(means).view(1, -1)
    stds_t = deltas.new_tensor(stds).view(1, -1)
    denorm = deltas * stds_t + means_t
    dxy, dwh = denorm[:, :2], denorm[:, 2:]
    rois_ = rois4.repeat(1, num_classes).reshape(-1, 4)
    pxy = (rois_[:, :2] + rois_[:, 2:]) * 0.5
    pwh = rois_[:, 2:] - rois_[:, :2]
    dxy_wh = pwh * dxy
    max_ratio = abs(math.log(wh_ratio_clip))
    dwh = dwh.clamp(min=-max_ratio, max=max_ratio)
    gxy = pxy + dxy_wh
    gwh = pwh * dwh.exp()
    x1y1 = gxy - (gwh * 0.5)
    x2y2 = gxy + (gwh * 0.5)
    bboxes = torch.cat([x1y1, x2y2], dim=-1)
    if max_shape is not None:
        bboxes[..., 0::2].clamp_(min=0, max=max_shape[1])
        bboxes[..., 1::2].clamp_(min=0, max=max_shape[0])
    return bboxes.reshape(num_bboxes, -1)


def nms_greedy(boxes: torch.Tensor, scores: torch.Tensor, iou_thr: float) -> torch.Tensor:
    """mmcv.ops.nms (offset=0) [3P] == torchvision.ops.nms: greedy suppression in (stable) descending score order,
    IoU = inter / (area_a + area_b - inter) > thr.  Plain loops (small cases; tests pin it to torchvision's op)."""
    order = torch.sort(scores, descending=True, stable=True).indices.tolist()
    b = boxes.tolist()
    f32 = np.float32
    area = [f32(f32(x2) - f32(x1)) * f32(f32(y2) - f32(y1)) for x1, y1, x2, y2 in b]
    dead, keep = set(), []
    for a_pos, i in enumerate(order):
        if i in dead:
            continue
        keep.append(i)
        ix1, iy1, ix2, iy2 = (f32(v) for v in b[i])
        for j in order[a_pos + 1:]:
            if j in dead:
                continue
            jx1, jy1, jx2, jy2 = (f32(v) for v in b[j])
            w = max(f32(0), f32(min(ix2, jx2) - max(ix1, jx1)))
            h = max(f32(0), f32(min(iy2, jy2) - max(iy1, jy1)))
            inter = f32(w * h)
            with np.errstate(invalid="ignore", divide="ignore"):
                ovr = f32(inter / f32(f32(area[i] + area[j]) - inter))
            if ovr > f32(iou_thr):
                dead.add(j)
    return torch.tensor(keep, dtype=torch.long)


def multiclass_nms(multi_bboxes: torch.Tensor, multi_scores: torch.Tensor, score_thr: float, iou_thr: float,
                   max_num: int = -1, nms_impl: str = "tv"):
    """mmdet 2.18 multiclass_nms + mmcv batched_nms (class_agnostic=False) [3P, unpinned]: candidates = every
    (RoI, class) with score > score_thr, flattened row-major; boxes shifted by label * (max coordinate + 1) so that
    one NMS call is class-aware; survivors in descending score order, first max_num.
    Returns (dets [D,5], labels [D], kept flat candidate indices r*N+n [D])."""
    num_classes = multi_scores.size(1) - 1
    bboxes = multi_bboxes.view(multi_scores.size(0), -1, 4)
    scores = multi_scores[:, :-1]
    labels = torch.arange(num_classes, dtype=torch.long).view(1, -1).expand_as(scores)
    flat = torch.arange(scores.numel(), dtype=torch.long).view_as(scores)
    bboxes, scores, labels, flat = bboxes.reshape(-1, 4), scores.reshape(-1), labels.reshape(-1), flat.reshape(-1)
    inds = (scores > score_thr).nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels, flat = bboxes[inds], scores[inds], labels[inds], flat[inds]
    if bboxes.numel() == 0:
        return multi_bboxes.new_zeros((0, 5)), torch.zeros((0,), dtype=torch.long), torch.zeros((0,), dtype=torch.long)
    max_coordinate = bboxes.max()
    offsets = labels.to(bboxes) * (max_coordinate + torch.tensor(1).to(bboxes))
    boxes_for_nms = bboxes + offsets[:, None]
    if nms_impl == "tv":
        import torchvision
        keep = torchvision.ops.nms(boxes_for_nms, scores, iou_thr)
    else:
        keep = nms_greedy(boxes_for_nms, scores, iou_thr)
    if max_num > 0:
        keep = keep[:max_num]
    return torch.cat([bboxes[keep], scores[keep, None]], -1), labels[keep], flat[keep]


def bbox_head_get_bboxes(rois: torch.Tensor, cls_score: torch.Tensor, bbox_pred: torch.Tensor, img_shape=None,
                         scale_factor=None, rescale: bool = False, score_thr: float = 0.05, iou_thr: float = 0.5,
                         max_per_img: int = 100, means=(0., 0., 0., 0.), stds=(0.1, 0.1, 0.2, 0.2), nms_impl: str = "tv"):
    """mmdet 2.18 BBoxHead.get_bboxes [3P, unpinned] for ONE image, as FGNBBoxHead.get_bboxes forwards to it
    (fgn_roi_head.py:170-178) with the config's bbox_coder (fgn_r50_c4_densecl.py:91-94) and test_cfg.rcnn (:181-185)."""
    scores = F.softmax(cls_score, dim=-1)
    bboxes = delta2bbox(rois[:, 1:], bbox_pred, means, stds, max_shape=img_shape)
    if rescale and bboxes.size(0) > 0:
        sf = bboxes.new_tensor(scale_factor)
        bboxes = (bboxes.view(bboxes.size(0), -1, 4) / sf).view(bboxes.size(0), -1)
    return multiclass_nms(bboxes, scores, score_thr, iou_thr, max_per_img, nms_impl)


# ---- RPN proposals: RPNHead.get_bboxes [3P] (called at fgn.py:229-235) ------------------------------------------
def anchor_base(base_size: float, scales, ratios) -> torch.Tensor:
    """mmdet 2.18 AnchorGenerator.gen_single_level_base_anchors [3P, unpinned] (scale_major, center_offset 0)."""
    sc, ra = torch.tensor(list(scales), dtype=torch.float32), torch.tensor(list(ratios), dtype=torch.float32)
    h_ratios = torch.sqrt(ra)
    w_ratios = 1 / h_ratios
    ws = (float(base_size) * w_ratios[:, None] * sc[None, :]).view(-1)
    hs = (float(base_size) * h_ratios[:, None] * sc[None, :]).view(-1)
    return torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], dim=-1)


def anchor_grid(base: torch.Tensor, h: int, w: int, stride: int) -> torch.Tensor:
    """mmdet 2.18 AnchorGenerator.single_level_grid_priors [3P]: [(y*W + x)*A + a, 4]."""
    shift_x = torch.arange(0, w, dtype=torch.float32) * stride
    shift_y = torch.arange(0, h, dtype=torch.float32) * stride
    xx = shift_x.repeat(h)
    yy = shift_y.view(-1, 1).repeat(1, w).view(-1)
    shifts = torch.stack([xx, yy, xx, yy], dim=-1)
    return (base[None, :, :] + shifts[:, None, :]).view(-1, 4)


def rpn_get_bboxes_single(cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor],
                          mlvl_anchors: Sequence[torch.Tensor], img_shape, nms_pre: int = 6000, iou_thr: float = 0.7,
                          max_per_img: int = 300, min_bbox_size: float = 0.0, means=(0., 0., 0., 0.),
                          stds=(1., 1., 1., 1.)):
    """mmdet 2.18 RPNHead._get_bboxes_single [3P, unpinned], sigmoid scores, ONE image: cls_scores[l] [A,H,W],
    bbox_preds[l] [4A,H,W].  The pre-NMS sort is made deterministic (stable, by logit: a refinement of the
    reference's unstable sort by score).  Returns (dets [D,5], level ids [D])."""
    import torchvision
    lv_scores, lv_preds, lv_anchors, lv_ids = [], [], [], []
    for idx, (c, r, anchors) in enumerate(zip(cls_scores, bbox_preds, mlvl_anchors)):
        logit = c.permute(1, 2, 0).reshape(-1)
        pred = r.permute(1, 2, 0).reshape(-1, 4)
        order = torch.sort(logit, descending=True, stable=True).indices
        if nms_pre > 0 and logit.shape[0] > nms_pre:
            order = order[:nms_pre]
        lv_scores.append(logit[order].sigmoid()); lv_preds.append(pred[order]); lv_anchors.append(anchors[order])
        lv_ids.append(torch.full((order.shape[0],), idx, dtype=torch.long))
    scores, anchors, preds, ids = torch.cat(lv_scores), torch.cat(lv_anchors), torch.cat(lv_preds), torch.cat(lv_ids)
    proposals = delta2bbox(anchors, preds, means, stds, max_shape=img_shape)
    if min_bbox_size >= 0:
        w, h = proposals[:, 2] - proposals[:, 0], proposals[:, 3] - proposals[:, 1]
        valid = (w > min_bbox_size) & (h > min_bbox_size)
        proposals, scores, ids = proposals[valid], scores[valid], ids[valid]
    if proposals.numel() == 0:
        return proposals.new_zeros((0, 5)), ids
    offsets = ids.to(proposals) * (proposals.max() + torch.tensor(1).to(proposals))
    keep = torchvision.ops.nms(proposals + offsets[:, None], scores, iou_thr)[:max_per_img]
    return torch.cat([proposals[keep], scores[keep, None]], -1), ids[keep]


# ------------------------------------------------------------------------------------------------------
# Mask pasting + RLE (second part): FCNMaskHead.get_seg_masks / _do_paste_mask [3P, mmdet
# 2.18] as called from fgn_roi_head.py:668-671, then mmdet.core.encode_mask_results -> pycocotools
# mask.encode [3P] as called from fgn.py:281.  mmdet and pycocotools are absent here ("parity unpinned" for
# the RLE string); the paste is pinned to torch CPU F.grid_sample (the op _do_paste_mask calls) in
# tests/test_oracle.py, the RLE to its own decoder and to run-length counts computed independently.
# ------------------------------------------------------------------------------------------------------
def paste_grid_coords(n_pix: int, b0: np.ndarray, b1: np.ndarray) -> np.ndarray:
    """_do_paste_mask: ``(arange(n) + 0.5 - b0) / (b1 - b0) * 2 - 1`` in fp32, inf -> 0.  [D, n_pix]."""
    p = np.arange(n_pix, dtype=np.float32)[None, :] + np.float32(0.5)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (p - b0[:, None].astype(np.float32)) / (b1 - b0)[:, None].astype(np.float32) * np.float32(2) - np.float32(1)
    g = g.astype(np.float32)
    g[np.isinf(g)] = 0
    return g


def paste_values(mask_logits: np.ndarray, boxes: np.ndarray, img_h: int, img_w: int) -> np.ndarray:
    """sigmoid + F.grid_sample(bilinear, zeros padding, align_corners=False) of every [M,M] mask over the whole
    image (skip_empty=False: the reference runs on the GPU), in the op order of torch's CUDA grid sampler:
    ``ix = ((g + 1) * M - 1) / 2``; weights ``nw = (ix_se - ix)(iy_se - iy)`` ...; ``out = nw_v*nw + ne_v*ne +
    sw_v*sw + se_v*se`` left to right, no FMA.  Returns [D, img_h, img_w] fp32."""
    f32 = np.float32
    m = np.asarray(mask_logits, dtype=f32)
    d, mm = m.shape[0], m.shape[-1]
    m = m.reshape(d, mm, mm)
    sig = (f32(1) / (f32(1) + np.exp(-m, dtype=f32))).astype(f32)
    boxes = np.asarray(boxes, dtype=f32)
    gx = paste_grid_coords(img_w, boxes[:, 0], boxes[:, 2])
    gy = paste_grid_coords(img_h, boxes[:, 1], boxes[:, 3])
    out = np.zeros((d, img_h, img_w), dtype=f32)
    for i in range(d):
        ix = ((gx[i] + f32(1)) * f32(mm) - f32(1)) / f32(2)
        iy = ((gy[i] + f32(1)) * f32(mm) - f32(1)) / f32(2)
        with np.errstate(invalid="ignore"):
            okx = (ix > -1) & (ix < mm)
            oky = (iy > -1) & (iy < mm)
        ixc = np.where(okx, ix, f32(0)).astype(f32)
        iyc = np.where(oky, iy, f32(0)).astype(f32)
        x0 = np.floor(ixc)
        y0 = np.floor(iyc)
        wx1 = (ixc - x0).astype(f32)                 # ix - ix_nw
        wx0 = ((x0 + f32(1)) - ixc).astype(f32)      # ix_se - ix
        wy1 = (iyc - y0).astype(f32)
        wy0 = ((y0 + f32(1)) - iyc).astype(f32)
        x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
        pad = np.zeros((mm + 2, mm + 2), dtype=f32)
        pad[1:-1, 1:-1] = sig[i]
        ya, yb = y0i + 1, y0i + 2                    # rows iy_nw, iy_sw in the padded map
        xa, xb = x0i + 1, x0i + 2
        nw = pad[np.ix_(ya, xa)] * (wy0[:, None] * wx0[None, :]).astype(f32)
        ne = pad[np.ix_(ya, xb)] * (wy0[:, None] * wx1[None, :]).astype(f32)
        sw = pad[np.ix_(yb, xa)] * (wy1[:, None] * wx0[None, :]).astype(f32)
        se = pad[np.ix_(yb, xb)] * (wy1[:, None] * wx1[None, :]).astype(f32)
        v = (((nw + ne).astype(f32) + sw).astype(f32) + se).astype(f32)
        v[~oky, :] = 0
        v[:, ~okx] = 0
        out[i] = v
    return out


def get_seg_masks(mask_logits, boxes, img_h: int, img_w: int, mask_thr_binary: float = 0.5) -> np.ndarray:
    """FCNMaskHead.get_seg_masks [3P] for a class-agnostic single-class mask head (fgn_r50_c4_densecl.py:123,127;
    labels are forced to 0, fgn_roi_head.py:716): ``paste >= thr`` as bool [D, img_h, img_w]."""
    return paste_values(mask_logits, boxes, img_h, img_w) >= np.float32(mask_thr_binary)


def rle_counts(mask: np.ndarray) -> List[int]:
    """pycocotools rleEncode [3P]: run lengths of the column-major (Fortran) scan, starting with a run of zeros
    (possibly empty)."""
    flat = np.asarray(mask, dtype=np.uint8).reshape(mask.shape[0], mask.shape[1]).T.reshape(-1)   # column-major
    if flat.size == 0:
        return []          # (pycocotools emits no run at all for an empty mask)
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    pos = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(pos).tolist()
    if flat[0] == 1:
        counts = [0] + counts
    return [int(c) for c in counts]


def rle_to_string(counts: Sequence[int]) -> bytes:
    """pycocotools rleToString [3P]: counts beyond the second are stored as the difference to the count two
    back; each value as 5-bit groups, low first, bit 5 = continuation, sign-extended stop rule, + 48."""
    s = bytearray()
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            s.append(ch + 48)
    return bytes(s)


def rle_from_string(s: bytes) -> List[int]:
    """pycocotools rleFrString [3P] (inverse of rle_to_string)."""
    counts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def rle_decode(counts: Sequence[int], h: int, w: int) -> np.ndarray:
    """pycocotools rleDecode [3P]: bool [h, w]."""
    flat = np.zeros(h * w, dtype=bool)
    pos, v = 0, False
    for c in counts:
        flat[pos:pos + c] = v
        pos += c
        v = not v
    return flat.reshape(w, h).T


def encode_mask_results(masks: np.ndarray) -> List[dict]:
    """mmdet.core.encode_mask_results [3P] for one class list (fgn.py:281): a COCO RLE dict per mask."""
    return [dict(size=[int(m.shape[0]), int(m.shape[1])], counts=rle_to_string(rle_counts(m))) for m in masks]
