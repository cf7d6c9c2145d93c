"""Attention-Guided RPN head -- host-side mirror of AGRPNHead (fgn_ag_rpn_head.py:14-118).

The attention part (class vectors :37-41, channel attention :44-46) and the best-class selection
(:87-108) run in libfgn_b200 kernels.  The RPN convolutions in between are mmdet RPNHead's [3P]
3x3 conv + two 1x1 convs (cuDNN through torch.nn); they are adjacent to the path, not part of it
(SURVEY section 8f).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class AGRPNHead(nn.Module):
    n_ways = 3
    k_shots = 3
    verbose = False
    subsampling_ratio = 16
    log_mode = False

    def __init__(self, in_channels: int = 1024, feat_channels: int = 1024, anchor_generator: Optional[dict] = None,
                 num_anchors: Optional[int] = None, loss_cls: Optional[dict] = None, n_ways: Optional[int] = None,
                 k_shots: Optional[int] = None, **kwargs):
        super().__init__()
        if num_anchors is None:
            ag = anchor_generator or dict(scales=[2, 4, 8, 16, 32], ratios=[0.5, 1.0, 2.0])
            num_anchors = len(ag["scales"]) * len(ag["ratios"])          # fgn_r50_c4_densecl.py:48-54
        use_sigmoid = True if loss_cls is None else loss_cls.get("use_sigmoid", True)
        if not use_sigmoid:
            raise NotImplementedError("AGRPNHead: the FGN configs use sigmoid objectness (one score per anchor)")
        self.in_channels, self.feat_channels, self.num_anchors = in_channels, feat_channels, num_anchors
        self.cls_out_channels = 1
        # mmdet RPNHead._init_layers [3P]
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, num_anchors * self.cls_out_channels, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, num_anchors * 4, 1)
        if n_ways is not None:
            self.n_ways = n_ways
        if k_shots is not None:
            self.k_shots = k_shots
        # proposal generation (get_bboxes): anchors, coder and test_cfg.rpn of the config (fgn_r50_c4_densecl.py:48-58,175-180)
        ag = anchor_generator or dict(scales=[2, 4, 8, 16, 32], ratios=[0.5, 1.0, 2.0], strides=[16])
        self.anchor_scales, self.anchor_ratios = list(ag["scales"]), list(ag["ratios"])
        self.anchor_strides = list(ag.get("strides", [16]))
        coder = dict(kwargs.get("bbox_coder") or {})
        self.bbox_coder = dict(target_means=list(coder.get("target_means", [0., 0., 0., 0.])),
                               target_stds=list(coder.get("target_stds", [1., 1., 1., 1.])))
        self.test_cfg = kwargs.get("test_cfg")
        # True: the channel attention is folded into rpn_conv's weights (qry_fmap_mod is never materialised);
        # False (default): the reference's own order of operations
        #   "auto": fold where it moves fewer bytes, (1+N)*H*W > N*kh*kw*feat_channels (large maps: FPN P2-P4)
        fa = kwargs.get("fold_attention", False)
        self.fold_attention = "auto" if fa == "auto" else bool(fa)
        # mmdet's RPNHead.loss [3P] (anchor targets, assigner, sampler, loss_cls / loss_bbox) is not part of this package:
        # a callable with its signature (cls_scores, bbox_preds, gt_bboxes, img_metas) -> dict can be plugged in here
        self.loss_fn = kwargs.get("loss_fn")

    def get_bboxes(self, cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor], img_metas=None,
                   cfg: Optional[dict] = None, rescale: bool = False):
        """mmdet RPNHead.get_bboxes [3P] as the reference calls it (fgn.py:229-235): per-level score / delta maps
        (already best-class-selected) -> per-image proposal tensors [D,5] (x1,y1,x2,y2,score), D <= max_per_img.
        One C-ABI call (ops.rpn_proposals -> fgn_rpn_proposals)."""
        cfg = cfg if cfg is not None else self.test_cfg
        if cfg is None:
            raise ValueError("AGRPNHead.get_bboxes needs cfg (test_cfg.rpn: nms_pre, nms.iou_threshold, max_per_img, min_bbox_size)")
        if rescale:
            raise NotImplementedError("RPN proposals stay at the test scale (the reference calls get_bboxes without rescale)")
        nl = len(cls_scores)
        strides = self.anchor_strides if len(self.anchor_strides) == nl else self.anchor_strides[:1] * nl
        anchors = torch.stack([ops.base_anchors(s, self.anchor_scales, self.anchor_ratios) for s in strides])
        shapes = [m["img_shape"] for m in img_metas] if img_metas is not None else None
        prop, _, cnt = ops.rpn_proposals(cls_scores, bbox_preds, strides, anchors, shapes, nms_pre=cfg["nms_pre"],
                                         iou_thr=cfg["nms"]["iou_threshold"], max_per_img=cfg["max_per_img"],
                                         min_bbox_size=cfg.get("min_bbox_size", 0),
                                         means=self.bbox_coder["target_means"], stds=self.bbox_coder["target_stds"])
        return [prop[i, :c] for i, c in enumerate(cnt.tolist())]       # ragged per-image lists: one host sync

    # mmdet RPNHead.forward_single [3P]
    def _rpn_forward_single(self, x: torch.Tensor):
        x = F.relu(self.rpn_conv(x), inplace=True)
        return self.rpn_cls(x), self.rpn_reg(x)

    def attention(self, qry_fmap: torch.Tensor, spp_fmaps: torch.Tensor):
        """fgn_ag_rpn_head.py:33-46 -> (spp_fvecs_cat_mean [B,N,C,1,1], qry_fmap_mod [B*N,C,H,W])."""
        batch, c = qry_fmap.shape[:2]
        if spp_fmaps.shape[0] != batch * self.n_ways * self.k_shots or spp_fmaps.shape[1] != c:
            raise ops.FgnError(f"spp_fmaps {tuple(spp_fmaps.shape)} is not [B*N*K={batch * self.n_ways * self.k_shots}, C={c}, h, w]")
        from . import autograd as A                      # the adjoint kernels take over when autograd is recording
        vec = A.attention_vectors(spp_fmaps, self.n_ways, self.k_shots)
        return vec, A.channel_attention(qry_fmap, vec)

    def folded_rpn_conv(self, qry_fmap: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
        """relu(rpn_conv(qry_fmap_mod)) without qry_fmap_mod: conv(q * v) = conv(q, W * v) -- one grouped cuDNN conv
        on the unmodified query map with the B*N weight sets of ops.fold_attention_weights.  [B*N,Cf,H,W]."""
        b, c, h, w = qry_fmap.shape
        n = vec.shape[1]
        wf = ops.fold_attention_weights(self.rpn_conv.weight, vec)                   # [B*N,Cf,C,3,3]
        cf = wf.shape[1]
        bias = self.rpn_conv.bias.repeat(b * n) if self.rpn_conv.bias is not None else None
        x = F.conv2d(qry_fmap.reshape(1, b * c, h, w), wf.view(b * n * cf, c, *wf.shape[-2:]), bias,
                     padding=self.rpn_conv.padding, groups=b)
        return F.relu(x.view(b * n, cf, h, w), inplace=True)

    def folded_weights_multilevel(self, spp_feats: Sequence[torch.Tensor]):
        """FPN mode: the class vectors of every level and rpn_conv's weights folded with them -- everything the
        attention contributes when the RPN conv consumes the unmodified pyramid.  (vec [L,B*N,C], w' [L,B*N,Cf,C,3,3])."""
        vec = ops.attention_vectors_multilevel(spp_feats, self.n_ways, self.k_shots)
        wf = ops.fold_attention_weights(self.rpn_conv.weight, vec)
        return vec, wf.view(vec.shape[0], vec.shape[1], *wf.shape[1:])

    def attention_multilevel(self, qry_feats: Sequence[torch.Tensor], spp_feats: Sequence[torch.Tensor]):
        """``attention`` for every pyramid level at once (FPN mode; three launches for channels_last inputs)."""
        return ops.attention_multilevel(qry_feats, spp_feats, self.n_ways, self.k_shots)

    def forward_single(self, qry_fmap: torch.Tensor, spp_fmaps: Optional[torch.Tensor] = None, qry_bboxes=None,
                       qry_cat_ids=None, img_metas_cpu: Optional[list] = None, train_mode: bool = False,
                       log_mode: bool = False):
        assert train_mode ^ (qry_bboxes is None and qry_cat_ids is None)
        batch = qry_fmap.shape[0]
        fold = False if train_mode else self.fold_attention       # training keeps the reference's order (and its autograd graph)
        if fold == "auto":
            kk = self.rpn_conv.kernel_size[0] * self.rpn_conv.kernel_size[1]
            fold = (1 + self.n_ways) * qry_fmap.shape[2] * qry_fmap.shape[3] > self.n_ways * kk * self.feat_channels
        if fold and not log_mode:
            vec = ops.attention_vectors(spp_fmaps, self.n_ways, self.k_shots)
            x = self.folded_rpn_conv(qry_fmap, vec)
            rpn_cls_score, rpn_bbox_pred = self.rpn_cls(x), self.rpn_reg(x)
        else:
            _, qry_fmap_mod = self.attention(qry_fmap, spp_fmaps)
            rpn_cls_score, rpn_bbox_pred = self._rpn_forward_single(qry_fmap_mod)
        if log_mode:
            self.qry_fmap_mod, self.rpn_cls_score, self.rpn_bbox_pred = qry_fmap_mod, rpn_cls_score, rpn_bbox_pred
        rpn_losses: dict = {}
        if train_mode:
            # fgn_ag_rpn_head.py:58-79: one GT-box list per (image, class) row of the attended batch -- the boxes of class j
            # in image i for row i*N + j (empty [0,4] when the class is absent), the image's meta repeated N times
            gts, metas = [], []
            for i in range(batch):
                for j in range(self.n_ways):
                    idx = torch.where(qry_cat_ids[i] == j)[0]
                    gts.append(qry_bboxes[i][idx].view(-1, 4) if len(idx) != 0 else qry_bboxes[i][:0])
                    metas.append(img_metas_cpu[i] if img_metas_cpu is not None else None)
            loss_inputs = ([rpn_cls_score], [rpn_bbox_pred], gts, metas)
            if self.loss_fn is not None:
                rpn_losses = self.loss_fn(*loss_inputs)
                rpn_losses["loss_rpn_cls"][0] = rpn_losses["loss_rpn_cls"][0] / self.n_ways     # :77-78 "balancer"
                rpn_losses["loss_rpn_bbox"][0] = rpn_losses["loss_rpn_bbox"][0] / self.n_ways
            else:
                rpn_losses = dict(loss_inputs=loss_inputs, balancer=1.0 / self.n_ways)
        if self.n_ways > 1:
            # the selected maps feed proposal generation only (mmdet detaches proposals): no adjoint needed
            with torch.no_grad():
                out = ops.best_class_select(rpn_cls_score.detach(), rpn_bbox_pred.detach(), batch, self.n_ways)
        else:
            c, h, w = rpn_cls_score.shape[-3:]
            c4 = rpn_bbox_pred.shape[1]
            out = (rpn_cls_score.reshape(batch, c, h, w), rpn_bbox_pred.reshape(batch, c4, h, w))
        return (out[0], out[1], rpn_losses) if train_mode else out

    def forward(self, qry_feats: Sequence[torch.Tensor], spp_feats: Sequence[torch.Tensor]):
        """FPN mode (SURVEY A.9, A-FPN): forward_single per level with that level's support maps."""
        outs = [self.forward_single(q, s) for q, s in zip(qry_feats, spp_feats)]
        return [o[0] for o in outs], [o[1] for o in outs]
