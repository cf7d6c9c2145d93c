"""Relation-Guided Detector / Attention-Guided FCN RoI head -- host-side mirror of FGNRoIHead
(fgn_roi_head.py:181-719) for the forward hot path.

Method names, argument meaning, tensor layouts and the attributes the reference leaves on ``self``
(``spp_fmaps_roi_aligned_cat_mean``, ``spp_fvecs_roi_aligned_cat_mean_mp``, ``spp_vecs_mask``) are
kept.  All RoIAlign / fusion arithmetic runs in libfgn_b200; ``shared_head`` (C4 res5, cuDNN via
torch.nn) and ``mask_head`` are adjacent modules the head merely hosts.
Losses, samplers and box post-processing (mmdet [3P]) are not part of this path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from . import autograd as A
from . import ops
from .roi_extractor import SingleRoIExtractor, bbox2roi


class _Bottleneck(nn.Module):
    """mmdet Bottleneck [3P] with expansion=2 and no downsample (fgn_roi_head.py:202-233).

    At inference (eval mode, CUDA, autograd not recording) all three convolutions run on the library's tcgen05 kernels
    over the NHWC RoI tiles -- BatchNorm folded into the weights, ReLU and the identity branch in the
    epilogues: the 1x1s on the contraction kernel (ops.conv1x1), the 3x3 as an implicit GEMM whose taps are shifted,
    zero-filled TMA boxes of the same tensor (ops.conv3x3).  ``tc_1x1=False`` / ``tc_3x3=False`` (or training) run the plain
    torch modules.  ``tc_3x3="auto"`` (default): the library's 3x3 whenever the convolutions run in TF32
    (``torch.backends.cudnn.allow_tf32``, PyTorch's default and the reference's); under strict fp32 the 3x3 stays on cuDNN,
    because the tensor core's fp32 accumulator truncates -- over K = 9*512 products the 3xTF32 route ends 1.2e-3 (7e-5 of
    the output scale) from fp64 after three bottlenecks where cuDNN fp32 ends 1.6e-5 (tools/diag_c4_head.py); ``True``
    takes the 3xTF32 route anyway (5x faster than cuDNN fp32, profiles/r02_heads_tcgen05_vs_cudnn.jsonl)."""

    def __init__(self, inplanes: int, planes: int, tc_1x1: bool = True, tc_3x3="auto"):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, inplanes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.tc_1x1, self.tc_3x3 = tc_1x1, tc_3x3
        self._folded = None

    def train(self, mode: bool = True):
        self._folded = None                                          # running statistics / weights may change
        return super().train(mode)

    @staticmethod
    def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d):
        """eval-mode BatchNorm folded into the convolution: w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps)."""
        scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
        w = conv.weight.detach() * scale.view(-1, 1, 1, 1)
        b = bn.bias.detach() - bn.running_mean.detach() * scale
        if conv.bias is not None:
            b = b + conv.bias.detach() * scale
        return w.contiguous(), b.contiguous()

    def forward(self, x):
        use_tc = (self.tc_1x1 and not self.training and x.is_cuda and x.dtype == torch.float32 and
                  not (torch.is_grad_enabled() and (x.requires_grad or self.conv1.weight.requires_grad)))
        if not use_tc:
            out = self.relu(self.bn1(self.conv1(x)))
            out = self.relu(self.bn2(self.conv2(out)))
            out = self.bn3(self.conv3(out))
            return self.relu(out + x)
        prec = "tf32" if torch.backends.cudnn.allow_tf32 else "fp32"
        if self._folded is None or self._folded["device"] != x.device or self._folded["prec"] != prec:
            w1, b1 = self._fold(self.conv1, self.bn1)
            w2, b2 = self._fold(self.conv2, self.bn2)
            w3, b3 = self._fold(self.conv3, self.bn3)
            t2 = ops.conv_taps(w2)
            self._folded = dict(device=x.device, prec=prec, w1=w1.flatten(1), b1=b1, t2=t2, b2=b2, w3=w3.flatten(1), b3=b3,
                                s2=ops.conv_split_weights(t2) if (prec == "fp32" and self.tc_3x3 is True) else None)
        f = self._folded
        xc = x if ops.storage_layout(x) == ops.LAYOUT_NHWC else ops.to_nhwc(x.contiguous())
        out = ops.conv1x1(xc, f["w1"], f["b1"], relu=True)
        if self.tc_3x3 is True or (self.tc_3x3 == "auto" and prec == "tf32"):
            out = ops.conv3x3(out, f["t2"], f["b2"], relu=True, precision=prec, w_split=f["s2"])
        else:
            out = self.relu(self.bn2(self.conv2(out)))
        return ops.conv1x1(out, f["w3"], f["b3"], residual=xc, relu=True)


def make_c4_shared_head(inplanes: int = 1024, planes: int = 512, num_blocks: int = 3) -> nn.Module:
    head = nn.Sequential(*[_Bottleneck(inplanes, planes) for _ in range(num_blocks)])
    for m in head.modules():                                     # fgn_roi_head.py:224-231
        if isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
    return head


class FGNBBoxHead(nn.Module):
    """Forward part of FGNBBoxHead(BBoxHead) with with_avg_pool=True, num_classes=1,
    reg_class_agnostic=False (fgn_r50_c4_densecl.py:76-93): holds fc_cls [2,C] and fc_reg [4,C];
    the avg-pool + FCs themselves run fused inside fgn_relation_fusion_fwd."""
    n_ways = 3
    k_shots = 3

    def __init__(self, in_channels: int = 1024, roi_feat_size: int = 7, num_classes: int = 1,
                 with_avg_pool: bool = True, reg_class_agnostic: bool = False, **kwargs):
        super().__init__()
        if not with_avg_pool or num_classes != 1 or reg_class_agnostic:
            raise NotImplementedError("FGNBBoxHead: the FGN config uses with_avg_pool, num_classes=1, class-specific reg")
        self.in_channels, self.roi_feat_size = in_channels, roi_feat_size
        self.fc_cls = nn.Linear(in_channels, num_classes + 1)
        self.fc_reg = nn.Linear(in_channels, 4 * num_classes)
        nn.init.xavier_normal_(self.fc_cls.weight)               # init_cfg Xavier/normal for Linear
        nn.init.xavier_normal_(self.fc_reg.weight)
        nn.init.zeros_(self.fc_cls.bias)
        nn.init.zeros_(self.fc_reg.bias)
        # DeltaXYWHBBoxCoder of the config (fgn_r50_c4_densecl.py:91-94), used by the test-time decode
        coder = dict(kwargs.get("bbox_coder") or {})
        self.bbox_coder = dict(target_means=list(coder.get("target_means", [0., 0., 0., 0.])),
                               target_stds=list(coder.get("target_stds", [0.1, 0.1, 0.2, 0.2])))


class FGNRoIHead(nn.Module):
    n_ways = 3
    k_shots = 3
    subsampling_ratio = 16
    spp_fmaps_roi_aligned_cat_mean: torch.Tensor
    spp_fvecs_roi_aligned_cat_mean_mp: torch.Tensor
    spp_vecs_mask: torch.Tensor

    def __init__(self, bbox_roi_extractor: Optional[dict] = None, bbox_head: Optional[dict] = None,
                 mask_roi_extractor: Optional[dict] = None, mask_head: Union[None, dict, nn.Module] = None,
                 shared_head: Union[None, str, nn.Module] = "c4", channels: int = 1024, n_ways: Optional[int] = None,
                 k_shots: Optional[int] = None, mutate_inputs: bool = True, precision: str = "fp32",
                 train_cfg=None, test_cfg=None, **kwargs):
        """``shared_head='c4'`` reproduces the reference, which always builds its res5 ResLayer and
        ignores the config's ``shared_head=None`` (fgn_roi_head.py:197-200); pass ``None`` for FPN mode.
        ``channels`` replaces the hard-coded 1024/2048 of fgn_roi_head.py:212,241-243."""
        super().__init__()
        bre = bbox_roi_extractor or dict(type="SingleRoIExtractor",
                                         roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                         out_channels=channels, featmap_strides=[16])
        bre = {k: v for k, v in bre.items() if k != "type"}
        self.bbox_roi_extractor = SingleRoIExtractor(**bre)
        if mask_roi_extractor is not None:
            mre = {k: v for k, v in mask_roi_extractor.items() if k != "type"}
            self.mask_roi_extractor = SingleRoIExtractor(**mre)
            self.share_roi_extractor = False
        else:                                                      # mmdet shares the bbox extractor [3P]
            self.mask_roi_extractor = self.bbox_roi_extractor
            self.share_roi_extractor = True
        bh = dict(bbox_head or {})
        bh.pop("type", None)
        bh.setdefault("in_channels", channels)
        self.bbox_head = FGNBBoxHead(**bh)
        if isinstance(mask_head, dict):                            # config form (fgn_r50_c4_densecl.py:115-129)
            from .mask_head import FCNMaskHead
            mh = {k: v for k, v in mask_head.items() if k not in ("type", "loss_mask", "init_cfg", "norm_cfg")}
            mask_head = FCNMaskHead(**mh)
        self.mask_head = mask_head
        self.fused_prologue = True        # count_spp as one launch where no shared_head intervenes (False: the four separate ops)
        self.fused_prologue_max_supports = 32
        self._class_term = None
        if shared_head == "c4":
            self.shared_head = make_c4_shared_head(channels, channels // 2, 3)
        else:
            self.shared_head = shared_head
        self.channels = channels
        self.init_cls_reg_shared_conv()
        if n_ways is not None:
            self.n_ways = n_ways
        if k_shots is not None:
            self.k_shots = k_shots
        self.mutate_inputs = mutate_inputs
        self.precision = precision
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self._params_cache = None

    # ---- construction (fgn_roi_head.py:240-251) -------------------------------------------------
    def init_cls_reg_shared_conv(self):
        c = self.channels
        self.cls_reg_shared_conv = nn.Conv2d(2 * c, c, kernel_size=(1, 1), stride=(1, 1), padding=0)
        self.cls_reg_shared_conv_norm = nn.GroupNorm(num_groups=32, num_channels=c, affine=True)
        nn.init.kaiming_normal_(self.cls_reg_shared_conv.weight, nonlinearity="relu")
        nn.init.ones_(self.cls_reg_shared_conv_norm.weight)
        nn.init.zeros_(self.cls_reg_shared_conv_norm.bias)

    @property
    def with_shared_head(self) -> bool:
        return getattr(self, "shared_head", None) is not None

    @property
    def with_bbox(self) -> bool:
        return self.bbox_head is not None

    @property
    def with_mask(self) -> bool:
        return self.mask_head is not None

    def shared_head_layer(self, x: torch.Tensor) -> torch.Tensor:
        return self.shared_head(x)

    def relation_params(self, refresh: bool = False) -> ops.RelationParams:
        """Packed device copies of the relation-head weights (re-packed when ``refresh`` or in training)."""
        if self._params_cache is None or refresh or self.training:
            n = self.cls_reg_shared_conv_norm
            self._params_cache = ops.RelationParams(
                self.cls_reg_shared_conv.weight, self.cls_reg_shared_conv.bias, n.weight, n.bias,
                self.bbox_head.fc_cls.weight, self.bbox_head.fc_cls.bias,
                self.bbox_head.fc_reg.weight, self.bbox_head.fc_reg.bias, n.num_groups, n.eps)
        return self._params_cache

    def _valid_class_term(self, params) -> Optional[torch.Tensor]:
        """count_spp's precomputed class half of the relation conv, if it still belongs to the class maps on ``self`` and to
        the weights about to be used (someone may have assigned ``spp_fmaps_roi_aligned_cat_mean`` by hand since)."""
        ct = self._class_term
        if ct is None or ct[1] is not self.spp_fmaps_roi_aligned_cat_mean or ct[2] is not params:
            return None
        return ct[0]

    def _recording(self, *tensors) -> bool:
        """True when autograd is recording and any of `tensors` or of this head's parameters requires grad."""
        if not torch.is_grad_enabled():
            return False
        flat = []
        for t in tensors:
            flat.extend(t if isinstance(t, (list, tuple)) else [t])
        return A._needs_grad(*flat) or any(p.requires_grad for p in self.parameters())

    @staticmethod
    def _as_levels(fmap) -> List[torch.Tensor]:
        # the reference emulates mmdet's tuple-of-levels with unsqueeze(0) (fgn_roi_head.py:330,365)
        return list(fmap) if isinstance(fmap, (list, tuple)) else [fmap]

    # ---- support branch (fgn_roi_head.py:419-449) -----------------------------------------------
    def count_spp(self, spp_fmaps, spp_bboxes: torch.Tensor, spp_isegmaps: torch.Tensor):
        """
        :param spp_fmaps: [B*N*K, C, H, W] (C4) or list of such levels (FPN mode)
        :param spp_bboxes: [B*N*K, 1, 4] XYXY px; divided in place by subsampling_ratio in C4 mode
        :param spp_isegmaps: [B*N*K, 1, H, W] bool
        """
        levels = self._as_levels(spp_fmaps)
        m = spp_bboxes.shape[0]
        # training (fgn_roi_head.py:451-529 reaches count_spp with a backbone that may require grad): the adjoint kernels
        # of fgn_b200.autograd take over wherever autograd is recording
        grad = A._needs_grad(*levels) or (self.with_shared_head and A._needs_grad(*self.shared_head.parameters()))
        roi_align = A.roi_align_multilevel if grad else ops.roi_align_multilevel
        self._class_term = None
        # (the one-launch form is latency-optimal for a handful of supports -- 1-way 1-shot: 50 -> 25 us, 4 launches -> 1;
        #  with a hundred of them (cfg4: N=20, K=5) its (class, bin) CTAs loop over the shots at two CTAs per SM and the four
        #  wide kernels win: 792 against 711 us per episode, profiles/r02_configs.jsonl -- hence the size rule)
        if (not grad and not self.with_shared_head and self.fused_prologue and spp_isegmaps.dtype in (torch.bool, torch.uint8)
                and (m <= self.fused_prologue_max_supports or self.fused_prologue == "always")
                and all(l.dtype == torch.float32 for l in levels)):               # (the bf16 variant keeps its own kernels)
            # inference without a shared_head between RoIAlign and the class mean: all of count_spp -- and the class half
            # of the relation convolution -- in ONE launch (fgn_support_prologue_fwd)
            ext = self.bbox_roi_extractor
            if len(levels) == 1:
                scales = [1.0 / self.subsampling_ratio]       # = boxes / 16 with scale 1 (:430-432): a power of two, same cells
            else:
                scales = [1.0 / s for s in ext.featmap_strides[: len(levels)]]
            pr = self.relation_params()
            cat_mean, mp, term = ops.support_prologue(levels, scales, spp_bboxes.reshape(m, 4), spp_isegmaps, self.n_ways,
                                                      self.k_shots, 7, float(ext.finest_scale), pr.conv_w, pr.conv_b)
            if len(levels) == 1 and self.mutate_inputs:
                spp_bboxes /= self.subsampling_ratio                                             # :430 (the reference's side effect)
            self.spp_fmaps_roi_aligned_cat_mean = cat_mean
            self.spp_fvecs_roi_aligned_cat_mean_mp = mp
            self._class_term = (term, cat_mean, pr)           # valid for exactly these class maps and these weights
            return
        mask_ra = ops.support_mask_pool(spp_isegmaps, spp_bboxes.reshape(m, 4), 7)          # :429 (masks carry no grad)
        idx = torch.arange(m, device=spp_bboxes.device, dtype=torch.float32).view(m, 1)
        # without a shared_head the class maps feed the relation GEMM directly: keep them channels_last
        # channels_last storage everywhere (logical shapes stay the reference's NCHW): the relation GEMM reads it directly,
        # and the res5 shared_head (tcgen05 1x1 convs + a cuDNN 3x3) prefers it -- so C4 mode runs the window kernel too
        fmt = "nhwc"
        if len(levels) == 1:
            if self.mutate_inputs:
                spp_bboxes /= self.subsampling_ratio                                             # :430
                boxes = spp_bboxes
            else:
                boxes = spp_bboxes / self.subsampling_ratio
            rois = torch.cat([idx, boxes.reshape(m, 4).float()], 1)
            feat_ra = roi_align(levels, rois, [1.0], 7, -1, aligned=False, out_format=fmt,
                                out_dtype=torch.float32)                                         # :432
        else:   # A-FPN: level from map_roi_levels on the pixel box, spatial_scale = 1/stride
            rois = torch.cat([idx, spp_bboxes.reshape(m, 4).float()], 1)
            ext = self.bbox_roi_extractor
            feat_ra = roi_align(levels, rois, [1.0 / s for s in ext.featmap_strides[: len(levels)]],
                                7, -1, aligned=False, finest_scale=float(ext.finest_scale), out_format=fmt,
                                out_dtype=torch.float32)
        if self.with_shared_head:
            feat_ra = self.shared_head_layer(feat_ra)                                            # :435-436
        cat_mean, mp = (A.support_pool if A._needs_grad(feat_ra) else ops.support_pool)(
            feat_ra, mask_ra, self.n_ways, self.k_shots)                                         # :439-447
        if cat_mean.dim() == 4:                                                                  # (autograd path: [B*N,C,P,P])
            c = cat_mean.shape[1]
            cat_mean = cat_mean.unflatten(0, (-1, self.n_ways))
            mp = mp.reshape(-1, self.n_ways, c, 1, 1)
        self.spp_fmaps_roi_aligned_cat_mean = cat_mean
        self.spp_fvecs_roi_aligned_cat_mean_mp = mp
        return

    # ---- relation-guided detector (fgn_roi_head.py:253-342) ------------------------------------
    def count_modified_cls_bbox(self, rois_amount: int, cls_score: torch.Tensor, bbox_pred: torch.Tensor):
        return ops.cls_bbox_reassemble(cls_score, bbox_pred, rois_amount, self.n_ways)

    def _bbox_forward(self, qry_fmap, rois: torch.Tensor, need_feats: Optional[bool] = None):
        """Box head forward (fgn_roi_head.py:328-342) -> dict(cls_score [R,N+1], bbox_pred [R,4N], bbox_feats).

        ``bbox_feats`` is materialised when a ``shared_head`` sits between RoIAlign and the fusion (C4),
        or when ``need_feats`` (default: ``self.training``, because the train-time mask branch re-reads
        it, :373-374); otherwise the FPN single-pass entry point is used and ``bbox_feats`` is None.
        """
        levels = self._as_levels(qry_fmap)[: self.bbox_roi_extractor.num_inputs]
        need_feats = self.training if need_feats is None else need_feats
        params = self.relation_params()
        if rois.shape[0] == 0:
            z = rois.new_zeros
            return dict(cls_score=z((0, self.n_ways + 1)), bbox_pred=z((0, 4 * self.n_ways)), bbox_feats=None)
        ext = self.bbox_roi_extractor
        layer = ext.roi_layers[0]
        if self._recording(levels, self.spp_fmaps_roi_aligned_cat_mean):
            # autograd is recording and something upstream (backbone maps, class maps, the shared_head or the relation
            # head's own parameters) requires grad: the differentiable wrappers (fgn_b200.autograd), RoI features materialised
            bbox_feats = ext(levels, rois, out_format="nhwc")
            if self.with_shared_head:
                bbox_feats = self.shared_head_layer(bbox_feats)
            n_ = self.cls_reg_shared_conv_norm
            cls, reg = A.relation_fusion(bbox_feats, rois[:, 0], self.spp_fmaps_roi_aligned_cat_mean, self.n_ways,
                                         self.cls_reg_shared_conv.weight, self.cls_reg_shared_conv.bias, n_.weight, n_.bias,
                                         self.bbox_head.fc_cls.weight, self.bbox_head.fc_cls.bias,
                                         self.bbox_head.fc_reg.weight, self.bbox_head.fc_reg.bias, n_.num_groups, n_.eps)
            return dict(cls_score=cls, bbox_pred=reg, bbox_feats=bbox_feats)
        if not self.with_shared_head and not need_feats and layer.output_size[0] == 7:
            cls, reg = ops.guided_roi_fused(levels, rois, [l.spatial_scale for l in ext.roi_layers][: len(levels)],
                                            self.spp_fmaps_roi_aligned_cat_mean, self.n_ways, params, 7,
                                            layer.sampling_ratio, layer.aligned, float(ext.finest_scale), self.precision,
                                            class_term=self._valid_class_term(params))
            return dict(cls_score=cls, bbox_pred=reg, bbox_feats=None)
        bbox_feats = ext(levels, rois, out_format="nhwc")
        if self.with_shared_head:
            bbox_feats = self.shared_head_layer(bbox_feats)
        cls, reg = ops.relation_fusion(bbox_feats, rois[:, 0], self.spp_fmaps_roi_aligned_cat_mean, self.n_ways,
                                       params, self.precision, class_term=self._valid_class_term(params))
        return dict(cls_score=cls, bbox_pred=reg, bbox_feats=bbox_feats)

    # ---- attention-guided FCN (fgn_roi_head.py:360-382) -----------------------------------------
    def gather_mask_vectors(self, labels: Sequence[torch.Tensor]) -> torch.Tensor:
        """Vector gather of simple_test (:707-714) / forward_train (:516-522): label + N * image index."""
        gather = torch.cat([labels[i] + self.n_ways * i for i in range(len(labels))])
        batch, n, c = self.spp_fvecs_roi_aligned_cat_mean_mp.shape[:3]
        assert batch == len(labels) and n == self.n_ways
        self._spp_vec_index = gather.to(torch.int32)
        self.spp_vecs_mask = self.spp_fvecs_roi_aligned_cat_mean_mp.view(batch * self.n_ways, c, 1, 1)[gather]
        return self.spp_vecs_mask

    def _mask_forward(self, qry_fmap, rois=None, pos_inds=None, bbox_feats=None):
        assert (rois is not None) ^ (pos_inds is not None and bbox_feats is not None)
        vec = self.spp_vecs_mask
        if rois is not None:
            levels = self._as_levels(qry_fmap)[: self.mask_roi_extractor.num_inputs]
            if not self.with_shared_head and A._needs_grad(*levels, vec):
                # training: the fused multiply has no adjoint of its own -- RoIAlign and the attention as two
                # differentiable steps
                mask_feats = self.mask_roi_extractor(levels, rois, out_format="nhwc")
                mask_feats = A.channel_attention(mask_feats, vec.reshape(vec.shape[0], 1, -1, 1, 1))
            elif not self.with_shared_head:
                # RoIAlign with the channel attention multiply in its epilogue (K12 fused into K1)
                # channels_last storage (logical shape unchanged): the fast RoIAlign kernel, and the layout cuDNN
                # prefers for the mask head's convs
                mask_feats = self.mask_roi_extractor(levels, rois, chan_scale=vec.reshape(vec.shape[0], -1),
                                                     out_format="nhwc")
            else:
                mask_feats = self.shared_head(self.mask_roi_extractor(levels, rois, out_format="nhwc"))
                mask_feats = A.channel_attention(mask_feats, vec.reshape(vec.shape[0], 1, -1, 1, 1))
        else:
            pos_inds_new = torch.nonzero(pos_inds).view(-1)
            mask_feats = A.channel_attention(bbox_feats[pos_inds_new], vec.reshape(vec.shape[0], 1, -1, 1, 1))
        assert mask_feats.shape[:2] == vec.shape[:2]
        mask_pred = self.mask_head(mask_feats) if self.mask_head is not None else None
        return dict(mask_pred=mask_pred, mask_feats=mask_feats)

    # ---- test-time driver (fgn_roi_head.py:531-616, 675-719) ----
    def simple_test_bboxes(self, x, img_metas, proposals: Sequence[torch.Tensor], rcnn_test_cfg=None, rescale=False):
        """fgn_roi_head.py:531-616.  With ``rcnn_test_cfg`` (dict with ``score_thr``, ``nms=dict(iou_threshold=..)``,
        ``max_per_img``, as fgn_r50_c4_densecl.py:181-185) and ``img_metas`` (``img_shape``, ``scale_factor``):
        per-image ``(det_bboxes [D,5], det_labels [D] long)`` lists, i.e. bbox_head.get_bboxes [3P] on the device
        (ops.det_postprocess).  With ``rcnn_test_cfg=None``: the raw per-image (cls_score, bbox_pred) splits."""
        rois = bbox2roi(proposals)
        n_per = tuple(len(p) for p in proposals)
        if rois.shape[0] == 0:                                   # reference :558-567
            if rcnn_test_cfg is None:
                return [rois.new_zeros((0, 4))] * len(proposals), [rois.new_zeros((0, self.n_ways + 1))] * len(proposals)
            return ([rois.new_zeros((0, 5))] * len(proposals),
                    [rois.new_zeros((0,), dtype=torch.long)] * len(proposals))
        res = self._bbox_forward(x, rois, need_feats=False)
        if rcnn_test_cfg is None:
            return res["cls_score"].split(n_per, 0), res["bbox_pred"].split(n_per, 0)
        cfg = rcnn_test_cfg.get("rcnn", rcnn_test_cfg) if isinstance(rcnn_test_cfg, dict) else rcnn_test_cfg
        shapes = [m["img_shape"] for m in img_metas] if img_metas is not None else None
        scales = [m["scale_factor"] for m in img_metas] if (rescale and img_metas is not None) else None
        coder = self.bbox_head.bbox_coder
        det, lab, cnt = ops.det_postprocess(rois, res["cls_score"], res["bbox_pred"], n_per, shapes, scales,
                                            score_thr=cfg["score_thr"], iou_thr=cfg["nms"]["iou_threshold"],
                                            max_per_img=cfg["max_per_img"], means=coder["target_means"],
                                            stds=coder["target_stds"])
        counts = cnt.tolist()                                    # the reference's outputs are ragged: one host sync
        return ([det[i, :c] for i, c in enumerate(counts)], [lab[i, :c].long() for i, c in enumerate(counts)])

    def simple_test_mask(self, x, det_bboxes: Sequence[torch.Tensor], det_labels: Sequence[torch.Tensor],
                         img_metas=None, rescale: bool = False):
        """fgn_roi_head.py:618-673 up to the mask head's input: support vectors gathered by (image, label)
        (:707-714), boxes scaled back to the test scale when they were rescaled (:642-652), RoIAlign with the
        AG-FCN multiply fused.  ``get_seg_masks`` (mask pasting) is mmdet's and stays downstream."""
        self.gather_mask_vectors(det_labels)
        boxes = [d[:, :4] for d in det_bboxes]
        if rescale and img_metas is not None:
            boxes = [bx * bx.new_tensor(m["scale_factor"]) for bx, m in zip(boxes, img_metas)]
        mask_rois = bbox2roi(boxes)
        if mask_rois.shape[0] == 0:
            return dict(mask_pred=None, mask_feats=mask_rois.new_zeros((0, self.channels, 7, 7)))
        res = self._mask_forward(x, mask_rois)
        if res["mask_pred"] is not None and img_metas is not None:
            # get_seg_masks + encode_mask_results (fgn_roi_head.py:668-671, fgn.py:281) for all images in one
            # launch: per image the list of COCO RLE dicts of its detections (class-agnostic head: class 0)
            res["segm_rles"] = self.get_seg_rles(res["mask_pred"], mask_rois, [len(bx) for bx in boxes], img_metas, rescale)
        return res

    def get_seg_rles(self, mask_pred: torch.Tensor, mask_rois: torch.Tensor, num_per_img: Sequence[int], img_metas,
                     rescale: bool = False) -> List[List[dict]]:
        """Batched get_seg_masks(..., encode=True): ``mask_rois`` [D,5] at test scale (bbox2roi of ``_bboxes``)."""
        cfg = self.test_cfg or {}
        cfg = cfg.get("rcnn", cfg) if isinstance(cfg, dict) else cfg
        thr = float(cfg.get("mask_thr_binary", 0.5))
        sfs = [[float(v) for v in m.get("scale_factor", (1.0, 1.0, 1.0, 1.0))] for m in img_metas]
        hw = []
        for m, sf in zip(img_metas, sfs):
            oh, ow = m.get("ori_shape", m.get("img_shape"))[:2]
            hw.append((int(oh), int(ow)) if rescale else
                      (int(np.round(oh * sf[1]).astype(np.int32)), int(np.round(ow * sf[0]).astype(np.int32))))
        det_img = mask_rois[:, 0].to(torch.int32)
        boxes = mask_rois[:, 1:5]
        if rescale:
            boxes = boxes / boxes.new_tensor(sfs)[det_img.long()]
        rles = ops.mask_paste_rle(mask_pred, boxes.contiguous(), hw, det_img=det_img, mask_thr_binary=thr)
        out, k = [], 0
        for n in num_per_img:
            out.append(rles[k:k + n])
            k += n
        return out

    def get_seg_masks(self, mask_pred: torch.Tensor, det_bboxes: torch.Tensor, det_labels: torch.Tensor,
                      rcnn_test_cfg=None, ori_shape=None, scale_factor=None, rescale: bool = False, encode: bool = False):
        """FCNMaskHead.get_seg_masks [3P] for the class-agnostic single-class mask head of this model
        (fgn_r50_c4_densecl.py:123,127; labels forced to 0, fgn_roi_head.py:716), as called from
        fgn_roi_head.py:668-671.  ``det_bboxes`` are the boxes at test scale (``_bboxes``); with ``rescale`` they are
        divided by ``scale_factor`` and pasted on ``ori_shape``, else on the scaled shape.  Returns the
        reference's ``cls_segms`` ([[bool [h,w] numpy] * D] for class 0); with ``encode`` the COCO RLE dicts
        encode_mask_results would make of them (fgn.py:281), produced without materialising the masks."""
        cfg = rcnn_test_cfg if rcnn_test_cfg is not None else (self.test_cfg or {})
        cfg = cfg.get("rcnn", cfg) if isinstance(cfg, dict) else cfg
        thr = float(cfg.get("mask_thr_binary", 0.5))
        sf = [float(v) for v in (scale_factor if scale_factor is not None else (1.0, 1.0, 1.0, 1.0))]
        boxes = det_bboxes[:, :4]
        if rescale:
            img_h, img_w = int(ori_shape[0]), int(ori_shape[1])
            boxes = boxes / boxes.new_tensor(sf)
        else:
            img_h = int(np.round(ori_shape[0] * sf[1]).astype(np.int32))
            img_w = int(np.round(ori_shape[1] * sf[0]).astype(np.int32))
        boxes = boxes.contiguous()
        if encode:
            return [ops.mask_paste_rle(mask_pred, boxes, [(img_h, img_w)], mask_thr_binary=thr)]
        dense = ops.mask_paste(mask_pred, boxes, img_h, img_w, thr).cpu().numpy()
        return [[dense[i] for i in range(dense.shape[0])]]

    def simple_test(self, qry_fmap, proposal_list, img_metas=None, proposals=None, rescale=False,
                    spp_fmaps=None, spp_bboxes=None, spp_isegmaps=None):
        """fgn_roi_head.py:675-719.  Without ``test_cfg``: the raw per-image (cls_score, bbox_pred) splits.
        With it: ``(det_bboxes, det_labels)`` and, when the head has a mask branch, the mask-branch result dict
        (``mask_feats`` = attended RoI features, ``mask_pred`` if a mask head was given) as third element."""
        assert self.with_bbox, "Bbox head must be implemented."
        self.count_spp(spp_fmaps, spp_bboxes, spp_isegmaps)
        out = self.simple_test_bboxes(qry_fmap, img_metas, proposal_list, self.test_cfg, rescale=rescale)
        if self.test_cfg is None or not self.with_mask:
            return out
        det_bboxes, det_labels = out
        return det_bboxes, det_labels, self.simple_test_mask(qry_fmap, det_bboxes, det_labels, img_metas, rescale)
