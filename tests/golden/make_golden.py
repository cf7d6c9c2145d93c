"""Mint the golden fixtures in tests/golden/ -- run in the BUILD container only (it reads
/root/reference, which does not exist on the GPU box; the tests read just the committed .npz).

    python tests/golden/make_golden.py

Two families of fixtures:

1. fgn_reference_*.npz -- outputs of the REFERENCE'S OWN, UNMODIFIED methods, imported from
   /root/reference/subprojects/sp02_omniiseg_fgn_mmdet/{fgn_roi_head,fgn_ag_rpn_head}.py.  Those
   files import mmdet/mmcv at module top (not installable here: mmcv-full 1.3.16 / mmdet 2.18.0,
   requirements.txt:45-46), so the third-party names are stubbed in sys.modules with minimal
   restatements of the [3P] semantics: bbox2roi, RPNHead.forward_single,
   SingleRoIExtractor (via torchvision's CPU roi_align), BBoxHead.forward (avg-pool + 2 FCs).
   The arithmetic FGN itself owns -- count_spp, count_one_roi_by_n_spp, count_modified_cls_bbox,
   _bbox_forward, _mask_forward + the vector gather of simple_test, AGRPNHead.forward_single --
   executes from the reference source, line for line.
2. roi_align_kat_*.npz -- known-answer vectors for RoIAlign / map_roi_levels produced by
   torch.ops.torchvision.roi_align on CPU (the op fgn_roi_head.py:429 calls) and by the torch
   expression of map_roi_levels, including the boundary cases of SURVEY section 8c.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision  # noqa: F401

REF_ROOT = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))     # repo root, for oracle.*


# ------------------------------------------------------------------------------------------------
# stubs for the third-party names the reference files import
# ------------------------------------------------------------------------------------------------
class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _bbox2roi(bbox_list):
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


class _RPNHead(nn.Module):
    """mmdet RPNHead [3P]: 3x3 conv + ReLU, 1x1 cls (A*1, sigmoid), 1x1 reg (A*4)."""

    def __init__(self, in_channels=16, feat_channels=16, num_anchors=15, **kw):
        super().__init__()
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, num_anchors * 1, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, num_anchors * 4, 1)

    def forward_single(self, x):
        x = F.relu(self.rpn_conv(x), inplace=True)
        return self.rpn_cls(x), self.rpn_reg(x)


class _SingleRoIExtractor(nn.Module):
    """mmdet SingleRoIExtractor [3P] over torchvision's CPU roi_align (aligned=True, sr=0)."""

    def __init__(self, strides, output_size=7, finest_scale=56):
        super().__init__()
        self.featmap_strides, self.output_size, self.finest_scale = list(strides), output_size, finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        target_lvls = torch.floor(torch.log2(scale / self.finest_scale + 1e-6))
        return target_lvls.clamp(min=0, max=num_levels - 1).long()

    def forward(self, feats, rois, roi_scale_factor=None):
        p = self.output_size
        num_levels = len(feats)
        out = feats[0].new_zeros(rois.size(0), feats[0].shape[1], p, p)
        ra = lambda f, r, s: torch.ops.torchvision.roi_align(f, r, 1.0 / s, p, p, 0, True)
        if num_levels == 1:
            return out if len(rois) == 0 else ra(feats[0], rois, self.featmap_strides[0])
        lv = self.map_roi_levels(rois, num_levels)
        for i in range(num_levels):
            inds = (lv == i).nonzero(as_tuple=False).squeeze(1)
            if inds.numel() > 0:
                out[inds] = ra(feats[i], rois[inds], self.featmap_strides[i])
        return out


class _BBoxHead(nn.Module):
    """mmdet BBoxHead [3P], with_avg_pool=True, num_classes=1, reg_class_agnostic=False."""

    def __init__(self, in_channels=16, roi_feat_size=7, **kw):
        super().__init__()
        self.avg_pool = nn.AvgPool2d(roi_feat_size)
        self.fc_cls = nn.Linear(in_channels, 2)
        self.fc_reg = nn.Linear(in_channels, 4)

    def forward(self, x):
        x = self.avg_pool(x)
        x = x.view(x.size(0), -1)
        return self.fc_cls(x), self.fc_reg(x)


class _StandardRoIHead(nn.Module):
    def __init__(self, **kw):
        super().__init__()

    @property
    def with_shared_head(self):
        return hasattr(self, "shared_head") and self.shared_head is not None

    @property
    def with_bbox(self):
        return hasattr(self, "bbox_head") and self.bbox_head is not None

    @property
    def with_mask(self):
        return hasattr(self, "mask_head") and self.mask_head is not None


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    ident = lambda *a, **k: (lambda f: f)
    mod("mmcv")
    mod("mmcv.runner", auto_fp16=ident, force_fp32=ident)
    mod("mmdet")
    dummy = type("Dummy", (), {})
    mod("mmdet.core", BitmapMasks=dummy, encode_mask_results=None, bbox2result=None, bbox2roi=_bbox2roi,
        build_assigner=None, build_sampler=None)
    mod("mmdet.core.bbox")
    mod("mmdet.core.bbox.samplers", RandomSampler=dummy)
    mod("mmdet.models")
    mod("mmdet.models.detectors", TwoStageDetector=nn.Module)
    mod("mmdet.models.builder", DETECTORS=_Registry(), HEADS=_Registry(), MODELS=_Registry())
    mod("mmdet.models.dense_heads", RPNHead=_RPNHead)
    mod("mmdet.models.roi_heads", BBoxHead=_BBoxHead, StandardRoIHead=_StandardRoIHead)
    mod("mmdet.models.backbones")
    mod("mmdet.models.backbones.resnet", Bottleneck=type("Bottleneck", (), {"expansion": 4}))
    mod("mmdet.models.utils", ResLayer=lambda **kw: None)      # shared_head built by the fixture, see below
    mod("printy", printy=print)


def import_reference():
    install_stubs()
    sys.path.insert(0, REF_ROOT)
    from subprojects.sp02_omniiseg_fgn_mmdet import fgn_roi_head, fgn_ag_rpn_head
    return fgn_roi_head, fgn_ag_rpn_head


# ------------------------------------------------------------------------------------------------
def synth_rois(g, n, img_h, img_w, batch, smin=8.0):
    """Random proposals: centre uniform, log-uniform size, log-uniform aspect in [1/3,3], clipped."""
    cx = torch.rand(n, generator=g) * img_w
    cy = torch.rand(n, generator=g) * img_h
    s = torch.exp(torch.rand(n, generator=g) * np.log(min(img_h, img_w) / smin)) * smin
    ar = torch.exp((torch.rand(n, generator=g) - 0.5) * 2 * np.log(3.0))
    w, h = s * torch.sqrt(ar), s / torch.sqrt(ar)
    x1, y1 = (cx - w / 2).clamp(0, img_w), (cy - h / 2).clamp(0, img_h)
    x2, y2 = (cx + w / 2).clamp(0, img_w), (cy + h / 2).clamp(0, img_h)
    b = torch.randint(0, batch, (n,), generator=g).float()
    return torch.stack([b, x1, y1, x2, y2], 1).float()


def synth_support(g, m, s):
    """Centred boxes of side 0.8*S with +-4 px jitter (fgn_train.py:40) and ellipse+speckle masks."""
    side = 0.8 * s
    c = s / 2 + (torch.rand(m, 2, generator=g) - 0.5) * 8
    boxes = torch.stack([c[:, 0] - side / 2, c[:, 1] - side / 2, c[:, 0] + side / 2, c[:, 1] + side / 2], 1)
    yy, xx = torch.meshgrid(torch.arange(s).float(), torch.arange(s).float(), indexing="ij")
    masks = torch.zeros(m, 1, s, s, dtype=torch.bool)
    for i in range(m):
        ax = side / 2 * (0.5 + 0.5 * torch.rand(1, generator=g))
        ay = side / 2 * (0.5 + 0.5 * torch.rand(1, generator=g))
        ell = ((xx - c[i, 0]) / ax) ** 2 + ((yy - c[i, 1]) / ay) ** 2 <= 1
        inside = (xx >= boxes[i, 0]) & (xx <= boxes[i, 2]) & (yy >= boxes[i, 1]) & (yy <= boxes[i, 3])
        speck = (torch.rand(s, s, generator=g) < 0.1) & inside
        masks[i, 0] = ell ^ speck
    return boxes.float().view(m, 1, 4), masks


def make_reference_fixture(name, seed, B, N, K, C, qh, qw, S, stride, R, shared):
    fgn_roi_head, fgn_ag_rpn_head = import_reference()
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)

    head = fgn_roi_head.FGNRoIHead.__new__(fgn_roi_head.FGNRoIHead)
    nn.Module.__init__(head)
    head.n_ways, head.k_shots, head.subsampling_ratio = N, K, stride
    # the reference hard-codes 2048->1024 / GN(32,1024) (fgn_roi_head.py:241-243); its own init
    # routine is reused for the initialisation scheme with the channel count made a parameter
    head.cls_reg_shared_conv = nn.Conv2d(2 * C, C, kernel_size=(1, 1))
    head.cls_reg_shared_conv_norm = nn.GroupNorm(32, C, affine=True)
    nn.init.kaiming_normal_(head.cls_reg_shared_conv.weight, nonlinearity="relu")
    nn.init.normal_(head.cls_reg_shared_conv.bias, std=0.1)
    nn.init.normal_(head.cls_reg_shared_conv_norm.weight, mean=1.0, std=0.2)
    nn.init.normal_(head.cls_reg_shared_conv_norm.bias, std=0.2)
    head.bbox_roi_extractor = _SingleRoIExtractor([stride])
    head.mask_roi_extractor = head.bbox_roi_extractor
    head.bbox_head = _BBoxHead(C)
    nn.init.xavier_normal_(head.bbox_head.fc_cls.weight)
    nn.init.xavier_normal_(head.bbox_head.fc_reg.weight)
    nn.init.normal_(head.bbox_head.fc_cls.bias, std=0.1)
    nn.init.normal_(head.bbox_head.fc_reg.bias, std=0.1)
    head.mask_head = nn.Identity()
    if shared:
        # stand-in for the C4 ResLayer (mmdet [3P]): any deterministic module between RoIAlign and
        # the fusion exercises the same reference lines (:333-334, :368-369, :435-436)
        head.shared_head = nn.Sequential(nn.Conv2d(C, C, 3, padding=1), nn.ReLU())
    else:
        head.shared_head = None
    head.eval()

    qry = torch.randn(B, C, qh // stride, qw // stride, generator=g)
    spp = torch.randn(B * N * K, C, S // stride, S // stride, generator=g)
    spp_bboxes, spp_masks = synth_support(g, B * N * K, S)
    rois = synth_rois(g, R, qh, qw, B)
    rois = rois[torch.argsort(rois[:, 0], stable=True)]          # bbox2roi order: grouped by image

    out = dict(seed=seed, B=B, N=N, K=K, C=C, S=S, stride=stride, qry=qry.numpy(), spp=spp.numpy(),
               spp_bboxes=spp_bboxes.numpy().copy(), spp_masks=spp_masks.numpy(), rois=rois.numpy())
    with torch.no_grad():
        # ---- FGNRoIHead.count_spp (fgn_roi_head.py:419-449), reference code
        boxes_in = spp_bboxes.clone()
        head.count_spp(spp.clone(), boxes_in, spp_masks.clone())
        out["spp_bboxes_after"] = boxes_in.numpy()                # in-place /= 16 side effect
        out["cat_mean"] = head.spp_fmaps_roi_aligned_cat_mean.numpy()
        out["masked_gap"] = head.spp_fvecs_roi_aligned_cat_mean_mp.numpy()
        # ---- FGNRoIHead._bbox_forward (fgn_roi_head.py:328-342), reference code
        if N in (1, 3):                                           # reference asserts N in {1,3}
            res = head._bbox_forward(qry, rois)
            out["cls_score"], out["bbox_pred"] = res["cls_score"].numpy(), res["bbox_pred"].numpy()
            out["bbox_feats"] = res["bbox_feats"].numpy()
            # pieces, for finer-grained checks
            n_r, fused = head.count_one_roi_by_n_spp(res["bbox_feats"], rois)
            raw_c, raw_r = head.bbox_head.forward(fused)
            out["raw_cls"], out["raw_reg"] = raw_c.numpy(), raw_r.numpy()
            out["fused_pooled"] = fused.mean(dim=(2, 3)).numpy()
            # ---- mask branch: vector gather of simple_test (:707-714) + _mask_forward (:360-382)
            D = min(R, 24)
            det_rois = rois[:D]
            labels = torch.randint(0, N, (D,), generator=g)
            det_labels = [labels[det_rois[:, 0] == b] for b in range(B)]
            gather = torch.cat([det_labels[i] + head.n_ways * i for i in range(B)])
            batch, n, c = head.spp_fvecs_roi_aligned_cat_mean_mp.shape[:3]
            head.spp_vecs_mask = head.spp_fvecs_roi_aligned_cat_mean_mp.view(batch * N, c, 1, 1)[gather]
            mres = head._mask_forward(qry, det_rois)
            out["det_rois"], out["det_labels"] = det_rois.numpy(), labels.numpy()
            out["mask_feats"] = mres["mask_feats"].numpy()
        # ---- AGRPNHead.forward_single (fgn_ag_rpn_head.py:26-118), reference code
        rpn = fgn_ag_rpn_head.AGRPNHead(in_channels=C, feat_channels=C, num_anchors=15)
        rpn.n_ways, rpn.k_shots = N, K
        rpn.eval()
        cls, reg = rpn.forward_single(qry, spp, log_mode=True)
        out["rpn_qry_fmap_mod"] = rpn.qry_fmap_mod.numpy()
        out["rpn_cls_raw"], out["rpn_reg_raw"] = rpn.rpn_cls_score.numpy(), rpn.rpn_bbox_pred.numpy()
        out["rpn_cls"], out["rpn_reg"] = cls.numpy(), reg.numpy()
        for k, v in rpn.state_dict().items():
            out["rpnw." + k] = v.numpy()
    for k, v in head.state_dict().items():
        out["w." + k] = v.numpy()
    path = os.path.join(OUT, f"fgn_reference_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_roi_align_kats():
    g = torch.Generator().manual_seed(77)
    cases = []
    # (name, B, C, H, W, scale, P, sampling_ratio, aligned, R)
    for name, B, C, H, W, scale, P, sr, aligned, R in [
        ("c4_aligned", 2, 8, 30, 30, 1 / 16, 7, 0, True, 64),
        ("c4_wide", 1, 4, 50, 84, 1 / 16, 7, 0, True, 64),
        ("support_tv", 3, 4, 16, 16, 1.0, 7, -1, False, 24),
        ("mask_tv", 3, 1, 128, 128, 1.0, 7, -1, False, 12),
        ("fpn_p14", 1, 4, 25, 42, 1 / 32, 14, 0, True, 32),
        ("fixed_sr2", 2, 4, 20, 20, 0.25, 7, 2, True, 32),
    ]:
        feat = torch.randn(B, C, H, W, generator=g)
        rois = synth_rois(g, R, H / scale, W / scale, B, smin=4.0)
        # boundary rows: zero area, x2<x1, fully outside, larger than the image, edges
        iw, ih = W / scale, H / scale
        special = torch.tensor([[0, 5., 5., 5., 5.], [0, 10., 10., 4., 4.], [0, -500., -500., -400., -400.],
                                [0, -50., -50., iw + 60, ih + 60], [0, 0., 0., iw, ih],
                                [0, iw - 1, ih - 1, iw, ih], [0, -1 / scale, -1 / scale, 0., 0.],
                                [0, iw, ih, iw + 3 / scale, ih + 3 / scale]])
        rois = torch.cat([special, rois], 0)
        out = torch.ops.torchvision.roi_align(feat, rois, scale, P, P, sr, aligned)
        cases.append((name, dict(feat=feat.numpy(), rois=rois.numpy(), scale=np.float32(scale), P=P, sr=sr,
                                 aligned=aligned, out=out.numpy())))
    path = os.path.join(OUT, "roi_align_kat.npz")
    flat = {}
    for name, d in cases:
        for k, v in d.items():
            flat[f"{name}.{k}"] = v
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")

    # map_roi_levels boundary vectors: sqrt(w*h) at {112,224,448}*(1 +- few ulp), plus degenerate boxes
    rows = []
    for t in (112.0, 224.0, 448.0, 56.0, 896.0):
        base = np.float32(t)
        for k in range(-6, 7):
            s = base
            for _ in range(abs(k)):
                s = np.nextafter(s, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
            rows.append([0, 0, 0, s, s])                        # square: sqrt(s*s) == s when exact
            rows.append([0, 10, 20, 10 + s * 2, 20 + s / 2])
    rows += [[0, 5, 5, 5, 5], [0, 10, 10, 4, 20], [0, 0, 0, 1e-3, 1e-3], [0, 0, 0, 5000, 5000], [0, 3, 3, 2, 2]]
    rois = torch.tensor(rows, dtype=torch.float32)
    extra = synth_rois(g, 4000, 800, 1344, 1, smin=4.0)
    rois = torch.cat([rois, extra], 0)
    lv = {}
    for L in (1, 2, 4, 5):
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        lv[f"L{L}"] = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=L - 1).long().numpy()
    path = os.path.join(OUT, "map_roi_levels_kat.npz")
    np.savez_compressed(path, rois=rois.numpy(), **lv)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make_roi_align_kats()
    #                      name        seed  B  N  K  C   qh   qw   S   stride R   shared
    make_reference_fixture("n1k1_c4",   101, 1, 1, 1, 32, 160, 160, 64, 16,   40, True)
    make_reference_fixture("n3k1_c4",   102, 2, 3, 1, 32, 192, 256, 64, 16,   48, False)
    make_reference_fixture("n3k3_c4",   103, 2, 3, 3, 64, 128, 160, 128, 16,  32, True)
    make_reference_fixture("n5k2_spp",  104, 1, 5, 2, 32, 128, 128, 64, 16,   16, False)
