"""FGN test-time driver: host-side mirror of the reference detector's ``simple_test``
(subprojects/sp02_omniiseg_fgn_mmdet/fgn.py:28-50,68-108,186-240) around the device path.

The backbone (ResNet-C4 / R50-FPN in the reference's configs) is the caller's module -- convolutions are not on
the path this package rebuilds -- everything after it runs through the C ABI:
AGRPNHead.forward_single -> AGRPNHead.get_bboxes -> FGNRoIHead.simple_test.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .ag_rpn_head import AGRPNHead
from .roi_head import FGNRoIHead


class FGN(nn.Module):
    """``FGN(n_ways, k_shots, backbone=..., rpn_head=..., roi_head=..., test_cfg=...)`` as fgn.py:41-50: the
    episode shape is pushed onto the heads.  ``test_cfg`` = dict(rpn=dict(nms_pre, nms=dict(iou_threshold),
    max_per_img, min_bbox_size), rcnn=dict(score_thr, nms=dict(iou_threshold), max_per_img))
    (fgn_r50_c4_densecl.py:173-186)."""

    def __init__(self, n_ways: int, k_shots: int, backbone: nn.Module, rpn_head: AGRPNHead, roi_head: FGNRoIHead,
                 test_cfg: Optional[dict] = None, train_cfg: Optional[dict] = None):
        super().__init__()
        self.backbone, self.rpn_head, self.roi_head = backbone, rpn_head, roi_head
        self.test_cfg, self.train_cfg = test_cfg, train_cfg
        self.n_ways, self.k_shots = n_ways, k_shots
        for m in (self.rpn_head, self.roi_head, self.roi_head.bbox_head):
            m.n_ways, m.k_shots = n_ways, k_shots
        if test_cfg is not None:
            self.roi_head.test_cfg = test_cfg.get("rcnn", test_cfg)

    def extract_feat(self, img: torch.Tensor):
        """fgn.py:68-80: backbone output (a tuple of maps; the C4 configs use element 0)."""
        x = self.backbone(img)
        return x if isinstance(x, (tuple, list)) else (x,)

    @staticmethod
    def get_img_metas(img_shape: Sequence) -> List[dict]:
        """fgn.py:110-123: one meta per query image, scale_factor 1."""
        metas = []
        for s in img_shape:
            t = tuple(int(v) for v in (s.tolist() if torch.is_tensor(s) else s))
            metas.append(dict(num_samples=1, pad_shape=t, img_shape=t, ori_shape=t, scale_factor=(1.0, 1.0, 1.0, 1.0)))
        return metas

    @torch.no_grad()
    def simple_test(self, qry_img: torch.Tensor, spp_imgs: torch.Tensor, spp_bboxes: torch.Tensor,
                    spp_isegmaps: torch.Tensor, img_shape: Sequence, rescale: bool = False, boxes_yxyx: bool = True):
        """fgn.py:186-240.  ``qry_img`` [B,3,H,W]; ``spp_imgs`` [B,N,K,3,S,S] (or already flattened
        [B*N*K,3,S,S]); ``spp_bboxes`` [...,4] in the dataset's YXYX order (``boxes_yxyx``, fgn.py:104-106) ;
        ``spp_isegmaps`` [...,S,S] bool.  Returns what FGNRoIHead.simple_test returns: per-image
        ``(det_bboxes, det_labels)`` lists (and the mask-branch dict when the head has a mask branch)."""
        if self.test_cfg is None:
            raise ValueError("FGN.simple_test needs test_cfg (rpn + rcnn)")
        if boxes_yxyx:
            spp_bboxes = spp_bboxes[..., [1, 0, 3, 2]]
        img_metas = self.get_img_metas(img_shape)
        qry_fmap = self.extract_feat(qry_img)[0]
        c, h, w = spp_imgs.shape[-3:]
        spp_fmaps = self.extract_feat(spp_imgs.reshape(-1, c, h, w))[0]
        spp_bboxes = spp_bboxes.reshape(-1, 1, 4).to(torch.float32).contiguous()
        h, w = spp_isegmaps.shape[-2:]
        spp_isegmaps = spp_isegmaps.reshape(-1, 1, h, w)
        self.rpn_head.log_mode = False
        rpn_cls_score, rpn_bbox_pred = self.rpn_head.forward_single(qry_fmap, spp_fmaps)
        proposal_cfg = self.test_cfg.get("rpn_proposal", self.test_cfg["rpn"])
        proposal_list = self.rpn_head.get_bboxes([rpn_cls_score], [rpn_bbox_pred], img_metas=img_metas, cfg=proposal_cfg)
        return self.roi_head.simple_test(qry_fmap, [p[:, :4] for p in proposal_list], img_metas, rescale=rescale,
                                         spp_fmaps=spp_fmaps, spp_bboxes=spp_bboxes, spp_isegmaps=spp_isegmaps)

    @staticmethod
    def format_results(outputs_all, inputs_all: Optional[dict] = None) -> List[dict]:
        """fgn.py:262-303: the per-image result dicts consumed by FSISEGEval (fsisegeval.py:51-104).
        ``outputs_all`` = what ``simple_test`` returned: (det_bboxes, det_labels[, mask-branch dict]);
        ``inputs_all`` = the batch-level inputs the reference copies through (fgn.py:242-260: ``qry_img_id``,
        ``qry_bboxes``, ``qry_cat_ids``, ``qry_isegmaps``, ``qry_child_idx``, ``cats_ids_to_sample_real``,
        ``spp_insts_ids`` ...), each indexable by image.  Per image: ``dt_scores`` [D], ``dt_bboxes`` [D,4] back in
        the dataset's YXYX order (fgn.py:275), ``dt_cat_ids`` [D], ``dt_isegmaps_rle`` (COCO RLE dicts, fgn.py:281),
        every input key of that image (fgn.py:284-285), tensors as numpy (fgn.py:287-291), and
        ``qry_isegmaps`` replaced by ``qry_isegmaps_rle`` (fgn.py:297-299; device masks are encoded on the device by
        ops.mask_rle_encode, never copied to the host as dense masks)."""
        from . import ops
        det_bboxes, det_labels = outputs_all[0], outputs_all[1]
        rles = outputs_all[2].get("segm_rles") if len(outputs_all) > 2 and isinstance(outputs_all[2], dict) else None
        out = []
        for i, (db, dl) in enumerate(zip(det_bboxes, det_labels)):
            db = db.detach().cpu().numpy()
            one = dict(dt_scores=db[:, -1].reshape(-1), dt_bboxes=db[:, [1, 0, 3, 2]].reshape(-1, 4),
                       dt_cat_ids=dl.detach().cpu().numpy().reshape(-1))
            if rles is not None:
                one["dt_isegmaps_rle"] = rles[i]
            for key in (inputs_all or {}):
                one[key] = inputs_all[key][i]
            if "qry_isegmaps" in one:
                qm = one.pop("qry_isegmaps")
                if torch.is_tensor(qm) and qm.is_cuda:
                    one["qry_isegmaps_rle"] = ops.mask_rle_encode(qm.reshape(-1, qm.shape[-2], qm.shape[-1]))
                else:                                  # host masks: move them once, encode on the device
                    qt = torch.as_tensor(qm)
                    dev = det_bboxes[i].device
                    one["qry_isegmaps_rle"] = ops.mask_rle_encode(qt.reshape(-1, qt.shape[-2], qt.shape[-1]).to(dev)) \
                        if dev.type == "cuda" else None
            for key, value in list(one.items()):
                if torch.is_tensor(value):
                    one[key] = value.detach().cpu().numpy()
            out.append(one)
        return out


class ChunkedResultWriter:
    """The evaluation hook's on-disk format (main.py:285-309): results accumulate over ``simple_test`` calls and are
    written as ``ResultsChunked/NN.pkl`` every ``chunk`` (1000) items and at the end.  ``add`` takes the list
    ``FGN.format_results`` returns; ``close`` flushes the remainder.  Returns / keeps the written paths."""

    def __init__(self, work_dir: str, chunk: int = 1000, subdir: str = "ResultsChunked"):
        import os
        self.dir = os.path.join(work_dir, subdir)
        os.makedirs(self.dir, exist_ok=True)
        self.chunk, self.results, self.counter, self.paths = int(chunk), [], 0, []

    def _flush(self):
        import os
        import pickle
        path = os.path.join(self.dir, f"{self.counter:02}.pkl")
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:                     # write_pkl_safe: never leave a half-written chunk behind
            pickle.dump(self.results, f, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, path)
        self.paths.append(path)
        self.counter += 1
        self.results = []

    def add(self, results) -> None:
        if isinstance(results, dict):
            results = [results]
        for r in results:
            self.results.append(r)
            if len(self.results) == self.chunk:
                self._flush()

    def close(self) -> List[str]:
        if self.results:
            self._flush()
        return self.paths
