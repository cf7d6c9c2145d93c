"""RoIAlign layer and single-level/multi-level RoI extractor with mmcv/mmdet's interface.

Mirrors (third-party): ``mmcv.ops.RoIAlign(output_size, spatial_scale,
sampling_ratio, pool_mode, aligned)`` and ``mmdet SingleRoIExtractor(roi_layer, out_channels,
featmap_strides, finest_scale)`` as instantiated by fgn_r50_c4_densecl.py:69-73 and called at
fgn_roi_head.py:331-332,366-367.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops


def bbox2roi(bbox_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """mmdet.core.bbox2roi [3P] (fgn_roi_head.py:348,390,556,654): list of [n_i,4+] -> [R,5]."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


class RoIAlign(nn.Module):
    """``mmcv.ops.RoIAlign`` [3P] signature; avg pooling only (the only mode the reference uses)."""

    def __init__(self, output_size, spatial_scale: float = 1.0, sampling_ratio: int = 0,
                 pool_mode: str = "avg", aligned: bool = True, use_torchvision: bool = False):
        super().__init__()
        if isinstance(output_size, (tuple, list)):
            if output_size[0] != output_size[1]:
                raise NotImplementedError("fgn_b200.RoIAlign: square output only")
            output_size = output_size[0]
        if pool_mode != "avg":
            raise NotImplementedError("fgn_b200.RoIAlign implements pool_mode='avg' (the FGN path) only")
        self.output_size = (int(output_size), int(output_size))
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = bool(aligned)

    def forward(self, input: torch.Tensor, rois: torch.Tensor, out_format: str = "nchw") -> torch.Tensor:
        from . import autograd as A                      # differentiable w.r.t. `input` when autograd is recording
        return A.roi_align_multilevel([input], rois, [self.spatial_scale], self.output_size[0],
                                      self.sampling_ratio, self.aligned, out_format=out_format)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, pool_mode={self.pool_mode}, aligned={self.aligned})")


class SingleRoIExtractor(nn.Module):
    """mmdet ``SingleRoIExtractor`` [3P]: one fused kernel does map_roi_levels + per-level RoIAlign.

    Args mirror mmdet: ``roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=0)``,
    ``out_channels``, ``featmap_strides``, ``finest_scale=56``.
    """

    def __init__(self, roi_layer: dict, out_channels: int, featmap_strides: Sequence[int],
                 finest_scale: int = 56, init_cfg=None):
        super().__init__()
        cfg = dict(roi_layer)
        layer_type = cfg.pop("type", "RoIAlign")
        if layer_type != "RoIAlign":
            raise NotImplementedError(f"roi_layer type {layer_type!r}: only RoIAlign is on the FGN path")
        self.roi_layers = nn.ModuleList([RoIAlign(spatial_scale=1.0 / s, **cfg) for s in featmap_strides])
        self.out_channels = out_channels
        self.featmap_strides = list(featmap_strides)
        self.finest_scale = finest_scale

    @property
    def num_inputs(self) -> int:
        return len(self.featmap_strides)

    def map_roi_levels(self, rois: torch.Tensor, num_levels: int) -> torch.Tensor:
        return ops.map_roi_levels(rois, num_levels, self.finest_scale)

    def forward(self, feats, rois: torch.Tensor, roi_scale_factor: Optional[float] = None,
                chan_scale: Optional[torch.Tensor] = None, scale_index: Optional[torch.Tensor] = None,
                out_format: str = "nchw", return_levels: bool = False):
        if roi_scale_factor is not None:
            raise NotImplementedError("roi_scale_factor is not used on the FGN path")
        feats = list(feats)[: self.num_inputs]
        layer = self.roi_layers[0]
        from . import autograd as A                      # differentiable w.r.t. `feats` when autograd is recording
        return A.roi_align_multilevel(
            feats, rois, [l.spatial_scale for l in self.roi_layers][: len(feats)], layer.output_size[0],
            layer.sampling_ratio, layer.aligned, float(self.finest_scale), chan_scale=chan_scale, scale_index=scale_index,
            out_format=out_format, return_levels=return_levels)
