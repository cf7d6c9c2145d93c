"""FCNMaskHead -- host-side mirror of mmdet 2.18 ``FCNMaskHead`` [3P] as the FGN config builds it
(fgn_r50_c4_densecl.py:115-129: num_convs=4, in_channels=1024, conv_out_channels=256, num_classes=1,
class_agnostic=True, no norm; called at fgn_roi_head.py:380), SURVEY 8f row 3.

Module / parameter names follow mmdet (``convs.{i}.conv``, ``upsample``, ``conv_logits``), so a reference checkpoint's
state dict loads.  At inference on CUDA the whole head runs in libfgn_b200 on the NHWC RoI tiles: each 3x3 convolution
as a tcgen05 implicit GEMM with bias + ReLU in the epilogue (ops.conv3x3), and the tail -- ConvTranspose2d(2, stride 2),
ReLU, 1x1 logits -- as ONE contraction whose epilogue reduces the logits straight out of tensor memory
(ops.deconv2x2_logits): num_convs + 1 launches, the upsampled [R,256,28,28] map is never written.  Training (autograd
recording) runs the plain torch modules; there is no CPU path for the fused route.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


class _ConvModule(nn.Module):
    """mmcv ConvModule(conv + ReLU) [3P] without norm: keeps the ``.conv`` attribute name of its state dict."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2)
        self.activate = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.activate(self.conv(x))


class FCNMaskHead(nn.Module):
    def __init__(self, num_convs: int = 4, roi_feat_size: int = 14, in_channels: int = 256, conv_kernel_size: int = 3,
                 conv_out_channels: int = 256, num_classes: int = 80, class_agnostic: bool = False,
                 upsample_cfg: Optional[dict] = None, tc: bool = True, precision: Optional[str] = None, **kwargs):
        super().__init__()
        up = dict(upsample_cfg or dict(type="deconv", scale_factor=2))
        if up.get("type") != "deconv" or up.get("scale_factor", 2) != 2 or conv_kernel_size != 3:
            raise NotImplementedError("FCNMaskHead: the FGN config uses 3x3 convs and a 2x deconv upsample")
        self.num_convs, self.in_channels, self.conv_out_channels = num_convs, in_channels, conv_out_channels
        self.num_classes, self.class_agnostic, self.roi_feat_size = num_classes, class_agnostic, roi_feat_size
        self.convs = nn.ModuleList(_ConvModule(in_channels if i == 0 else conv_out_channels, conv_out_channels, 3)
                                   for i in range(num_convs))
        up_in = conv_out_channels if num_convs > 0 else in_channels
        self.upsample = nn.ConvTranspose2d(up_in, conv_out_channels, 2, stride=2)
        self.conv_logits = nn.Conv2d(conv_out_channels, 1 if class_agnostic else num_classes, 1)
        self.relu = nn.ReLU(inplace=True)
        self.tc, self.precision = tc, precision
        self._prepared = None
        for m in (self.upsample, self.conv_logits):                       # mmdet FCNMaskHead.init_weights
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(m.bias, 0)

    def train(self, mode: bool = True):
        self._prepared = None                                              # weights may change
        return super().train(mode)

    def _prepare(self, device):
        """tap-major weights (and their TF32 split when the fp32-parity passes will run) -- once per weight load"""
        prec = self.precision or ("tf32" if torch.backends.cudnn.allow_tf32 else "fp32")
        if self._prepared is not None and self._prepared["device"] == device and self._prepared["prec"] == prec:
            return self._prepared
        split = (lambda t: ops.conv_split_weights(t)) if prec == "fp32" else (lambda t: None)
        taps = [ops.conv_taps(m.conv.weight.detach()) for m in self.convs]
        up = ops.conv_taps(self.upsample.weight.detach(), transposed=True)
        self._prepared = dict(device=device, prec=prec, taps=taps, splits=[split(t) for t in taps],
                              biases=[m.conv.bias.detach() for m in self.convs], up=up, up_split=split(up))
        return self._prepared

    def forward(self, x):
        use_tc = (self.tc and not self.training and x.is_cuda and x.dtype == torch.float32 and
                  not (torch.is_grad_enabled() and (x.requires_grad or self.conv_logits.weight.requires_grad)))
        if not use_tc:
            for conv in self.convs:
                x = conv(x)
            x = self.relu(self.upsample(x))
            return self.conv_logits(x)
        p = self._prepare(x.device)
        for i in range(self.num_convs):
            x = ops.conv3x3(x, p["taps"][i], p["biases"][i], relu=True, precision=p["prec"], w_split=p["splits"][i])
        return ops.deconv2x2_logits(x, p["up"], self.upsample.bias.detach(), self.conv_logits.weight.detach(),
                                    self.conv_logits.bias.detach(), precision=p["prec"], w_split=p["up_split"])
