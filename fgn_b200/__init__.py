"""fgn_b200 -- B200 (sm_100a) implementation of FGN's guided RoIAlign + support-guided fusion path.

Host-side mirror of the reference's operator/plugin interface for this path only:

  reference (subprojects/sp02_omniiseg_fgn_mmdet/)          here
  -------------------------------------------------         ----------------------------------
  mmcv.ops.RoIAlign / SingleRoIExtractor [3P]               fgn_b200.RoIAlign / SingleRoIExtractor
  AGRPNHead.forward_single   (fgn_ag_rpn_head.py:26)        fgn_b200.AGRPNHead.forward_single
  FGNRoIHead.count_spp/_bbox_forward/_mask_forward/...      fgn_b200.FGNRoIHead (same names)
  (fgn_roi_head.py:253-449, 675-719)
  FGN.simple_test            (fgn.py:186-240)               fgn_b200.FGN.simple_test (backbone = caller's module)

All device work goes through the C ABI in include/fgn_b200.h (libfgn_b200.so, hand-written CUDA).
There is no CPU fallback.
"""
from ._lib import FgnError, load as load_library  # noqa: F401
from . import ops  # noqa: F401
from .roi_extractor import RoIAlign, SingleRoIExtractor, bbox2roi  # noqa: F401
from .ag_rpn_head import AGRPNHead  # noqa: F401
from .mask_head import FCNMaskHead  # noqa: F401
from .roi_head import FGNBBoxHead, FGNRoIHead  # noqa: F401
from .detector import FGN  # noqa: F401

__version__ = "0.1.0"
