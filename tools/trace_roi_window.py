"""Per-CTA timeline of the rotating-window RoIAlign kernel (development tool; FGN_RA_DEBUG bit 5)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FGN_RA_DEBUG"] = str(32 | int(sys.argv[1]) if len(sys.argv) > 1 else 32)
import numpy as np
import torch
from fgn_b200 import _lib, ops
from fgn_b200.episodes import CONFIGS, episode_to_device, make_episode
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
ep = episode_to_device(make_episode(cfg, seed=0), dev)
n_ext = len(cfg.strides)
for i in range(3):
    ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], [1.0 / s for s in cfg.strides], 7, 0, True, out_format="nhwc")
torch.cuda.synchronize()
lib = _lib.load()
N = 296 * 16 * 12
buf = (ctypes.c_ulonglong * N)()
lib.fgn_debug_roi_window_trace.restype = None
lib.fgn_debug_roi_window_trace(buf, N)
a = np.frombuffer(buf, dtype=np.uint64).reshape(296, 16, 12).astype(np.int64)
t0 = a[:, :, :11][a[:, :, :11] > 0].min()
names = ["ticket", "yrange", "yslot", "ydone", "xrange", "xdone", "prod0", "prodN", "cons0", "consR", "consE", "nst"]
print(" ".join(f"{n:>7s}" for n in ["cta", "item"] + names))
for cta in (0, 1, 100, 295):
    for it in range(8):
        row = a[cta, it]
        if row[8] == 0:
            continue
        print(" ".join(f"{v:7d}" for v in [cta, it] + [int(x - t0) if (x > 0 and j < 11) else int(x) for j, x in enumerate(row)]))
# aggregates over all CTAs / items
valid = a[:, :, 8] > 0
def stat(x, what):
    x = x[valid & (x > -10**8) & (x < 10**8)]
    print(f"{what:28s} mean {x.mean():8.0f} ns  p50 {np.median(x):8.0f}  p90 {np.percentile(x, 90):8.0f}  n={x.size}")
stat(a[:, :, 1] - a[:, :, 0], "ticket -> y ranges")
stat(a[:, :, 2] - a[:, :, 1], "y wait for slot")
stat(a[:, :, 3] - a[:, :, 2], "y tables")
stat(a[:, :, 8] - a[:, :, 3], "plan done -> consumer has it")
stat(a[:, :, 9] - a[:, :, 8], "consumer rows")
stat(a[:, :, 10] - a[:, :, 9], "consumer tail release")
stat(a[:, :, 7] - a[:, :, 6], "producer issue span")
stat(a[:, :, 6] - a[:, :, 3], "plan done -> producer start")
stat(a[:, :, 3] - a[:, :, 0], "plan total (ticket -> published)")
stat(a[:, :, 10] - a[:, :, 8], "consumer item total")
end = a[:, :, 10].max(axis=1)
print("kernel span (first stamp -> last consumer end): %.1f us; per-CTA end p10/p50/p90: %s" % (
    (end.max() - t0) / 1e3, np.percentile(end - t0, [10, 50, 90]) / 1e3))
print("items per CTA (traced, <=16):", valid.sum(axis=1).mean(), " stages/item mean", a[:, :, 11][valid].mean())
