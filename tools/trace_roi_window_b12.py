"""Per-CTA timeline of the rotating-window RoIAlign kernel on a 12-image launch (FGN_RA_DEBUG bit 5 + argv[1]):
who waits for whom over the first 16 items of every CTA.  usage: trace_roi_window_b12.py [debug bits] [bf16]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FGN_RA_DEBUG"] = str(32 | int(sys.argv[1]) if len(sys.argv) > 1 else 32)
import numpy as np
import torch
from fgn_b200 import _lib, ops
from fgn_b200.episodes import CONFIGS, batch_episodes, episode_to_device, make_episode
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
ep = batch_episodes([episode_to_device(make_episode(cfg, seed=i % 4), dev) for i in range(12)])
n_ext = len(cfg.strides)
q = [x.contiguous(memory_format=torch.channels_last) for x in ep["qry"][:n_ext]]
if len(sys.argv) > 2:
    q = [x.bfloat16().contiguous(memory_format=torch.channels_last) for x in q]
for i in range(3):
    ops.roi_align_multilevel(q, ep["rois"], [1.0 / s for s in cfg.strides], 7, 0, True, out_format="nhwc")
torch.cuda.synchronize()
lib = _lib.load()
N = 296 * 16 * 12
buf = (ctypes.c_ulonglong * N)()
lib.fgn_debug_roi_window_trace.restype = None
lib.fgn_debug_roi_window_trace(buf, N)
a = np.frombuffer(buf, dtype=np.uint64).reshape(296, 16, 12).astype(np.int64)
ok = (a[:, :, 8] > 0) & (a[:, :, 10] > 0) & (a[:, :, 3] > 0)
t0 = a[:, :, 0][a[:, :, 0] > 0].min()
def stat(x, m, what):
    x = x[m]
    print(f"{what:44s} mean {x.mean():7.0f} ns  p10 {np.percentile(x, 10):7.0f}  p50 {np.median(x):7.0f}  p90 {np.percentile(x, 90):7.0f}  n={x.size}")
print("debug bits", os.environ["FGN_RA_DEBUG"], "bf16" if len(sys.argv) > 2 else "fp32")
stat(a[:, :, 1] - a[:, :, 0], ok, "planner: ticket -> ranges known")
stat(a[:, :, 2] - a[:, :, 1], ok, "planner: wait for a free plan slot")
stat(a[:, :, 3] - a[:, :, 2], ok, "planner: weight tables")
stat(a[:, :, 7] - a[:, :, 6], ok & (a[:, :, 7] > 0) & (a[:, :, 6] > 0), "producer: item's copies issued")
stat(a[:, :, 9] - a[:, :, 8], ok, "consumer warp 0: rows of the item")
stat(a[:, :, 10] - a[:, :, 9], ok, "consumer warp 0: trailing stores")
nxt = ok[:, 1:] & ok[:, :-1]
stat(a[:, 1:, 8] - a[:, :-1, 10], nxt, "consumer warp 0: gap to the next item")
stat(a[:, 1:, 8] - a[:, 1:, 3], nxt, "plan published -> consumer picks it up")
stat(a[:, 1:, 10] - a[:, :-1, 10], nxt, "item period (consumer end to end)")
stat(a[:, 1:, 3] - a[:, :-1, 3], nxt, "plan period")
span = a[:, 15, 10] - a[:, 0, 0]
m = (a[:, 15, 10] > 0) & (a[:, 0, 0] > 0)
print(f"first 16 items: span per CTA mean {span[m].mean() / 1e3:.1f} us  ({m.sum()} CTAs)")
