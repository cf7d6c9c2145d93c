"""One warm-up + a few roi_align_multilevel launches on cfg3 (for ncu captures).
usage: run_roi_once.py [cfg] [fmt] [distinct episodes] [images per call]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, batch_episodes, episode_to_device, make_episode
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3_coco2voc_n1k1_fpn"]
fmt = sys.argv[2] if len(sys.argv) > 2 else "nhwc"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
bcall = int(sys.argv[4]) if len(sys.argv) > 4 else 1
eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(n)]
if bcall > 1:      # calls of `bcall` images, as bench.py issues them (distinct data by perturbation on the device)
    def var(ep, k):
        e = dict(ep)
        e["qry"] = [q + 0.01 * k for q in ep["qry"]]
        return e
    eps = [batch_episodes([var(eps[(c + i) % n], c * bcall + i) for i in range(bcall)]) for c in range(2)]
    n = 2
n_ext = len(cfg.strides)
for i in range(6):
    ep = eps[i % n]
    ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], [1.0 / s for s in cfg.strides], 7, 0, True, out_format=fmt)
torch.cuda.synchronize()
print("ok")
