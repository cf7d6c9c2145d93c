"""Does the state of the L2 at kernel start explain the gap between the ncu time (caches flushed) and the back-to-back time of
the 12-image RoIAlign launch?  Times single launches with events after (a) another call's launch, (b) a 512 MB memset."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, batch_episodes, episode_to_device, make_episode
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
base = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(4)]


def var(ep, k):
    e = dict(ep); e["qry"] = [q + 0.01 * k for q in ep["qry"]]; return e


calls = [batch_episodes([var(base[(c * B + i) % 4], c * B + i) for i in range(B)]) for c in range(2)]
n_ext = len(cfg.strides); scales = [1.0 / s for s in cfg.strides]
outs = [torch.empty((B * cfg.num_rois, 7, 7, cfg.channels), device=dev) for _ in range(2)]
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def launch(c):
    ops.roi_align_multilevel(calls[c]["qry"][:n_ext], calls[c]["rois"], scales, 7, 0, True, out_format="nhwc")


def time_one(c, before):
    before()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(c); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


for _ in range(3):
    launch(0); launch(1)
res = {}
res["after_other_call"] = sorted(time_one(i % 2, lambda: launch((i + 1) % 2)) for i in range(8))[3]
res["after_memset_512MB"] = sorted(time_one(i % 2, lambda: flush.zero_()) for i in range(8))[3]
res["after_same_call"] = sorted(time_one(0, lambda: launch(0)) for i in range(8))[3]
print(json.dumps({"images_per_launch": B, "us_single_launch_median": {k: round(v, 1) for k, v in res.items()}}))
