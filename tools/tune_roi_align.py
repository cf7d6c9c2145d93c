"""Times roi_align_multilevel on the cfg3 episode stream for the tuning knobs the library reads from
the environment (FGN_RA_NB, FGN_RA_CB).  Development tool, not part of the product path."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, episode_to_device, make_episode

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_coco2voc_n1k1_fpn"
out_fmt = sys.argv[2] if len(sys.argv) > 2 else "nhwc"
cfg = CONFIGS[name]
dev = torch.device("cuda:0")
E = int(os.environ.get("TUNE_E", "8"))
base = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(min(2, E))]
eps = []
for i in range(E):
    ep = dict(base[i % len(base)])
    if i >= 2:
        ep["qry"] = [q + 0.01 * i for q in ep["qry"]]
    eps.append(ep)
n_ext = len(cfg.strides)
scales = [1.0 / s for s in cfg.strides]


MASK = os.environ.get("TUNE_MASK") == "1"        # the mask-branch launch instead: cfg.mask_rois RoIs with the AG-FCN multiply
if MASK:
    for ep in eps:
        ep["cs"] = torch.rand(ep["det_rois"].shape[0], cfg.channels, device=dev)


def run():
    for ep in eps:
        if MASK:
            ops.roi_align_multilevel(ep["qry"][:n_ext], ep["det_rois"], scales, cfg.mask_size, 0, True, chan_scale=ep["cs"],
                                     out_format=out_fmt)
        else:
            ops.roi_align_multilevel(ep["qry"][:n_ext], ep["rois"], scales, 7, 0, True, out_format=out_fmt)


def timeit(reps=10):
    """The E launches are captured in one CUDA graph and the graph is replayed: the number is the
    device's, not the Python wrapper's (about 30 us of host work per call)."""
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * E)


knobs = {"FGN_RA_NB": [1, 2, 3, 4, 6, 8], "FGN_RA_CB": [128, 256]}
extra = [a for a in sys.argv[3:]]
for e in extra:           # e.g. FGN_RA_X=1,2
    k, v = e.split("=")
    knobs[k] = [x for x in v.split(";")] if k == "FGN_RA_THR" else [int(x) for x in v.split(",")]
keys = list(knobs)
for combo in itertools.product(*[knobs[k] for k in keys]):
    for k, v in zip(keys, combo):
        os.environ[k] = str(v)
    us = timeit()
    print(json.dumps({"cfg": name, "out": out_fmt, **dict(zip(keys, combo)), "us_per_launch": round(us, 2)}), flush=True)
