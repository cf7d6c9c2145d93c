"""A few eager passes of the whole guided path on one config (for ncu launch lists).  usage: run_episode_once.py [cfg] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode, run_guided_path
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg4_coco2voc_n20k5_fpn"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(2)]
rpn, head = build_heads(cfg, dev, seed=0, shared_head=None)
with torch.no_grad():
    for i in range(n):
        run_guided_path(rpn, head, eps[i % 2])
torch.cuda.synchronize()
print("ok")
