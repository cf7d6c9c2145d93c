"""A few calls of the one-launch support branch at the cfg3 / cfg4 shapes (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(2)]
_, head = build_heads(cfg, dev, seed=0, shared_head=None)
head.relation_params()
n_ext = len(cfg.strides)
with torch.no_grad():
    for i in range(4):
        ep = eps[i % 2]
        head.count_spp(ep["spp"][:n_ext], ep["spp_bboxes"].clone(), ep["spp_masks"])
torch.cuda.synchronize()
print("ok")
