"""A few FPN single-pass calls (RoIAlign -> contraction -> epilogue) on distinct cfg3 episodes, back to back on one stream
(for the chained ncu capture: how much of the RoI features the contraction finds in L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200.episodes import CONFIGS, build_heads, episode_to_device, make_episode
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
eps = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(3)]
_, head = build_heads(cfg, dev, seed=0, shared_head=None)
n_ext = len(cfg.strides)
with torch.no_grad():
    for i in range(6):
        ep = eps[i % 3]
        head.count_spp(ep["spp"][:n_ext], ep["spp_bboxes"].clone(), ep["spp_masks"])
        out = head._bbox_forward(ep["qry"][:n_ext], ep["rois"])
torch.cuda.synchronize()
print("ok", float(out["cls_score"].abs().sum()))
