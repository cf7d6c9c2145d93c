"""Where the overlapped step's time goes: the bench.py step (16 resident cfg3 episodes, one CUDA graph each, round-robin
on S streams) with the AG-RPN attention and / or the mask branch left out, plus the host time to enqueue a step.
Reporting tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200.episodes import CONFIGS, EpisodeRunner, build_heads, episode_to_device, make_episode

cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
E = 16
host = [make_episode(cfg, seed=i) for i in range(2)]
eps = []
for i in range(E):
    ep = episode_to_device(host[i % 2], dev, channels_last=True)
    if i >= 2:
        ep["qry"] = [q + 0.01 * i for q in ep["qry"]]
    eps.append(ep)
rpn, head = build_heads(cfg, dev, seed=0, shared_head=None)
for streams in (1, 4, 8, 16):
    for att, msk in ((1, 1), (0, 1), (1, 0), (0, 0), ("fold", 1)):
        if streams in (4, 16) and (att, msk) != (1, 1):
            continue
        r = EpisodeRunner(rpn, head, eps, use_graphs=True, with_attention=att if att == "fold" else bool(att), with_mask=bool(msk), n_streams=streams)

        def step():
            r.begin()
            for i in range(E):
                r.run(i)
            r.end()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (20 * E)
        print(json.dumps({"streams": streams, "attention": att, "mask_branch": msk, "us_per_episode": round(us, 1),
                          "host_enqueue_us_per_episode": round(t_host * 1e6 / (20 * E), 1),
                          "kernels_per_episode": r.launches_per_episode[0]}), flush=True)
        del r
