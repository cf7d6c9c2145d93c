"""RoIAlign on bf16 cells: rotating-window kernel vs the one-CTA-per-RoI kernel (FGN_RA_IMPL=2), cfg3, graph-replayed.
usage: probe_roi_bf16.py [images per launch ...]   -> one JSON line per (images per launch, kernel)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, batch_episodes, episode_to_device, make_episode

cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
n_ext = len(cfg.strides)
scales = [1.0 / s for s in cfg.strides]
base = [episode_to_device(make_episode(cfg, seed=i), dev) for i in range(4)]
for bcall in [int(a) for a in sys.argv[1:]] or [1, 12]:
    n_calls = 8 if bcall == 1 else 3
    calls = []
    for c in range(n_calls):
        eps = [dict(base[(c + i) % 4], qry=[q + 0.01 * (c * bcall + i) for q in base[(c + i) % 4]["qry"]]) for i in range(bcall)]
        ep = batch_episodes(eps) if bcall > 1 else eps[0]
        calls.append(([q.bfloat16().contiguous(memory_format=torch.channels_last) for q in ep["qry"][:n_ext]], ep["rois"],
                      [q.contiguous(memory_format=torch.channels_last) for q in ep["qry"][:n_ext]]))
    R = calls[0][1].shape[0]
    # FGN_RA_DEBUG ablations (results are garbage, timing only): bit 0 = no copies, bit 1 = no cell math
    variants = [("bf16 window", {}, 0), ("bf16 window-ns2", {"FGN_RA_NS": "2"}, 0), ("bf16 stream", {"FGN_RA_IMPL": "2"}, 0),
                ("bf16 window, no copies", {"FGN_RA_DEBUG": "1"}, 0), ("bf16 window, no cell math", {"FGN_RA_DEBUG": "2"}, 0),
                ("bf16 window, neither", {"FGN_RA_DEBUG": "3"}, 0), ("bf16 window, 1 CTA/SM", {"FGN_RA_CTAS": "1"}, 0),
                ("fp32 window", {}, 2), ("fp32 window, no copies", {"FGN_RA_DEBUG": "1"}, 2),
                ("fp32 window, no cell math", {"FGN_RA_DEBUG": "2"}, 2), ("fp32 window, neither", {"FGN_RA_DEBUG": "3"}, 2),
                ("fp32 window, 1 CTA/SM", {"FGN_RA_CTAS": "1"}, 2)]
    for name, env, qi in variants:
        os.environ.update(env)
        try:
            def run():
                for c in calls:
                    ops.roi_align_multilevel(c[qi], c[1], scales, 7, 0, True, out_format="nhwc")
            run(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            for _ in range(3):
                g.replay()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); t0.record()
            for _ in range(10):
                g.replay()
            t1.record(); torch.cuda.synchronize()
            us = t0.elapsed_time(t1) * 1e3 / (10 * n_calls)
            print(json.dumps({"images_per_launch": bcall, "rois": R, "kernel": name, "us_per_launch": round(us, 2),
                              "us_per_1000_rois": round(us * 1000 / R, 2)}), flush=True)
        finally:
            for k in env:
                del os.environ[k]
