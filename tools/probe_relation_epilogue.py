"""Relation fusion (contraction + epilogue) on 12 000 RoIs, N = 1: the persistent ring-fed epilogue, the one-CTA-per-RoI
single-class epilogue (FGN_EPI_RING=0) and the general one (FGN_EPI_ONE=0 too), graph-replayed, same process.  usage: [RoIs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import CONFIGS, build_heads
cfg = CONFIGS["cfg3_coco2voc_n1k1_fpn"]
dev = torch.device("cuda:0")
rpn, head = build_heads(cfg, dev, seed=0)
params = head.relation_params()
g = torch.Generator(device="cpu").manual_seed(0)
R, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 12000), cfg.channels
IM = max(1, R // 1000)
feats = [torch.randn(R, 7, 7, C, generator=g).to(dev).permute(0, 3, 1, 2) for _ in range(3)]
rb = (torch.arange(R) * IM // R).to(dev)
spp = torch.randn(IM, 1, C, 7, 7, generator=g).to(dev)
outs = {}
for name, env in (("ring epilogue", {}), ("single-class epilogue", {"FGN_EPI_RING": "0"}),
                  ("general epilogue", {"FGN_EPI_RING": "0", "FGN_EPI_ONE": "0"}), ("ring epilogue (again)", {})):
    os.environ.update(env)
    try:
        def run():
            for f in feats:
                run.out = ops.relation_fusion(f, rb, spp, 1, params)
        run(); torch.cuda.synchronize()
        if os.environ.get("PROBE_VERBOSE"): print("eager ok", name, flush=True)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        for _ in range(3):
            gr.replay()
        if os.environ.get("PROBE_VERBOSE"): torch.cuda.synchronize(); print("replay ok", name, flush=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record()
        for _ in range(10):
            gr.replay()
        t1.record(); torch.cuda.synchronize()
        outs[name] = [o.clone() for o in run.out]
        print(json.dumps({"kernel": name, "rois": R, "us_per_call": round(t0.elapsed_time(t1) * 1e3 / 30, 2)}), flush=True)
    finally:
        for k in env:
            del os.environ[k]
ref = outs["general epilogue"]
print(json.dumps({"bitwise_equal": {k: all(torch.equal(x, y) for x, y in zip(v, ref)) for k, v in outs.items()}}))
