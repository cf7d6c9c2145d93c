"""Times ops.relation_fusion (class-term contraction + query contraction + GroupNorm/ReLU/pool/FC epilogue) at the cfg3
and cfg4 shapes, fused tcgen05 kernel (FGN_REL_FUSED=1, default) against the separate kernels (=0).  Development tool."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
from fgn_b200.episodes import make_weights
dev = torch.device("cuda:0")
for N in (1,):
    C, R = 256, 1000
    g = torch.Generator().manual_seed(7)
    w = make_weights(C, 3)
    params = ops.RelationParams(*[w[k].to(dev) for k in ("conv_w", "conv_b", "gn_w", "gn_b", "fc_cls_w", "fc_cls_b", "fc_reg_w", "fc_reg_b")])
    feats = [torch.randn(R, C, 7, 7, generator=g).to(dev).contiguous(memory_format=torch.channels_last) for _ in range(4)]
    cat = torch.randn(1, N, C, 7, 7, generator=g).to(dev)
    rb = torch.zeros(R, device=dev)
    res = {}
    for fused, dbg in ((1, 0), (1, 1), (0, 0)):
        os.environ["FGN_REL_FUSED"] = str(fused)
        os.environ["FGN_REL_DEBUG"] = str(dbg)
        def run():
            for f in feats:
                ops.relation_fusion(f, rb, cat, N, params)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        res[fused] = ops.relation_fusion(feats[0], rb, cat, N, params)
        print(json.dumps({"N": N, "C": C, "R": R, "fused": fused, "debug": dbg, "us_per_call": round(e0.elapsed_time(e1) * 1e3 / 40, 2)}), flush=True)
    d = max(float((res[1][0] - res[0][0]).abs().max()), float((res[1][1] - res[0][1]).abs().max()))
    print(json.dumps({"N": N, "max_abs_diff_fused_vs_separate": d}), flush=True)
