"""Times the relation contraction (ops.gemm_nt) on the CTA-pair and the single-CTA kernel (FGN_TC_2SM); development tool."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
dev = torch.device("cuda:0")
shapes = [(49000, 256, 256), (49 * 300, 1024, 1024)]
for M, N, K in shapes:
    g = torch.Generator(device="cpu").manual_seed(1)
    a = [torch.randn(M, K, generator=g).to(dev) for _ in range(4)]
    w = (torch.randn(N, 2 * K, generator=g) * (1.0 / K) ** 0.5).to(dev)
    want = (a[0].double() @ w[:, :K].double().t()).float()
    wq = w[:, :K].contiguous()
    for bk, pair in ((16, 0), (16, 2)):
        for prec in ("fp32", "tf32"):
            os.environ["FGN_TC_2SM"] = "1" if pair else "0"       # 0 = the single-CTA kernel of conv_tc.cu
            got = ops.gemm_nt(a[0], wq, None, prec)
            err = float((got - want).abs().max())
            for _ in range(3):
                for x in a:
                    ops.gemm_nt(x, wq, None, prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                for x in a:
                    ops.gemm_nt(x, wq, None, prec)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * len(a))
            flops = 2.0 * M * N * K * (3 if prec == "fp32" else 1)
            print(json.dumps({"M": M, "N": N, "K": K, "bk": bk, "pair": pair, "precision": prec, "us": round(us, 1),
                              "tensor_tflops": round(flops / us / 1e6, 1), "max_abs_err": err}), flush=True)
