"""Where the CTA-pair contraction's time goes: FGN_TC_DEBUG ablations (bit 0 no epilogue, bit 1 no operand split). Development tool."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
dev = torch.device("cuda:0")
for M, N, K in [(49000, 256, 256), (49 * 12000, 256, 256), (49 * 300, 1024, 1024)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    a = [torch.randn(M, K, generator=g).to(dev) for _ in range(4)]
    wq = (torch.randn(N, K, generator=g) * (1.0 / K) ** 0.5).to(dev)
    for prec in ("fp32", "tf32"):
        want = (a[0].double() @ wq.double().t()).float()
        for dbg in (0, 1, 2, 3):
            if prec == "tf32" and dbg >= 2:
                continue
            os.environ["FGN_TC_DEBUG"] = str(dbg)
            for _ in range(3):
                for x in a:
                    ops.gemm_nt(x, wq, None, prec)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                for x in a:
                    ops.gemm_nt(x, wq, None, prec)
            e1.record()
            torch.cuda.synchronize()
            err = float((ops.gemm_nt(a[0], wq, None, prec) - want).abs().max()) if dbg == 0 else None
            print(json.dumps({"M": M, "N": N, "K": K, "precision": prec, "debug": dbg, "us": round(e0.elapsed_time(e1) * 1e3 / 40, 1),
                              "max_abs_err_vs_fp64": err}), flush=True)
