"""Timing of the fused mask paste + RLE kernel (fgn_mask_paste_rle) on cfg3's test-time shape (100 detections,
28x28 masks, 800x1344 image) next to (a) the materialised form the reference uses -- dense [D,H,W] masks on the
device (fgn_mask_paste) copied to the host -- and (b) the CPU restatement (paste + RLE) on a bounded sample.
Reporting tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fgn_b200 import _lib, ops
from fgn_b200.episodes import synth_rois
from oracle import fgn_oracle as O

dev = torch.device("cuda:0")
H, W, D, M = 800, 1344, 100, 28
g = torch.Generator().manual_seed(9)
boxes = synth_rois(g, D, H, W, 1)[:, 1:].contiguous()
# blob-like logits: a smooth bump per mask (real mask heads give one or two connected components)
yy, xx = torch.meshgrid(torch.linspace(-1, 1, M), torch.linspace(-1, 1, M), indexing="ij")
logits = torch.stack([6 * (0.6 + 0.3 * torch.rand(1, generator=g) - (xx * xx + yy * yy)) + 0.5 * torch.randn(M, M, generator=g)
                      for _ in range(D)])[:, None].contiguous()
lt, bt = logits.to(dev), boxes.to(dev)
cap = 8192
hw_t = torch.tensor([[H, W]], dtype=torch.int32, device=dev)
counts = torch.empty((D, cap), device=dev, dtype=torch.int32)
ncounts = torch.empty((D,), device=dev, dtype=torch.int32)
sbuf = torch.empty((D, 2 * cap), device=dev, dtype=torch.uint8)
slen = torch.empty((D,), device=dev, dtype=torch.int32)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


wsb = int(lib.fgn_mask_paste_rle_workspace_bytes(D, cap, H, W))
ws = torch.empty((wsb,), device=dev, dtype=torch.uint8)


def fused():
    _lib.check(lib.fgn_mask_paste_rle(lt.data_ptr(), bt.data_ptr(), 4, None, hw_t.data_ptr(), D, M, 0.5, H, W,
                                      counts.data_ptr(), ncounts.data_ptr(), sbuf.data_ptr(), slen.data_ptr(), cap, 2 * cap,
                                      ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream), "fgn_mask_paste_rle")


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


us_fused = timeit(fused)
runs = ncounts.tolist()
sl = slen.tolist()
us_dense = timeit(lambda: ops.mask_paste(lt, bt, H, W))
host = torch.empty((D, H, W), dtype=torch.bool, pin_memory=True)
t0 = time.perf_counter()
for _ in range(5):
    host.copy_(ops.mask_paste(lt, bt, H, W)); torch.cuda.synchronize()
ms_dense_d2h = (time.perf_counter() - t0) * 1e3 / 5
for _ in range(2):
    rles = ops.mask_paste_rle(lt, bt, [(H, W)], cap=cap)
t0 = time.perf_counter()
for _ in range(10):
    rles = ops.mask_paste_rle(lt, bt, [(H, W)], cap=cap)
ms_fused_e2e = (time.perf_counter() - t0) * 1e3 / 10
ns = 5
t0 = time.perf_counter()
mk = O.get_seg_masks(logits[:ns].numpy(), boxes[:ns].numpy(), H, W)
enc = O.encode_mask_results(mk)
cpu_ms = (time.perf_counter() - t0) * 1e3 * D / ns
assert [r["counts"] for r in rles[:ns]] == [e["counts"] for e in enc], "fused RLE differs from the oracle's on the sample"
print(json.dumps({"case": f"cfg3 test-time masks: D={D}, M={M}, image {H}x{W}",
                  "fused_paste_rle_kernel_us": round(us_fused, 1), "runs_per_mask_mean": float(np.mean(runs)),
                  "rle_bytes_total": int(sum(sl)), "dense_paste_kernel_us": round(us_dense, 1),
                  "dense_bytes": D * H * W, "dense_paste_plus_d2h_ms": round(ms_dense_d2h, 2),
                  "fused_paste_rle_to_host_dicts_ms": round(ms_fused_e2e, 2),
                  "cpu_oracle_ms_extrapolated": round(cpu_ms, 1), "cpu_sample": f"{ns} of {D} masks",
                  "cpu_threads": torch.get_num_threads()}), flush=True)
