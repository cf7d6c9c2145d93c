"""Achievable read bandwidth from L2 vs HBM with a plain streaming reduction (torch.sum)."""
import torch, json
dev = torch.device("cuda:0")
for mb in (16, 32, 64, 96, 256, 1024):
    x = torch.ones(mb * 1024 * 1024 // 4, device=dev)
    for _ in range(5):
        x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(json.dumps({"MB": mb, "us": round(us, 2), "GB/s": round(mb * 1.048576 / us * 1e3, 1)}), flush=True)
# copy kernel (read + write)
for mb in (32, 512):
    a = torch.ones(mb * 1024 * 1024 // 4, device=dev); b = torch.empty_like(a)
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(json.dumps({"copy MB": mb, "us": round(us, 2), "GB/s (r+w)": round(2 * mb * 1.048576 / us * 1e3, 1)}), flush=True)
