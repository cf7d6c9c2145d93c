"""Pinned host -> device copy rate of this box: one large copy, and the bench's pattern (many copies on several streams)."""
import json, torch, time
dev = torch.device("cuda:0")
for mb in (97, 1024):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for streams in (1, 2, 4, 8):
        ss = [torch.cuda.Stream() for _ in range(streams)]
        parts = h.chunk(streams); dparts = d.chunk(streams)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            for s, a, b in zip(ss, parts, dparts):
                with torch.cuda.stream(s):
                    b.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"buffer_MB": mb, "streams": streams, "GB_per_s": round(8 * h.numel() / dt / 1e9, 2)}), flush=True)
