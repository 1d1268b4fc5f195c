"""H2D / D2H bandwidth of pinned host memory on this box (context for bench.py's e2e number)."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for mb in (6, 50, 200):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out["h2d_%dMB_GBs" % mb] = round(10 * n / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
    a.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    out["d2h_%dMB_GBs" % mb] = round(10 * n / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
print(json.dumps(out))
