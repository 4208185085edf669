#!/usr/bin/env python
"""Raw pinned device->host and host->device copy bandwidth of the box (cudaMemcpyAsync through torch, CUDA events):
the ceiling against which bench.py's e2e_wfs_to_host is read."""
import json, torch
out = {}
for mb in (64, 256):
    n = mb * 1024 * 1024
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, (src, dst) in (("d2h", (d, h)), ("h2d", (h, d))):
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["%s_%dMB_GBps" % (name, mb)] = n / best / 1e6
print(json.dumps(out))
