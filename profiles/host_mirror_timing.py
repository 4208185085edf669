#!/usr/bin/env python
"""Where the time of DeviceStore.host() goes for the 1024 x 1024 Haldane array (67 MB): permuted-view copy as
torch does it, vs contiguous() + copy, vs raw contiguous copy."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from tests import models as M
w = tb.wf_array(M.haldane(tb, delta=0.0), [1025, 1025])
w.solve_on_grid([-0.5, -0.5])
st = w._store if hasattr(w, "_store") else None
out = {}
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3
dev = st._dev
host_t = torch.zeros(st.shape, dtype=torch.complex128, pin_memory=True)
out["dev_is_contiguous"] = bool(dev.is_contiguous())
out["copy_permuted_view_ms"] = wall(lambda: host_t.copy_(dev, non_blocking=False))
out["contiguous_ms"] = wall(lambda: dev.contiguous())
c = dev.contiguous()
out["copy_contiguous_ms"] = wall(lambda: host_t.copy_(c, non_blocking=False))
out["copy_contiguous_nonblocking_ms"] = wall(lambda: host_t.copy_(c, non_blocking=True))
phys = st._phys
hp = torch.zeros(phys.shape, dtype=torch.complex128, pin_memory=True)
out["copy_phys_ms"] = wall(lambda: hp.copy_(phys, non_blocking=True))
out["bytes"] = dev.numel() * 16
def store_host():
    st.state = "device"; st.host()
out["store_host_ms"] = wall(store_host)
print(json.dumps(out))
