#!/usr/bin/env python
"""solve_all WITH eigenvectors of the 8-band silicon model (2^18 k-points, device results) and its solve_on_grid on a
64^3 mesh: register / tensor-pipe kernel vs the tile solver (TBK_REG_EIGVALS=0)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from pythtb_b200 import _engine, _lib
import bench_extras as BX
eng = _engine.get_engine()
m = BX.silicon_model(tb)
handle, plan = eng.model_handle(m)
n = plan.nsta
nk = 1 << 18
k = torch.rand((nk, 3), dtype=torch.float64, device=eng.device)
def timed(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = {"n": n, "nk": nk}
out["solve_all_vectors_ms"] = timed(lambda: eng.solve_all_device(m, k, nk, True))
w = tb.wf_array(m, [65, 65, 65])
out["solve_on_grid_64cube_ms"] = timed(lambda: w._solve_on_grid_device(np.zeros(3)))
out["kpts_per_s_vectors"] = nk / out["solve_all_vectors_ms"] * 1e3
out["kpts_per_s_grid"] = 64 ** 3 / out["solve_on_grid_64cube_ms"] * 1e3
print(json.dumps(out))
