#!/usr/bin/env python
"""Config 3 (silicon, n = 8): how much of solve_all is Hamiltonian assembly, how much the eigensolver?  Times
tbk_solve_k (fused), tbk_gen_ham alone and tbk_eigh_batched on prebuilt matrices for 2^20 k-points (CUDA events)."""
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
nk = 1 << 20
k = torch.rand((nk, 3), dtype=torch.float64, device=eng.device)
ev = torch.empty((n, nk), dtype=torch.float64, device=eng.device)
ham = torch.empty((nk, n, n), dtype=torch.complex128, device=eng.device)
ws = eng.workspace(max(eng.lib.tbk_solve_workspace(n, nk, 0), eng.lib.tbk_eigh_workspace(n, nk, 0), 1024))
P = _engine._ptr
def timed(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = {"n": n, "nk": nk, "nph": int(plan.nph) if hasattr(plan, "nph") else None}
out["solve_k_eigenvalues_ms"] = timed(lambda: _lib.check(eng.lib.tbk_solve_k(handle, P(k), nk, P(ev), nk, 1, P(None), 0, 0, P(ws), ws.numel(), eng.stream())))
out["gen_ham_ms"] = timed(lambda: _lib.check(eng.lib.tbk_gen_ham(handle, P(k), nk, P(ham), eng.stream())))
ev2 = torch.empty((nk, n), dtype=torch.float64, device=eng.device)
out["eigh_batched_eigenvalues_ms"] = timed(lambda: _lib.check(eng.lib.tbk_eigh_batched(P(ham), n, nk, P(ev2), P(None), P(ws), ws.numel(), eng.stream())))
out["kpts_per_s_fused"] = nk / out["solve_k_eigenvalues_ms"] * 1e3
print(json.dumps(out))
