# large-n wf_array.solve_on_grid + berry_phase timing with CUDA events and the per-stage counters (TBK_PROF=1)
import os, sys, time, ctypes, json
os.environ["TBK_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from pythtb_b200 import _engine
from tests import models as M
eng = _engine.get_engine()
out = {}
def ev_time(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best
cases = [("ribbon_n200", M.bn_ribbon(tb, 100), [1777]), ("ribbon_n400", M.bn_ribbon(tb, 200), [297]), ("ribbon_n400_4w", M.bn_ribbon(tb, 200), [1185]),
         ("slab_n199", M.cubic_slab(tb, 100), [18, 18]), ("slab_n499", M.cubic_slab(tb, 250), [13, 13])]
for name, model, mesh in cases:
    n = model._nsta
    w = tb.wf_array(model, mesh)
    start = [0.0] * len(mesh)
    buf = (ctypes.c_uint64 * 8)()
    w._solve_on_grid_device(np.array(start)); torch.cuda.synchronize()
    eng.lib.tbk_debug_profile(buf, 1)
    t = ev_time(lambda: w._solve_on_grid_device(np.array(start)), reps=1)
    eng.lib.tbk_debug_profile(buf, 1)
    c = [int(x) for x in buf]
    npts = int(np.prod([m - 1 for m in mesh]))
    tot = max(1, sum(c[:4]))
    occ = list(range(n // 2))
    tb_ = ev_time(lambda: w.berry_phase(occ, 0, contin=False), reps=1)
    nlinks = npts
    out[name] = {"n": n, "kpts": npts, "solve_s": t, "kpts_per_s": npts / t, "cycles_ms_per_matrix": 1e3 * tot / 1.9e9 / max(1, c[4] // 2),
                 "slowest_matrix_ms": 1e3 * c[6] / 1.9e9, "share": {s: round(c[i] / tot, 3) for i, s in enumerate(("hetrd", "bisect", "invit", "backtr"))},
                 "fallbacks": c[5], "berry_phase_s": tb_, "links_per_s": nlinks / tb_, "nocc": len(occ)}
print(json.dumps(out, indent=1))
