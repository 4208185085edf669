#!/usr/bin/env python
"""CUDA-event timings of the large-matrix kernel families at the config 4 / 5 shapes (A/B of library variants:
PYTHTB_B200_LIB=profiles/ab/libtbk_X.so python profiles/time_large.py).  Prints one JSON line."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pythtb_b200 as tb
from tests import models as M


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


out = {"lib": os.environ.get("PYTHTB_B200_LIB", "default")}
for tag, model, mesh, d in (("ribbon_n200", M.bn_ribbon(tb, 100), [593], 0), ("ribbon_n400", M.bn_ribbon(tb, 200), [297], 0),
                            ("slab_n499", M.cubic_slab(tb, 250), [9, 33], 1)):
    n = model._nsta
    nocc = (n + 1) // 2 if tag.startswith("slab") else n // 2
    occ = list(range(nocc))
    w = tb.wf_array(model, mesh)
    rec = {"n": n, "nocc": nocc}
    rec["solve_on_grid_ms"] = timed(lambda: w._solve_on_grid_device(np.zeros(len(mesh))), reps=2)
    npts = int(np.prod([m - 1 for m in mesh]))
    nlinks = int(np.prod(mesh)) // mesh[d] * (mesh[d] - 1)
    rec["kpts_per_s"] = npts / (rec["solve_on_grid_ms"] * 1e-3)
    rec["berry_det_ms"] = timed(lambda: w.berry_phase(occ, d, contin=False))
    rec["links_per_s_det"] = nlinks / (rec["berry_det_ms"] * 1e-3)
    rec["overlap_lu_tflops"] = nlinks * (8.0 * nocc * nocc * n + (8.0 / 3.0) * nocc ** 3) / (rec["berry_det_ms"] * 1e-3) / 1e12
    rec["berry_evals_ms"] = timed(lambda: w.berry_phase(occ, d, contin=False, berry_evals=True))
    rec["links_per_s_evals"] = nlinks / (rec["berry_evals_ms"] * 1e-3)
    pdir = 1 if tag.startswith("ribbon") else 2
    rec["position_hwf_all_ms"] = timed(lambda: w.position_hwf_all(occ, pdir, hwf_evec=True))
    rec["hwf_kpts_per_s"] = int(np.prod(mesh)) / (rec["position_hwf_all_ms"] * 1e-3)
    out[tag] = rec
    del w
    torch.cuda.empty_cache()
print(json.dumps(out))
