#!/usr/bin/env python
"""One grid solve with eigenvectors per config-scale shape (for an ncu launch list: per-kernel durations of the staged solver)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from tests import models as M
which = os.environ.get("PROF_WHICH", "ribbon_n200,ribbon_n400,slab_n499").split(",")
shapes = {"ribbon_n200": (lambda: M.bn_ribbon(tb, 100), [593]), "ribbon_n400": (lambda: M.bn_ribbon(tb, 200), [297]),
          "slab_n499": (lambda: M.cubic_slab(tb, 250), [9, 33])}
for tag in which:
    mk, mesh = shapes[tag]
    w = tb.wf_array(mk(), mesh)
    for _ in range(int(os.environ.get("PROF_REPS", "2"))):
        w._solve_on_grid_device(np.zeros(len(mesh)))
        torch.cuda.synchronize()
    del w
    torch.cuda.empty_cache()
