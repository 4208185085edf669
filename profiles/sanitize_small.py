#!/usr/bin/env python
"""Small problems through the kernels added in round 2, for compute-sanitizer (memcheck / racecheck):
staged blocked eigensolver (DMMA rank-2nb update, compact-WY back-transformation with its cp.async ring),
the register / tensor-pipe sweep kernels (warp-synchronous protocol), the n = 4 direct solver on a grid with images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from tests import models as M
from oracle import pythtb_oracle as orc
worst = 0.0
rib = M.bn_ribbon(tb, 35)                                   # n = 70: three reflector blocks, two 64-row chunks
k = np.array([[0.0], [0.11], [0.37], [0.5], [0.73]])
ev, vec = rib.solve_all(k, eig_vectors=True)
worst = max(worst, float(np.max(np.abs(ev - orc.solve_all(rib, k)))))
w = tb.wf_array(rib, [7]); w.solve_on_grid([0.0]); ph = w.berry_phase(list(range(35)), 0, contin=False)
si = M.random_model(tb, norb=8, dim=3, nhop=40, nspin=1, seed=5)
kk = np.random.RandomState(1).rand(300, 3)
worst = max(worst, float(np.max(np.abs(si.solve_all(kk) - orc.solve_all(si, kk)))))
m6 = M.random_model(tb, norb=3, dim=2, nhop=9, nspin=2, seed=6)
kk2 = np.random.RandomState(2).rand(200, 2)
worst = max(worst, float(np.max(np.abs(m6.solve_all(kk2) - orc.solve_all(m6, kk2)))))
km = M.kane_mele(tb, "odd")
wk = tb.wf_array(km, [49, 50]); g = wk.solve_on_grid([-0.5, -0.5]); f = wk.berry_flux([0, 1])
torch.cuda.synchronize()
print("sanitize_small ok, worst eigenvalue deviation %.2e, KM flux %.3e" % (worst, f))
