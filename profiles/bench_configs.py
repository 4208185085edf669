#!/usr/bin/env python
"""Secondary measurements (not the headline bench): BASELINE configs 3-5 at reduced sizes through the
public API, with the numpy oracle timed beside each on one host core.  Usage: python profiles/bench_configs.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pythtb_b200 as tb
from tests import models as M, oracle_api
from oracle import pythtb_oracle as orc


def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


out = {}
# config 3: silicon-like 8-orbital 3-D model (random w90-like, 300 hoppings), eigenvalues only on a 64^3 mesh
m8 = M.random_model(tb, norb=8, dim=3, nhop=300, nspin=1, seed=8)
k = m8.k_uniform_mesh([64, 64, 64])
t = timeit(lambda: m8.solve_all(k))
m8o = M.random_model(oracle_api, norb=8, dim=3, nhop=300, nspin=1, seed=8)
t0 = time.perf_counter(); orc.solve_all(m8o, k[:4096]); tc = time.perf_counter() - t0
out["cfg3_n8_nhop300_solve_all_64^3"] = {"gpu_kpts_per_s": len(k) / t, "cpu1_kpts_per_s": 4096 / tc}
# config 3 at full scale: the same model on the 256^3 mesh of BASELINE configs[2] (16.8 M k-points), k generated on the device
for nhop, tag in ((300, "nhop300"), (2972, "nhop2972_silicon_size")):
    mm = M.random_model(tb, norb=8, dim=3, nhop=nhop, nspin=1, seed=8)
    lazy = mm.k_uniform_mesh([256, 256, 256], lazy=True)
    td = timeit(lambda: mm.solve_all(lazy, device_result=True), reps=2)
    th = timeit(lambda: mm.solve_all(lazy), reps=1)
    out["cfg3_n8_%s_solve_all_256^3" % tag] = {"nk": len(lazy), "gpu_kpts_per_s_device_result": len(lazy) / td,
                                                "gpu_kpts_per_s_host_result": len(lazy) / th, "seconds_device_result": td}
# config 4: BN ribbon: solve_on_grid + berry_phase of the lower half (determinant branch)
for ncell, nk in ((100, 2369), (200, 1185)):
    rib = M.bn_ribbon(tb, ncell)
    n = rib._nsta
    w = tb.wf_array(rib, [nk])
    t = timeit(lambda: w._solve_on_grid_device(np.array([0.0])))
    tb_ = timeit(lambda: w.berry_phase(range(n // 2), 0))
    ribo = M.bn_ribbon(oracle_api, ncell)
    t0 = time.perf_counter(); orc.solve_all(ribo, np.linspace(0, 1, 8)[:, None], eig_vectors=True); tc = time.perf_counter() - t0
    out["cfg4_ribbon_n%d" % n] = {"nk": nk - 1, "gpu_kpts_per_s_eigh": (nk - 1) / t, "gpu_links_per_s_berry_nocc%d" % (n // 2): (nk - 1) / tb_,
                                  "cpu1_kpts_per_s_eigh": 8 / tc}
# config 5: cubic slab: solve_on_grid, all-band Wilson loop (berry_evals), batched HWF + single-band HWF Berry phases
for nl, mesh in ((50, [25, 25]), (100, [18, 18]), (250, [9, 9])):
    slab = M.cubic_slab(tb, nl)
    n = slab._nsta
    w = tb.wf_array(slab, mesh)
    npts = (mesh[0] - 1) * (mesh[1] - 1)
    t = timeit(lambda: w._solve_on_grid_device(np.array([0.0, 0.0])), reps=1)
    tw = timeit(lambda: w.berry_phase(range(nl), 0, contin=False), reps=1)
    te = timeit(lambda: w.berry_phase(range(nl), 0, contin=False, berry_evals=True), reps=1)
    th = timeit(lambda: w.position_hwf_all(list(range(nl)), 2, hwf_evec=True), reps=1)
    nlinks = (mesh[0] - 1) * mesh[1]
    out["cfg5_slab_n%d" % n] = {"mesh": mesh, "gpu_kpts_per_s_eigh": npts / t, "gpu_links_per_s_det_nocc%d" % nl: nlinks / tw,
                                "gpu_links_per_s_wilson_evals_nocc%d" % nl: nlinks / te,
                                "gpu_kpts_per_s_position_hwf_all": mesh[0] * mesh[1] / th}
print(json.dumps(out, indent=1))
