#!/usr/bin/env python
"""Secondary measurements (not the headline bench): BASELINE configs 3-5 at reduced sizes through the
public API, with the numpy oracle timed beside each on the host.  Usage: python profiles/bench_configs.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pythtb_b200 as tb
from tests import models as M, oracle_api
from oracle import pythtb_oracle as orc


def timeit(fn, reps=3):
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
# config 4: BN ribbon norb=200: solve_on_grid on 2001 k + berry_phase of the lower half
for ncell in (100, 200):
    rib = M.bn_ribbon(tb, ncell)
    n = rib._nsta
    nk = 1001 if ncell == 100 else 257
    w = tb.wf_array(rib, [nk])
    t = timeit(lambda: w.solve_on_grid([0.0]), reps=2)
    tb_ = timeit(lambda: w.berry_phase(range(n // 2), 0), reps=2)
    ribo = M.bn_ribbon(oracle_api, ncell)
    t0 = time.perf_counter(); orc.solve_all(ribo, np.linspace(0, 1, 16)[:, None], eig_vectors=True); tc = time.perf_counter() - t0
    out["cfg4_ribbon_n%d" % n] = {"gpu_kpts_per_s_eigh": (nk - 1) / t, "gpu_links_per_s_berry": (nk - 1) / tb_,
                                  "cpu_kpts_per_s_eigh": 16 / tc}
# config 5: cubic slab norb 99 / 199: solve_on_grid on 17x17 + all-band Wilson loop
for nl in (50, 100):
    slab = M.cubic_slab(tb, nl)
    n = slab._nsta
    w = tb.wf_array(slab, [17, 17])
    t = timeit(lambda: w.solve_on_grid([0.0, 0.0]), reps=2)
    tw = timeit(lambda: w.berry_phase(range(nl), 0, contin=False), reps=2)
    out["cfg5_slab_n%d" % n] = {"gpu_kpts_per_s_eigh": 256 / t, "gpu_links_per_s_berry_nocc%d" % nl: 16 * 17 / tw}
print(json.dumps(out, indent=1))
