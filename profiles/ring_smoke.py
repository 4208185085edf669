#!/usr/bin/env python
"""Small, fast check of the blocked solver (eigenvalues + vectors) against numpy on a few matrix sizes — the first thing
to run after touching the tridiagonalisation (seconds; run it under a short `timeout`)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from tests import models as M
out = {}
for ncell, nk in ((20, 7), (50, 40), (100, 300), (200, 150)):
    m = M.bn_ribbon(tb, ncell)
    k = (np.arange(nk) + 0.25) / nk
    t0 = time.perf_counter()
    ev, vec = m.solve_all(k[:, None], eig_vectors=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    worst = 0.0
    for ik in (0, nk // 2, nk - 1):
        h = m._gen_ham([k[ik]])
        ref = np.linalg.eigvalsh(h)
        worst = max(worst, float(np.max(np.abs(ev[:, ik] - ref))))
        v = vec[:, ik, :]
        res = float(np.max(np.abs(h @ v.T - v.T * ev[:, ik][None, :])))
        orth = float(np.max(np.abs(v.conj() @ v.T - np.eye(v.shape[0]))))
        worst = max(worst, res, orth)
    out["n%d" % m._nsta] = {"worst": worst, "s": round(dt, 3)}
    print(json.dumps(out), flush=True)
