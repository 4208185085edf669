# per-stage cycle breakdown of the blocked eigensolver (run with TBK_PROF=1 on the GPU box)
import os, sys, time, ctypes, json
os.environ["TBK_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from pythtb_b200 import _engine
from tests import models as M
eng = _engine.get_engine()
out = {}
for name, model, nk in (("ribbon_n200", M.bn_ribbon(tb, 100), 1776), ("ribbon_n400", M.bn_ribbon(tb, 200), 592),
                        ("slab_n99", M.cubic_slab(tb, 50), None), ("slab_n499", M.cubic_slab(tb, 250), None)):
    n = model._nsta
    if nk is None:
        nk = 148 * 2
        k = np.random.RandomState(1).rand(nk, 2)
    else:
        k = np.linspace(0, 1, nk, endpoint=False)[:, None]
    for vec in (True, False):
        model.solve_all(k[:8], eig_vectors=vec); torch.cuda.synchronize()
        buf = (ctypes.c_uint64 * 8)()
        eng.lib.tbk_debug_profile(buf, 1)
        t0 = time.perf_counter(); model.solve_all(k, eig_vectors=vec); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        eng.lib.tbk_debug_profile(buf, 1)
        c = [int(x) for x in buf]
        tot = max(1, sum(c[:4]))
        out["%s_vec%d" % (name, vec)] = {"n": n, "nk": nk, "kpts_per_s": nk / dt, "ms_per_matrix_per_cta": 1e3 * sum(c[:4]) / 1.9e9 / max(1, c[4]),
                                        "share": {s: round(c[i] / tot, 3) for i, s in enumerate(("hetrd", "bisect", "invit", "backtr"))},
                                        "matrices": c[4], "fallbacks": c[5]}
print(json.dumps(out, indent=1))
