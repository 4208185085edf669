# phase breakdown of the tridiagonalisation (library built with -DTBK_HETRD_PROF=1; TBK_PROF=1)
import os, sys, ctypes, json
os.environ["TBK_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from pythtb_b200 import _engine
from tests import models as M
eng = _engine.get_engine()
out = {}
for name, model, mesh in (("ribbon_n200", M.bn_ribbon(tb, 100), [593]), ("ribbon_n400", M.bn_ribbon(tb, 200), [297]),
                          ("slab_n499", M.cubic_slab(tb, 250), [9, 33])):
    w = tb.wf_array(model, mesh)
    w._solve_on_grid_device(np.zeros(len(mesh))); torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 8)()
    eng.lib.tbk_debug_profile(buf, 1)
    w._solve_on_grid_device(np.zeros(len(mesh))); torch.cuda.synchronize()
    eng.lib.tbk_debug_profile(buf, 1)
    c = [int(x) for x in buf]
    tot = c[0] + c[1] + c[2] + c[3] + c[7]
    out[name] = {"n": model._nsta, "matrices": c[4], "ms_per_matrix_per_cta": 1e3 * tot / 1.965e9 / max(1, c[4]), "slowest_matrix_ms": 1e3 * c[6] / 1.965e9,
                 "share": {"matvec": round(c[0] / tot, 3), "rank2k": round(c[3] / tot, 3), "hetrd_other": round(c[7] / tot, 3),
                           "bisect": round(c[1] / tot, 3), "invit": round(c[2] / tot, 3)}}
    del w
print(json.dumps(out, indent=1))
