# The UNMODIFIED reference (/root/reference/pythtb.py 1.8.0, per-k Python loop) on the configs[1] workload, one core of the BUILD container
# (the reference tree is not on the GPU box).  Context for the cpu_baseline of bench.py, which is the vectorised numpy port.
# Run: python profiles/time_reference_loop.py > profiles/reference_loop_timing.json
import sys, time
sys.path.insert(0, '/root/reference')
import warnings; warnings.simplefilter("ignore")
import numpy as np
import pythtb as ref
sys.path.insert(0, '/root/repo')
from tests import models as M
out = {}
for name, model, occ in (("haldane", M.haldane(ref, 0.0), [0]), ("kane_mele", M.kane_mele(ref, "odd"), [0, 1])):
    nmesh = 65
    w = ref.wf_array(model, [nmesh, nmesh])
    t0 = time.perf_counter(); w.solve_on_grid([-0.5, -0.5]); t1 = time.perf_counter()
    f = w.berry_flux(occ); t2 = time.perf_counter()
    n = (nmesh - 1) ** 2
    out[name] = dict(mesh=nmesh, solve_kpts_per_s=n / (t1 - t0), flux_plaq_per_s=n / (t2 - t1), step_kpts_per_s=n / (t2 - t0), chern=f / (2 * np.pi))
import json; print(json.dumps(out, indent=1))
