# cProfile of the e2e step on the GPU box
import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo')
import numpy as np
import pythtb_b200 as tb
from tests import models as M
m = M.haldane(tb, 0.0)
w = tb.wf_array(m, [1025, 1025])
for _ in range(20):
    w.solve_on_grid([-0.5, -0.5]); w.berry_flux([0])
pr = cProfile.Profile(); pr.enable()
for _ in range(2000):
    w.solve_on_grid([-0.5, -0.5]); w.berry_flux([0])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:6000])
