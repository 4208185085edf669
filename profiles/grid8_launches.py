import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
import bench_extras as BX
m = BX.silicon_model(tb)
w = tb.wf_array(m, [65, 65, 65])
for _ in range(2):
    w._solve_on_grid_device(np.zeros(3)); torch.cuda.synchronize()
