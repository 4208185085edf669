# One launch of every large-matrix kernel family, for an `ncu --set full` capture (profiles/run_large_ncu.sh):
#   solve_blocked_kernel  (n = 200 ribbon grid solve with eigenvectors: blocked Hermitian eigensolver)
#   link_matrix_kernel    (berry_phase, nocc = 100: overlap GEMM on the FP64 tensor pipe + LU determinant)
#   Wilson spectrum       (berry_evals=True: Newton-Schulz polar factors on the DMMA GEMM, products, eigenphases)
#   position_matrix_dmma / hwf_to_orbital_dmma  (position_hwf_all, nocc = 100)
# Prints the algorithmic flop counts of DESIGN.md / SURVEY.md 8(d) for the shapes used, so that achieved
# FLOP/s = flops / (ncu duration) can be formed per kernel.
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pythtb_b200 as tb
from tests import models as M

ncell = int(os.environ.get("PROF_NCELL", "100"))
nk = int(os.environ.get("PROF_NK", "297"))
rib = M.bn_ribbon(tb, ncell)
n = rib._nsta
nocc = n // 2
w = tb.wf_array(rib, [nk])
w._solve_on_grid_device(np.array([0.0]))
torch.cuda.synchronize()
occ = list(range(nocc))
w.berry_phase(occ, 0, contin=False)
torch.cuda.synchronize()
w.berry_phase(occ, 0, contin=False, berry_evals=True)
torch.cuda.synchronize()
w.position_hwf_all(occ, 1, hwf_evec=True)
torch.cuda.synchronize()
npts, nlinks = nk - 1, nk - 1
print(json.dumps({
    "n": n, "nocc": nocc, "matrices": npts, "links": nlinks,
    "flops_eigh_with_vectors": npts * (40.0 / 3.0) * n ** 3,          # (16/3 + 8) n^3 per matrix
    "flops_overlap": nlinks * 8.0 * nocc ** 2 * n,                    # complex GEMM per link
    "flops_lu": nlinks * (8.0 / 3.0) * nocc ** 3,
    "flops_position_matrix": npts * 8.0 * nocc ** 2 * n,
    "bytes_wfs": nk * n * n * 16,
}))
