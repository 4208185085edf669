# debug helper: full-mesh solve_on_grid vs the numpy oracle, point by point (gauge-invariant projector of band 0)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pythtb_b200 as tb
from tests import models as M, oracle_api
from oracle import pythtb_oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
m = M.haldane(tb, 0.0)
w = tb.wf_array(m, [n, n])
w.solve_on_grid([-0.5, -0.5])
wf = np.array(w._wfs) if hasattr(w, "_wfs") else None
mo = M.haldane(oracle_api, 0.0)
ref, _ = orc.solve_on_grid(mo, [n, n], [-0.5, -0.5])
print(wf.shape, ref.shape)
nrm = np.abs(np.sum(np.abs(wf) ** 2, axis=-1) - 1).max(axis=-1)
print("non-normalised points:", np.argwhere(nrm > 1e-12)[:20], (nrm > 1e-12).sum())
ov = np.abs(np.einsum("ijo,ijo->ij", wf[:, :, 0].conj(), ref[:, :, 0]))
bad = np.argwhere(np.abs(ov - 1) > 1e-9)
print("bad points:", len(bad), bad[:40])
print("max dev", np.abs(ov - 1).max())
# image consistency
print("pbc axis0", np.abs(wf[-1] - wf[0] * np.exp(-2j * np.pi * np.array(m._orb)[:, 0])[None, None, :]).max())
print("pbc axis1", np.abs(wf[:, -1] - wf[:, 0] * np.exp(-2j * np.pi * np.array(m._orb)[:, 1])[None, None, :]).max())
fl = w.berry_flux([0], individual_phases=True)
fr = orc.berry_flux(ref, 2, [0], individual_phases=True)
d = (fl - fr + np.pi) % (2 * np.pi) - np.pi
print("plaq max dev", np.abs(d).max(), np.argwhere(np.abs(d) > 1e-9)[:30], (np.abs(d) > 1e-9).sum())
print("sum", fl.sum() / 2 / np.pi, fr.sum() / 2 / np.pi, w.berry_flux([0]) / 2 / np.pi)
