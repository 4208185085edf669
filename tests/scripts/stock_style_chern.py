"""A script written the way PythTB users write them — ``from pythtb import *`` and nothing else from this
repository — used by tests/test_backend_flag.py to show that the PYTHTB_BACKEND flag alone picks the
implementation.  (The physics follows the reference's haldane_bp example: Haldane model at delta = 0,
31 x 31 mesh, Berry phases along k_x and the Chern number of the lower band; and its kane_mele example:
hybrid Wannier centres of the Z2-odd phase.)  Prints one JSON line."""
import json
import sys

import numpy as np

from pythtb import *  # noqa: F401,F403


def haldane():
    lat = [[1.0, 0.0], [0.5, np.sqrt(3.0) / 2.0]]
    orb = [[1.0 / 3.0, 1.0 / 3.0], [2.0 / 3.0, 2.0 / 3.0]]
    model = tb_model(2, 2, lat, orb)          # noqa: F405
    t, t2 = -1.0, 0.15 * np.exp(1.0j * np.pi / 2.0)
    model.set_onsite([0.0, 0.0])
    model.set_hop(t, 0, 1, [0, 0])
    model.set_hop(t, 1, 0, [1, 0])
    model.set_hop(t, 1, 0, [0, 1])
    model.set_hop(t2, 0, 0, [1, 0])
    model.set_hop(t2, 1, 1, [1, -1])
    model.set_hop(t2, 1, 1, [0, 1])
    model.set_hop(t2.conjugate(), 1, 1, [1, 0])
    model.set_hop(t2.conjugate(), 0, 0, [1, -1])
    model.set_hop(t2.conjugate(), 0, 0, [0, 1])
    return model


def main():
    model = haldane()
    arr = wf_array(model, [31, 31])           # noqa: F405
    gaps = arr.solve_on_grid([-0.5, -0.5])
    phi_a = arr.berry_phase([0], 0, contin=True)
    flux_a = arr.berry_flux([0])
    # second route of the reference example: fill the array point by point
    kx = np.linspace(-0.5, 0.5, num=31)
    arr2 = wf_array(model, [31, 31])          # noqa: F405
    for i in range(31):
        for j in range(31):
            (_, evec) = model.solve_one([kx[i], kx[j]], eig_vectors=True)
            arr2[i, j] = evec
    arr2.impose_pbc(0, 0)
    arr2.impose_pbc(1, 1)
    flux_a2 = arr2.berry_flux([0])
    path = model.k_path([[0.0, 0.0], [2.0 / 3.0, 1.0 / 3.0], [0.5, 0.5]], 21, report=False)[0]
    evals = model.solve_all(path)
    out = dict(backend=sys.modules["pythtb"].get_backend() if hasattr(sys.modules["pythtb"], "get_backend") else "stock",
               tb_model_module=tb_model.__module__,     # noqa: F405
               gaps=np.asarray(gaps).tolist(), phi_a1=np.asarray(phi_a).tolist(), flux_a1=float(flux_a),
               flux_a2=float(flux_a2), evals=np.asarray(evals).tolist())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
