"""The oracle, driven through the product's host classes, must reproduce the
outputs of the unmodified reference stored in tests/golden/ (which were in
turn cross-checked against the reference's own golden_outputs/*.npy)."""
import io
import os
import contextlib

import numpy as np
import pytest

from tests import cases, compare, oracle_api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(cases.ALL_CASES))
def test_oracle_matches_reference(name):
    want = np.load(os.path.join(GOLD, name + ".npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        got = cases.ALL_CASES[name](oracle_api)
    bad = compare.compare_case(name, got, want)
    assert not bad, "\n".join(bad)


def test_golden_log_pins_reference():
    import json
    log = json.load(open(os.path.join(GOLD, "golden_log.json")))
    assert log["pythtb"] == "1.8.0"
    assert log["worst_cross_check_dev"] < 1e-10
    assert len(log["cross_check"]) >= 20


def test_position_hwf_all_host_logic():
    """Host logic of the batched extension wf_array.position_hwf_all, served by the numpy oracle."""
    from tests import models as M, oracle_api as api
    nl = 5
    slab = M.cubic_slab(api, nl)
    bloch = api.wf_array(slab, [4, 5])
    bloch.solve_on_grid([0.0, 0.0])
    occ = list(range(nl))
    hwfc, hwf = bloch.position_hwf_all(occ, 2, hwf_evec=True)
    assert hwfc.shape == (4, 5, nl) and hwf._wfs.shape[:3] == (4, 5, nl)
    val, vec = bloch.position_hwf([1, 2], occ=occ, dir=2, hwf_evec=True, basis="orbital")
    assert np.max(np.abs(hwfc[1, 2] - val)) < 1e-12
    assert np.max(np.abs(hwf._wfs[1, 2] - vec)) < 1e-12
