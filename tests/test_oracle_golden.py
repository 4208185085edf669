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
