"""The backend flag (BASELINE north_star: "existing scripts switch paths with a flag").

``shim/pythtb`` is an ``import pythtb`` that re-exports either ``pythtb_b200`` (PYTHTB_BACKEND=b200, the
default) or the stock PythTB (PYTHTB_BACKEND=reference).  Covered here:

* CPU: what the shim exports, run-time switching, loud failures; and — in the build container, where
  /root/reference exists — the reference's own ``tests/test_examples/*/*/run.py`` executed UNMODIFIED under
  both values of the flag (the b200 host classes served by the numpy oracle, since there is no GPU here),
  outputs compared.
* GPU (``-m gpu``): a stock-style script (``from pythtb import *``) run under PYTHTB_BACKEND=b200 against the
  committed reference fixtures; the process must have loaded libtbk_b200.so.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import compare

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim")
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
HAVE_REF = os.path.exists(os.path.join(REF, "pythtb.py"))


def _run(code_or_path, backend, extra_path=(), args=(), is_code=True, timeout=600):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SHIM, ROOT] + list(extra_path))
    env["PYTHTB_BACKEND"] = backend
    env.pop("PYTHTB_REFERENCE", None)
    cmd = [sys.executable, "-W", "ignore"] + (["-c", code_or_path] if is_code else [code_or_path]) + list(args)
    return subprocess.run(cmd, env=env, cwd="/tmp", stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)


def test_shim_exports_and_switches():
    res = _run("import pythtb, pythtb_b200\n"
               "assert pythtb.get_backend() == 'b200' and pythtb.tb_model is pythtb_b200.tb_model\n"
               "assert pythtb.wf_array is pythtb_b200.wf_array and pythtb.w90 is pythtb_b200.w90\n"
               "from pythtb import *\n"
               "assert tb_model is pythtb_b200.tb_model\n"
               "try:\n    pythtb.set_backend('nope')\nexcept ValueError: print('valueerror')\n"
               "try:\n    pythtb.set_backend('reference')\nexcept ImportError: print('importerror')\n"
               "assert pythtb.get_backend() == 'b200'\n", "b200")
    assert res.returncode == 0, res.stderr
    assert "valueerror" in res.stdout and "importerror" in res.stdout
    res = _run("import pythtb", "reference")          # no stock PythTB on this path: loud
    assert res.returncode != 0 and "no stock PythTB" in res.stderr


_DRIVER = r"""
import io, json, sys, contextlib, runpy
import numpy as np
import pythtb
if pythtb.get_backend() == "b200" and sys.argv[2] == "oracle":
    # CPU-only container: the product's host classes with the numpy oracle as their engine (test seam)
    from tests import oracle_api
    import pythtb_b200.model as _m
    _m.tb_model._engine_factory = staticmethod(lambda: oracle_api._ENGINE)
with contextlib.redirect_stdout(io.StringIO()):
    ns = runpy.run_path(sys.argv[1])
    out = ns["run"]()
out = out if isinstance(out, tuple) else (out,)
def conv(x):
    if isinstance(x, (tuple, list)):
        return [conv(y) for y in x]
    return None if x is None else np.asarray(x, dtype=float).tolist()
print(json.dumps(dict(backend=pythtb.get_backend(), module=pythtb.tb_model.__module__, out=[conv(x) for x in out])))
"""


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("rel,kind", [
    ("haldane/haldane_bp/run.py", "phase"),
    ("haldane/haldane/run.py", "value"),
    ("checkerboard/checkerboard/run.py", "value"),
    ("graphene/cone/run.py", "phase"),
    ("kane_mele/kane_mele/run.py", "phase"),
    ("slab/cubic_slab_hwf/run.py", "phase"),
    ("boron_nitride/bn_ribbon_berry/run.py", "phase"),
])
def test_reference_run_py_unmodified_under_the_flag(rel, kind):
    path = os.path.join(REF, "tests", "test_examples", rel)
    got = {}
    for backend in ("reference", "b200"):
        res = _run(_DRIVER, backend, extra_path=[REF] if backend == "reference" else [], args=[path, "oracle"])
        assert res.returncode == 0, res.stderr[-2000:]
        got[backend] = json.loads(res.stdout.strip().splitlines()[-1])
        assert got[backend]["backend"] == backend
    assert got["reference"]["module"] != got["b200"]["module"] and got["b200"]["module"].startswith("pythtb_b200")

    a, b = got["reference"]["out"], got["b200"]["out"]          # one entry per value run() returned
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if x is None:
            continue
        x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
        assert x.shape == y.shape
        if x.dtype == object or x.size == 0:
            continue
        if kind == "phase":
            dev = np.minimum(np.abs(compare.circ_diff(x, y, 2 * np.pi)), np.abs(x - y))
            assert np.max(dev) < 1e-8
        else:
            assert np.max(np.abs(x - y)) < 1e-10 * max(1.0, np.max(np.abs(x)))


@pytest.mark.gpu
def test_stock_style_script_runs_on_the_b200_backend_by_flag():
    script = os.path.join(ROOT, "tests", "scripts", "stock_style_chern.py")
    res = _run(script, "b200", is_code=False)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["backend"] == "b200" and out["tb_model_module"].startswith("pythtb_b200")
    want = np.load(os.path.join(GOLD, "haldane_bp.npz"))
    assert np.max(np.abs(np.array(out["gaps"]) - want["gaps"])) < 1e-10
    assert np.max(np.abs(compare.circ_diff(out["phi_a1"], want["phi_a1"], 2 * np.pi))) < 1e-8
    for key in ("flux_a1", "flux_a2"):
        assert abs(compare.circ_diff(out[key], float(want[key]), 2 * np.pi)) < 1e-8
    assert abs(abs(out["flux_a1"]) / (2 * np.pi) - 1.0) < 1e-9          # Chern number +-1
    bands = np.load(os.path.join(GOLD, "haldane_bands.npz"))
    assert np.asarray(out["evals"]).shape[0] == 2 and "evals" in bands.files
