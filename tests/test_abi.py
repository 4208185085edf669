"""CPU checks of the drop-in boundary: libtbk_b200.so loads and exports every
symbol include/tbk.h declares; the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "tbk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tbk_[a-z0-9_]+)\s*\(", txt)))


def _lib_path():
    from pythtb_b200 import build
    return build.build()


def test_library_exports_header_symbols():
    lib = ctypes.CDLL(_lib_path())
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "libtbk_b200.so does not export " + s


def test_binding_table_matches_header():
    from pythtb_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _header_symbols()
    lib = _lib.load()
    assert lib.tbk_version() >= 100
    assert lib.tbk_last_error() is not None


def test_host_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    from pythtb_b200 import _lib
    lib = _lib.load()
    out = ctypes.c_void_p(0)
    assert lib.tbk_model_create(None, ctypes.byref(out)) == -1
    assert b"null" in lib.tbk_last_error()
    assert lib.tbk_gen_ham(None, None, 1, None, None) == -1
    assert lib.tbk_model_destroy(None) == 0


def test_new_entry_points_validate_arguments_without_gpu():
    """Prepared calls, the device k-mesh and the peer helpers reject bad arguments / are no-ops on NULL."""
    from pythtb_b200 import _lib
    lib = _lib.load()
    out = ctypes.c_void_p(0)
    assert lib.tbk_solve_grid_prepare(None, None, None, 2, 0, 4, 1, None, None, None, None, 0, 0, None, ctypes.byref(out)) == -1
    assert lib.tbk_flux_plane_prepare(None, None, 1, 4, 4, 4, 1, None, None, None, 0, None, ctypes.byref(out)) == -1
    assert lib.tbk_prepared_run(None, None, 0) == -1
    assert lib.tbk_prepared_destroy(None) == 0
    assert lib.tbk_kmesh_uniform(None, 2, None, None) == -1
    assert lib.tbk_peer_barrier(None, None) == 0 and lib.tbk_peer_flush(None, None) == 0 and lib.tbk_peer_defer(None, 1) == 0


def test_lazy_k_mesh_is_the_reference_mesh():
    """KMesh (k_uniform_mesh(..., lazy=True)) materialises to exactly the eager array (pythtb.py:1848-1857)."""
    import numpy as np
    import pythtb_b200
    from pythtb_b200.model import KMesh
    from tests import models as M
    m = M.haldane(pythtb_b200)
    eager = m.k_uniform_mesh([5, 7])
    lazy = m.k_uniform_mesh([5, 7], lazy=True)
    assert isinstance(lazy, KMesh) and lazy.shape == eager.shape == (35, 2) and len(lazy) == 35
    assert np.array_equal(np.asarray(lazy), eager)
    assert np.array_equal(eager[8], [1 / 5.0, 1 / 7.0])
    with pytest.raises(Exception, match="Incorrect size"):
        m.k_uniform_mesh([5, 7, 3], lazy=True)


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pythtb_b200
    from tests import models as M
    m = M.haldane(pythtb_b200)            # model building is host-only and works
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.solve_all([[0.0, 0.0]])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pythtb_b200.wf_array(m, [5, 5])


def test_no_oracle_in_product():
    """The product package must never import the oracle or the host emulation."""
    pkg = os.path.join(ROOT, "pythtb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "hostemu" not in txt.replace("tests/hostemu", ""), f


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/tbk.h compiles as strict C99 and a pure-C program links against the library and gets the
    argument-check answers of the entry points without a GPU (tests/cabi/abi_smoke.c)."""
    import shutil
    import subprocess
    so = _lib_path()
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the toolchain of this image"
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(ROOT, "tests", "cabi", "abi_smoke.c")
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           src, "-o", exe, so, "-Wl,-rpath," + os.path.dirname(so)])
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    assert "abi_smoke: ok" in res.stdout
