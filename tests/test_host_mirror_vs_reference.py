"""Host mirror against the LIVE reference (build container only).

``pythtb_b200.tb_model`` re-implements the reference's host-side bookkeeping — ``set_onsite/set_hop`` modes,
``cut_piece``, ``reduce_dim``, ``make_supercell``, ``change_nonperiodic_vector``, ``remove_orb``,
``k_path``, ``k_uniform_mesh`` (pythtb.py:186-560, 1105-2026) — because that state IS the kernel input.  When
``/root/reference`` is present (it is not on the GPU box; these tests are then skipped) the same call sequences
are driven through both classes and the resulting model state must agree entry by entry: lattice, orbitals,
periodic directions, site energies and the hopping list in order.  No GPU, no numerics beyond host arithmetic.
"""
import io
import os
import sys
import contextlib

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "pythtb.py")),
                                reason="the reference tree is only present in the build container")


def _mods():
    import warnings
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pythtb as ref
    import pythtb_b200 as mine
    return ref, mine


def _state(m):
    hops = [(np.array(h[0], dtype=complex), int(h[1]), int(h[2]),
             None if m._dim_k == 0 else np.array(h[3], dtype=int)) for h in m._hoppings]
    return dict(dim_k=m._dim_k, dim_r=m._dim_r, nspin=m._nspin, norb=m._norb, nsta=m._nsta,
                per=list(m._per), lat=np.array(m._lat), orb=np.array(m._orb),
                site=np.array(m._site_energies), hops=hops)


def _same(a, b, what):
    for key in ("dim_k", "dim_r", "nspin", "norb", "nsta", "per"):
        assert a[key] == b[key], (what, key, a[key], b[key])
    for key in ("lat", "orb", "site"):
        assert a[key].shape == b[key].shape, (what, key)
        assert np.max(np.abs(a[key] - b[key]), initial=0.0) < 1e-13, (what, key)
    assert len(a["hops"]) == len(b["hops"]), (what, len(a["hops"]), len(b["hops"]))
    for n, (x, y) in enumerate(zip(a["hops"], b["hops"])):
        assert x[1:3] == y[1:3], (what, n)
        assert np.max(np.abs(x[0] - y[0])) < 1e-13, (what, n)
        if x[3] is not None:
            assert np.array_equal(x[3], y[3]), (what, n)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _build_zoo(mod):
    from tests import models as M
    zoo = {
        "haldane": M.haldane(mod, 0.2),
        "kane_mele": M.kane_mele(mod, "odd"),
        "cubic": M.cubic_bulk(mod),
        "random_spin": M.random_model(mod, norb=3, dim=3, nhop=14, nspin=2, seed=9),
        "random": M.random_model(mod, norb=4, dim=2, nhop=12, nspin=1, seed=10),
    }
    # every set_onsite / set_hop mode, spinor value formats, conjugate pairs (pythtb.py:186-515)
    m = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.8, 0.2, 3.5], [-0.1, -3.1, -1.2]],
                     [[0.3, 0.2, 0.1], [0.1, 0.3, 0.8], [0.2, 0.4, 0.3]], per=[0, 1], nspin=2)
    m.set_onsite([0.1, [0.2, 0.0, 0.1, 0.3], [[0.5, 0.1j], [-0.1j, -0.5]]])
    m.set_onsite(0.4, 1, mode="add")
    m.set_onsite([0.0, 0.1, 0.0, 0.0], 2, mode="reset")
    m.set_hop(0.24, 0, 1, [1, 2, 0])
    m.set_hop([0.1, 0.2, 0.3, 0.4], 0, 1, [3, 2, 0], mode="reset")
    m.set_hop([[0.1, 0.2j], [0.3, -0.4]], 1, 2, [2, 3, 0])
    m.set_hop(0.5j, 1, 2, [2, 3, 0], mode="add")
    m.set_hop((-0.34 + 0.3j) * 0.7, 2, 0, [-1, 2, 0], allow_conjugate_pair=True)
    m.set_hop((-0.34 - 0.3j) * 0.3, 0, 2, [1, -2, 0], allow_conjugate_pair=True)
    zoo["modes_spin"] = m
    return zoo


def test_model_definition_and_surgery_match_the_reference():
    ref, mine = _mods()
    za, zb = _quiet(_build_zoo, ref), _quiet(_build_zoo, mine)
    for name in za:
        a, b = za[name], zb[name]
        _same(_state(a), _state(b), name)
        for d in range(a._dim_k):
            pd = a._per[d]
            for glue in (False, True):
                if glue and a._dim_k == 1 and a._norb * 3 < 3:
                    continue
                _same(_state(_quiet(a.cut_piece, 3, pd, glue_edgs=glue)),
                      _state(_quiet(b.cut_piece, 3, pd, glue_edgs=glue)), (name, "cut_piece", pd, glue))
            _same(_state(_quiet(a.reduce_dim, pd, 0.37)), _state(_quiet(b.reduce_dim, pd, 0.37)), (name, "reduce_dim", pd))
        _same(_state(_quiet(a.remove_orb, [0])), _state(_quiet(b.remove_orb, [0])), (name, "remove_orb"))
        if a._dim_k == a._dim_r == 2:
            for sc in ([[2, 1], [-1, 2]], [[1, 0], [0, 3]], [[2, 0], [1, 1]]):
                for home in (True, False):
                    _same(_state(_quiet(a.make_supercell, sc, to_home=home, to_home_suppress_warning=True)),
                          _state(_quiet(b.make_supercell, sc, to_home=home, to_home_suppress_warning=True)),
                          (name, "make_supercell", sc, home))
            for who in (a, b):                          # non-integer super-lattice: both refuse (pythtb.py:1508-1509)
                with pytest.raises(Exception, match="must be integers"):
                    _quiet(who.make_supercell, [[2.5, 0], [0, 1]])
        if a._dim_k == a._dim_r == 3:
            sc = [[1, 1, 0], [0, 2, 0], [0, 1, 2]]
            _same(_state(_quiet(a.make_supercell, sc, to_home=True, to_home_suppress_warning=True)),
                  _state(_quiet(b.make_supercell, sc, to_home=True, to_home_suppress_warning=True)),
                  (name, "make_supercell3"))
        # a ribbon, then another non-periodic vector (pythtb.py:1313-1438)
        if a._dim_k == 2 and a._dim_r == 2:
            ra, rb = _quiet(a.cut_piece, 4, 1), _quiet(b.cut_piece, 4, 1)
            for vec in (None, [-1.3, 4.8]):
                _same(_state(_quiet(ra.change_nonperiodic_vector, 1, vec, to_home_suppress_warning=True)),
                      _state(_quiet(rb.change_nonperiodic_vector, 1, vec, to_home_suppress_warning=True)),
                      (name, "change_nonperiodic_vector", vec))


def test_k_helpers_match_the_reference():
    ref, mine = _mods()
    za, zb = _quiet(_build_zoo, ref), _quiet(_build_zoo, mine)
    for name in ("haldane", "cubic", "random"):
        a, b = za[name], zb[name]
        mesh = [4, 3, 5][:a._dim_k]
        assert np.array_equal(a.k_uniform_mesh(mesh), b.k_uniform_mesh(mesh)), name
        nodes = np.random.RandomState(3).rand(4, a._dim_k).tolist()      # the reference compares kpts with strings: lists only
        for x, y in zip(_quiet(a.k_path, nodes, 23, report=False), _quiet(b.k_path, nodes, 23, report=False)):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) < 1e-13, name
    from tests import models as M
    ra, rb = M.bn_ribbon(ref, 3), M.bn_ribbon(mine, 3)
    for spec in ("full", "fullc", "half", [[-0.3], [0.1], [0.6]]):
        for x, y in zip(_quiet(ra.k_path, spec, 17, report=False), _quiet(rb.k_path, spec, 17, report=False)):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) < 1e-13, spec


def test_error_behaviour_matches_the_reference():
    """Same exceptions for the same misuse (messages are the reference's)."""
    ref, mine = _mods()
    def bad_calls(mod):
        from tests import models as M
        m = M.haldane(mod)
        m3 = mod.tb_model(2, 3, [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]], [[0.0, 0.0, 0.0]], per=[0, 1])
        out = []
        for fn in (lambda: m.set_hop(1.0, 0, 1, [0, 0]),            # already set
                   lambda: m.set_hop(1.0, 0, 0, [0, 0]),            # on-site as hopping
                   lambda: m.set_hop(1.0, 0, 5, [1, 0]),            # index out of scope
                   lambda: m.set_onsite([1.0]),                     # wrong length
                   lambda: m.cut_piece(0, 0),                       # non-positive num
                   lambda: m3.cut_piece(2, 2),                      # not a periodic direction
                   lambda: m.reduce_dim(3, 0.1),
                   lambda: mod.tb_model(3, 2),                      # dim_r < dim_k
                   lambda: mod.tb_model(2, 2, [[1, 0], [0, 1]], [[0, 0]], nspin=3),
                   lambda: mod.wf_array(m, [1, 5]),                 # mesh extent < 2
                   lambda: m.k_uniform_mesh([3]),
                   lambda: m.remove_orb([0, 0])):
            try:
                _quiet(fn)
                out.append(None)
            except Exception as e:       # noqa: BLE001
                out.append((type(e).__name__, " ".join(str(e).split())))
        return out
    a, b = bad_calls(ref), bad_calls(mine)
    for n, (x, y) in enumerate(zip(a, b)):
        assert (x is None) == (y is None), (n, x, y)
        if x is not None:
            assert x[0] == y[0], (n, x, y)
            assert x[1] == y[1], (n, x, y)


def test_wf_array_error_behaviour_matches_the_reference():
    """wf_array argument checks (pythtb.py:2448-2459, 2649-2660, 2998, 3027, 3128-3130) — the product's host
    class with the numpy oracle injected as its engine, against the reference class."""
    ref, _ = _mods()
    from tests import oracle_api
    from tests import models as M

    def bad_calls(mod):
        m = M.haldane(mod)
        w = mod.wf_array(m, [5, 4])
        _quiet(w.solve_on_grid, [0.0, 0.0])
        w1 = mod.wf_array(M.three_site(mod), [6])
        _quiet(w1.solve_on_grid, [0.0])
        out = []
        for fn in (lambda: w[7, 0], lambda: w[0], lambda: w[0, 1, 2], lambda: w[0.5, 1],
                   lambda: w.berry_phase([0], dir=2), lambda: w.berry_phase([0]),
                   lambda: w1.berry_flux([0]), lambda: w.berry_flux([0], dirs=[0, 0]),
                   lambda: w.impose_pbc(4, 0), lambda: w.impose_pbc(0, 2), lambda: w.impose_loop(4),
                   lambda: mod.wf_array(m, [5]).solve_on_grid([0.0]),
                   lambda: mod.wf_array(m, [5, 5], nsta_arr=1).solve_on_grid([0.0, 0.0]),
                   lambda: w.choose_states([[0]]), lambda: w.position_matrix([0, 0], [0], 0)):
            try:
                _quiet(fn)
                out.append(None)
            except Exception as e:       # noqa: BLE001
                out.append((type(e).__name__, " ".join(str(e).split())))
        ok = [np.asarray(w.berry_phase([0], 1)).shape, np.asarray(w.berry_flux([0, 1])).shape,
              w.choose_states([1])._wfs.shape, w.empty_like(nsta_arr=1)._wfs.shape, np.asarray(w1.berry_phase([0, 1])).shape]
        return out, ok

    # (mesh_dir beyond the array's own rank but <= 3 is NOT compared: the reference then silently overwrites
    # the state / orbital axes, pythtb.py:2738-2747; this package raises "Wrong value of mesh_dir.")
    (a, oka), (b, okb) = bad_calls(ref), bad_calls(oracle_api)
    assert oka == okb
    for n, (x, y) in enumerate(zip(a, b)):
        assert (x is None) == (y is None), (n, x, y)
        if x is not None:
            assert x == y, (n, x, y)


@pytest.mark.parametrize("spec", [dict(norb=3, dim=2, nhop=8, nspin=1, seed=31), dict(norb=2, dim=2, nhop=6, nspin=2, seed=32),
                                  dict(norb=5, dim=2, nhop=14, nspin=1, seed=33), dict(norb=3, dim=3, nhop=12, nspin=1, seed=34),
                                  dict(norb=4, dim=1, nhop=7, nspin=2, seed=35)])
def test_random_models_berry_quantities_oracle_vs_live_reference(spec):
    """The oracle (through the product's host classes) against the live reference on seeded random models:
    every wf_array result the goldens pin on physical models, here on generic ones — gaps, Berry phases in
    both branches along every axis, fluxes over every pair of axes, per plaquette and summed."""
    ref, _ = _mods()
    from tests import compare, models as M, oracle_api
    res = []
    for mod in (ref, oracle_api):
        m = _quiet(M.random_model, mod, **spec)
        dim = spec["dim"]
        mesh = [6, 5, 4][:dim]
        w = mod.wf_array(m, mesh)
        out = dict(gaps=_quiet(w.solve_on_grid, [0.1, -0.2, 0.3][:dim]))
        nocc = max(1, m._nsta // 2)
        for occ in ([0], list(range(nocc))):
            tag = "_%d" % len(occ)
            for d in range(dim):
                out["phase%d%s" % (d, tag)] = w.berry_phase(occ, d, contin=False)
                if len(occ) > 1:
                    out["wilson%d%s" % (d, tag)] = w.berry_phase(occ, d, contin=False, berry_evals=True)
            for d0 in range(dim):
                for d1 in range(dim):
                    if d0 != d1:
                        out["plaq%d%d%s" % (d0, d1, tag)] = w.berry_flux(occ, dirs=[d0, d1], individual_phases=True)
                        out["flux%d%d%s" % (d0, d1, tag)] = np.array(w.berry_flux(occ, dirs=[d0, d1]))
        res.append(out)
    bad = compare.compare_case("random_live", res[1], res[0])
    assert not bad, "\n".join(bad)
