"""``wf_array.solve_on_slice`` — the batched, device-resident replacement of the per-point fill loop that
parametric (k, lambda) arrays need (examples/3site_cycle.py:48-90).  The
result must equal the reference's pattern ``w[i_k, i_lambda] = evec[:, i_k]`` (pythtb.py:2662-2672) in
every gauge-invariant quantity, and the golden fixtures of the 3-site pump."""
import os

import numpy as np
import pytest

from tests import compare, models as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_mod():
    from tests import oracle_api
    return oracle_api


def _gpu_mod():
    import pythtb_b200
    return pythtb_b200


MODS = [pytest.param(_oracle_mod, id="oracle-engine"),
        pytest.param(_gpu_mod, id="b200", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("get", MODS)
def test_three_site_pump_filled_by_slices(get):
    """tests/cases.py::case_three_site with the double loop replaced by one call per lambda."""
    mod = get()
    nk, nl = 31, 21
    lam = np.linspace(0, 1, nl, endpoint=True)
    w = mod.wf_array(M.three_site(mod, 0.0), [nk, nl])
    evals = np.zeros((nl, nk, 3))
    for il in range(nl):
        m = M.three_site(mod, lam[il])
        k_vec, _, _ = m.k_path([[-0.5], [0.5]], nk, report=False)
        evals[il] = w.solve_on_slice({1: il}, k_vec, model=m)
    # eigenvalues come back as [free..., band]
    m = M.three_site(mod, lam[5])
    k_vec, _, _ = m.k_path([[-0.5], [0.5]], nk, report=False)
    assert np.max(np.abs(evals[5] - m.solve_all(k_vec).T)) < 1e-12
    w.impose_pbc(0, 0)
    want = np.load(os.path.join(GOLD, "three_site.npz"))
    got = dict(wann_center=w.berry_phase([0], 0) / (2.0 * np.pi), final=np.array(w.berry_flux([0])))
    w.impose_loop(1)
    got["phase_lambda"] = w.berry_phase([0], 1, contin=True)
    got["flux_01"] = np.array(w.berry_flux([0, 1]))
    bad = compare.compare_case("three_site", got, want)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("get", MODS)
def test_slices_of_a_3d_array_match_the_point_loop(get):
    """Two free axes (several launches), a leading pinned axis, negative indices, a spinor model, and a
    slice written into an array that already holds host-side data."""
    mod = get()
    m = M.kane_mele(mod, "odd")
    shape = [3, 4, 5]
    rng = np.random.RandomState(2)
    kpts = rng.rand(4, 5, 2) - 0.5
    a = mod.wf_array(m, shape)
    b = mod.wf_array(m, shape)
    for i in range(4):
        for j in range(5):
            (_, evec) = m.solve_one(kpts[i, j], eig_vectors=True)
            for s in range(3):
                a[s, i, j] = evec
    b[0, 0, 0] = a[0, 0, 0]                      # host-side write first: the slice fill must keep the rest
    ev = b.solve_on_slice({0: 1}, kpts)
    assert ev.shape == (4, 5, 4)
    assert np.max(np.abs(ev[2, 3] - m.solve_one(kpts[2, 3]))) < 1e-12
    for j in range(5):
        b.solve_on_slice({0: -1, 2: j}, kpts[:, j])
        b.solve_on_slice({0: 0, 2: j}, kpts[:, j])
    two_pi = 2.0 * np.pi
    for occ in ([0, 1], [0, 1, 2]):
        for dirs in ([1, 2], [0, 1]):
            pa = a.berry_flux(occ, dirs=dirs, individual_phases=True)
            pb = b.berry_flux(occ, dirs=dirs, individual_phases=True)
            assert np.max(np.abs(compare.circ_diff(pa, pb, two_pi))) <= compare.TOL_PHASE
    a.impose_loop(2)                             # closed strings: the only gauge-invariant ones
    b.impose_loop(2)
    pa = a.berry_phase([0, 1], 2, contin=False)
    pb = b.berry_phase([0, 1], 2, contin=False)
    assert np.max(np.abs(compare.circ_diff(pa, pb, two_pi))) <= compare.TOL_PHASE


@pytest.mark.parametrize("get", [MODS[0]])
def test_solve_on_slice_argument_checks(get):
    mod = get()
    m = M.haldane(mod)
    w = mod.wf_array(m, [4, 5])
    with pytest.raises(Exception):
        w.solve_on_slice({0: 1}, np.zeros((4, 2)))             # free axis 1 has 5 points
    with pytest.raises(IndexError):
        w.solve_on_slice({0: 7}, np.zeros((5, 2)))
    with pytest.raises(Exception):
        w.solve_on_slice({0: 1, 1: 1}, np.zeros((2,)))         # nothing left free
    with pytest.raises(Exception):
        w.solve_on_slice({0: 1}, np.zeros((5, 2)), model=M.kane_mele(mod))   # other state count


@pytest.mark.parametrize("get", MODS)
def test_lazy_reduce_dim_matches_the_reference_algorithm(get):
    """``reduce_dim(remove_k, value, lazy=True)`` (the reduced model's plan derived from the parent's plan,
    pythtb_b200/_plan.py::reduce_plan) against the eager reference algorithm (pythtb.py:1233-1311) and the oracle:
    eigenvalues, gauge-invariant projectors, a Berry phase along the remaining direction — for scalar, spinor and
    3-D models, both phase conventions, nested reductions, and the materialised Python description."""
    from oracle import pythtb_oracle as orc
    mod = get()
    k1 = np.linspace(-0.5, 0.5, 13)[:, None]
    specs = [(M.haldane(mod, 0.2), 1, 0.37), (M.haldane(mod, 0.2), 0, -0.12), (M.kane_mele(mod, "odd"), 0, 0.21),
             (M.random_model(mod, norb=3, dim=2, nhop=8, nspin=1, seed=31), 1, 0.4),
             (M.random_model(mod, norb=7, dim=2, nhop=20, nspin=1, seed=25), 0, 0.77)]
    for model, rk, val in specs:
        for conv in (1, 2):
            model.set_convention(conv)
            lazy = model.reduce_dim(rk, val, lazy=True)
            eager = model.reduce_dim(rk, val)
            assert lazy._dim_k == eager._dim_k == 1 and list(lazy._per) == list(eager._per)
            on_gpu = mod.__name__ == "pythtb_b200"                 # (the oracle engine reads _hoppings: it materialises)
            assert lazy.__dict__.get("_lazy") is not None          # nothing has been materialised so far
            ev_l, vec_l = lazy.solve_all(k1, eig_vectors=True)
            ev_ref = orc.solve_all(eager, k1)
            vec_ref = eager.solve_all(k1, eig_vectors=True)[1]     # same engine, eagerly reduced model (either convention)
            scale = max(1.0, np.max(np.abs(ev_ref)))
            assert np.max(np.abs(ev_l - ev_ref)) <= compare.TOL_EVAL * scale
            n = model._nsta
            gaps = np.min(np.diff(ev_ref, axis=0))
            if gaps > 1e-3:
                a = vec_l.reshape(n, len(k1), n)
                b = np.asarray(vec_ref).reshape(n, len(k1), n)
                pa = np.einsum("bki,bkj->bkij", a, a.conj())
                pb = np.einsum("bki,bkj->bkij", b, b.conj())
                assert np.max(np.abs(pa - pb)) < compare.TOL_PROJ
            # a Berry phase of the reduced model (examples/haldane_hwf-style sweeps)
            wl = mod.wf_array(lazy, [21])
            we = mod.wf_array(eager, [21])
            wl.solve_on_grid([0.0])
            we.solve_on_grid([0.0])
            assert abs(compare.circ_diff(wl.berry_phase([0]), we.berry_phase([0]), 2 * np.pi)) < compare.TOL_PHASE
            assert lazy.__dict__.get("_lazy") is not None or not on_gpu
            # the Python description, on demand, is the reference algorithm's
            assert len(lazy._hoppings) == len(eager._hoppings) and lazy.__dict__.get("_lazy") is None
            assert np.allclose(np.asarray(lazy._site_energies), np.asarray(eager._site_energies))
        model.set_convention(1)
    # nested: 3-D -> 1-D, and down to 0-D
    m3 = M.random_model(mod, norb=3, dim=3, nhop=9, nspin=1, seed=2)
    lz = m3.reduce_dim(2, 0.3, lazy=True).reduce_dim(0, -0.2, lazy=True)
    eg = m3.reduce_dim(2, 0.3).reduce_dim(0, -0.2)
    assert np.max(np.abs(lz.solve_all(k1) - orc.solve_all(eg, k1))) < 1e-10
    lz0 = lz.reduce_dim(1, 0.45, lazy=True)
    assert lz0._dim_k == 0 and np.max(np.abs(lz0.solve_all() - orc.solve_all(eg.reduce_dim(1, 0.45)))) < 1e-10
    # a parent edited after the reduction is detected when the Python description is asked for
    h = M.haldane(mod, 0.2)
    lz = h.reduce_dim(0, 0.2, lazy=True)
    h.set_onsite([0.1, -0.1], mode="reset")
    with pytest.raises(Exception, match="modified since"):
        lz._hoppings
