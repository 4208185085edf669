"""The reference's own unit and equivalence tests, restated against this package:

  /root/reference/tests/test_pythtb.py:20-72          (0-D known answers, k_path shapes, determinism)
  /root/reference/tests/test_tbmodel/test_dimr_dimk_different.py:8-83
  /root/reference/tests/test_tbmodel/test_different_modes.py:10-50
  /root/reference/tests/test_tbmodel/test_spin.py:9-52
  /root/reference/tests/test_tbmodel/test_non_periodic.py:9-166

Each test runs twice: on the CPU through the product's host classes with the numpy oracle
injected as the engine (``tests/oracle_api.py``; host logic only), and — marked ``gpu`` —
through ``pythtb_b200`` -> C ABI -> sm_100a kernels.  The physics is built in several
equivalent ways (dim_r != dim_k, set/reset/add modes, spinor vs doubled spinless model,
supercell vs finite stack) and must give the same Berry phases, energies and charge centres.
"""
import io
import contextlib

import numpy as np
import pytest


def _oracle_mod():
    from tests import oracle_api
    return oracle_api


def _gpu_mod():
    import pythtb_b200
    return pythtb_b200


MODS = [pytest.param(_oracle_mod, id="oracle-engine"),
        pytest.param(_gpu_mod, id="b200", marks=pytest.mark.gpu)]


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------ tests/test_pythtb.py
@pytest.mark.parametrize("get", MODS)
def test_known_answers_and_shapes(get):
    mod = get()
    import pythtb_b200
    assert isinstance(pythtb_b200.__version__, str) and pythtb_b200.__version__
    m = mod.tb_model(0, 1, [[1.0]], [[0.0]])
    m.set_onsite([2.5])
    ev = m.solve_all()
    assert ev.shape == (1,) and np.allclose(ev, [2.5])
    m = mod.tb_model(0, 1, [[1.0]], [[0.0], [0.5]])
    m.set_onsite([0.0, 0.0])
    m.set_hop(3.0, 0, 1)
    ev = np.sort(m.solve_all())
    assert ev.shape == (2,) and np.allclose(ev, [-3.0, 3.0])
    m = mod.tb_model(1, 1, [[1.0]], [[0.0]])
    m.set_onsite([0.0])
    k_vec, k_dist, k_node = _quiet(m.k_path, [[0.0], [0.5]], 5)
    assert k_vec.shape == (5, 1) and k_dist.shape == (5,) and len(k_node) == 2
    assert k_dist[0] == pytest.approx(0.0) and k_dist[-1] > 0.0
    m = mod.tb_model(0, 1, [[1.0]], [[0.0]])
    m.set_onsite([-1.0])
    assert np.array_equal(m.solve_all(), m.solve_all())


# -------------------------------------------- tests/test_tbmodel: the shared driver (:44-83)
def _generic_test_of_models(mod, models, use_dir, use_occ):
    val = []
    for ii, m in enumerate(models):
        arr = mod.wf_array(m, [11, 11])
        arr.solve_on_grid([-0.5, -0.5])
        val.append(arr.berry_phase(use_occ[ii], 1, contin=True))
    for v in val[1:]:
        assert np.all(np.isclose(val[0], v))
    val = [m.solve_one([0.123, 0.523]) for m in models]
    for v in val[1:]:
        assert np.all(np.isclose(val[0], v))
    val = []
    for ii, m in enumerate(models):
        cut = m.cut_piece(4, use_dir[ii], glue_edgs=False)
        ev, evec = cut.solve_one([0.214], eig_vectors=True)
        pos = np.asarray(cut.position_expectation(evec, use_dir[ii]))
        # per-state expectations inside an exactly degenerate level depend on the solver's basis (the cut
        # pieces are stacks of decoupled, hence degenerate, cells; the reference's own test_spin fails for
        # this reason, SURVEY.md section 4, and its other tests pass only because LAPACK happens to return
        # the same basis for all models): compare the sums over groups of numerically equal eigenvalues
        groups = np.concatenate([[0], np.cumsum(np.diff(ev) > 1.0e-8)])
        pos = np.array([pos[groups == g].sum() for g in range(groups[-1] + 1)])
        val.append(pos)
    for v in val[1:]:
        assert val[0].shape == v.shape and np.all(np.isclose(val[0], v))


@pytest.mark.parametrize("get", MODS)
def test_dimr_dimk_different(get):
    mod = get()
    m0 = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.1, 3.1, 1.2], [0.8, 0.2, 3.5]],
                      [[0.3, 0.1, 0.2], [0.1, 0.8, 0.3], [0.2, 0.3, 0.4]], per=[0, 2])
    m0.set_onsite([-2.3, 0.5, 0.1])
    m0.set_hop(0.24, 0, 1, [1, 0, 2])
    m0.set_hop(0.42, 0, 1, [3, 0, 2])
    m0.set_hop(-0.12, 1, 2, [2, 0, 3])
    m0.set_hop(-0.34, 2, 0, [-1, 0, 2])
    m1 = mod.tb_model(2, 2, [[3.0, 0.4], [0.8, 3.5]], [[0.3, 0.2], [0.1, 0.3], [0.2, 0.4]])
    m1.set_onsite([-2.3, 0.5, 0.1])
    m1.set_hop(0.24, 0, 1, [1, 2])
    m1.set_hop(0.42, 0, 1, [3, 2])
    m1.set_hop(-0.12, 1, 2, [2, 3])
    m1.set_hop(-0.34, 2, 0, [-1, 2])
    m2 = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.8, 0.2, 3.5], [-0.1, -3.1, -1.2]],
                      [[0.3, 0.2, 0.1], [0.1, 0.3, 0.8], [0.2, 0.4, 0.3]], per=[0, 1])
    m2.set_onsite([-2.3, 0.5, 0.1])
    m2.set_hop(0.24, 0, 1, [1, 2, 0])
    m2.set_hop(0.42, 0, 1, [3, 2, 0])
    m2.set_hop(-0.12, 1, 2, [2, 3, 0])
    m2.set_hop(-0.34, 2, 0, [-1, 2, 0])
    _generic_test_of_models(mod, [m0, m1, m2], use_dir=[2, 1, 1], use_occ=[[0], [0], [0]])


@pytest.mark.parametrize("get", MODS)
def test_different_modes(get):
    mod = get()
    m0 = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.1, 3.1, 1.2], [0.8, 0.2, 3.5]],
                      [[0.3, 0.1, 0.2], [0.1, 0.8, 0.3], [0.2, 0.3, 0.4]], per=[0, 2])
    m0.set_onsite([-2.3, 0.5, 0.1])
    m0.set_hop(0.24, 0, 1, [1, 0, 2], mode="set")
    m0.set_hop(0.42, 0, 1, [3, 0, 2])
    m0.set_hop(-0.12, 1, 2, [2, 0, 3])
    m0.set_hop(-0.34 + 0.3j, 2, 0, [-1, 0, 2])
    m1 = mod.tb_model(2, 2, [[3.0, 0.4], [0.8, 3.5]], [[0.3, 0.2], [0.1, 0.3], [0.2, 0.4]])
    m1.set_onsite(-2.3, 0)
    m1.set_onsite(0.5, 1)
    m1.set_onsite(9.1, 2, mode="reset")
    m1.set_onsite(0.07, 2, mode="reset")
    m1.set_onsite(0.03, 2, mode="add")
    m1.set_hop(99.24, 0, 1, [1, 2], mode="set")
    m1.set_hop(0.04, 0, 1, [1, 2], mode="reset")
    m1.set_hop(0.08, 0, 1, [1, 2], mode="add")
    m1.set_hop(0.12, 0, 1, [1, 2], mode="add")
    m1.set_hop(0.42, 0, 1, [3, 2])
    m1.set_hop(-0.12, 1, 2, [2, 3])
    m1.set_hop(-0.34 + 0.3j, 2, 0, [-1, 2])
    m2 = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.8, 0.2, 3.5], [-0.1, -3.1, -1.2]],
                      [[0.3, 0.2, 0.1], [0.1, 0.3, 0.8], [0.2, 0.4, 0.3]], per=[0, 1])
    m2.set_onsite([-2.3, 0.5, 0.1])
    m2.set_hop(0.24, 0, 1, [1, 2, 0])
    m2.set_hop(99.42, 0, 1, [3, 2, 0], mode="reset")
    m2.set_hop(0.42, 0, 1, [3, 2, 0], mode="reset")
    m2.set_hop(-0.12, 1, 2, [2, 3, 0])
    m2.set_hop((-0.34 + 0.3j) * 0.7, 2, 0, [-1, 2, 0], allow_conjugate_pair=True)
    m2.set_hop((-0.34 - 0.3j) * 0.3, 0, 2, [1, -2, 0], allow_conjugate_pair=True)
    _generic_test_of_models(mod, [m0, m1, m2], use_dir=[2, 1, 1], use_occ=[[0], [0], [0]])


@pytest.mark.parametrize("get", MODS)
def test_spin(get):
    mod = get()
    orb6 = [[0.3, 0.1, 0.2]] * 2 + [[0.1, 0.8, 0.3]] * 2 + [[0.2, 0.3, 0.4]] * 2
    m0 = mod.tb_model(2, 3, [[3.0, 0.1, 0.4], [0.1, 3.1, 1.2], [0.8, 0.2, 3.5]], orb6, nspin=1, per=[0, 2])
    m0.set_onsite([-2.3, -2.3, 0.5, 0.5, 0.1, 0.1])
    m0.set_hop(0.11 + 0.41, 0, 2, [1, 0, 2])
    m0.set_hop(0.11 - 0.41, 1, 3, [1, 0, 2])
    m0.set_hop(0.21 - 0.31j, 0, 3, [1, 0, 2])
    m0.set_hop(0.21 + 0.31j, 1, 2, [1, 0, 2])
    m0.set_hop(0.42, 0, 2, [3, 0, 2])
    m0.set_hop(0.42, 1, 3, [3, 0, 2])
    m0.set_hop(-0.12, 2, 4, [2, 0, 3])
    m0.set_hop(-0.12, 3, 5, [2, 0, 3])
    m0.set_hop(-0.34 + 0.29, 4, 0, [-1, 0, 2])
    m0.set_hop(-0.34 - 0.29, 5, 1, [-1, 0, 2])
    m0.set_hop(0.21 + 0.14j, 4, 1, [-1, 0, 2])
    m0.set_hop(0.21 - 0.14j, 5, 0, [-1, 0, 2])
    m1 = mod.tb_model(2, 2, [[3.0, 0.4], [0.8, 3.5]], [[0.3, 0.2], [0.1, 0.3], [0.2, 0.4]], nspin=2)
    m1.set_onsite([-2.3, 0.5, 0.1])
    m1.set_hop([[0.11 + 0.41, 0.21 - 0.31j], [0.21 + 0.31j, 0.11 - 0.41]], 0, 1, [1, 2])
    m1.set_hop(0.42, 0, 1, [3, 2])
    m1.set_hop(-0.12, 1, 2, [2, 3])
    m1.set_hop([-0.34, 0.21, -0.14, 0.29], 2, 0, [-1, 2])
    _generic_test_of_models(mod, [m0, m1], use_dir=[2, 1], use_occ=[[0, 1], [0, 1]])


# ------------------------------------------------ tests/test_tbmodel/test_non_periodic.py
@pytest.mark.parametrize("get", MODS)
def test_non_periodic_charge_centres(get):
    mod = get()
    bulk = mod.tb_model(2, 2, [[2.3, -0.2], [1.9, 2.4]], [[0.15, 0.34], [0.29, 0.65]], per=[0, 1])
    t_first, t_second, delta = 0.8 + 0.6j, 2.0, -0.8
    bulk.set_onsite([-delta, delta])
    bulk.set_hop(t_second, 0, 0, [1, 0])
    bulk.set_hop(t_second, 1, 1, [1, 0])
    bulk.set_hop(t_first, 0, 1, [0, 0])
    bulk.set_hop(t_first, 1, 0, [1, 0])
    numk, num_wire = 21, 3

    arr = mod.wf_array(bulk, [numk, 100])
    arr.solve_on_grid([0.0, 0.0])
    ph0 = np.mean(arr.berry_phase([0], dir=0, contin=True)[:-1])
    ph1 = np.mean(arr.berry_phase([0], dir=1, contin=True)[:-1])
    loc = (ph0 / (2 * np.pi)) * bulk._lat[0] + (ph1 / (2 * np.pi)) * bulk._lat[1] + bulk._lat[1]
    loc_three = (loc + (loc + bulk._lat[1]) + (loc + 2 * bulk._lat[1])) / float(num_wire)

    sc = bulk.make_supercell([[1, 0], [0, num_wire]], to_home=False, to_home_suppress_warning=True)
    sarr = mod.wf_array(sc, [numk, 100])
    sarr.solve_on_grid([0.0, 0.0])
    s0 = np.mean(sarr.berry_phase(range(num_wire), dir=0, contin=True)[:-1])
    s1 = np.mean(sarr.berry_phase(range(num_wire), dir=1, contin=True)[:-1])
    sloc = (s0 / (2 * np.pi)) * sc._lat[0] + (s1 / (2 * np.pi)) * sc._lat[1] + sc._lat[0] + 2 * sc._lat[1]
    assert np.allclose(loc_three, sloc / float(num_wire), rtol=1.0e-5)

    def centres_01(m, num_bands):
        wfa = mod.wf_array(m, [numk])
        wfa.solve_on_grid([0.0])
        p0 = wfa.berry_phase(range(num_bands), dir=0, contin=True)
        pos1 = np.mean([np.sum(wfa.position_expectation([i], range(num_bands), dir=1)) for i in range(numk - 1)])
        return (p0 / (2 * np.pi)) * m._lat[0] + pos1 * m._lat[1], m._lat[0]

    fin = bulk.cut_piece(num=num_wire, fin_dir=1, glue_edgs=False)
    c, per0 = centres_01(fin, num_wire)
    assert np.allclose(loc_three, (c + per0) / float(num_wire), rtol=1.0e-5)
    fin_orth = fin.change_nonperiodic_vector(np_dir=1, new_latt_vec=None, to_home_suppress_warning=True)
    c, per0 = centres_01(fin_orth, num_wire)
    assert np.allclose(loc_three, (c + 5 * per0) / float(num_wire), rtol=1.0e-3)
    fin_arb = fin.change_nonperiodic_vector(np_dir=1, new_latt_vec=[-1.3, 4.8], to_home_suppress_warning=True)
    c, per0 = centres_01(fin_arb, num_wire)
    assert np.allclose(loc_three, (c + 6 * per0) / float(num_wire), rtol=1.0e-3)
