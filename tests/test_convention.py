"""Convention II (doc/formalism/pythtb-formalism.tex:341-364) — the second Bloch-phase
convention the north star asks for.  PythTB 1.8.0 implements Convention I only
(pythtb.py:912-916), so no reference fixture exists; the oracle's Convention-II branch is
pinned here on the formalism's identity with the (golden-pinned) Convention-I branch,

    H~_ij(k) = e^{ik.(tau_i - tau_j)} H_ij(k),   C~_j(k) = e^{ik.tau_j} C_j(k),   C~(k+G) = C~(k),

and the product (``tb_model.set_convention(2)``) is compared with that oracle: CPU tests go
through the host layer with the oracle engine injected, ``-m gpu`` tests through the C ABI.
"""
import numpy as np
import pytest

from oracle import pythtb_oracle as orc
from tests import compare, models as M


def _models(mod):
    return {
        "haldane": M.haldane(mod, delta=0.0),                         # n = 2 register kernels
        "kane_mele": M.kane_mele(mod, "odd"),                         # spinor, n = 4
        "random3": M.random_model(mod, norb=3, dim=2, nhop=7, nspin=1, seed=5),
        "random_spin": M.random_model(mod, norb=5, dim=2, nhop=12, nspin=2, seed=6),   # n = 10 tile solver
        "bn_ribbon": M.bn_ribbon(mod, 20),                            # dim_k 1 of dim_r 2, n = 40 blocked solver
    }


def _oracle_mod():
    from tests import oracle_api
    return oracle_api


def _gpu_mod():
    import pythtb_b200
    return pythtb_b200


@pytest.mark.parametrize("name", ["haldane", "kane_mele", "random3", "random_spin", "bn_ribbon"])
def test_oracle_convention_ii_is_the_gauge_transform_of_convention_i(name):
    """tex:355-364: the direct Convention-II sum equals D H_I D^H; same spectrum."""
    m = _models(_oracle_mod())[name]
    k = np.random.RandomState(3).rand(17, m._dim_k) * 2.0 - 1.0
    h1 = orc.gen_ham(m, k)
    m.set_convention(2)
    h2 = orc.gen_ham(m, k)
    d = orc.convention_gauge(m, k)
    want = d[:, :, None] * h1 * d[:, None, :].conj()
    assert np.max(np.abs(h2 - want)) < 1e-13 * max(1.0, np.max(np.abs(h1)))
    assert np.max(np.abs(orc.sol_ham(h1) - orc.sol_ham(h2))) < 1e-12 * max(1.0, np.max(np.abs(h1)))


def test_oracle_convention_ii_mesh_is_periodic_and_keeps_the_chern_number():
    """C~(k+G) = C~(k): the closing rows are plain copies; the Chern integer does not
    depend on the convention, and Convention-II wavefunctions are D x Convention-I ones
    up to a phase per band (compared through the gauge-invariant plaquette fluxes of the
    transformed Convention-I array)."""
    mod = _oracle_mod()
    m = M.haldane(mod, delta=0.0)
    mesh = [25, 25]
    w1, _ = orc.solve_on_grid(m, mesh, [-0.5, -0.5])
    m.set_convention(2)
    w2, _ = orc.solve_on_grid(m, mesh, [-0.5, -0.5])
    assert np.array_equal(w2[-1], w2[0]) and np.array_equal(w2[:, -1], w2[:, 0])
    c1 = orc.berry_flux(w1, 2, [0]) / (2 * np.pi)
    c2 = orc.berry_flux(w2, 2, [0]) / (2 * np.pi)
    assert abs(c1 - round(c1)) < 1e-9 and abs(c2 - c1) < 1e-9 and round(c1) != 0
    ax = [-0.5 + np.arange(n) / float(n - 1) for n in mesh]
    kk = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 2)
    d = orc.convention_gauge(m, kk).reshape(mesh[0], mesh[1], 1, m._nsta)
    p_from_1 = orc.berry_flux(w1 * d, 2, [0], individual_phases=True)
    p2 = orc.berry_flux(w2, 2, [0], individual_phases=True)
    assert np.max(np.abs(compare.circ_diff(p2, p_from_1, 2 * np.pi))) < 1e-10


def _run_convention_ii(mod, name):
    """Drive the public API on a Convention-II model and return gauge-invariant results."""
    m = _models(mod)[name]
    m.set_convention(2)
    assert m.get_convention() == 2
    out = {}
    k = np.random.RandomState(11).rand(9, m._dim_k) * 2.0 - 1.0
    out["evals"] = m.solve_all(k)
    ev, vec = m.solve_all(k, eig_vectors=True)
    out["evals_vec"] = ev
    nocc = m._nsta // 2
    v = vec[:nocc].reshape(nocc, len(k), m._nsta)
    out["proj"] = np.einsum("bki,bkj->kij", v, v.conj())
    out["ham"] = np.array([np.asarray(m._gen_ham(kk)).reshape(m._nsta, m._nsta) for kk in k[:3]])
    occ = list(range(nocc))
    if m._dim_k == 2:
        w = mod.wf_array(m, [17, 13])
        out["gaps"] = w.solve_on_grid([-0.5, -0.5])
        out["flux"] = np.array(w.berry_flux(occ))
        out["plaq"] = w.berry_flux(occ, individual_phases=True)
        out["phi0"] = w.berry_phase(occ, 0, contin=False)
        out["phi1"] = w.berry_phase(occ, 1, contin=True)
        if nocc > 1:
            out["wilson"] = w.berry_phase(occ, 1, contin=False, berry_evals=True)
        # a manually filled array closed with impose_pbc must agree with the fused grid solve
        w2 = mod.wf_array(m, [5, 4])
        for i in range(5):
            for j in range(4):
                w2.solve_on_one_point([-0.5 + i / 4.0, -0.5 + j / 3.0], [i, j])
        w2.impose_pbc(0, 0)
        w2.impose_pbc(1, 1)
        out["plaq_manual"] = w2.berry_flux(occ, individual_phases=True)
    else:
        w = mod.wf_array(m, [21])
        out["gaps"] = w.solve_on_grid([0.0])
        out["phi0"] = np.array(w.berry_phase(occ, 0))
        out["wilson"] = w.berry_phase(occ, 0, contin=False, berry_evals=True)
    return m, k, out


def _check_against_oracle(m, k, out):
    """``m`` carries _convention == 2, so the oracle takes its Convention-II branch."""
    scale = max(1.0, float(np.max(np.abs(out["evals"]))))
    ev_ref, vec_ref = orc.solve_all(m, k, eig_vectors=True)
    assert np.max(np.abs(out["evals"] - ev_ref)) <= compare.TOL_EVAL * scale
    assert np.max(np.abs(out["evals_vec"] - ev_ref)) <= compare.TOL_EVAL * scale
    nocc = m._nsta // 2
    v = vec_ref[:nocc].reshape(nocc, len(k), m._nsta)
    assert np.max(np.abs(out["proj"] - np.einsum("bki,bkj->kij", v, v.conj()))) <= compare.TOL_PROJ
    assert np.max(np.abs(out["ham"] - orc.gen_ham(m, k[:3]))) <= compare.TOL_HAM * scale
    occ = list(range(nocc))
    two_pi = 2.0 * np.pi
    if m._dim_k == 2:
        wfs, gaps = orc.solve_on_grid(m, [17, 13], [-0.5, -0.5])
        assert np.max(np.abs(out["gaps"] - gaps)) <= compare.TOL_EVAL * scale
        assert abs(compare.circ_diff(out["flux"], orc.berry_flux(wfs, 2, occ), two_pi)) <= compare.TOL_PHASE
        ref_plaq = orc.berry_flux(wfs, 2, occ, individual_phases=True)
        assert np.max(np.abs(compare.circ_diff(out["plaq"], ref_plaq, two_pi))) <= compare.TOL_PHASE
        for key, d, contin in (("phi0", 0, False), ("phi1", 1, True)):
            ref = orc.berry_phase(wfs, 2, occ, d, contin=contin)
            assert np.max(np.abs(compare.circ_diff(out[key], ref, two_pi))) <= compare.TOL_PHASE, key
        if nocc > 1:
            ref = orc.berry_phase(wfs, 2, occ, 1, contin=False, berry_evals=True)
            ok, dev = compare.sets_close(out["wilson"], ref, two_pi, compare.TOL_PHASE)
            assert ok, dev
        wfs2, _ = orc.solve_on_grid(m, [5, 4], [-0.5, -0.5])
        ref = orc.berry_flux(wfs2, 2, occ, individual_phases=True)
        assert np.max(np.abs(compare.circ_diff(out["plaq_manual"], ref, two_pi))) <= compare.TOL_PHASE
    else:
        wfs, gaps = orc.solve_on_grid(m, [21], [0.0])
        assert np.max(np.abs(out["gaps"] - gaps)) <= compare.TOL_EVAL * scale
        assert abs(compare.circ_diff(out["phi0"], orc.berry_phase(wfs, 1, occ, 0), two_pi)) <= compare.TOL_PHASE
        ref = orc.berry_phase(wfs, 1, occ, 0, contin=False, berry_evals=True)
        ok, dev = compare.sets_close(out["wilson"], ref, two_pi, compare.TOL_PHASE)
        assert ok, dev


@pytest.mark.parametrize("name", ["haldane", "kane_mele", "bn_ribbon"])
def test_host_layer_convention_ii(name):
    """Host logic (plan convention flag, pbc factors, set_convention invalidating the cached
    plan) with the oracle engine injected."""
    m, k, out = _run_convention_ii(_oracle_mod(), name)
    _check_against_oracle(m, k, out)
    m.set_convention(1)
    assert m._plan().convention == 1
    m.set_convention(2)
    assert m._plan().convention == 2
    with pytest.raises(Exception):
        m.set_convention(3)


def test_convention_survives_model_surgery():
    mod = _oracle_mod()
    m = M.haldane(mod, delta=0.2)
    m.set_convention(2)
    assert m.cut_piece(3, 1).get_convention() == 2
    assert m.make_supercell([[2, 0], [0, 1]]).get_convention() == 2
    assert m.reduce_dim(1, 0.25).get_convention() == 2
    # the gauge differs between the conventions, the ribbon spectrum does not
    rib2 = m.cut_piece(4, 1)
    m1 = M.haldane(mod, delta=0.2)
    rib1 = m1.cut_piece(4, 1)
    k = [[0.13], [0.4]]
    assert np.max(np.abs(rib1.solve_all(k) - rib2.solve_all(k))) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["haldane", "kane_mele", "random3", "random_spin", "bn_ribbon"])
def test_gpu_convention_ii(name):
    """Every kernel family with the Convention-I gauge switched off (plan convention = 2):
    register mesh kernels (n = 2, 3, 4), tile solver (n = 10), blocked solver (n = 40)."""
    m, k, out = _run_convention_ii(_gpu_mod(), name)
    _check_against_oracle(m, k, out)


@pytest.mark.gpu
def test_gpu_chern_number_is_convention_independent():
    mod = _gpu_mod()
    res = []
    for conv in (1, 2):
        m = M.haldane(mod, delta=0.0)
        m.set_convention(conv)
        w = mod.wf_array(m, [257, 257])
        w.solve_on_grid([-0.5, -0.5])
        res.append(w.berry_flux([0]) / (2 * np.pi))
    assert abs(res[0] - round(res[0])) < 1e-9 and round(res[0]) != 0
    assert abs(res[0] - res[1]) < 1e-9
