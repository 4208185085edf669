"""GPU parity at the sizes of BASELINE configs 4 and 5 (run on the B200 with ``-m gpu``).

The small-matrix headline kernels are covered by tests/test_gpu_parity.py; here the large-matrix
families are compared with the oracle at the orbital counts the configs name:

* ``solve_blocked_kernel`` WITH eigenvectors at n = 200 / 400 / 499: eigenvalues, residual
  |H u - E u|, orthonormality of all n vectors, and the gauge-invariant occupied projector against
  numpy/LAPACK — including a flat-band model whose occupied states form one 60-fold degenerate cluster
  (the inverse-iteration re-orthogonalisation is hardest there; pythtb.py:927-953 is the reference).
* ``link_matrix_kernel`` beyond a single 64 x 64 DMMA tile: ``berry_phase`` (determinant branch AND
  ``berry_evals=True``: multi-tile overlap GEMM, LU at n = 250, Newton-Schulz polar factors, tree
  product, unitary eigenphases) at nocc = 100 / 130 / 250 on the SAME ``_wfs`` the oracle sees
  (pythtb.py:3798-3838), general-nocc ``berry_flux``, and ``position_hwf`` at nocc = 250
  (pythtb.py:2162-2279).
"""
import numpy as np
import pytest

from tests import compare, models as M

pytestmark = pytest.mark.gpu
TWO_PI = 2.0 * np.pi


def _mod():
    import pythtb_b200
    return pythtb_b200


def _dimer_flat_bands(mod, ncell=60, t=0.8, eps=0.3):
    """1-D chain of `ncell` decoupled dimers per cell: two perfectly flat, ncell-fold degenerate bands at
    eps -+ t.  No k dependence of the spectrum, maximal degeneracy: the worst case for inverse iteration."""
    orb = [[(i + 0.25 * (j + 1)) / ncell] for i in range(ncell) for j in range(2)]
    m = mod.tb_model(1, 1, [[1.0]], orb)
    m.set_onsite([eps] * (2 * ncell))
    for i in range(ncell):
        m.set_hop(t * np.exp(0.3j * i), 2 * i, 2 * i + 1, [0])
    return m


def _check_vectors(model, kpts, nocc, tol_res=1e-10, tol_orth=1e-11, tol_proj=1e-9):
    from oracle import pythtb_oracle as orc
    kpts = np.asarray(kpts, dtype=float)
    n = model._nsta
    ev, evec = model.solve_all(kpts, eig_vectors=True)
    ham = orc.gen_ham(model, kpts)
    ev_ref, vec_ref = orc.sol_ham(ham, True)                      # [k, band], [k, band, orb]
    scale = max(1.0, float(np.max(np.abs(ev_ref))))
    assert np.max(np.abs(ev.T - ev_ref)) <= compare.TOL_EVAL * scale
    worst = dict(res=0.0, orth=0.0, proj=0.0)
    for i in range(len(kpts)):
        v = evec[:, i].reshape(n, n)                              # rows = eigenvectors (pythtb.py:947)
        worst["res"] = max(worst["res"], float(np.max(np.abs(ham[i] @ v.T - v.T * ev[:, i][None, :]))))
        worst["orth"] = max(worst["orth"], float(np.max(np.abs(v.conj() @ v.T - np.eye(n)))))
        p = v[:nocc].T @ v[:nocc].conj()
        r = vec_ref[i].reshape(n, n)
        p_ref = r[:nocc].T @ r[:nocc].conj()
        worst["proj"] = max(worst["proj"], float(np.max(np.abs(p - p_ref))))
    assert worst["res"] <= tol_res * scale, worst
    assert worst["orth"] <= tol_orth, worst
    assert worst["proj"] <= tol_proj, worst
    return worst


@pytest.mark.parametrize("which", ["ribbon200", "ribbon400", "slab499", "flat120", "flat240"])
def test_blocked_solver_vectors_config_scale(which):
    mod = _mod()
    if which.startswith("ribbon"):
        n = int(which[6:])
        m = M.bn_ribbon(mod, n // 2)
        kpts, nocc = [[0.0], [0.123], [0.5], [0.3333]], n // 2
    elif which == "slab499":
        m = M.cubic_slab(mod, 250)
        kpts, nocc = [[0.1, 0.2], [0.0, 0.5], [0.25, 0.25]], 250
    else:
        ncell = int(which[4:]) // 2
        m = _dimer_flat_bands(mod, ncell)
        kpts, nocc = [[0.0], [0.37]], ncell
    assert m._nsta == (499 if which == "slab499" else int("".join(c for c in which if c.isdigit())))
    _check_vectors(m, kpts, nocc)
    from pythtb_b200 import _engine
    assert _engine.get_engine().last_solve_kernel == "solve_blocked_kernel"


def test_ribbon_berry_phase_nocc100():
    """Config 4 in miniature: BN ribbon cut_piece(100, 1), norb 200, 41 k-points, occupied half."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rib = M.bn_ribbon(mod, 100)
    nocc = 100
    w = mod.wf_array(rib, [41])
    w.solve_on_grid([0.0])
    wfs = np.array(w._wfs)
    got = w.berry_phase(range(nocc), 0)
    ref = orc.berry_phase(wfs, 1, list(range(nocc)), 0)
    assert abs(compare.circ_diff(got, ref, TWO_PI)) < compare.TOL_PHASE
    got_ev = w.berry_phase(range(nocc), 0, berry_evals=True, contin=False)
    ref_ev = orc.berry_phase(wfs, 1, list(range(nocc)), 0, contin=False, berry_evals=True)
    ok, dev = compare.sets_close(got_ev, ref_ev, TWO_PI, compare.TOL_PHASE)
    assert ok, dev
    assert abs(compare.circ_diff(np.sum(got_ev), got, TWO_PI)) < 1e-7
    # after moving the non-periodic vector (examples/bn_ribbon_berry.py:50) the phase is unchanged mod 2 pi
    rib2 = rib.change_nonperiodic_vector(1, to_home_suppress_warning=True)
    w2 = mod.wf_array(rib2, [41])
    w2.solve_on_grid([0.0])
    ref2 = orc.berry_phase(np.array(w2._wfs), 1, list(range(nocc)), 0)
    assert abs(compare.circ_diff(w2.berry_phase(range(nocc), 0), ref2, TWO_PI)) < compare.TOL_PHASE


@pytest.mark.parametrize("nl", [130, 250])
def test_slab_wilson_loops_and_hwf_config_scale(nl):
    """Config 5 in miniature: the cubic slab with nl layers (norb 2 nl - 1) on a 5 x 5 mesh — all-band Berry
    phases (LU determinant at n = nl), all-band Wilson spectra, general-nocc flux, and position_hwf."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    slab = M.cubic_slab(mod, nl)
    occ = list(range(nl))
    mesh = [5, 5]
    w = mod.wf_array(slab, mesh)
    w.solve_on_grid([0.0, 0.0])
    wfs = np.array(w._wfs)
    for d in (0, 1):
        got = w.berry_phase(occ, d, contin=False)
        ref = orc.berry_phase(wfs, 2, occ, d, contin=False)
        assert np.max(np.abs(compare.circ_diff(got, ref, TWO_PI))) < compare.TOL_PHASE, d
    got_ev = w.berry_phase(occ, 0, contin=False, berry_evals=True)
    ref_ev = orc.berry_phase(wfs, 2, occ, 0, contin=False, berry_evals=True)
    assert np.shape(got_ev) == np.shape(ref_ev) == (5, nl)
    ok, dev = compare.sets_close(got_ev, ref_ev, TWO_PI, compare.TOL_PHASE)
    assert ok, dev
    got_pl = w.berry_flux(occ, individual_phases=True)
    ref_pl = orc.berry_flux(wfs, 2, occ, None, True)
    assert np.max(np.abs(compare.circ_diff(got_pl, ref_pl, TWO_PI))) < compare.TOL_PHASE
    # hybrid Wannier centres along the finite direction at two k-points (examples/cubic_slab_hwf.py:63-79)
    for key in ([0, 0], [2, 3]):
        evec = wfs[key[0], key[1]][:nl]
        c_ref = orc.position_hwf(slab, evec, 2)
        c = w.position_hwf(key, occ, 2)
        assert np.max(np.abs(c - c_ref)) < 1e-9
        c2, hwf = w.position_hwf(key, occ, 2, hwf_evec=True, basis="orbital")
        hw = hwf.reshape(nl, -1)
        assert np.max(np.abs(hw.conj() @ hw.T - np.eye(nl))) < 1e-10
        pos = np.asarray(slab._orb)[:, 2]
        assert np.max(np.abs(np.einsum("io,o,jo->ij", hw.conj(), pos, hw) - np.diag(c2))) < 1e-9
    # batched variant on the whole mesh against the per-point call
    call = w.position_hwf_all(occ, 2)
    assert np.max(np.abs(call[2, 3] - w.position_hwf([2, 3], occ, 2))) < 1e-10


def test_wilson_polar_factor_of_unnormalised_and_broken_arrays():
    """berry_evals=True on user-filled arrays: a non-normalised (but regular) fill — spectral norm of the link
    overlaps 2.5 > sqrt(3), outside the raw Newton-Schulz convergence region — is rescaled and matches the oracle's
    SVD; an array holding NaN raises instead of returning phases of garbage."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rib = M.bn_ribbon(mod, 12)
    n = rib._nsta
    w = mod.wf_array(rib, [9])
    w.solve_on_grid([0.0])
    host = w._wfs
    host[3] = 2.5 * host[3]
    occ = list(range(n // 2))
    ref = orc.berry_phase(np.array(host), 1, occ, 0, contin=False, berry_evals=True)
    got = w.berry_phase(occ, 0, contin=False, berry_evals=True)
    ok, dev = compare.sets_close(got, ref, TWO_PI, compare.TOL_PHASE)
    assert ok, dev
    host = w._wfs
    host[5, 1, 0] = np.nan
    with pytest.raises(Exception, match="not finite"):
        w.berry_phase(occ, 0, contin=False, berry_evals=True)
