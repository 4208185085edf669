"""Parity cases: each function drives the PUBLIC PythTB API of ``mod`` and
returns a dict of gauge-invariant results.

The same function runs against
  * the unmodified reference (``tests/golden/make_golden.py`` -> fixtures),
  * the numpy oracle through ``tests/oracle_api.py`` (CPU tests),
  * ``pythtb_b200`` on a B200 (``-m gpu`` tests),
so the three are compared on identical inputs.  The cases restate the
reference's own regression scripts ``tests/test_examples/*/*/run.py`` (cited
per case) plus a few wider ones.  Raw eigenvectors are never returned (gauge).
"""
import numpy as np
from . import models as M


def _projector(evec, occ):
    """Gauge-invariant band projector sum_n |u_n><u_n| over ``occ`` for
    ``evec[band, ..., orb(,spin)]`` at one k-point (flattened orbitals)."""
    v = np.asarray(evec)[occ].reshape(len(occ), -1)
    return v.T @ v.conj()


def case_haldane_bands(mod):
    """tests/test_examples/haldane/haldane/run.py:28-50 (config 1)."""
    m = M.haldane(mod, delta=0.2)
    path = [[0.0, 0.0], [2.0 / 3.0, 1.0 / 3.0], [0.5, 0.5], [1.0 / 3.0, 2.0 / 3.0], [0.0, 0.0]]
    k_vec, k_dist, k_node = m.k_path(path, 101, report=False)
    evals = m.solve_all(k_vec)
    kpts = [[i / 20.0, j / 20.0] for i in range(20) for j in range(20)]
    evals_dos = m.solve_all(kpts).flatten()
    ev2, evec = m.solve_all(k_vec, eig_vectors=True)
    proj = np.array([_projector(evec[:, i], [0]) for i in range(0, 101, 10)])
    return dict(k_vec=k_vec, k_dist=k_dist, k_node=k_node, evals=evals,
                evals_dos=evals_dos, evals_with_vec=ev2, proj_band0=proj)


def case_haldane_bp(mod):
    """tests/test_examples/haldane/haldane_bp/run.py:29-58."""
    m = M.haldane(mod, delta=0.0)
    w1 = mod.wf_array(m, [31, 31])
    gaps = w1.solve_on_grid([-0.5, -0.5])
    out = dict(gaps=gaps)
    out["phi_a1"] = w1.berry_phase([0], 0, contin=True)
    out["phi_b1"] = w1.berry_phase([1], 0, contin=True)
    out["phi_c1"] = w1.berry_phase([0, 1], 0, contin=True)
    out["phi_a1_dir1"] = w1.berry_phase([0], 1, contin=False)
    out["flux_a1"] = np.array(w1.berry_flux([0]))
    out["flux_b1"] = np.array(w1.berry_flux([1]))
    out["plaq_a1"] = w1.berry_flux([0], individual_phases=True)
    out["flux_a1_swapped"] = np.array(w1.berry_flux([0], dirs=[1, 0]))
    kx = np.linspace(-0.5, 0.5, num=31)
    w2 = mod.wf_array(m, [31, 31])
    for i in range(31):
        for j in range(31):
            (_, evec) = m.solve_one([kx[i], kx[j]], eig_vectors=True)
            w2[i, j] = evec
    w2.impose_pbc(0, 0)
    w2.impose_pbc(1, 1)
    out["flux_a2"] = np.array(w2.berry_flux([0]))
    return out


def case_kane_mele(mod):
    """tests/test_examples/kane_mele/kane_mele/run.py:65-91 (spinor model,
    SVD + eigvals Wilson-loop branch)."""
    out = {}
    path = [[0, 0], [2 / 3, 1 / 3], [1 / 2, 1 / 2], [1 / 3, 2 / 3], [0, 0]]
    for top in ("even", "odd"):
        m = M.kane_mele(mod, top)
        w = mod.wf_array(m, [41, 41])
        out["gaps_" + top] = w.solve_on_grid([-0.5, -0.5])
        k_vec, _, _ = m.k_path(path, 101, report=False)
        out["evals_" + top] = m.solve_all(k_vec)
        wc = w.berry_phase([0, 1], dir=1, contin=False, berry_evals=True)
        out["wan_cent_" + top] = wc / (2.0 * np.pi)
        out["wan_cent_contin_" + top] = w.berry_phase([0, 1], dir=1, contin=True, berry_evals=True)
        out["phase_tot_" + top] = w.berry_phase([0, 1], dir=1, contin=True)
        out["flux01_" + top] = np.array(w.berry_flux([0, 1]))
        out["plaq01_" + top] = w.berry_flux([0, 1], individual_phases=True)
    return out


def case_cone(mod):
    """tests/test_examples/graphene/cone/run.py:14-64 (manual fills, open
    meshes, solve_on_one_point)."""
    m = M.graphene(mod, delta=-0.1)
    n = 31
    center = np.array([1.0 / 3.0, 2.0 / 3.0])
    wc = mod.wf_array(m, [n])
    for i in range(n):
        ang = 2.0 * np.pi * float(i) / float(n - 1)
        wc.solve_on_one_point(np.array([np.cos(ang), np.sin(ang)]) * 0.05 + center, i)
    wc[-1] = wc[0]
    out = dict(bphase_circ0=np.array(wc.berry_phase([0], 0)),
               bphase_circ1=np.array(wc.berry_phase([1], 0)),
               bphase_circ01=np.array(wc.berry_phase([0, 1], 0)))
    ws = mod.wf_array(m, [n, n])
    for i in range(n):
        for j in range(n):
            kpt = np.array([0.1 * (-0.5 + float(i) / float(n - 1)),
                            0.1 * (-0.5 + float(j) / float(n - 1))]) + center
            (_, evec) = m.solve_one(kpt, eig_vectors=True)
            ws[i, j] = evec
    out["bflux_square_0"] = np.array(ws.berry_flux([0]))
    out["bflux_square_1"] = np.array(ws.berry_flux([1]))
    out["bflux_square_01"] = np.array(ws.berry_flux([0, 1]))
    out["plaq"] = ws.berry_flux([0], individual_phases=True)
    return out


def case_bn_ribbon(mod):
    """tests/test_examples/boron_nitride/bn_ribbon_berry/run.py:18-45
    (config 4 in miniature) plus a wider ribbon."""
    out = {}
    for tag, ncell in (("", 3), ("_w12", 12)):
        orig = M.bn_ribbon(mod, ncell)
        k_vec, _, _ = orig.k_path([[-0.5], [0.5]], 41, report=False)
        out["evals" + tag] = orig.solve_all(k_vec)
        nocc = orig.get_num_orbitals() // 2
        w = mod.wf_array(orig, [41])
        out["gaps" + tag] = w.solve_on_grid([0.0])
        out["berry_phase_orig" + tag] = np.array(w.berry_phase(range(nocc), dir=0))
        out["wilson_evals" + tag] = w.berry_phase(range(nocc), dir=0, berry_evals=True)
        perp = orig.change_nonperiodic_vector(1, to_home_suppress_warning=True)
        w2 = mod.wf_array(perp, [41])
        w2.solve_on_grid([0.0])
        out["berry_phase_perp" + tag] = np.array(w2.berry_phase(range(nocc), dir=0))
    return out


def case_cubic_slab(mod):
    """tests/test_examples/slab/cubic_slab_hwf/run.py:16-66 (config 5 in
    miniature: slab solve, position_hwf, HWF Wilson loops)."""
    nl = 9
    slab = M.cubic_slab(mod, nl)
    k1 = np.linspace(0.0, 1.0, 10, endpoint=False)
    kpts = [[kx, ky] for kx in k1 for ky in k1]
    out = dict(evals=slab.solve_all(kpts))
    nk = 9
    bloch = mod.wf_array(slab, [nk, nk])
    out["gaps"] = bloch.solve_on_grid([0.0, 0.0])
    hwf_arr = bloch.empty_like(nsta_arr=nl)
    hwfc = np.zeros([nk, nk, nl])
    for ix in range(nk):
        for iy in range(nk):
            (val, vec) = bloch.position_hwf([ix, iy], occ=list(range(nl)), dir=2,
                                            hwf_evec=True, basis="orbital")
            hwfc[ix, iy] = val
            hwf_arr[ix, iy] = vec
    hwf_arr.impose_pbc(0, 0)
    hwf_arr.impose_pbc(1, 1)
    out["hwfc"] = hwfc
    px = np.zeros((nl, nk))
    for n in range(nl):
        px[n, :] = hwf_arr.berry_phase(dir=0, occ=[n]) / (2.0 * np.pi)
    out["px"] = px
    out["pos_exp"] = bloch.position_expectation([2, 3], occ=list(range(nl)), dir=2).sum()
    out["wilson_all"] = bloch.berry_phase(occ=list(range(nl)), dir=0, berry_evals=True, contin=False)
    out["flux_occ"] = np.array(bloch.berry_flux(occ=list(range(nl))))
    return out


def case_three_site(mod):
    """tests/test_examples/three_site/3site_cycle/run.py:18-39 with
    t=-1, delta=2 (test.py:18): mixed (k, lambda) array, impose_pbc on one
    axis only."""
    nk, nl = 31, 21
    lam = np.linspace(0, 1, nl, endpoint=True)
    w = mod.wf_array(M.three_site(mod, 0.0), [nk, nl])
    for il in range(nl):
        m = M.three_site(mod, lam[il])
        k_vec, _, _ = m.k_path([[-0.5], [0.5]], nk, report=False)
        _, evec = m.solve_all(k_vec, eig_vectors=True)
        for ik in range(nk):
            w[ik, il] = evec[:, ik, :]
    w.impose_pbc(0, 0)
    out = dict(wann_center=w.berry_phase([0], 0) / (2.0 * np.pi),
               final=np.array(w.berry_flux([0])))
    w.impose_loop(1)
    out["phase_lambda"] = w.berry_phase([0], 1, contin=True)
    out["flux_01"] = np.array(w.berry_flux([0, 1]))
    return out


def case_misc_bands(mod):
    """checkerboard / 0-D molecule / 1-D chain eigenvalues
    (tests/test_examples/checkerboard, zero_dim/0dim, tests/test_pythtb.py:20-45)."""
    out = {}
    cb = M.checkerboard(mod)
    k_vec, _, _ = cb.k_path([[0.0, 0.0], [0.0, 0.5], [0.5, 0.5], [0.0, 0.0]], 301, report=False)
    out["checkerboard"] = cb.solve_all(k_vec)
    mol = M.molecule(mod)
    out["molecule"] = mol.solve_all()
    ev, evec = mol.solve_all(eig_vectors=True)
    out["molecule_proj"] = _projector(evec, [0, 1])
    ts = M.three_site(mod, 0.3)
    out["three_site_one"] = ts.solve_one([0.123])
    hal = M.haldane(mod, 0.2)
    out["haldane_one"] = hal.solve_one([0.123, 0.523])
    fin = hal.cut_piece(4, 0).cut_piece(4, 1, glue_edgs=True)
    out["haldane_fin"] = fin.solve_all()
    return out


def case_random(mod):
    """Seeded random models (complex hoppings, generic tau, spinors, dim_k
    1..3): H(k) elementwise through ``_gen_ham`` and eigenvalues."""
    out = {}
    specs = [("r1", dict(norb=3, dim=1, nhop=6, nspin=1, seed=1)),
             ("r2", dict(norb=6, dim=2, nhop=20, nspin=1, seed=2)),
             ("r3", dict(norb=5, dim=3, nhop=24, nspin=1, seed=3)),
             ("s2", dict(norb=2, dim=2, nhop=7, nspin=2, seed=4)),
             ("s3", dict(norb=4, dim=3, nhop=15, nspin=2, seed=5)),
             ("r40", dict(norb=40, dim=2, nhop=160, nspin=1, seed=6))]
    for tag, kw in specs:
        m = M.random_model(mod, **kw)
        rng = np.random.RandomState(100 + kw["seed"])
        k = rng.rand(7, kw["dim"]) * 2.0 - 1.0
        out["k_" + tag] = k
        out["ham_" + tag] = np.array([np.array(m._gen_ham(kk)).reshape(m._nsta, m._nsta) for kk in k])
        ev, evec = m.solve_all(k, eig_vectors=True)
        out["evals_" + tag] = ev
        nocc = max(1, m._nsta // 2)
        out["proj_" + tag] = np.array([_projector(evec[:, i], list(range(nocc))) for i in range(7)])
    return out


def case_grid3d(mod):
    """3-D wf_array: berry_phase along every axis and berry_flux over every
    pair of axes (pythtb.py:3000-3027, 3153-3202)."""
    m = M.cubic_bulk(mod, delta=1.0, ta=0.4, tb=0.7)
    w = mod.wf_array(m, [7, 6, 5])
    out = dict(gaps=w.solve_on_grid([0.0, 0.1, -0.2]))
    for d in range(3):
        out["phase_dir%d" % d] = w.berry_phase([0], d, contin=True)
        out["wilson_dir%d" % d] = w.berry_phase([0, 1], d, contin=True, berry_evals=True)
    for dirs in ([0, 1], [1, 2], [2, 0]):
        tag = "%d%d" % tuple(dirs)
        out["flux_" + tag] = w.berry_flux([0], dirs=dirs)
        out["plaq_" + tag] = w.berry_flux([0], dirs=dirs, individual_phases=True)
    return out


def case_position(mod):
    """position_matrix / expectation / hwf on a (gapped) BN ribbon
    (examples/haldane_hwf.py pattern; pythtb.py:2034-2279)."""
    rib = M.bn_ribbon(mod, 6)
    nocc = rib.get_num_orbitals() // 2
    out = {}
    ks = [[0.0], [0.17], [0.5]]
    ev, evec = rib.solve_all(ks, eig_vectors=True)
    out["evals"] = ev
    out["pos_trace"] = np.array([rib.position_expectation(evec[:nocc, i], 1).sum() for i in range(3)])
    out["hwfc"] = np.array([rib.position_hwf(evec[:nocc, i], 1) for i in range(3)])
    hw = [rib.position_hwf(evec[:nocc, i], 1, hwf_evec=True, basis="orbital") for i in range(3)]
    out["hwfc_vec"] = np.array([h[0] for h in hw])
    # X restricted to the occupied subspace is gauge invariant as an operator
    out["pos_op"] = np.array([evec[:nocc, i].T @ rib.position_matrix(evec[:nocc, i], 1) @ evec[:nocc, i].conj()
                              for i in range(3)])
    return out


def case_more_bands(mod):
    """The remaining band-structure regression scripts of the reference:
    tests/test_examples/buckling/buckled_layer/run.py:21-40 (dim_k 2 of dim_r 3),
    buckling/trestle/run.py:17-35 (per=[0], complex hoppings, k_path("fullc")),
    graphene/graphene/run.py:17-34 (touching bands at K),
    supercell/supercell/run.py:8-27 (make_supercell + cut_piece, k_path("full")),
    zero_dim/0dim/run.py:8-33 (0-D, four orbitals)."""
    out = {}
    m = mod.tb_model(2, 3, [[1.0, 0.0, 0.0], [0.0, 1.25, 0.0], [0.0, 0.0, 3.0]],
                     [[0.0, 0.0, -0.15], [0.5, 0.5, 0.15]])
    m.set_onsite([-1.1, 1.1])
    for R in ([0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]):
        m.set_hop(0.6, 1, 0, R)
    k_vec, _, _ = m.k_path([[0.0, 0.0], [0.0, 0.5], [0.5, 0.5], [0.0, 0.0]], 81, report=False)
    out["evals_buckled"] = m.solve_all(k_vec)
    m = mod.tb_model(1, 2, [[2.0, 0.0], [0.0, 1.0]], [[0.0, 0.0], [0.5, 1.0]], per=[0])
    m.set_hop(2.0, 0, 0, [1, 0])
    m.set_hop(2.0, 1, 1, [1, 0])
    m.set_hop(0.8 + 0.6j, 0, 1, [0, 0])
    m.set_hop(0.8 + 0.6j, 1, 0, [1, 0])
    k_vec, _, _ = m.k_path("fullc", 100, report=False)
    out["evals_trestle"] = m.solve_all(k_vec)
    g = M.graphene(mod, delta=0.0)
    k_vec, _, _ = g.k_path([[0.0, 0.0], [2.0 / 3.0, 1.0 / 3.0], [0.5, 0.5], [0.0, 0.0]], 121, report=False)
    out["evals_graphene"] = g.solve_all(k_vec)
    sc = g.make_supercell([[2, 1], [-1, 2]], to_home=True)
    slab = sc.cut_piece(6, 1, glue_edgs=False)
    k_vec, _, _ = slab.k_path("full", 100, report=False)
    out["evals_supercell"] = slab.solve_all(k_vec)
    sq32 = np.sqrt(3.0) / 2.0
    m = mod.tb_model(0, 3, np.identity(3).tolist(),
                     [[(2.0 / 3.0) * sq32, 0.0, 0.0], [(-1.0 / 3.0) * sq32, 0.5, 0.0],
                      [(-1.0 / 3.0) * sq32, -0.5, 0.0], [0.0, 0.0, 1.0]])
    m.set_onsite([-0.5, -0.5, -0.5, 0.5])
    for i, j in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):
        m.set_hop(1.0, i, j)
    out["evals_0dim"] = m.solve_all()
    return out


def case_haldane_finite(mod):
    """Finite Haldane flakes, 0-D models of 200 and 800 orbitals:
    tests/test_examples/haldane/edge/run.py:35-50 (10x10, with eigenvectors) and
    haldane/haldane_fin/run.py:37-52 (20x20, glued or not).  The raw eigenvectors of
    the reference golden are gauge dependent; the projector on the lower half of the
    spectrum stands in for them."""
    hal = M.haldane(mod, delta=0.0)
    out = {}
    # occupied sets end below a real gap: the glued flake has its two mid-gap levels 5.7e-6 apart
    # (levels 99, 100), so the projector on 100 states would amplify rounding by 1e5
    for tag, glue, nocc in (("", False, 100), ("_half", True, 99)):
        fin = hal.cut_piece(10, 0, glue_edgs=glue).cut_piece(10, 1, glue_edgs=False)
        ev, evec = fin.solve_all(eig_vectors=True)
        out["evals_edge" + tag] = ev
        out["proj_edge" + tag] = _projector(evec, list(range(nocc)))[:48, :48]    # a corner block keeps the fixture small
    out["evals_fin_false"] = hal.cut_piece(20, 0, glue_edgs=False).cut_piece(20, 1, glue_edgs=False).solve_all().flatten()
    out["evals_fin_true"] = hal.cut_piece(20, 0, glue_edgs=True).cut_piece(20, 1, glue_edgs=True).solve_all().flatten()
    return out


def _mask_close(vals, evals, tol=1.0e-3):
    """Per-state expectation values are only defined up to the mixing inside a
    (near-)degenerate level: zero the entries whose level has a neighbour closer than tol
    (the reference's own test_spin fails for exactly this reason)."""
    ev = np.asarray(evals, dtype=float)
    gap = np.full(ev.shape, np.inf)
    d = np.diff(ev, axis=0)
    gap[1:] = np.minimum(gap[1:], d)
    gap[:-1] = np.minimum(gap[:-1], d)
    return np.where(gap > tol, np.asarray(vals, dtype=float), 0.0)


def case_haldane_hwf(mod):
    """tests/test_examples/haldane/haldane_hwf/run.py:37-94: Berry phase flow of the bulk
    against the hybrid Wannier centres / position expectations of a ribbon, with an occupied
    set that changes along k (the edge state crosses the Fermi level)."""
    m = mod.tb_model(2, 2, M._HEX_LAT, M._HEX_ORB)
    delta, t, t2 = -0.2, -1.0, 0.05 - 0.15j
    m.set_onsite([-delta, delta])
    m.set_hop(t, 0, 1, [0, 0])
    m.set_hop(t, 1, 0, [1, 0])
    m.set_hop(t, 1, 0, [0, 1])
    for amp, i, R in ((t2, 0, [1, 0]), (t2, 1, [1, -1]), (t2, 1, [0, 1]),
                      (t2.conjugate(), 1, [1, 0]), (t2.conjugate(), 0, [1, -1]), (t2.conjugate(), 0, [0, 1])):
        m.set_hop(amp, i, i, R)
    len_0, len_1, efermi = 100, 10, 0.25
    w = mod.wf_array(m, [len_0, len_1])
    out = dict(gaps=w.solve_on_grid([0.0, 0.0]))
    out["phi1"] = w.berry_phase(occ=[0], dir=1, contin=True)
    rib = m.cut_piece(len_1, fin_dir=1, glue_edgs=False)
    k_vec, _, _ = rib.k_path([0.0, 0.5, 1.0], len_0, report=False)
    rib_eval, rib_evec = rib.solve_all(k_vec, eig_vectors=True)
    rib_eval = rib_eval - efermi
    out["rib_eval"] = rib_eval
    nocc = np.sum(rib_eval < 0.0, axis=0)
    out["jump_k"] = np.array([i for i in range(len_0 - 1) if nocc[i] != nocc[i + 1]], dtype=float)
    pos = np.array([rib.position_expectation(rib_evec[:, i], dir=1) for i in range(len_0)]).T
    out["pos_exp_sum"] = pos.sum(axis=0)
    out["pos_exp_masked"] = _mask_close(pos, rib_eval)
    out["hwfc_flat"] = np.concatenate([rib.position_hwf(rib_evec[rib_eval[:, i] < 0.0, i], 1) for i in range(len_0)])
    return out


def case_three_site_fin(mod):
    """tests/test_examples/three_site/3site_cycle_fin/run.py:35-79 with t=-1.3, delta=2
    (regen_golden_data.py:22):
    Berry fluxes on a (lambda, k) array filled by hand (parametric axis, no pbc imposed),
    and the finite chain of 10 cells along the cycle (every 4th of the 241 lambda steps)."""
    nl, nk = 21, 31
    lam = np.linspace(0.0, 1.0, nl, endpoint=True)
    t = -1.3
    m0 = M.three_site(mod, 0.0, t=t)
    k_vec, _, _ = m0.k_path([[-0.5], [0.5]], nk, report=False)
    w = mod.wf_array(m0, [nl, nk])
    for il in range(nl):
        _, evec = M.three_site(mod, lam[il], t=t).solve_all(k_vec, eig_vectors=True)
        for ik in range(nk):
            w[il, ik] = evec[:, ik, :]
    out = dict(fluxes=np.array([w.berry_flux(o) for o in ([0], [1], [2], [0, 1], [0, 1, 2])]))
    lam2 = np.linspace(0.0, 1.0, 241)[::4]
    ch_eval = np.zeros((30, len(lam2)))
    ch_xexp = np.zeros((30, len(lam2)))
    for i, lmbd in enumerate(lam2):
        fin = M.three_site(mod, lmbd, t=t).cut_piece(10, 0)
        ev, evec = fin.solve_all(eig_vectors=True)
        ch_eval[:, i] = ev
        ch_xexp[:, i] = fin.position_expectation(evec, 0)
    out["evals_chain"] = ch_eval
    out["pos_exp_sum"] = ch_xexp.sum(axis=0)
    out["pos_exp_masked"] = _mask_close(ch_xexp, ch_eval)
    return out


ALL_CASES = {
    "haldane_bands": case_haldane_bands,
    "haldane_bp": case_haldane_bp,
    "kane_mele": case_kane_mele,
    "cone": case_cone,
    "bn_ribbon": case_bn_ribbon,
    "cubic_slab": case_cubic_slab,
    "three_site": case_three_site,
    "misc_bands": case_misc_bands,
    "random": case_random,
    "grid3d": case_grid3d,
    "position": case_position,
    "more_bands": case_more_bands,
    "haldane_finite": case_haldane_finite,
    "haldane_hwf": case_haldane_hwf,
    "three_site_fin": case_three_site_fin,
}
