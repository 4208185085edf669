"""GPU parity tests (run on the B200 with ``-m gpu``): every case of
tests/cases.py through the product path (pythtb_b200 -> ctypes C-ABI ->
sm_100a kernels) against the fixtures produced by the unmodified reference,
plus oracle comparisons on seeded inputs and size-independent properties at
the BASELINE.json sizes."""
import io
import os
import contextlib

import numpy as np
import pytest

from tests import cases, compare, models as M

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mod():
    import pythtb_b200
    return pythtb_b200


@pytest.mark.parametrize("name", sorted(cases.ALL_CASES))
def test_case_matches_reference(name):
    want = np.load(os.path.join(GOLD, name + ".npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        got = cases.ALL_CASES[name](_mod())
    bad = compare.compare_case(name, got, want)
    assert not bad, "\n".join(bad)


def test_native_library_is_loaded():
    """The tests above must have gone through libtbk_b200.so."""
    mod = _mod()
    M.haldane(mod).solve_all([[0.1, 0.2]])
    with open("/proc/self/maps") as f:
        assert "libtbk_b200.so" in f.read()


def _residual(model, k, ev, evec):
    from oracle import pythtb_oracle as orc
    ham = orc.gen_ham(model, k)
    worst = 0.0
    for i in range(len(k)):
        v = evec[:, i].reshape(model._nsta, -1)
        worst = max(worst, np.max(np.abs(ham[i] @ v.T - v.T * ev[:, i][None, :])))
        worst = max(worst, np.max(np.abs(v.conj() @ v.T - np.eye(model._nsta))))
    return worst


@pytest.mark.parametrize("spec", [
    dict(norb=2, dim=2, nhop=5, nspin=1, seed=21), dict(norb=3, dim=1, nhop=5, nspin=1, seed=22),
    dict(norb=4, dim=2, nhop=9, nspin=1, seed=23), dict(norb=2, dim=3, nhop=8, nspin=2, seed=24),
    dict(norb=7, dim=2, nhop=20, nspin=1, seed=25), dict(norb=6, dim=1, nhop=14, nspin=2, seed=26),
    dict(norb=24, dim=2, nhop=70, nspin=1, seed=27), dict(norb=33, dim=1, nhop=90, nspin=1, seed=28),
    dict(norb=60, dim=2, nhop=200, nspin=1, seed=29), dict(norb=120, dim=1, nhop=400, nspin=1, seed=30),
])
def test_random_models_against_oracle(spec):
    """Every eigensolver family (registers / tile / CTA-smem / CTA-workspace)."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    m = M.random_model(mod, **spec)
    k = np.random.RandomState(spec["seed"]).rand(37, spec["dim"]) * 2 - 1
    ev_ref = orc.solve_all(m, k)
    ev = m.solve_all(k)
    scale = max(1.0, np.max(np.abs(ev_ref)))
    assert np.max(np.abs(ev - ev_ref)) <= compare.TOL_EVAL * scale
    ev2, evec = m.solve_all(k, eig_vectors=True)
    assert np.max(np.abs(ev2 - ev_ref)) <= compare.TOL_EVAL * scale
    assert _residual(m, k, ev2, evec) <= 1e-11 * scale
    ham = np.array([np.asarray(m._gen_ham(kk)).reshape(m._nsta, m._nsta) for kk in k[:3]])
    assert np.max(np.abs(ham - orc.gen_ham(m, k[:3]))) <= compare.TOL_HAM * scale


@pytest.mark.parametrize("spec", [dict(norb=3, dim=2, nhop=8, nspin=1, seed=31), dict(norb=2, dim=2, nhop=6, nspin=2, seed=32),
                                  dict(norb=5, dim=2, nhop=14, nspin=1, seed=33), dict(norb=3, dim=3, nhop=12, nspin=1, seed=34),
                                  dict(norb=4, dim=1, nhop=7, nspin=2, seed=35), dict(norb=12, dim=2, nhop=40, nspin=1, seed=36)])
def test_random_models_berry_quantities_against_oracle(spec):
    """Generic (random, complex, non-symmetric) models through solve_on_grid + berry_phase (both branches, every
    axis) + berry_flux (every ordered pair of axes, per plaquette and summed) against the oracle; the same specs
    are checked oracle-vs-live-reference in tests/test_host_mirror_vs_reference.py."""
    from tests import oracle_api
    res = []
    for mod in (oracle_api, _mod()):
        with contextlib.redirect_stdout(io.StringIO()):
            m = M.random_model(mod, **spec)
        dim = spec["dim"]
        mesh = [6, 5, 4][:dim]
        w = mod.wf_array(m, mesh)
        out = dict(gaps=w.solve_on_grid([0.1, -0.2, 0.3][:dim]))
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
    bad = compare.compare_case("random_berry", res[1], res[0])
    assert not bad, "\n".join(bad)


def test_large_ribbon_and_slab_eigenvalues():
    """Configs 4/5 sizes on a few k-points: norb 200 ribbon, norb 499 slab."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rib = M.bn_ribbon(mod, 100)
    k = [[0.0], [0.123], [0.5]]
    assert np.max(np.abs(rib.solve_all(k) - orc.solve_all(rib, k))) <= 1e-10 * 4
    ev, evec = rib.solve_all(k, eig_vectors=True)
    assert _residual(rib, np.array(k), ev, evec) <= 1e-10
    slab = M.cubic_slab(mod, 250)
    k2 = [[0.1, 0.2], [0.0, 0.5]]
    assert np.max(np.abs(slab.solve_all(k2) - orc.solve_all(slab, k2))) <= 1e-10 * 4


def test_chern_number_full_mesh():
    """BASELINE config 2 at full size: 1024x1024 Haldane mesh, Chern integer exact,
    total flux equal to the sum of the plaquette phases, gaps match a subsample."""
    mod = _mod()
    m = M.haldane(mod, delta=0.0)
    w = mod.wf_array(m, [1025, 1025])
    gaps = w.solve_on_grid([-0.5, -0.5])
    flux = w.berry_flux([0])
    chern = flux / (2 * np.pi)
    assert abs(chern - round(chern)) < 1e-9 and abs(round(chern)) == 1
    plaq = w.berry_flux([0], individual_phases=True)
    assert plaq.shape == (1024, 1024)
    assert abs(plaq.sum() - flux) < 1e-8
    flux1 = w.berry_flux([1])
    assert abs(flux + flux1) < 1e-8            # the two bands carry opposite Chern numbers
    assert gaps.shape == (1,) and 0.5 < gaps[0] < 2.0
    # periodic images: last row/column equal the first up to the pbc phase
    wfs = w._wfs
    ph = np.exp(-2j * np.pi * m._orb[:, 0])
    assert np.max(np.abs(wfs[-1, 5] - wfs[0, 5] * ph)) < 1e-14
    # Kane-Mele: Z2-odd phase has zero total Chern number for the occupied pair
    km = M.kane_mele(mod, "odd")
    wk = mod.wf_array(km, [257, 257])
    wk.solve_on_grid([-0.5, -0.5])
    assert abs(wk.berry_flux([0, 1])) < 1e-7


def test_full_mesh_eigenvectors_satisfy_the_eigenproblem():
    """BASELINE config 2 at full size, both models: at 4000 random mesh points (images included) the stored
    rows are orthonormal eigenvectors of the oracle's H(k) — a size-independent property of every point —
    and the band gaps of those points are bounded below by the minimal gaps the kernel reduced."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rng = np.random.RandomState(12)
    for m in (M.haldane(mod, delta=0.0), M.kane_mele(mod, "odd")):
        n = m._nsta
        w = mod.wf_array(m, [1025, 1025])
        gaps = w.solve_on_grid([-0.5, -0.5])
        wfs = w._wfs.reshape(1025, 1025, n, n)
        ij = rng.randint(0, 1025, size=(4000, 2))
        ij[:40, 0] = 1024                                   # rows / columns written as periodic images
        ij[40:80, 1] = 1024
        k = -0.5 + ij / 1024.0
        ham = orc.gen_ham(m, k)
        u = wfs[ij[:, 0], ij[:, 1]]                         # [pts, band, orb]
        hu = np.einsum("pij,pbj->pbi", ham, u)
        ev = np.einsum("pbi,pbi->pb", u.conj(), hu).real
        assert np.max(np.abs(hu - ev[:, :, None] * u)) < 1e-11
        assert np.max(np.abs(np.einsum("pbi,pci->pbc", u.conj(), u) - np.eye(n))) < 1e-12
        assert np.all(np.diff(ev, axis=1) >= -1e-12)        # ascending
        ev_ref = orc.sol_ham(ham, False)
        assert np.max(np.abs(ev - ev_ref)) < 1e-10 * max(1.0, np.max(np.abs(ev_ref)))
        assert np.all((ev[:, 1:] - ev[:, :-1]).min(axis=0) >= gaps - 1e-12)


def test_solve_on_grid_subsample_vs_oracle():
    """129x129 Haldane / Kane-Mele grids against the oracle (gauge-invariant)."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    for m, occ in ((M.haldane(mod, 0.0), [0]), (M.kane_mele(mod, "even"), [0, 1])):
        w = mod.wf_array(m, [129, 129])
        gaps = w.solve_on_grid([-0.5, -0.5])
        wfs_ref, gaps_ref = orc.solve_on_grid(m, [129, 129], [-0.5, -0.5])
        assert np.max(np.abs(gaps - gaps_ref)) < 1e-10
        ref_plaq = orc.berry_flux(wfs_ref, 2, occ, None, True)
        got_plaq = w.berry_flux(occ, individual_phases=True)
        assert np.max(np.abs(compare.circ_diff(got_plaq, ref_plaq, 2 * np.pi))) < 1e-8
        ref_ph = orc.berry_phase(wfs_ref, 2, occ, 1, contin=False)
        got_ph = w.berry_phase(occ, 1, contin=False)
        assert np.max(np.abs(compare.circ_diff(got_ph, ref_ph, 2 * np.pi))) < 1e-8


def test_w90_silicon_fixture():
    """Config 3 in miniature: the silicon Wannier90 model (arrays from the
    reference's parse) on the fixture k-points."""
    mod = _mod()
    z = np.load(os.path.join(GOLD, "w90.npz"))
    for tag in ("full", "small"):
        pre = "silicon_%s_" % tag
        m = mod.tb_model(3, 3, z[pre + "lat"], z[pre + "orb"])
        m.set_onsite(z[pre + "site_energies"].real)
        m._bulk_set_hops(z[pre + "hop_amp"], z[pre + "hop_i"], z[pre + "hop_j"], z[pre + "hop_R"])
        ev = m.solve_all(z["silicon_k"])
        want = z["silicon_%s_evals" % tag]
        assert np.max(np.abs(ev - want)) <= 1e-10 * max(1.0, np.max(np.abs(want)))
    # Wannier90's own interpolation (text precision of silicon_band.dat)
    full = mod.tb_model(3, 3, z["silicon_full_lat"], z["silicon_full_orb"])
    full.set_onsite(z["silicon_full_site_energies"].real)
    full._bulk_set_hops(z["silicon_full_hop_amp"], z["silicon_full_hop_i"], z["silicon_full_hop_j"], z["silicon_full_hop_R"])
    assert np.max(np.abs(full.solve_all(z["silicon_band_kpts"]) - z["silicon_band_ene"])) < 1e-4


def test_edge_cases():
    mod = _mod()
    one = mod.tb_model(1, 1, [[1.0]], [[0.0]])                       # single orbital
    one.set_onsite([0.3])
    one.set_hop(-1.0, 0, 0, [1])
    k = np.linspace(0, 1, 7)[:, None]
    assert np.allclose(one.solve_all(k)[0], 0.3 - 2 * np.cos(2 * np.pi * k[:, 0]), atol=1e-13)
    ev, evec = one.solve_all(k, eig_vectors=True)
    assert evec.shape == (1, 7, 1) and np.allclose(np.abs(evec), 1.0)
    w = mod.wf_array(one, [8])
    assert w.solve_on_grid([0.0]) is None
    assert abs(w.berry_phase([0])) < 1e-12 or abs(abs(w.berry_phase([0])) - 2 * np.pi) < 1e-12
    mol = M.molecule(mod)                                            # 0-D: no k argument
    assert mol.solve_all().shape == (3,)
    with pytest.raises(Exception):
        M.haldane(mod).solve_all([[0.1, 0.2, 0.3]])                  # wrong k shape
    with pytest.raises(Exception):
        mod.wf_array(M.haldane(mod), [5]).solve_on_grid([0.0])       # dim mismatch
    empty = M.haldane(mod).solve_all(np.zeros((0, 2)))               # empty k list
    assert empty.shape == (2, 0)


def _band_projectors(wfs, nsta):
    """Gauge-invariant per-band projectors |u><u| of a _wfs array [..., band, orb(,spin)]."""
    v = wfs.reshape((-1, nsta, nsta))
    return np.einsum("kbi,kbj->kbij", v, v.conj())


@pytest.mark.parametrize("spec,mesh", [
    (dict(norb=2, dim=1, nhop=3, nspin=1, seed=41), [203]),            # 1-D mesh, n=2
    (dict(norb=3, dim=2, nhop=5, nspin=1, seed=42), [6, 131]),          # n=3, NPH<=4? (random) / Jacobi
    (dict(norb=2, dim=2, nhop=7, nspin=2, seed=43), [5, 70]),           # spinor n=4, nph up to 8
    (dict(norb=2, dim=3, nhop=4, nspin=1, seed=44), [4, 5, 66]),        # 3-D mesh
    (dict(norb=4, dim=2, nhop=4, nspin=1, seed=45), [9, 50]),           # n=4 scalar
    (dict(norb=2, dim=2, nhop=16, nspin=1, seed=46), [5, 64]),          # nph > 8 -> generic kernel
    (dict(norb=2, dim=2, nhop=4, nspin=1, seed=47), [70, 5]),           # short fastest axis -> generic kernel
])
def test_mesh_kernels_against_oracle(spec, mesh):
    """wf_array.solve_on_grid through every mesh kernel variant: gaps, per-band
    projectors at every mesh point (incl. periodic images), plaquette phases."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    m = M.random_model(mod, **spec)
    start = [0.13, -0.27, 0.4][:len(mesh)]
    w = mod.wf_array(m, mesh)
    gaps = w.solve_on_grid(start)
    wfs_ref, gaps_ref = orc.solve_on_grid(m, mesh, start)
    assert np.max(np.abs(gaps - gaps_ref)) < 1e-10
    wfs = np.array(w._wfs)
    assert wfs.shape == wfs_ref.shape
    n = m._nsta
    if np.min(gaps_ref) > 1e-3:      # projectors are only defined for non-degenerate bands
        pa, pb = _band_projectors(wfs, n), _band_projectors(wfs_ref, n)
        assert np.max(np.abs(pa - pb)) < 1e-9
    if len(mesh) >= 2:
        got = w.berry_flux([0], individual_phases=True)
        ref = orc.berry_flux(wfs_ref, len(mesh), [0], None, True)
        assert np.max(np.abs(compare.circ_diff(got, ref, 2 * np.pi))) < 1e-8
        assert abs(compare.circ_diff(np.sum(w.berry_flux([0])), ref.sum(), 2 * np.pi)) < 1e-7
    else:
        got = w.berry_phase([0])
        ref = orc.berry_phase(wfs_ref, 1, [0])
        assert abs(compare.circ_diff(got, ref, 2 * np.pi)) < 1e-8


@pytest.mark.parametrize("which", ["haldane", "kane_mele", "random7"])
def test_shard_slabs_bit_identical(which):
    """Multi-GPU slabs emulated on one GPU: solving rows [row0, row0+nrows] with the
    closing row recomputed in-launch (wrap0=2) gives exactly the rows of the unsharded array."""
    mod = _mod()
    from pythtb_b200 import _engine
    eng = _engine.get_engine()
    if which == "haldane":
        m, mesh = M.haldane(mod, 0.0), [41, 130]
    elif which == "kane_mele":
        m, mesh = M.kane_mele(mod, "odd"), [23, 67]
    else:
        m, mesh = M.random_model(mod, norb=7, dim=2, nhop=12, nspin=1, seed=3), [13, 9]
    full = mod.wf_array(m, mesh)
    gaps_full = full.solve_on_grid([-0.5, -0.5])
    ref = np.array(full._wfs)
    world = 3
    gaps_min = None
    for rank in range(world):
        w = mod.wf_array(m, mesh, shard=(rank, world), halo="recompute")
        sh = w._shard
        g = eng.solve_grid(w._model, w._store, w._mesh_arr, np.array([-0.5, -0.5]), row0=sh.row0, nrows=sh.nrows, wrap0=2)
        g = g.cpu().numpy()
        gaps_min = g if gaps_min is None else np.minimum(gaps_min, g)
        got = np.array(w._wfs)
        assert got.shape[0] == sh.nrows + 1
        assert np.array_equal(got, ref[sh.row0:sh.row0 + sh.nrows + 1]), (which, rank)
        # halo-exchange variant: the closing row is left untouched by the launch
        w2 = mod.wf_array(m, mesh, shard=(rank, world), halo="exchange")
        eng.solve_grid(w2._model, w2._store, w2._mesh_arr, np.array([-0.5, -0.5]), row0=sh.row0, nrows=sh.nrows, wrap0=0)
        got2 = np.array(w2._wfs)
        assert np.array_equal(got2[:-1], ref[sh.row0:sh.row0 + sh.nrows])
        assert np.all(got2[-1] == 0)
    assert np.array_equal(gaps_min, gaps_full)


@pytest.mark.parametrize("which,mesh", [("haldane", [300, 200]), ("kane_mele", [70, 161]), ("haldane", [9, 64])])
def test_flux_ring_kernel_matches_register_kernel(which, mesh):
    """The cp.async.bulk ring variant of the plaquette kernel (TBK_FLUX_RING=1) against the
    default register-prefetch kernel and the oracle: per-plaquette phases and totals."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    m, occ = (M.haldane(mod, 0.0), [0]) if which == "haldane" else (M.kane_mele(mod, "odd"), [0, 1])
    w = mod.wf_array(m, mesh)
    # the ring kernel streams whole rows of k-points ([k..., state, orb] storage): give this array the reference
    # layout on the device instead of the state-major default of 2..4-band arrays
    w._store = w._model._engine().new_store(w._store.shape)
    w.solve_on_grid([-0.5, -0.5])
    base_plaq = w.berry_flux(occ, individual_phases=True)
    base_tot = w.berry_flux(occ)
    os.environ["TBK_FLUX_RING"] = "1"
    try:
        ring_plaq = w.berry_flux(occ, individual_phases=True)
        ring_tot = w.berry_flux(occ)
        from pythtb_b200 import _engine, _lib
        assert _lib.last_kernel(_engine.get_engine().lib) == "flux_ring_kernel"
    finally:
        os.environ["TBK_FLUX_RING"] = "0"
    assert np.max(np.abs(compare.circ_diff(ring_plaq, base_plaq, 2 * np.pi))) < 1e-12
    assert abs(ring_tot - base_tot) < 1e-9
    ref = orc.berry_flux(np.array(w._wfs), 2, occ, None, True)
    assert np.max(np.abs(compare.circ_diff(ring_plaq, ref, 2 * np.pi))) < 1e-8


@pytest.mark.parametrize("ncell", [12, 40])
def test_position_operator_tensor_path(ncell):
    """position_matrix / position_hwf with nocc >= 16 (DMMA GEMM kernels; nocc = 40 also takes the blocked
    eigensolver for the position matrix) against the oracle on the same eigenvectors."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rib = M.bn_ribbon(mod, ncell)
    n = rib._nsta
    nocc = n // 2
    ks = [[0.11], [0.37]]
    ev, evec = rib.solve_all(ks, eig_vectors=True)
    for i in range(len(ks)):
        occ = np.ascontiguousarray(evec[:nocc, i])
        x = rib.position_matrix(occ, 1)
        x_ref = orc.position_matrix(rib, occ, 1)
        assert np.max(np.abs(x - x_ref)) < 1e-11
        hwfc = rib.position_hwf(occ, 1)
        hwfc_ref = orc.position_hwf(rib, occ, 1)
        assert np.max(np.abs(hwfc - hwfc_ref)) < 1e-9
        c2, hwf = rib.position_hwf(occ, 1, hwf_evec=True, basis="orbital")
        assert np.max(np.abs(c2 - hwfc_ref)) < 1e-9
        # each hybrid Wannier function is a unit vector in the occupied subspace and an eigenvector of P X P
        hw = hwf.reshape(nocc, -1)
        assert np.max(np.abs(np.sum(np.abs(hw) ** 2, axis=1) - 1.0)) < 1e-10
        pos = np.asarray(rib._orb)[:, 1]
        xw = np.einsum("io,o,jo->ij", hw.conj(), pos, hw)
        assert np.max(np.abs(xw - np.diag(c2))) < 1e-9


@pytest.mark.parametrize("which", ["ribbon12", "ribbon40", "slab10"])
def test_wilson_loop_spectrum_large_nocc(which):
    """berry_phase(..., berry_evals=True) for nocc >= 8: Newton-Schulz polar factors on the DMMA GEMM, tree
    product along the string, eigenphases through the Hermitian solver — against the oracle's SVD / eigvals
    (pythtb.py:3821-3838) on the same wave functions."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    if which.startswith("ribbon"):
        m = M.bn_ribbon(mod, int(which[6:]))
        mesh, start, d = [37], [0.0], 0
    else:
        m = M.cubic_slab(mod, 10)
        mesh, start, d = [9, 11], [0.0, 0.0], 0
    nocc = m._nsta // 2
    w = mod.wf_array(m, mesh)
    w.solve_on_grid(start)
    got = w.berry_phase(range(nocc), d, contin=False, berry_evals=True)
    ref = orc.berry_phase(np.array(w._wfs), len(mesh), list(range(nocc)), d, contin=False, berry_evals=True)
    assert np.shape(got) == np.shape(ref)
    ok, dev = compare.sets_close(got, ref, 2 * np.pi, 1e-8)
    assert ok, dev
    # total phase of the spectrum = Berry phase of the determinant branch
    tot = w.berry_phase(range(nocc), d, contin=False)
    assert np.max(np.abs(compare.circ_diff(np.sum(got, axis=-1), tot, 2 * np.pi))) < 1e-7


def test_position_hwf_all_matches_per_point_calls():
    """wf_array.position_hwf_all (one batched launch, device resident) against the reference-style loop of
    per-point position_hwf calls that fills a second wf_array (examples/cubic_slab_hwf.py:63-91), including the
    Berry phases of the individual hybrid Wannier bands."""
    mod = _mod()
    nl = 10
    slab = M.cubic_slab(mod, nl)
    mesh = [6, 7]
    bloch = mod.wf_array(slab, mesh)
    bloch.solve_on_grid([0.0, 0.0])
    occ = list(range(nl))
    hwfc_all, hwf_all = bloch.position_hwf_all(occ, 2, hwf_evec=True)
    manual = mod.wf_array(slab, mesh, nsta_arr=nl)
    hwfc = np.zeros(tuple(mesh) + (nl,))
    for ix in range(mesh[0]):
        for iy in range(mesh[1]):
            (val, vec) = bloch.position_hwf([ix, iy], occ=occ, dir=2, hwf_evec=True, basis="orbital")
            hwfc[ix, iy] = val
            manual[ix, iy] = vec
    assert np.max(np.abs(hwfc_all - hwfc)) < 1e-10
    assert np.max(np.abs(bloch.position_hwf_all(occ, 2) - hwfc)) < 1e-10
    for arr in (hwf_all, manual):
        arr.impose_pbc(0, 0)
        arr.impose_pbc(1, 1)
    for band in (0, nl // 2, nl - 1):
        a = hwf_all.berry_phase([band], dir=0, contin=False)
        b = manual.berry_phase([band], dir=0, contin=False)
        assert np.max(np.abs(compare.circ_diff(a, b, 2 * np.pi))) < 1e-8


def test_replayed_calls_follow_their_arguments():
    """The replay records of wf_array.solve_on_grid / berry_flux (tbk_prepared_run) are used only while
    the call is the same; a different start_k, occupied set, model parameter or a user write to the
    array goes back through the full path.  Every answer is checked against the oracle."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    mesh = [65, 66]

    def reference(model, start, occ):
        wfs_ref, gaps_ref = orc.solve_on_grid(model, mesh, start)
        return gaps_ref, orc.berry_flux(wfs_ref, 2, occ)

    m = M.kane_mele(mod, "odd")
    w = mod.wf_array(m, mesh)
    for start, occ in (([-0.5, -0.5], [0, 1]), ([-0.5, -0.5], [0, 1]), ([-0.5, -0.5], [0, 1]), ([0.1, 0.2], [0, 1]),
                       ([0.1, 0.2], [0, 1]), ([0.1, 0.2], [2, 3]), ([0.1, 0.2], [2, 3]), ([-0.5, -0.5], [0, 1])):
        gaps = w.solve_on_grid(start)
        flux = w.berry_flux(occ)
        gaps_ref, flux_ref = reference(m, start, occ)
        assert np.max(np.abs(gaps - gaps_ref)) < 1e-10, (start, occ)
        assert abs(compare.circ_diff(flux, flux_ref, 2 * np.pi)) < 1e-8, (start, occ)
    assert w._rp_sg is not None and w._rp_fx is not None        # the last calls were replayable
    # results are fresh copies, not views of the pinned buffer the next call overwrites
    g1 = w.solve_on_grid([-0.5, -0.5])
    g1_keep = g1.copy()
    w.solve_on_grid([0.1, 0.2])
    w.solve_on_grid([0.1, 0.2])
    assert np.array_equal(g1, g1_keep)
    # a write through the host view must be seen by the next flux (no replay over stale device data)
    w.solve_on_grid([-0.5, -0.5])
    f0 = w.berry_flux([0, 1])
    f0b = w.berry_flux([0, 1])
    assert f0 == f0b
    host = w._wfs
    host[3, 4] = host[3, 4] * np.exp(0.3j)                      # a gauge change: the flux must not move
    host[5, 6, 0] = (host[5, 6, 0] + 0.5 * host[5, 6, 2]) / np.sqrt(1.25)   # a real change: it must
    f1 = w.berry_flux([0, 1])
    ref = orc.berry_flux(np.array(w._wfs), 2, [0, 1])
    assert abs(compare.circ_diff(f1, ref, 2 * np.pi)) < 1e-8        # (the total is 2 pi C whatever was written:
    p1 = w.berry_flux([0, 1], individual_phases=True)               #  the edit shows in the plaquettes around it)
    pref = orc.berry_flux(np.array(w._wfs), 2, [0, 1], None, True)
    assert np.max(np.abs(compare.circ_diff(p1, pref, 2 * np.pi))) < 1e-8
    w2 = mod.wf_array(m, mesh)
    w2.solve_on_grid([-0.5, -0.5])
    assert np.max(np.abs(compare.circ_diff(p1, w2.berry_flux([0, 1], individual_phases=True), 2 * np.pi))) > 1e-3
    # a grid solve overwrites every element: it neither uploads nor is misled by a host mirror that was handed
    # out (and scribbled on) since the last solve — replayed or not
    fresh = np.array(w2._wfs)
    for _ in range(2):
        host = w._wfs
        host[...] = 0.0
        w.solve_on_grid([-0.5, -0.5])
        assert w._store.state == "device"
        assert np.array_equal(np.array(w._wfs), fresh)
    # a model edit rebuilds the plan: the solve must follow it
    h = M.haldane(mod, delta=0.0)
    wh = mod.wf_array(h, mesh)
    a = wh.solve_on_grid([-0.5, -0.5]); wh.solve_on_grid([-0.5, -0.5])
    wh._model.set_onsite([-0.7, 0.7], mode="reset")
    b = wh.solve_on_grid([-0.5, -0.5])
    _, gaps_ref = orc.solve_on_grid(wh._model, mesh, [-0.5, -0.5])
    assert np.max(np.abs(b - gaps_ref)) < 1e-10 and abs(a[0] - b[0]) > 1e-3


def test_five_to_eight_band_register_paths_edge_cases():
    """The one-k-point-per-thread kernels of 5..8-band models on their edge cases: a finite (dim_k = 0) cluster (no
    phases, no dense coefficient table -> scalar assembly), a single k-point (127 idle lanes), a model with on-site terms
    only, eigenvalues only and with eigenvectors, against the oracle."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    rng = np.random.RandomState(77)
    mol = mod.tb_model(0, 2, [[1.0, 0.0], [0.0, 1.0]], rng.rand(6, 2).tolist())
    mol.set_onsite(rng.randn(6).tolist())
    for i in range(6):
        for j in range(i + 1, 6):
            if rng.rand() < 0.7:
                mol.set_hop(complex(rng.randn(), rng.randn()), i, j)
    ev = mol.solve_all()
    assert np.max(np.abs(ev - orc.solve_all(mol))) < 1e-10
    ev2, vec = mol.solve_all(eig_vectors=True)
    assert np.max(np.abs(ev2 - ev)) < 1e-12
    h = orc.gen_ham(mol).reshape(6, 6)
    assert np.max(np.abs(h @ vec.T - vec.T * ev2[None, :])) < 1e-11
    for norb in (5, 6, 7, 8):
        m = M.random_model(mod, norb=norb, dim=2, nhop=3 * norb, nspin=1, seed=100 + norb)
        for k in ([[0.123, -0.4]], np.random.RandomState(norb).rand(129, 2)):
            k = np.asarray(k)
            ref = orc.solve_all(m, k)
            assert np.max(np.abs(m.solve_all(k) - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
            evv, vecv = m.solve_all(k, eig_vectors=True)
            assert np.max(np.abs(evv - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
            assert _residual(m, k, evv, vecv) < 1e-11 * max(1.0, np.max(np.abs(ref)))
    flat = mod.tb_model(1, 1, [[1.0]], [[0.1 * i] for i in range(5)])
    flat.set_onsite([0.3, -0.2, 0.3, 1.5, -0.2])               # no hoppings at all: degenerate pairs, no phases
    kk = [[0.0], [0.3]]
    assert np.max(np.abs(flat.solve_all(kk) - orc.solve_all(flat, kk))) < 1e-12
    evf, vecf = flat.solve_all(kk, eig_vectors=True)
    assert _residual(flat, np.array(kk), evf, vecf) < 1e-12


def test_pipelined_host_results_equal_the_single_shot():
    """solve_all with host results on a long k-list is cut into chunks whose copies overlap the next chunk's kernels
    (engine.solve_all_host); with a small chunk size the chunked sweep must equal the single-shot one bit for bit —
    eigenvalues and eigenvectors, last partial chunk included."""
    mod = _mod()
    from pythtb_b200 import _engine
    eng = _engine.get_engine()
    old = eng.CHUNK_K
    try:
        for model in (M.haldane(mod, delta=0.2), M.random_model(mod, norb=8, dim=3, nhop=40, nspin=1, seed=5),
                      M.random_model(mod, norb=3, dim=2, nhop=8, nspin=2, seed=9)):
            k = np.random.RandomState(3).rand(3500, model._dim_k)
            eng.CHUNK_K = 1 << 21
            ev0 = model.solve_all(k)
            ev0v, vec0 = model.solve_all(k, eig_vectors=True)
            eng.CHUNK_K = 1000
            ev1 = model.solve_all(k)
            ev1v, vec1 = model.solve_all(k, eig_vectors=True)
            assert ev1.shape == ev0.shape and np.array_equal(ev0, ev1)
            assert vec1.shape == vec0.shape and np.array_equal(ev0v, ev1v) and np.array_equal(vec0, vec1)
    finally:
        eng.CHUNK_K = old


def test_solve_all_on_a_device_generated_mesh():
    """solve_all(k_uniform_mesh(mesh, lazy=True)): k-points generated on the device (tbk_kmesh_uniform) give
    exactly what the host list gives, for values, vectors (spinor layout included) and device-resident results."""
    from oracle import pythtb_oracle as orc
    mod = _mod()
    for model, mesh in ((M.haldane(mod, delta=0.2), [12, 9]), (M.kane_mele(mod, "odd"), [7, 8]),
                        (M.random_model(mod, norb=8, dim=3, nhop=40, nspin=1, seed=5), [5, 4, 6]),
                        (M.random_model(mod, norb=3, dim=1, nhop=5, nspin=1, seed=22), [33])):
        eager = model.k_uniform_mesh(mesh)
        lazy = model.k_uniform_mesh(mesh, lazy=True)
        ev_e, vec_e = model.solve_all(eager, eig_vectors=True)
        ev_l, vec_l = model.solve_all(lazy, eig_vectors=True)
        assert np.array_equal(ev_e, ev_l) and np.array_equal(vec_e, vec_l) and vec_l.shape == vec_e.shape
        assert np.max(np.abs(model.solve_all(lazy) - orc.solve_all(model, eager))) < 1e-10
        ev_d = model.solve_all(lazy, device_result=True)
        # (eigenvalues-only calls of 5..8-band models run the register solver, calls with eigenvectors the tile solver:
        # bit-identical among themselves, equal to rounding between them)
        assert ev_d.is_cuda and np.array_equal(ev_d.cpu().numpy(), model.solve_all(eager))
        assert np.max(np.abs(ev_d.cpu().numpy() - ev_e)) < 1e-12 * max(1.0, np.max(np.abs(ev_e)))
    with pytest.raises(Exception, match="wrong shape"):
        M.haldane(mod).solve_all(M.kane_mele(mod, "odd").k_uniform_mesh([3, 3], lazy=True)[:0] if False else
                                 M.random_model(mod, norb=3, dim=1, nhop=5, nspin=1, seed=22).k_uniform_mesh([4], lazy=True))
