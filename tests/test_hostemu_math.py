"""CPU unit tests of the DEVICE arithmetic: the TBK_HD headers of pythtb_b200/csrc compiled by
g++ (tests/hostemu) and checked against numpy/LAPACK — eigh2, register Jacobi (n = 3, 4), the group
Householder+QL solver, link determinants, polar factors and unitary eigenvalues.  Lets the math of
the CUDA kernels be verified in the GPU-less build container; the product never uses this build."""
import ctypes

import numpy as np
import pytest

from tests import hostemu

DP = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(DP)


def _rand_herm(rng, n, degenerate=False):
    a = rng.randn(n, n) + 1j * rng.randn(n, n)
    h = a + a.conj().T
    if degenerate:                       # exactly repeated eigenvalues (Kramers-like pairs)
        q, _ = np.linalg.qr(a)
        lam = np.repeat(rng.randn((n + 1) // 2), 2)[:n]
        h = (q * lam) @ q.conj().T
        h = 0.5 * (h + h.conj().T)
    return h


def _check_eig(h, ev, w, tol=1e-12):
    """rows of w are eigenvectors (not conjugated): H w[b]^T = ev[b] w[b]^T."""
    n = h.shape[0]
    scale = max(1.0, np.max(np.abs(h)))
    assert np.all(np.diff(ev) >= 0)
    assert np.max(np.abs(ev - np.linalg.eigvalsh(h))) < 1e-12 * scale * n
    assert np.max(np.abs(h @ w.T - w.T * ev[None, :])) < tol * scale * n
    assert np.max(np.abs(w.conj() @ w.T - np.eye(n))) < tol * n


def test_eigh2_closed_form():
    lib = hostemu.lib()
    rng = np.random.RandomState(1)
    cases = [(rng.randn(), rng.randn(), complex(rng.randn(), rng.randn())) for _ in range(200)]
    cases += [(1.0, 1.0, 0j), (2.0, -1.0, 0j), (-1.0, 2.0, 0j), (0.3, 0.3, 1e-9 + 0j), (1.0, 1.0 + 1e-13, 1e-3j),
              (1e8, -1e8, 1.0 + 1j), (0.0, 0.0, 1e-140 + 0j)]  # below ~1e-150 |h10|^2 is denormal: documented limit
    for h00, h11, h10 in cases:
        ev = np.zeros(2)
        w = np.zeros((2, 2), dtype=complex)
        lib.emu_eigh2(ctypes.c_double(h00), ctypes.c_double(h11), ctypes.c_double(h10.real), ctypes.c_double(h10.imag),
                      _p(ev), _p(w.view(np.float64)))
        h = np.array([[h00, np.conj(h10)], [h10, h11]])
        _check_eig(h, ev, w)


def test_eigh2_fast_branch_free_variant():
    """The mesh kernels' branch-free 2x2 solver (tbk_eig_small.cuh: eigh2_fast) on generic, diagonal and
    EXACTLY degenerate matrices — H proportional to the identity is reachable on a mesh because the phases come
    from sincospi (cospi(0.5) == 0 at k = 0.5 of a two-site chain, a Dirac point of graphene on the mesh)."""
    lib = hostemu.lib()
    rng = np.random.RandomState(2)
    cases = [(rng.randn(), rng.randn(), complex(rng.randn(), rng.randn())) for _ in range(200)]
    cases += [(1.0, 1.0, 0j), (0.0, 0.0, 0j), (-3.5, -3.5, 0j), (2.0, -1.0, 0j), (-1.0, 2.0, 0j), (0.3, 0.3, 1e-9 + 0j),
              (1.0, 1.0 + 1e-13, 1e-3j), (1e8, -1e8, 1.0 + 1j), (0.0, 0.0, 1e-140 + 0j), (1.0, 1.0, 1e-200 + 0j),
              (1.0 + 1e-200, 1.0, 0j)]
    for h00, h11, h10 in cases:
        ev = np.zeros(2)
        w = np.zeros((2, 2), dtype=complex)
        lib.emu_eigh2_fast(ctypes.c_double(h00), ctypes.c_double(h11), ctypes.c_double(h10.real), ctypes.c_double(h10.imag),
                           _p(ev), _p(w.view(np.float64)))
        h = np.array([[h00, np.conj(h10)], [h10, h11]])
        _check_eig(h, ev, w)
        assert np.max(np.abs(np.sum(np.abs(w) ** 2, axis=1) - 1.0)) < 1e-14, (h00, h11, h10)


@pytest.mark.parametrize("n", [3, 4])
def test_register_jacobi(n):
    lib = hostemu.lib()
    rng = np.random.RandomState(10 + n)
    mats = [_rand_herm(rng, n) for _ in range(200)] + [_rand_herm(rng, n, degenerate=True) for _ in range(50)]
    mats += [np.diag(rng.randn(n)).astype(complex), np.zeros((n, n), dtype=complex), np.eye(n, dtype=complex) * 3.0]
    for h in mats:
        ev = np.zeros(n)
        w = np.zeros((n, n), dtype=complex)
        hc = np.ascontiguousarray(h)
        assert lib.emu_jacobi(n, _p(hc.view(np.float64)), _p(ev), _p(w.view(np.float64))) == 0
        _check_eig(h, ev, w)


@pytest.mark.parametrize("n", [3, 4])
def test_small_direct_solver(n):
    """Householder + implicit QL in registers (eigh_small_ql), the solver of the n = 3, 4 mesh kernels."""
    lib = hostemu.lib()
    rng = np.random.RandomState(30 + n)
    mats = [_rand_herm(rng, n) for _ in range(400)] + [_rand_herm(rng, n, degenerate=True) for _ in range(100)]
    mats += [np.diag(rng.randn(n)).astype(complex), np.zeros((n, n), dtype=complex), np.eye(n, dtype=complex) * 3.0]
    mats += [1e-9 * _rand_herm(rng, n) + np.eye(n), 1e6 * _rand_herm(rng, n)]
    tri = np.diag(rng.randn(n)).astype(complex) + np.diag(rng.randn(n - 1) + 1j * rng.randn(n - 1), -1)
    mats.append(tri + np.tril(tri, -1).conj().T)
    kram = np.kron(_rand_herm(rng, 2), np.eye(2))[:n, :n]          # Kramers-like exact pairs
    mats.append(kram)
    for h in mats:
        ev = np.zeros(n)
        w = np.zeros((n, n), dtype=complex)
        hc = np.ascontiguousarray(h)
        assert lib.emu_small_ql(n, _p(hc.view(np.float64)), _p(ev), _p(w.view(np.float64))) == 1
        _check_eig(h, ev, w)


@pytest.mark.parametrize("n", [3, 4, 5, 6, 7, 8])
def test_small_eigenvalues_only_solver(n):
    """eigvals_small<N>: the one-matrix-per-thread register solver of eigenvalue-only sweeps (n = 5..8 Wannier models)."""
    lib = hostemu.lib()
    rng = np.random.RandomState(70 + n)
    mats = [_rand_herm(rng, n) for _ in range(300)] + [_rand_herm(rng, n, degenerate=True) for _ in range(60)]
    mats += [np.diag(rng.randn(n)).astype(complex), np.zeros((n, n), dtype=complex), np.eye(n, dtype=complex) * 3.0,
             1e-9 * _rand_herm(rng, n) + np.eye(n), 1e6 * _rand_herm(rng, n)]
    tri = np.diag(rng.randn(n)).astype(complex) + np.diag(rng.randn(n - 1) + 1j * rng.randn(n - 1), -1)
    mats.append(tri + np.tril(tri, -1).conj().T)
    for h in mats:
        ev = np.zeros(n)
        hc = np.ascontiguousarray(h, dtype=complex)
        assert lib.emu_eigvals_small(n, _p(hc.view(np.float64)), _p(ev)) == 1
        scale = max(1.0, np.max(np.abs(h)))
        assert np.all(np.diff(ev) >= 0)
        assert np.max(np.abs(ev - np.linalg.eigvalsh(h))) < 1e-12 * scale * n


@pytest.mark.parametrize("n", [3, 5, 6, 7, 8])
def test_small_solver_with_vectors_through_memory(n):
    """eigh_small_mem<N>: reflectors in registers, the QL rotation matrix in a strided memory column, eigenvectors
    back-transformed one at a time (the n = 5..8 path with eigenvectors)."""
    lib = hostemu.lib()
    rng = np.random.RandomState(90 + n)
    mats = [_rand_herm(rng, n) for _ in range(200)] + [_rand_herm(rng, n, degenerate=True) for _ in range(60)]
    mats += [np.diag(rng.randn(n)).astype(complex), np.zeros((n, n), dtype=complex), np.eye(n, dtype=complex) * 3.0,
             1e-9 * _rand_herm(rng, n) + np.eye(n), 1e6 * _rand_herm(rng, n)]
    for h in mats:
        ev = np.zeros(n)
        w = np.zeros((n, n), dtype=complex)
        hc = np.ascontiguousarray(h, dtype=complex)
        assert lib.emu_eigh_small_mem(n, _p(hc.view(np.float64)), _p(ev), _p(w.view(np.float64))) == 1
        _check_eig(h, ev, w)


def test_small_direct_solver_n4():
    """eigh4_direct, the n = 4 solver of the mesh kernels: closed-form roots of the tridiagonal's quartic + Newton on the
    Sturm recurrence + adjugate-column eigenvectors, with the implicit-QL lane behind it for close or multiple roots.
    Both lanes must meet the same tolerances; generic spectra must stay on the fast lane."""
    lib = hostemu.lib()
    rng = np.random.RandomState(44)

    def with_spectrum(lam):
        a = rng.randn(4, 4) + 1j * rng.randn(4, 4)
        q, _ = np.linalg.qr(a)
        h = (q * np.asarray(lam)) @ q.conj().T
        return 0.5 * (h + h.conj().T)

    def run(h):
        ev = np.zeros(4)
        w = np.zeros((4, 4), dtype=complex)
        hc = np.ascontiguousarray(h, dtype=complex)
        lane = lib.emu_eigh4_direct(_p(hc.view(np.float64)), _p(ev), _p(w.view(np.float64)))
        assert lane in (0, 1)
        _check_eig(h, ev, w)
        return lane

    generic = [_rand_herm(rng, 4) for _ in range(2000)]
    lanes = [run(h) for h in generic]
    assert np.mean(lanes) < 0.03                                   # almost always the fast lane
    for gap in (1e-1, 1e-2, 2e-3, 1e-3, 5e-4, 1e-4, 1e-6, 1e-9, 1e-12, 0.0):
        for _ in range(100):
            a, b, c = rng.randn(3)
            run(with_spectrum([a, a + gap, b, c]))
            run(with_spectrum([a, a + gap, b, b + gap]))
            run(with_spectrum([a, a + gap, a + 2 * gap, b]))
    special = [np.diag(rng.randn(4)).astype(complex), np.zeros((4, 4), dtype=complex), np.eye(4, dtype=complex) * 3.0,
               1e-9 * _rand_herm(rng, 4) + np.eye(4), 1e6 * _rand_herm(rng, 4), 1e-9 * _rand_herm(rng, 4),
               _rand_herm(rng, 4) + 1e3 * np.eye(4), np.kron(_rand_herm(rng, 2), np.eye(2)), np.kron(np.eye(2), _rand_herm(rng, 2))]
    tri = np.diag(rng.randn(4)).astype(complex) + np.diag(rng.randn(3) + 1j * rng.randn(3), -1)
    special.append(tri + np.tril(tri, -1).conj().T)
    blk = np.zeros((4, 4), dtype=complex)
    blk[:2, :2] = _rand_herm(rng, 2); blk[2:, 2:] = _rand_herm(rng, 2)
    special.append(blk)
    for h in special:
        run(h)
    # the Kane-Mele model of the headline workload on a mesh through the TRIM points (exact Kramers pairs there)
    import pythtb_b200 as tb
    from tests import models as M
    from oracle import pythtb_oracle as orc
    km = M.kane_mele(tb, "odd")
    nk = 33
    kpts = np.array([[i / (nk - 1) - 0.5, j / (nk - 1) - 0.5] for i in range(nk) for j in range(nk)])
    hams = orc.gen_ham(km, kpts)
    lanes = [run(hams[i].reshape(4, 4)) for i in range(len(kpts))]
    assert np.mean(lanes) < 0.02


@pytest.mark.parametrize("n", [1, 2, 5, 8, 17, 32, 40, 64])
def test_group_heev(n):
    lib = hostemu.lib()
    rng = np.random.RandomState(100 + n)
    for trial in range(6):
        h = _rand_herm(rng, n, degenerate=(trial == 5 and n > 2))
        lda = n | 1
        a = np.zeros((n, lda), dtype=complex)           # a[c, r] = A(r, c) column-major
        a[:, :n] = np.tril(h).T                          # lower triangle only; upper left as garbage-free zeros
        ev = np.zeros(n)
        vec = np.zeros((n, n), dtype=complex)
        info = lib.emu_heev_group(n, _p(a.view(np.float64)), lda, 1, _p(ev), _p(vec.view(np.float64)))
        assert info == 0
        _check_eig(h, ev, vec, tol=1e-12)
        a[:, :n] = np.tril(h).T
        ev2 = np.zeros(n)
        assert lib.emu_heev_group(n, _p(a.view(np.float64)), lda, 0, _p(ev2), None) == 0
        assert np.max(np.abs(ev2 - ev)) < 1e-12 * max(1.0, np.max(np.abs(ev))) * n


def test_link_det_polar_eigvals():
    lib = hostemu.lib()
    rng = np.random.RandomState(7)
    for nocc, n in ((1, 2), (2, 4), (3, 7), (6, 10), (12, 30)):
        a = rng.randn(nocc, n) + 1j * rng.randn(nocc, n)
        b = a + 0.3 * (rng.randn(nocc, n) + 1j * rng.randn(nocc, n))
        out = np.zeros(3)
        lib.emu_link_det(nocc, n, _p(np.ascontiguousarray(a).view(np.float64)), _p(np.ascontiguousarray(b).view(np.float64)), _p(out))
        det = np.linalg.det(a.conj() @ b.T)
        assert abs(complex(out[0], out[1]) - det / abs(det)) < 1e-11
        assert abs(out[2] - np.log(abs(det))) < 1e-10
        m = np.ascontiguousarray(a.conj() @ b.T)
        u, _, vh = np.linalg.svd(m)
        pol = m.copy()
        assert lib.emu_polar(nocc, _p(pol.view(np.float64))) > 0      # iteration count (-1 = singular)
        assert np.max(np.abs(pol - u @ vh)) < 1e-10
        uni = np.ascontiguousarray(u @ vh)
        want = np.sort(np.angle(np.linalg.eigvals(uni)))
        ev = np.zeros(nocc, dtype=complex)
        work = uni.copy()
        assert lib.emu_eigvals(nocc, _p(work.view(np.float64)), _p(ev.view(np.float64))) == 0
        got = np.sort(np.angle(ev))
        assert np.max(np.abs(np.exp(1j * got) - np.exp(1j * want))) < 1e-9
        assert np.max(np.abs(np.abs(ev) - 1.0)) < 1e-10


def _blocked_matrix(rng, n, kind):
    a = rng.randn(n, n) + 1j * rng.randn(n, n)
    h = a + a.conj().T
    if kind == "deg":                      # exactly repeated eigenvalues (pairs)
        q, _ = np.linalg.qr(a)
        lam = np.repeat(rng.randn((n + 1) // 2), 2)[:n]
        h = (q * lam) @ q.conj().T
    elif kind == "cluster":                # a third of the spectrum within 1e-9
        q, _ = np.linalg.qr(a)
        lam = np.concatenate([1.0 + 1e-9 * rng.randn(n // 3), rng.randn(n - n // 3)])
        h = (q * lam) @ q.conj().T
    elif kind == "diag":
        h = np.diag(rng.randn(n)).astype(complex)
    elif kind == "flat":                   # two flat bands: every eigenvalue n/2-fold degenerate (slab at k = (1/2, 1/2))
        q, _ = np.linalg.qr(a)
        lam = np.where(np.arange(n) % 2 == 0, -1.0, 1.0)
        h = (q * lam) @ q.conj().T
    elif kind == "flatdiag":
        h = np.diag(np.where(np.arange(n) % 2 == 0, -1.0, 1.0)).astype(complex) + 3.7e-17 * (np.eye(n, k=1) + np.eye(n, k=-1))
    elif kind == "ribbon":                 # bipartite chain with a staggered potential (banded, like a ribbon H(k))
        h = np.zeros((n, n), complex)
        for i in range(n - 1):
            h[i, i + 1] = -1.0 * (1 + np.exp(0.7j) * (i % 2))
        h = h + h.conj().T + np.diag(0.4 * (-1.0) ** np.arange(n))
    return 0.5 * (h + h.conj().T)


@pytest.mark.parametrize("n,nb", [(2, 8), (9, 8), (33, 8), (40, 16), (64, 8), (100, 16), (130, 7), (200, 8)])
def test_blocked_heev(n, nb):
    """tbk_eig_blocked.cuh (blocked tridiagonalisation, bisection, inverse iteration with cluster
    re-orthogonalisation, staged back-transformation) on random, exactly degenerate, tightly
    clustered, diagonal and banded matrices; nb odd exercises the per-column back-transformation."""
    lib = hostemu.lib()
    rng = np.random.RandomState(500 + n)
    for kind in ("rand", "deg", "cluster", "diag", "flat", "flatdiag", "ribbon"):
        lib.emu_set_hetrd_sym(0 if kind in ("deg", "ribbon") else 1)     # both tridiagonalisation variants
        h = _blocked_matrix(rng, n, kind)
        lda = n | 1
        a = np.zeros((n, lda), dtype=complex)
        a[:, :n] = np.tril(h).T
        ev = np.zeros(n)
        vec = np.zeros((n, n), dtype=complex)
        tri = np.zeros(2 * n)
        rc = lib.emu_heev_blocked(n, _p(a.view(np.float64)), lda, nb, 1, _p(ev), _p(vec.view(np.float64)), _p(tri))
        assert rc == 0, (kind, "fallback requested")
        scale = max(1.0, np.max(np.abs(h)))
        assert np.max(np.abs(ev - np.linalg.eigvalsh(h))) < 2e-13 * scale, kind
        # (a flat band is one n/2-fold cluster: what its last members keep after Gram-Schmidt is amplified by 1 / keep,
        #  tbk_eig_blocked.cuh phase 3 — still three orders below the 1e-10 parity tolerance)
        assert np.max(np.abs(h @ vec.T - vec.T * ev[None, :])) < (1e-12 if kind == "flat" else 5e-14) * scale, kind
        assert np.max(np.abs(vec.conj() @ vec.T - np.eye(n))) < 5e-12, kind
        a[:, :n] = np.tril(h).T
        ev2 = np.zeros(n)
        assert lib.emu_heev_blocked(n, _p(a.view(np.float64)), lda, nb, 0, _p(ev2), None, None) == 0
        assert np.array_equal(ev2, ev)


# ---------------------------------------------------------------------------------------------
# Hamiltonian assembly: the model compiler (pythtb_b200/_plan.py) + the device arithmetic of
# csrc/tbk_plan.cuh (phase table, element-major accumulation, Convention-I gauge), emulated on
# the host, against the oracle's restatement of tb_model._gen_ham (pythtb.py:874-925).
# ---------------------------------------------------------------------------------------------
class _PlanView(ctypes.Structure):
    _fields_ = [("dim_k", ctypes.c_int), ("nsta", ctypes.c_int), ("nph", ctypes.c_int), ("nel", ctypes.c_int),
                ("nterm", ctypes.c_int), ("convention", ctypes.c_int),
                ("ph_R", ctypes.c_void_p), ("tau", ctypes.c_void_p), ("el_ptr", ctypes.c_void_p),
                ("el_row", ctypes.c_void_p), ("el_col", ctypes.c_void_p), ("t_ph", ctypes.c_void_p),
                ("t_amp", ctypes.c_void_p), ("pm_ptr", ctypes.c_void_p), ("pm_el", ctypes.c_void_p),
                ("pm_amp", ctypes.c_void_p)]


def _emu_gen_ham(model, k):
    plan = model._plan()
    pv = _PlanView(plan.dim_k, plan.nsta, plan.nph, plan.nel, plan.nterm, plan.convention)
    for name in ("ph_R", "tau", "el_ptr", "el_row", "el_col", "t_ph", "t_amp", "pm_ptr", "pm_el", "pm_amp"):
        setattr(pv, name, getattr(plan, name).ctypes.data)
    k = np.ascontiguousarray(k, dtype=float).reshape(-1, max(plan.dim_k, 1))
    nk = k.shape[0]
    ham = np.zeros((nk, plan.nsta, plan.nsta), dtype=complex)
    fn = hostemu.lib().emu_gen_ham
    fn.restype = None
    fn.argtypes = [ctypes.POINTER(_PlanView), DP, ctypes.c_int64, DP]
    fn(ctypes.byref(pv), _p(k), nk, _p(ham.view(np.float64)))
    return ham, plan


def _phase_major_ham(plan, k):
    """The same H_II from the phase-major CSR the mesh kernels stream (numpy, Convention II)."""
    nk = k.shape[0]
    ham = np.zeros((nk, plan.nsta, plan.nsta), dtype=complex)
    amp = plan.pm_amp[:, 0] + 1j * plan.pm_amp[:, 1]
    for p in range(plan.nph + 1):
        ph = np.exp(2j * np.pi * (k @ plan.ph_R[p, :plan.dim_k])) if p < plan.nph else np.ones(nk)
        for t in range(plan.pm_ptr[p], plan.pm_ptr[p + 1]):
            e = int(plan.pm_el[t]) & ((1 << 30) - 1)
            z = ph.conj() if int(plan.pm_el[t]) & (1 << 30) else ph
            ham[:, plan.el_row[e], plan.el_col[e]] += amp[t] * z
    low = np.tril(np.ones((plan.nsta, plan.nsta), dtype=bool), -1)
    ham = ham + np.where(low, ham, 0).conj().transpose(0, 2, 1)
    idx = np.arange(plan.nsta)
    ham[:, idx, idx] = ham[:, idx, idx].real
    return ham


@pytest.mark.parametrize("convention", [1, 2])
def test_plan_compiler_and_device_assembly(convention):
    from oracle import pythtb_oracle as orc
    from tests import models as M, oracle_api as api
    zoo = [M.haldane(api, 0.2), M.kane_mele(api, "odd"), M.bn_ribbon(api, 7), M.cubic_slab(api, 5),
           M.three_site(api, 0.3), M.checkerboard(api), M.cubic_bulk(api),
           M.random_model(api, norb=5, dim=3, nhop=24, nspin=1, seed=3),
           M.random_model(api, norb=4, dim=3, nhop=15, nspin=2, seed=5),
           M.random_model(api, norb=9, dim=1, nhop=30, nspin=2, seed=8)]
    rng = np.random.RandomState(40 + convention)
    for m in zoo:
        m.set_convention(convention)
        k = rng.rand(11, m._dim_k) * 4.0 - 2.0
        got, plan = _emu_gen_ham(m, k)
        want = orc.gen_ham(m, k)
        scale = max(1.0, np.max(np.abs(want)))
        assert plan.convention == convention
        assert np.max(np.abs(got - want)) < 1e-13 * scale
        assert np.max(np.abs(got - got.conj().transpose(0, 2, 1))) < 1e-15 * scale   # the kernels also zero Im H_ii
        # the phase-major list is the same operator (always Convention II: the gauge is applied afterwards)
        m.set_convention(2)
        assert np.max(np.abs(_phase_major_ham(plan, k) - orc.gen_ham(m, k))) < 1e-13 * scale
    mol = M.molecule(api)                                     # dim_k = 0: no phases at all
    got, _ = _emu_gen_ham(mol, np.zeros((1, 1)))
    assert np.max(np.abs(got[0] - orc.gen_ham(mol, None)[0])) < 1e-14


# ---------------------------------------------------------------------------------------------
# The PARALLEL decomposition of the group algorithms: a CTA emulated by T host threads (one per CUDA
# thread, pthread barriers for __syncthreads / __syncwarp, sub-teams of S threads as warps) runs the
# same SPMD code as solve_blocked_kernel / solve_tile_kernel / solve_block_kernel.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,nb,T,S", [(24, 4, 16, 8), (45, 8, 64, 32), (33, 16, 32, 8), (70, 8, 96, 32)])
def test_blocked_heev_by_a_team_of_threads(n, nb, T, S):
    lib = hostemu.lib()
    rng = np.random.RandomState(900 + n)
    for kind in ("rand", "deg", "cluster", "flat", "ribbon"):
        lib.emu_set_hetrd_sym(0 if kind in ("cluster", "ribbon") else 1)  # both tridiagonalisation variants
        h = _blocked_matrix(rng, n, kind)
        lda = n | 1
        a = np.zeros((n, lda), dtype=complex)
        a[:, :n] = np.tril(h).T
        ev = np.zeros(n)
        vec = np.zeros((n, n), dtype=complex)
        rc = lib.emu_heev_blocked_team(n, _p(a.view(np.float64)), lda, nb, 1, _p(ev), _p(vec.view(np.float64)), T, S)
        assert rc == 0, (kind, rc)
        scale = max(1.0, np.max(np.abs(h)))
        assert np.max(np.abs(ev - np.linalg.eigvalsh(h))) < 2e-13 * scale, kind
        assert np.max(np.abs(h @ vec.T - vec.T * ev[None, :])) < (1e-12 if kind == "flat" else 5e-14) * scale, kind
        assert np.max(np.abs(vec.conj() @ vec.T - np.eye(n))) < 5e-12, kind
    lib.emu_set_hetrd_sym(1)


@pytest.mark.parametrize("n,T", [(5, 8), (13, 16), (32, 32), (48, 64)])
def test_group_heev_by_a_team_of_threads(n, T):
    lib = hostemu.lib()
    rng = np.random.RandomState(950 + n)
    for degenerate in (False, True):
        h = _rand_herm(rng, n, degenerate)
        lda = n | 1
        a = np.zeros((n, lda), dtype=complex)
        a[:, :n] = np.tril(h).T
        ev = np.zeros(n)
        vec = np.zeros((n, n), dtype=complex)
        assert lib.emu_heev_group_team(n, _p(a.view(np.float64)), lda, 1, _p(ev), _p(vec.view(np.float64)), T) == 0
        _check_eig(h, ev, vec)


def test_solvers_have_no_unsynchronised_accesses(tmp_path):
    """The same team emulation under ThreadSanitizer (tests/hostemu/tsan_driver.cpp): an access to the shared
    panels / workspace that no barrier orders — a missing __syncthreads() or __syncwarp() in the kernels — is a
    reported data race.  Random, exactly degenerate, banded and flat-band spectra; blocked and group solvers."""
    import os
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    src = os.path.join(os.path.dirname(hostemu.__file__), "tsan_driver.cpp")
    exe = str(tmp_path / "tsan_driver")
    build = subprocess.run([gxx, "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-ffp-contract=off",
                            "-I", hostemu.CSRC, src, "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if build.returncode != 0 and "tsan" in build.stdout.lower():
        pytest.skip("ThreadSanitizer runtime not available: " + build.stdout[-200:])
    assert build.returncode == 0, build.stdout
    runs = ["24 4 16 8 0 1", "30 8 32 8 1 2", "41 8 32 16 2 3", "36 8 32 8 3 4", "70 8 64 32 0 5", "50 16 64 32 3 6",
            "7 0 8 8 0 1", "12 0 16 16 1 2", "31 0 32 32 2 4", "40 0 64 64 3 5"]
    for args in runs:
        res = subprocess.run([exe] + args.split(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
        assert "ThreadSanitizer" not in res.stdout, args + "\n" + res.stdout[:3000]
        assert res.returncode == 0, args + "\n" + res.stdout[-500:]
