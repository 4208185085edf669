"""CPU oracle: a numpy restatement of PythTB 1.8.0's k-mesh hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``) may import it.  ``pythtb_b200`` never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (``/root/reference/pythtb.py``) in the build container, checks the
reference against its own ``tests/test_examples/*/*/golden_outputs/*.npy`` and
stores the reference outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures.  The arithmetic itself lives in numpy -> LAPACK (``zheevd``,
``zgesdd``, ``zgetrf``, ``zgeev``), exactly as in the reference (numpy is an
unpinned dependency of the reference, ``setup.py:17``).

Each function cites the reference lines it restates (``pythtb.py:<lines>``).
The restatement is vectorised over k-points / links / plaquettes (the
reference loops in Python), so it is a *faster* CPU baseline than the
reference itself, with the same operations per k-point.

A model is passed as any object exposing the reference's attribute names
(``_dim_k,_dim_r,_nspin,_norb,_nsta,_per,_orb,_site_energies,_hoppings``) —
both the reference ``tb_model`` and ``pythtb_b200.tb_model`` do.
"""
import numpy as np

__all__ = [
    "gen_ham", "sol_ham", "solve_all", "solve_one", "grid_kpoints",
    "solve_on_grid", "impose_pbc", "impose_loop", "one_berry_loop",
    "one_flux_plane", "berry_phase", "berry_flux", "no_2pi",
    "one_phase_cont", "array_phases_cont", "position_matrix",
    "position_expectation", "position_hwf", "convention_gauge",
]


# ----------------------------------------------------------------------------
# Hamiltonian assembly and diagonalisation
# ----------------------------------------------------------------------------
def gen_ham(model, k_list=None):
    """H(k) for a list of k-points; restates ``tb_model._gen_ham``
    (pythtb.py:874-925), Convention I.  A model carrying ``_convention == 2``
    (the ``pythtb_b200`` extension ``set_convention``; the reference has no
    such attribute) gets the Convention-II matrix of
    doc/formalism/pythtb-formalism.tex:341-344, i.e. the same sum without the
    orbital positions in the phase.  UNPINNED by reference fixtures (1.8.0 does
    not implement it); pinned instead on the identity tex:355-364 against the
    Convention-I result, see ``convention_gauge``.

    Returns complex128 ``[nk, nsta, nsta]`` (spinor index = 2*orb+spin, the
    reshape of pythtb.py:933).  For dim_k == 0 pass ``k_list=None`` and get
    ``[1, nsta, nsta]``.
    """
    norb, nspin, nsta = model._norb, model._nspin, model._nsta
    if model._dim_k == 0:
        kpts = np.zeros((1, 0))
    else:
        kpts = np.asarray(k_list, dtype=float).reshape(-1, model._dim_k)
    nk = kpts.shape[0]
    ham = np.zeros((nk, norb, nspin, norb, nspin), dtype=complex)
    # diagonal, pythtb.py:894-898
    for i in range(norb):
        if nspin == 1:
            ham[:, i, 0, i, 0] = model._site_energies[i]
        else:
            ham[:, i, :, i, :] = model._site_energies[i]
    per = list(model._per)
    # hoppings in list order, pythtb.py:900-924
    for hop in model._hoppings:
        amp = np.array(hop[0], dtype=complex).reshape(nspin, nspin)
        i, j = hop[1], hop[2]
        if model._dim_k > 0:
            ind_R = np.array(hop[3], dtype=float)
            rv = -model._orb[i, :] + model._orb[j, :] + ind_R   # :912
            if getattr(model, "_convention", 1) == 2:
                rv = ind_R                                      # tex:341-344
            rv = rv[per]                                        # :914
            phase = np.exp((2.0j) * np.pi * (kpts @ rv))        # :916
        else:
            phase = np.ones(nk, dtype=complex)
        blk = phase[:, None, None] * amp[None, :, :]            # :917
        ham[:, i, :, j, :] += blk                               # :920/:923
        ham[:, j, :, i, :] += blk.conj().transpose(0, 2, 1)     # :921/:924
    return ham.reshape(nk, nsta, nsta)


def convention_gauge(model, k_list):
    """D[k, j] = exp(2 pi i k.tau_j) per state (both spin components of an
    orbital share tau_j): H~ = D H D^H element-wise H~_ij = D_i H_ij conj(D_j)
    and C~_j = D_j C_j (doc/formalism/pythtb-formalism.tex:355-364) relate
    Convention II (tilde) to the reference's Convention I."""
    kpts = np.asarray(k_list, dtype=float).reshape(-1, model._dim_k)
    tau = np.asarray(model._orb, dtype=float)[:, list(model._per)]
    d = np.exp(2.0j * np.pi * (kpts @ tau.T))
    return np.repeat(d, model._nspin, axis=1)


def sol_ham(ham, eig_vectors=False):
    """Batched ``tb_model._sol_ham`` (pythtb.py:927-953) + ``_nicefy_eig``
    (pythtb.py:3765-3775).  ``ham`` is ``[nk, n, n]``.

    Returns ``eval[nk, n]`` and, if asked, ``evec[nk, band, n]`` (rows are
    eigenvectors, not conjugated; pythtb.py:947).
    """
    if np.max(ham - ham.conj().transpose(0, 2, 1)) > 1.0e-9:    # :935
        raise Exception("\n\nHamiltonian matrix is not hermitian?!")
    if not eig_vectors:
        ev = np.linalg.eigvalsh(ham)                            # :939
        return np.sort(np.array(ev.real, dtype=float), axis=-1)
    ev, vec = np.linalg.eigh(ham)                               # :944
    vec = vec.transpose(0, 2, 1)                                # :947
    ev = np.array(ev.real, dtype=float)
    order = np.argsort(ev, axis=-1)                             # :3770
    ev = np.take_along_axis(ev, order, axis=-1)
    vec = np.take_along_axis(vec, order[:, :, None], axis=1)
    return ev, vec


def solve_all(model, k_list=None, eig_vectors=False):
    """``tb_model.solve_all`` (pythtb.py:955-1079): ``eval[band,k]``,
    ``evec[band,k,orb(,spin)]``; the k axis is dropped for dim_k == 0."""
    ham = gen_ham(model, k_list)
    nk = ham.shape[0]
    if not eig_vectors:
        ev = sol_ham(ham, False)
        return ev[0] if k_list is None else ev.T.copy()
    ev, vec = sol_ham(ham, True)
    if model._nspin == 2:
        vec = vec.reshape(nk, model._nsta, model._norb, 2)      # :951-952
    if k_list is None:
        return ev[0], vec[0]
    return ev.T.copy(), np.ascontiguousarray(np.swapaxes(vec, 0, 1))


def solve_one(model, k_point=None, eig_vectors=False):
    """``tb_model.solve_one`` (pythtb.py:1081-1103)."""
    if k_point is None:
        return solve_all(model, None, eig_vectors)
    if not eig_vectors:
        return solve_all(model, [k_point], False)[:, 0]
    ev, vec = solve_all(model, [k_point], True)
    return ev[:, 0], vec[:, 0]


# ----------------------------------------------------------------------------
# wf_array: grid solve and boundary conditions
# ----------------------------------------------------------------------------
def grid_kpoints(start_k, mesh_arr):
    """k = start_k + i/(N-1) on the (N-1)^d solved points, C order
    (pythtb.py:2477, 2490-2491, 2502-2504, 2516-2519)."""
    mesh_arr = np.asarray(mesh_arr, dtype=int)
    axes = [np.array([float(start_k[d]) + float(i) / float(mesh_arr[d] - 1)
                      for i in range(mesh_arr[d] - 1)]) for d in range(len(mesh_arr))]
    grids = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=-1)


def solve_on_grid(model, mesh_arr, start_k):
    """``wf_array.solve_on_grid`` (pythtb.py:2421-2532).

    Returns ``(wfs, gaps)`` with ``wfs[k1..kd, state, orb(,spin)]`` as stored
    in ``wf_array._wfs`` (pythtb.py:2419) after ``impose_pbc`` on every axis,
    and the minimal direct gaps (``None`` if nsta <= 1).
    """
    mesh_arr = np.asarray(mesh_arr, dtype=int)
    dim = len(mesh_arr)
    if dim != model._dim_k:
        raise Exception("\n\nIf using solve_on_grid method, dimension of wf_array must equal"
                        "\ndim_k of the tight-binding model!")
    kpts = grid_kpoints(start_k, mesh_arr)
    ev, vec = sol_ham(gen_ham(model, kpts), True)
    nsta, norb, nspin = model._nsta, model._norb, model._nspin
    tail = (nsta, norb) if nspin == 1 else (nsta, norb, 2)
    wfs = np.zeros(tuple(mesh_arr) + tail, dtype=complex)
    inner = tuple(slice(0, m - 1) for m in mesh_arr)
    wfs[inner] = vec.reshape(tuple(mesh_arr - 1) + tail)
    gaps = None
    if nsta > 1:
        gaps = (ev[:, 1:] - ev[:, :-1]).min(axis=0)             # :2484, :2529-2530
    for d in range(dim):                                        # :2486, :2496-2497
        if getattr(model, "_convention", 1) == 2:
            impose_loop(wfs, d)                                 # C~(k+G) = C~(k), tex:341-364
        else:
            impose_pbc(wfs, model._orb, model._nspin, d, model._per[d])
    return wfs, gaps


def impose_pbc(wfs, orb, nspin, mesh_dir, k_dir):
    """In-place ``wf_array.impose_pbc`` (pythtb.py:2674-2749): last slice =
    first slice * exp(-2 pi i tau_j[k_dir]), same factor on both spins."""
    ffac = np.exp(-2.0j * np.pi * np.asarray(orb)[:, k_dir])    # :2729
    phase = ffac if nspin == 1 else np.stack([ffac, ffac], axis=-1)
    idx_last = [slice(None)] * mesh_dir + [-1]
    idx_first = [slice(None)] * mesh_dir + [0]
    wfs[tuple(idx_last)] = wfs[tuple(idx_first)] * phase        # :2740-2747
    return wfs


def impose_loop(wfs, mesh_dir):
    """In-place ``wf_array.impose_loop`` (pythtb.py:2751-2791)."""
    idx_last = [slice(None)] * mesh_dir + [-1]
    idx_first = [slice(None)] * mesh_dir + [0]
    wfs[tuple(idx_last)] = wfs[tuple(idx_first)]
    return wfs


# ----------------------------------------------------------------------------
# Berry phases and fluxes
# ----------------------------------------------------------------------------
def _overlaps(wf_a, wf_b):
    """M[...,m,n] = <a_m|b_n>, conjugating the first argument
    (``_wf_dpr`` pythtb.py:3793-3796 inside the loop :3813-3817).
    ``wf_*`` are ``[..., nocc, flat]``."""
    return np.einsum("...mo,...no->...mn", wf_a.conj(), wf_b)


def one_berry_loop(wf, berry_evals=False):
    """``_one_berry_loop`` (pythtb.py:3798-3838).  ``wf[kpnt, band, orb(,spin)]``."""
    wf = np.asarray(wf, dtype=complex)
    nk, nocc = wf.shape[0], wf.shape[1]
    flat = wf.reshape(nk, nocc, -1)
    ovr = _overlaps(flat[:-1], flat[1:])                        # all links at once
    prd = np.identity(nocc, dtype=complex)
    if not berry_evals:
        for i in range(nk - 1):
            prd = np.dot(prd, ovr[i])                           # :3821
        return (-1.0) * np.angle(np.linalg.det(prd))            # :3829-3831
    mat_u, _, mat_v = np.linalg.svd(ovr)                        # :3825
    uni = mat_u @ mat_v
    for i in range(nk - 1):
        prd = np.dot(prd, uni[i])                               # :3826
    evals = np.linalg.eigvals(prd)                              # :3834
    return np.sort((-1.0) * np.angle(evals))                    # :3835-3838


def one_flux_plane(wfs2d):
    """``_one_flux_plane`` (pythtb.py:3840-3865): for each plaquette the loop
    (i,j)->(i+1,j)->(i+1,j+1)->(i,j+1)->(i,j), phase = -arg det(M1 M2 M3 M4)."""
    w = np.asarray(wfs2d, dtype=complex)
    nk0, nk1, nocc = w.shape[0], w.shape[1], w.shape[2]
    f = w.reshape(nk0, nk1, nocc, -1)
    a, b, c, d = f[:-1, :-1], f[1:, :-1], f[1:, 1:], f[:-1, 1:]
    prd = _overlaps(a, b) @ _overlaps(b, c) @ _overlaps(c, d) @ _overlaps(d, a)
    return (-1.0) * np.angle(np.linalg.det(prd))


def no_2pi(x, clos):
    """pythtb.py:3867-3874."""
    while abs(clos - x) > np.pi:
        if clos - x > np.pi:
            x += 2.0 * np.pi
        elif clos - x < -1.0 * np.pi:
            x -= 2.0 * np.pi
    return x


def one_phase_cont(pha, clos):
    """pythtb.py:3876-3888."""
    ret = np.copy(pha)
    for i in range(len(ret)):
        cmpr = clos if i == 0 else ret[i - 1]
        ret[i] = no_2pi(ret[i], cmpr)
    return ret


def array_phases_cont(arr_pha, clos):
    """pythtb.py:3890-3921 (greedy nearest-phase matching, ``<=`` tie-break)."""
    ret = np.zeros_like(arr_pha)
    for i in range(arr_pha.shape[0]):
        cmpr = clos if i == 0 else ret[i - 1, :]
        avail = list(range(arr_pha.shape[1]))
        for j in range(cmpr.shape[0]):
            min_dist, best_k = 1.0e10, None
            for k in avail:
                cur = np.abs(np.exp(1.0j * cmpr[j]) - np.exp(1.0j * arr_pha[i, k]))
                if cur <= min_dist:
                    min_dist, best_k = cur, k
            avail.pop(avail.index(best_k))
            ret[i, j] = no_2pi(arr_pha[i, best_k], cmpr[j])
    return ret


def _occ_array(occ, nsta_arr):
    if (isinstance(occ, str) and occ == "All") or occ is None:  # :2959-2963
        return np.arange(nsta_arr, dtype=int)
    occ = np.array(occ, dtype=int)
    if occ.ndim != 1:
        raise Exception("\n\nParameter occ must be a one-dimensional array or string \"All\" or None.")
    return occ


def berry_phase(wfs, dim_arr, occ="All", dir=None, contin=True, berry_evals=False):
    """``wf_array.berry_phase`` (pythtb.py:2863-3066) on a raw ``_wfs`` array
    with ``dim_arr`` mesh axes."""
    wfs = np.asarray(wfs)
    occ = _occ_array(occ, wfs.shape[dim_arr])
    if dim_arr == 1:
        ret = one_berry_loop(wfs[:, occ], berry_evals)          # :2979-2983
    elif dim_arr in (2, 3):
        if dir is None or dir < 0 or dir >= dim_arr:
            raise Exception("\n\nWrong direction for Berry phase calculation!")
        moved = np.moveaxis(wfs, dir, dim_arr - 1)              # strings last of mesh axes
        other = moved.shape[:dim_arr - 1]
        ret = np.empty(other, dtype=object)
        for idx in np.ndindex(*other):                          # :2985-3027
            ret[idx] = one_berry_loop(moved[idx][:, occ], berry_evals)
        ret = np.array(ret.tolist(), dtype=float)
    else:
        raise Exception("\n\nWrong dimensionality!")
    if dim_arr > 1 or berry_evals:
        ret = np.array(ret, dtype=float)                        # :3032-3033
    if contin:                                                  # :3036-3065
        if not berry_evals:
            if dim_arr == 2:
                ret = one_phase_cont(ret, ret[0])
            elif dim_arr == 3:
                for i in range(ret.shape[1]):
                    clos = ret[0, 0] if i == 0 else ret[0, i - 1]
                    ret[:, i] = one_phase_cont(ret[:, i], clos)
        else:
            if dim_arr == 2:
                ret = array_phases_cont(ret, ret[0, :])
            elif dim_arr == 3:
                for i in range(ret.shape[1]):
                    clos = ret[0, 0, :] if i == 0 else ret[0, i - 1, :]
                    ret[:, i] = array_phases_cont(ret[:, i], clos)
    return ret


def berry_flux(wfs, dim_arr, occ="All", dirs=None, individual_phases=False):
    """``wf_array.berry_flux`` (pythtb.py:3068-3205)."""
    wfs = np.asarray(wfs)
    occ = _occ_array(occ, wfs.shape[dim_arr])
    if dirs is None:
        dirs = [0, 1]
    if dirs[0] == dirs[1]:
        raise Exception("Need to specify two different directions for Berry flux calculation.")
    if min(dirs) < 0 or max(dirs) >= dim_arr:
        raise Exception("Direction for Berry flux calculation out of bounds.")
    order = list(range(wfs.ndim))
    order[0], order[1] = dirs[0], dirs[1]                       # :3135-3138
    if dim_arr == 2:
        plane = wfs.transpose(order)[:, :, occ]
        allp = one_flux_plane(plane)
        return allp if individual_phases else allp.sum()        # :3147-3150
    if dim_arr not in (3, 4):
        raise Exception("\n\nWrong dimensionality!")
    rest = [d for d in range(dim_arr) if d not in dirs]         # :3161-3163
    for n, d in enumerate(rest):
        order[2 + n] = d                                        # :3168-3172
    use = wfs.transpose(order)
    mesh = wfs.shape[:dim_arr]
    shp = tuple(mesh[d] for d in rest) + (mesh[dirs[0]] - 1, mesh[dirs[1]] - 1)
    out = np.zeros(shp, dtype=float)
    for idx in np.ndindex(*shp[:len(rest)]):                    # :3178-3196
        sel = (slice(None), slice(None)) + idx
        out[idx] = one_flux_plane(use[sel][:, :, occ])
    return out if individual_phases else out.sum(axis=(-2, -1))  # :3199-3202


# ----------------------------------------------------------------------------
# Position operator / hybrid Wannier functions
# ----------------------------------------------------------------------------
def position_matrix(model, evec, dir):
    """``tb_model.position_matrix`` (pythtb.py:2034-2113)."""
    if dir in model._per:
        raise Exception("Can not compute position matrix elements along periodic direction!")
    if dir < 0 or dir >= model._dim_r:
        raise Exception("Direction out of range!")
    evec = np.asarray(evec)
    pos = model._orb[:, dir]
    if model._nspin == 2:
        pos = np.tile(pos, (2, 1)).transpose().flatten()        # :2096
    flat = evec.reshape(evec.shape[0], -1)
    mat = np.einsum("io,o,jo->ij", flat.conj(), pos, flat)      # :2104-2107
    if np.max(mat - mat.T.conj()) > 1.0e-9:
        raise Exception("\n\n Position matrix is not hermitian?!")
    return mat


def position_expectation(model, evec, dir):
    """``tb_model.position_expectation`` (pythtb.py:2115-2160)."""
    return np.array(np.real(position_matrix(model, evec, dir).diagonal()), dtype=float)


def position_hwf(model, evec, dir, hwf_evec=False, basis="orbital"):
    """``tb_model.position_hwf`` (pythtb.py:2162-2279)."""
    mat = position_matrix(model, evec, dir)
    if not hwf_evec:
        return sol_ham(mat[None], False)[0]                     # :2246-2249
    ev, vec = sol_ham(mat[None], True)                          # :2251-2257
    hwfc, hwf = ev[0], vec[0]
    b = basis.lower().strip()
    if b in ("wavefunction", "bloch"):
        return hwfc, hwf
    if b != "orbital":
        raise Exception("\n\nBasis must be either 'wavefunction', 'bloch', or 'orbital'")
    evec = np.asarray(evec)
    flat = evec.reshape(evec.shape[0], -1)
    out = hwf @ flat                                            # :2262-2274
    if model._nspin == 2:
        out = out.reshape(hwf.shape[0], model._norb, 2)
    return hwfc, out
