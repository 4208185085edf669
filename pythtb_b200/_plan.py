"""Model compiler: flattens a ``tb_model`` into the SoA buffers the kernels
stream ("plan").  Layout is documented in ``csrc/tbk_plan.cuh`` and DESIGN.md.

The plan restates the accumulation of ``tb_model._gen_ham``
(/root/reference/pythtb.py:894-924): the on-site block first, then every
hopping in list order contributing ``amp*phase`` to (i,j) and its conjugate to
(j,i).  Only lower-triangle elements are kept (numpy's ``eigh`` reads
UPLO='L'), and the Bloch phase is factored as  exp(2 pi i k.R) * gauge(tau)
so that phases are shared by all hoppings with the same lattice vector R.
"""
import numpy as np

PH_CONJ = 1 << 30


class Plan(object):
    """Host-side compiled model (numpy arrays, C-contiguous)."""

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in
                   ("ph_R", "tau", "el_ptr", "el_row", "el_col", "t_ph", "t_amp", "pm_ptr", "pm_el", "pm_amp"))


def _canon(rper):
    """Canonical representative of +-R: first non-zero component positive.
    Returns (key, conj_flag) or (None, False) for R = 0."""
    for x in rper:
        if x > 0:
            return rper, False
        if x < 0:
            return tuple(-y for y in rper), True
    return None, False


# models with at least this many hoppings take the array-native term builder (large Wannier models)
VECTORISE_FROM = 4096


def _terms_vectorised(model):
    """The term list of ``compile_plan`` for a periodic model, built with array operations: the same terms in
    the same order as the loop (on-site blocks first, then per hopping and per spin pair (s, s') its
    (i s, j s') term followed by the conjugate (j s', i s) term; upper-triangle and zero entries dropped;
    lattice vectors numbered in first-seen order, canonical sign = first non-zero component positive)."""
    norb, ns = model._norb, model._nspin
    per = list(model._per)
    hops = model._hoppings
    nhop = len(hops)
    amp = np.array([np.asarray(h[0], dtype=complex).reshape(ns, ns) for h in hops], dtype=complex).reshape(nhop, ns, ns)
    hi = np.fromiter((h[1] for h in hops), dtype=np.int64, count=nhop)
    hj = np.fromiter((h[2] for h in hops), dtype=np.int64, count=nhop)
    rper = np.array([h[3] for h in hops], dtype=np.int64).reshape(nhop, -1)[:, per]
    nz = rper != 0
    has = nz.any(axis=1)
    first = np.argmax(nz, axis=1)
    sgn = np.sign(rper[np.arange(nhop), first])              # 0 for R = 0
    canon = rper * np.where(sgn < 0, -1, 1)[:, None]
    idx = np.full(nhop, -1, dtype=np.int64)
    table = {}
    if has.any():
        uniq, first_seen, inv = np.unique(canon[has], axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first_seen, kind="stable")
        rank = np.empty(len(order), dtype=np.int64)
        rank[order] = np.arange(len(order))
        idx[has] = rank[np.asarray(inv).reshape(-1)]
        for pos, u in enumerate(order):
            table[tuple(int(x) for x in uniq[u])] = pos
    ph_f = np.where(has, idx | np.where(sgn < 0, PH_CONJ, 0), -1)
    ph_c = np.where(has, idx | np.where(sgn > 0, PH_CONJ, 0), -1)
    # on-site blocks, pythtb.py:894-898 (LAPACK ignores the imaginary part of the diagonal)
    if ns == 1:
        site = np.array([complex(x).real for x in model._site_energies], dtype=complex).reshape(norb, 1, 1)
    else:
        site = np.array(model._site_energies, dtype=complex).reshape(norb, 2, 2).copy()
        site[:, 0, 0] = site[:, 0, 0].real
        site[:, 1, 1] = site[:, 1, 1].real
    sgrid, spgrid = np.meshgrid(np.arange(ns), np.arange(ns), indexing="ij")
    o_rows = (np.arange(norb)[:, None, None] * ns + sgrid[None]).reshape(-1)
    o_cols = (np.arange(norb)[:, None, None] * ns + spgrid[None]).reshape(-1)
    o_amps = site.reshape(-1)
    # hoppings: [hop, s, s', forward/conjugate] in C order = the loop's order
    f_rows = hi[:, None, None] * ns + sgrid[None]
    f_cols = hj[:, None, None] * ns + spgrid[None]
    rows = np.stack([f_rows, f_cols], axis=-1).reshape(-1)
    cols = np.stack([f_cols, f_rows], axis=-1).reshape(-1)
    amps = np.stack([amp, amp.conj()], axis=-1).reshape(-1)
    phs = np.stack([np.broadcast_to(ph_f[:, None, None], amp.shape), np.broadcast_to(ph_c[:, None, None], amp.shape)],
                   axis=-1).reshape(-1)
    rows = np.concatenate([o_rows, rows])
    cols = np.concatenate([o_cols, cols])
    amps = np.concatenate([o_amps, amps])
    phs = np.concatenate([np.full(len(o_rows), -1, dtype=np.int64), phs])
    keep = (rows >= cols) & (amps != 0.0)
    return rows[keep], cols[keep], amps[keep], phs[keep], table


def compile_plan(model, convention=1):
    """Build the plan from any object exposing the reference's attribute names
    (``_dim_k,_nspin,_norb,_nsta,_per,_orb,_site_energies,_hoppings``)."""
    if convention not in (1, 2):
        raise Exception("\n\nconvention must be 1 (PythTB, orbital positions in the phase) or 2")
    dim_k, nspin, norb, nsta = model._dim_k, model._nspin, model._norb, model._nsta
    per = list(model._per)
    rows, cols, amps, phs = [], [], [], []
    table = {}

    def phase_id(rper):
        key, cj = _canon(rper)
        if key is None:
            return -1
        idx = table.get(key)
        if idx is None:
            idx = len(table)
            table[key] = idx
        return idx | (PH_CONJ if cj else 0)

    def add(row, col, val, ph):
        if row < col:
            return
        val = complex(val)
        if val == 0.0:
            return
        rows.append(row)
        cols.append(col)
        amps.append(val)
        phs.append(ph)

    vectorised = dim_k > 0 and len(model._hoppings) >= VECTORISE_FROM
    if vectorised:
        rows, cols, amps, phs, table = _terms_vectorised(model)
    # on-site block, pythtb.py:894-898
    for i in range(norb if not vectorised else 0):
        if nspin == 1:
            add(i, i, complex(model._site_energies[i]).real, -1)
        else:
            blk = np.array(model._site_energies[i], dtype=complex).reshape(2, 2)
            for s in range(2):
                for sp in range(2):
                    val = blk[s, sp]
                    if s == sp:
                        val = val.real      # LAPACK ignores the imaginary part of the diagonal
                    add(2 * i + s, 2 * i + sp, val, -1)
    # hoppings in list order, pythtb.py:900-924
    for hop in (model._hoppings if not vectorised else ()):
        blk = np.array(hop[0], dtype=complex).reshape(nspin, nspin)
        i, j = int(hop[1]), int(hop[2])
        if dim_k > 0:
            rvec = np.array(hop[3])
            rper = tuple(int(rvec[p]) for p in per)
        else:
            rper = ()
        ph_f = phase_id(rper)
        ph_c = phase_id(tuple(-x for x in rper))
        for s in range(nspin):
            for sp in range(nspin):
                add(i * nspin + s, j * nspin + sp, blk[s, sp], ph_f)
                add(j * nspin + sp, i * nspin + s, np.conj(blk[s, sp]), ph_c)

    table_R = np.zeros((len(table), max(dim_k, 1)), dtype=np.float64)
    for key, idx in table.items():
        table_R[idx, :dim_k] = key
    tau = np.zeros((max(nsta, 1), max(dim_k, 1)), dtype=np.float64)
    if dim_k > 0:
        tau[:nsta, :dim_k] = np.repeat(np.asarray(model._orb, dtype=float)[:, per], nspin, axis=0)
    return _finish_plan(rows, cols, amps, phs, table_R, tau, dim_k, nsta, norb, nspin, convention)


def _finish_plan(rows, cols, amps, phs, table_R, tau, dim_k, nsta, norb, nspin, convention):
    """Term lists (reference accumulation order) -> Plan with its two CSR orderings."""
    nterm = len(rows)
    rows = np.array(rows, dtype=np.int64)
    cols = np.array(cols, dtype=np.int64)
    amps = np.array(amps, dtype=complex)
    phs = np.array(phs, dtype=np.int64)
    p = Plan()
    p.dim_k, p.nsta, p.norb, p.nspin, p.convention = dim_k, nsta, norb, nspin, convention
    p.nph = int(table_R.shape[0])
    p.ph_R = np.zeros((max(p.nph, 1), max(dim_k, 1)), dtype=np.float64)
    p.ph_R[:p.nph, :dim_k] = table_R[:, :dim_k]
    p.tau = np.ascontiguousarray(tau, dtype=np.float64)
    # ---- element-major CSR (stable: keeps the reference's accumulation order)
    key = rows * nsta + cols
    order = np.argsort(key, kind="stable")
    skey = key[order]
    uniq, first = np.unique(skey, return_index=True)
    p.nel = len(uniq)
    p.nterm = nterm
    p.el_ptr = np.append(first, nterm).astype(np.int32)
    p.el_row = (uniq // nsta).astype(np.int32)
    p.el_col = (uniq % nsta).astype(np.int32)
    p.t_ph = phs[order].astype(np.int32)
    p.t_amp = np.ascontiguousarray(amps[order].view(np.float64).reshape(-1, 2))
    # ---- phase-major CSR
    el_of_term = np.empty(nterm, dtype=np.int64)
    el_of_term[order] = np.searchsorted(uniq, skey)
    bucket = np.where(phs < 0, p.nph, phs & (PH_CONJ - 1))
    order2 = np.argsort(bucket, kind="stable")
    counts = np.bincount(bucket, minlength=p.nph + 1)
    p.pm_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    cj = np.where(phs >= 0, phs & PH_CONJ, 0)
    p.pm_el = (el_of_term | cj)[order2].astype(np.int32)
    p.pm_amp = np.ascontiguousarray(amps[order2].view(np.float64).reshape(-1, 2))
    if nterm == 0:
        p.t_amp = np.zeros((1, 2))
        p.pm_amp = np.zeros((1, 2))
        p.t_ph = np.zeros(1, dtype=np.int32)
        p.pm_el = np.zeros(1, dtype=np.int32)
    if p.nel == 0:
        p.el_row = np.zeros(1, dtype=np.int32)
        p.el_col = np.zeros(1, dtype=np.int32)
    return p


def reduce_plan(p, axis, value):
    """Plan of ``model.reduce_dim(per[axis], value)`` (pythtb.py:1233-1311) obtained from the PARENT's plan with
    array operations — no Python model is rebuilt, no hopping list is walked:

        H_red(k') = H(k', k_axis = value).

    The factor exp(+-2 pi i value R_axis) of every term moves from the phase into its amplitude, the phase table
    loses a column (vectors that coincide afterwards are merged, vectors that vanish turn their terms into
    constants), and the constant diagonal gauge Phi = diag(exp(2 pi i value tau_o[axis])) is folded in as
    Phi^H H Phi, so that the kernels' eigenvectors carry exactly the phases of the reduced reference model
    (whose hopping amplitudes contain exp(2 pi i value (tau_j - tau_i + R)[axis]))."""
    if p.dim_k < 1 or not (0 <= axis < p.dim_k):
        raise Exception("\n\nSpecified wrong dimension to reduce!")
    nterm, nel, nph, dk = p.nterm, p.nel, p.nph, p.dim_k
    counts = np.diff(p.el_ptr[:nel + 1].astype(np.int64)) if nel else np.zeros(0, dtype=np.int64)
    rows = np.repeat(p.el_row[:nel].astype(np.int64), counts)
    cols = np.repeat(p.el_col[:nel].astype(np.int64), counts)
    phs = p.t_ph[:nterm].astype(np.int64)
    amps = np.ascontiguousarray(p.t_amp[:nterm]).view(complex).reshape(-1).copy()
    has = phs >= 0
    idx = np.where(has, phs & (PH_CONJ - 1), 0)
    cj = has & ((phs & PH_CONJ) != 0)
    R = p.ph_R[:nph, :dk] if nph else np.zeros((0, dk))
    r_axis = np.where(has, R[idx, axis] if nph else 0.0, 0.0)
    r_axis = np.where(cj, -r_axis, r_axis)
    fac = np.exp(2.0j * np.pi * value * r_axis)
    # (either convention: the reference algorithm puts the orbital positions along the removed direction into the
    #  amplitudes, so the reduced model's Convention-II matrix is Phi^H H_II(k', value) Phi as well)
    phi = np.exp(2.0j * np.pi * value * p.tau[:p.nsta, axis])
    fac = fac * np.conj(phi[rows]) * phi[cols]
    amps = amps * fac
    # ---- the phase table without the fixed component: canonical sign, merged duplicates, first-seen order
    Rn = np.rint(np.delete(R, axis, axis=1)).astype(np.int64)              # [nph, dk-1]
    new_id = np.full(nph, -1, dtype=np.int64)
    flip = np.zeros(nph, dtype=bool)
    table_R = np.zeros((0, max(dk - 1, 1)))
    if dk - 1 > 0 and nph > 0:
        nz = Rn != 0
        hasnz = nz.any(axis=1)
        first = np.argmax(nz, axis=1)
        sgn = np.sign(Rn[np.arange(nph), first])
        flip = sgn < 0
        canon = Rn * np.where(flip, -1, 1)[:, None]
        if hasnz.any():
            uniq, first_seen, inv = np.unique(canon[hasnz], axis=0, return_index=True, return_inverse=True)
            order = np.argsort(first_seen, kind="stable")                  # ids in first-seen order
            rank = np.empty(len(order), dtype=np.int64)
            rank[order] = np.arange(len(order))
            new_id[hasnz] = rank[np.asarray(inv).reshape(-1)]
            table_R = uniq[order].astype(np.float64)
    nid = np.where(has, new_id[idx] if nph else -1, -1)
    ncj = cj ^ (flip[idx] if nph else False)
    phs2 = np.where(nid >= 0, nid | np.where(ncj, PH_CONJ, 0), -1)
    tau = np.delete(p.tau, axis, axis=1) if dk - 1 > 0 else np.zeros((max(p.nsta, 1), 1))
    return _finish_plan(rows, cols, amps, phs2, table_R, tau, dk - 1, p.nsta, p.norb, p.nspin, p.convention)
