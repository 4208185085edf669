"""TEST INFRASTRUCTURE: the PythTB public API served by the numpy oracle.

``tests/cases.py`` drives ``mod.tb_model`` / ``mod.wf_array``.  This module is
such a ``mod`` whose numerical seams are answered by ``oracle/pythtb_oracle.py``
instead of the CUDA engine, so the oracle can be checked against the golden
fixtures on a CPU-only box through exactly the host logic (model building,
plan-independent bookkeeping, continuity post-processing) that the product
uses.  The product package never imports this file or the oracle.
"""
import numpy as np

from oracle import pythtb_oracle as orc
from pythtb_b200 import model as _model
from pythtb_b200 import wfarray as _wfarray


class _HostStore(object):
    def __init__(self, shape):
        self.shape = tuple(int(x) for x in shape)
        self.arr = np.zeros(self.shape, dtype=complex)
        self.state = "host"

    def host(self):
        return self.arr

    def replace_host(self, arr):
        self.arr = np.array(arr, dtype=complex)
        self.shape = self.arr.shape


class OracleEngine(object):
    """Same method set as pythtb_b200._engine.B200Engine, numpy arithmetic."""

    def new_store(self, shape, state_axis=None):
        return _HostStore(shape)                  # (the device layout option of the product engine does not apply)

    def gen_ham(self, model, klist):
        return orc.gen_ham(model, klist if model._dim_k > 0 else None)

    def eigh(self, ham, eig_vectors):
        if not eig_vectors:
            return orc.sol_ham(ham, False), None
        return orc.sol_ham(ham, True)

    def solve_all(self, model, klist, eig_vectors):
        if model._dim_k == 0:
            res = orc.solve_all(model, None, eig_vectors)
            if not eig_vectors:
                return res[:, None]
            return res[0][:, None], res[1][:, None]
        return orc.solve_all(model, klist, eig_vectors)

    def solve_slice(self, model, store, dim_arr, fixed, free, kpts):
        """Same contract as B200Engine.solve_slice."""
        fshape = tuple(store.arr.shape[d] for d in free)
        ev, vec = orc.solve_all(model, kpts.reshape(-1, model._dim_k), True)
        key = tuple(fixed[d] if d in fixed else slice(None) for d in range(dim_arr))
        tail = store.arr.shape[dim_arr:]
        store.arr[key] = np.swapaxes(vec, 0, 1).reshape(fshape + tuple(tail))
        return ev.T.reshape(fshape + (model._nsta,))

    def pbc_phases(self, orb, nspin, k_dirs, convention=1):
        if convention == 2:
            return np.ones((len(k_dirs), np.asarray(orb).shape[0] * nspin), dtype=complex)
        return np.array([np.repeat(np.exp(-2.0j * np.pi * np.asarray(orb)[:, kd]), nspin) for kd in k_dirs])

    def solve_grid(self, model, store, mesh_arr, start_k, row0=0, nrows=None, wrap0=1, want_gaps=True, host_result=False,
                   reduce_ranks=None, defer_reduce=False):
        """Same contract as B200Engine.solve_grid: fills the local rows
        [row0, row0+nrows] of a shard (the closing row only for wrap0 in (1, 2))."""
        wfs, _ = orc.solve_on_grid(model, mesh_arr, start_k)
        n0 = int(mesh_arr[0])
        if nrows is None:
            nrows = n0 - 1
        store.arr[:nrows] = wfs[row0:row0 + nrows]
        if wrap0 in (1, 2, True):
            store.arr[nrows] = wfs[row0 + nrows]
        if model._nsta <= 1 or not want_gaps:
            return None
        # minimal gaps over the rows solved here only (so that the cross-rank min is exercised)
        kpts = orc.grid_kpoints(start_k, mesh_arr).reshape(tuple(np.asarray(mesh_arr) - 1) + (len(mesh_arr),))
        ev = orc.sol_ham(orc.gen_ham(model, kpts[row0:row0 + nrows].reshape(-1, len(mesh_arr))), False)
        gaps = (ev[:, 1:] - ev[:, :-1]).min(axis=0)
        return self.allreduce(gaps, "min") if reduce_ranks is not None else gaps

    # ---- multi-rank plumbing over torch.distributed (gloo on CPU in the tests)
    def halo_ring_shift(self, store, dim_arr, phase, rank, nranks):
        import torch
        import torch.distributed as dist
        row = np.array(store.arr[0], copy=True)
        if phase is not None:
            row = row * np.asarray(phase).reshape(store.arr.shape[dim_arr + 1:])
        send = torch.from_numpy(np.ascontiguousarray(row).view(np.float64))
        recv = torch.empty_like(send)
        reqs = [dist.isend(send, (rank - 1) % nranks), dist.irecv(recv, (rank + 1) % nranks)]
        for r in reqs:
            r.wait()
        store.arr[-1] = recv.numpy().view(np.complex128).reshape(store.arr.shape[1:])

    def allreduce(self, x, op):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.array(x, dtype=np.float64, copy=True).reshape(-1))
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN)
        return t.numpy().reshape(np.shape(x))

    def allgather_rows(self, local, nranks, n0):
        import torch.distributed as dist
        parts = [None] * nranks
        dist.all_gather_object(parts, np.asarray(local))
        return parts

    def flux_total(self, store, dim_arr, occ, dirs, host_result=False, reduce_ranks=None, defer_reduce=False):
        tot = np.asarray(orc.berry_flux(store.arr, dim_arr, occ, dirs, individual_phases=False)).reshape(-1)
        return self.allreduce(tot, "sum") if reduce_ranks is not None else tot

    def impose_boundary(self, store, dim_arr, mesh_dir, phase):
        if phase is None:
            orc.impose_loop(store.arr, mesh_dir)
        else:
            tail = store.arr.shape[dim_arr + 1:]
            idx_last = [slice(None)] * mesh_dir + [-1]
            idx_first = [slice(None)] * mesh_dir + [0]
            store.arr[tuple(idx_last)] = store.arr[tuple(idx_first)] * phase.reshape(tail)

    def berry_strings(self, store, dim_arr, occ, dir, berry_evals):
        return np.asarray(orc.berry_phase(store.arr, dim_arr, occ, dir, contin=False, berry_evals=berry_evals))

    def wilson_phases_across_ranks(self, store, dim_arr, occ, dir, nranks):
        """numpy restatement of the split Wilson loop: local ordered product of the SVD polar factors of the
        local links (pythtb.py:3813-3826), gathered in rank order, multiplied, eigenphases sorted (3834-3838)."""
        import torch.distributed as dist
        wfs = np.moveaxis(store.arr, dir, 0)                      # [npts, other..., state, orb...]
        npts = wfs.shape[0]
        other = wfs.shape[1:dim_arr]
        nsta = wfs.shape[dim_arr]
        flat = wfs.reshape((npts, int(np.prod(other)) if other else 1, nsta, -1))[:, :, list(occ)]
        nstr, nocc = flat.shape[1], len(occ)
        prods = np.zeros((nstr, nocc, nocc), dtype=complex)
        for s in range(nstr):
            prd = np.identity(nocc, dtype=complex)
            for t in range(npts - 1):
                ovr = flat[t, s].conj() @ flat[t + 1, s].T
                u, _, vh = np.linalg.svd(ovr)
                prd = prd @ (u @ vh)
            prods[s] = prd
        parts = [None] * nranks
        dist.all_gather_object(parts, prods)
        out = np.zeros((nstr, nocc))
        for s in range(nstr):
            prd = np.identity(nocc, dtype=complex)
            for r in range(nranks):
                prd = prd @ parts[r][s]
            out[s] = np.sort(-np.angle(np.linalg.eigvals(prd)))
        return out.reshape(tuple(other) + (nocc,))

    # ---- streamed 1-D strings (same contract as B200Engine.stream_links / stream_gaps / wilson_finish)
    def _string_points(self, model, npts, start_k, l0, l1):
        """Eigenvector blocks of the points l0 .. l1 of the closed string (the last mesh point is the image of the first)."""
        wfs, _ = orc.solve_on_grid(model, [npts], start_k)
        return wfs[l0:l1 + 1].reshape(l1 - l0 + 1, model._nsta, -1)

    def stream_links(self, model, npts, start_k, occ, l0, l1, berry_evals, want_gaps=False, chunk=None):
        pts = self._string_points(model, npts, start_k, l0, l1)[:, list(occ)]
        gaps = self.stream_gaps(model, npts, start_k, l0, l1 + 1 if l1 < npts - 1 else l1) if want_gaps else None
        nocc = len(occ)
        if not berry_evals:
            tot = 0.0
            for t in range(pts.shape[0] - 1):
                tot += -np.angle(np.linalg.det(pts[t].conj() @ pts[t + 1].T))
            return float(tot), gaps
        prd = np.identity(nocc, dtype=complex)
        for t in range(pts.shape[0] - 1):
            u, _, vh = np.linalg.svd(pts[t].conj() @ pts[t + 1].T)
            prd = prd @ (u @ vh)
        return prd.reshape(1, nocc, nocc), gaps

    def stream_gaps(self, model, npts, start_k, l0, l1):
        if model._nsta <= 1:
            return None
        k = float(np.asarray(start_k, dtype=float).reshape(-1)[0]) + np.arange(l0, l1, dtype=float) / float(npts - 1)
        ev = orc.sol_ham(orc.gen_ham(model, k.reshape(-1, 1)), False)
        return (ev[:, 1:] - ev[:, :-1]).min(axis=0)

    def wilson_finish(self, prod, nranks):
        parts = [np.asarray(prod)]
        if nranks > 1:
            import torch.distributed as dist
            parts = [None] * nranks
            dist.all_gather_object(parts, np.asarray(prod))
        nstr, nocc = parts[0].shape[0], parts[0].shape[1]
        out = np.zeros((nstr, nocc))
        for s in range(nstr):
            prd = np.identity(nocc, dtype=complex)
            for part in parts:
                prd = prd @ part[s]
            out[s] = np.sort(-np.angle(np.linalg.eigvals(prd)))
        return out

    def flux(self, store, dim_arr, occ, dirs, individual):
        return np.asarray(orc.berry_flux(store.arr, dim_arr, occ, dirs, individual_phases=individual))

    def position_hwf_store(self, model, store, dim_arr, occ, dir, hwf_evec, out_store=None):
        mesh = store.arr.shape[:dim_arr]
        nocc = len(occ)
        hwfc = np.zeros(tuple(mesh) + (nocc,))
        for idx in np.ndindex(*mesh):
            ev = store.arr[idx][list(occ)]
            ev = ev.reshape(nocc, -1)
            if hwf_evec:
                c, v = orc.position_hwf(model, ev, dir, True, "orbital")
                out_store.arr[idx] = np.asarray(v).reshape(out_store.arr[idx].shape)
            else:
                c = orc.position_hwf(model, ev, dir)
            hwfc[idx] = c
        return hwfc

    def position_matrix(self, model, evec, dir):
        return np.array([orc.position_matrix(model, e, dir) for e in evec])

    def position_hwf(self, model, evec, dir, hwf_evec, orbital_basis):
        if not hwf_evec:
            return np.array([orc.position_hwf(model, e, dir) for e in evec]), None
        res = [orc.position_hwf(model, e, dir, True, "orbital" if orbital_basis else "bloch") for e in evec]
        return np.array([r[0] for r in res]), np.array([np.asarray(r[1]).reshape(r[1].shape[0], -1) for r in res])


_ENGINE = OracleEngine()


class tb_model(_model.tb_model):
    _engine_factory = staticmethod(lambda: _ENGINE)


class wf_array(_wfarray.wf_array):
    pass
