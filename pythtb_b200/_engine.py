"""B200 engine: the only numerical backend of ``pythtb_b200``.

PyTorch is used for device memory, pinned host buffers and streams; all
arithmetic happens in the hand-written sm_100a kernels of libtbk_b200.so,
called through the C ABI of ``include/tbk.h``.  If the library or a CUDA
device is missing every numerical call raises — there is no CPU fallback.

The engine methods mirror the seams of the reference
(``_gen_ham``/``_sol_ham``/``solve_all``/``solve_on_grid``/``_one_berry_loop``/
``_one_flux_plane``/``position_*``; file:line in the docstrings).
"""
import collections
import ctypes
import os
import weakref

import numpy as np

from . import _lib

_engine = None


def get_engine():
    global _engine
    if _engine is None:
        _engine = B200Engine()
    return _engine


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class DeviceStore(object):
    """Wavefunction storage of a ``wf_array``: a device tensor plus a lazily
    synchronised pinned host mirror (the reference keeps ``_wfs`` in host
    memory and lets users index/mutate it, pythtb.py:2419, 2662-2672).

    state: 'empty'  nothing materialised (all zeros, like the reference's np.zeros)
           'host'   the host mirror is authoritative (it has been handed out)
           'device' the device tensor is authoritative
    """

    def __init__(self, engine, shape, state_axis=None):
        """state_axis: index of the state (band) axis of `shape` when the DEVICE copy is to be stored state-major,
        [state][mesh...][orb(,spin)] — the layout the mesh kernels of 2..4-band models write and the Berry kernels
        read (a band's data is then contiguous: berry_flux([0]) of a two-band model fetches half the bytes).  The
        host mirror and the tensor `dev()` returns always have the reference's logical shape; `dev()` is then a
        permuted view of the state-major allocation."""
        self.engine = engine
        self.shape = tuple(int(x) for x in shape)
        self.state_axis = state_axis
        self._host_t = None
        self._host = None
        self._dev = None
        self._phys = None
        self.state = "empty"

    def _alloc_dev(self):
        torch = self.engine.torch
        ax = self.state_axis
        if ax is None:
            self._phys = torch.zeros(self.shape, dtype=torch.complex128, device=self.engine.device)
            return self._phys
        phys_shape = (self.shape[ax],) + self.shape[:ax] + self.shape[ax + 1:]
        self._phys = torch.zeros(phys_shape, dtype=torch.complex128, device=self.engine.device)
        perm = list(range(1, ax + 1)) + [0] + list(range(ax + 1, len(self.shape)))
        return self._phys.permute(perm)               # logical shape, state-major strides

    def host(self):
        """Host ndarray (a view the caller may mutate)."""
        torch = self.engine.torch
        if self._host is None:
            self._host_t = torch.zeros(self.shape, dtype=torch.complex128, pin_memory=True)
            self._host = self._host_t.numpy()
        if self.state == "device":
            self._host_t.copy_(self._dev, non_blocking=False)
        self.state = "host"
        return self._host

    def replace_host(self, arr):
        torch = self.engine.torch
        arr = np.ascontiguousarray(arr, dtype=complex)
        self.shape = arr.shape
        self._host_t = torch.empty(self.shape, dtype=torch.complex128, pin_memory=True)
        self._host = self._host_t.numpy()
        self._host[...] = arr
        self._dev = self._phys = None
        self.state = "host"

    def dev(self, will_write, discard=False):
        """Device tensor, uploaded if the host mirror is newer — unless the caller is about to overwrite
        every element (``discard``: 67 MB of H2D per step for a script that reads ``_wfs`` between two
        grid solves of the 1024 x 1024 mesh)."""
        torch = self.engine.torch
        if self._dev is None or tuple(self._dev.shape) != self.shape:
            self._dev = self._alloc_dev()
            if self.state == "device":
                self.state = "empty"
        if self.state == "host" and not (discard and will_write):
            self._dev.copy_(self._host_t, non_blocking=True)
        if will_write:
            self.state = "device"
        # after a read-only use the host mirror stays authoritative: the caller may
        # still hold (and mutate) the view it was given
        return self._dev


class _Prepared(object):
    """Owner of a ``tbk_prepared`` handle (include/tbk.h): destroyed with the object."""
    __slots__ = ("handle", "__weakref__")


class B200Engine(object):
    def __init__(self):
        self.lib = _lib.load()          # raises if the .so is missing
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pythtb_b200: no CUDA device visible. This package runs only on a GPU "
                               "(built for B200 / sm_100a); there is no CPU fallback.")
        self.torch = torch
        self.device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))
        torch.cuda.set_device(self.device)
        self._ws = None
        self._host_results = {}
        self._peer = None
        self._pending_keep = collections.deque(maxlen=16)   # result tensors of posted, not yet completed reductions
        self._dev_index = self.device.index
        self._raw_stream = torch._C._cuda_getCurrentRawStream

    # ------------------------------------------------------------------ helpers
    @property
    def launches(self):
        """Kernels launched by libtbk_b200.so in this process (counted inside the library)."""
        return int(self.lib.tbk_launch_count())

    def stream(self):
        # raw cudaStream_t of torch's current stream (one C call; torch.cuda.current_stream() costs ~10 us)
        return ctypes.c_void_p(self._raw_stream(self._dev_index))

    def host_result(self, n):
        """A small pinned host buffer the kernels write their results into directly
        (pinned memory is device-addressable under UVA): a result costs one stream
        synchronisation instead of a cudaMemcpy.  Returns (tensor, numpy view)."""
        buf = self._host_results.get(n)
        if buf is None:
            t = self.torch.zeros(max(int(n), 1), dtype=self.torch.float64, pin_memory=True)
            buf = self._host_results[n] = (t, t.numpy())
        return buf

    def sync(self):
        # cudaStreamSynchronize on torch's current raw stream: one C call
        # (torch.cuda.current_stream(...).synchronize() costs ~10 us of Python on top)
        _lib.check(self.lib.tbk_stream_sync(self.stream()))

    def _prepare(self, fn, args, peer):
        """Bind the arguments of a solve_grid / flux_plane call in the library (tbk_*_prepare)."""
        out = ctypes.c_void_p(0)
        _lib.check(fn(*args, peer, ctypes.byref(out)))
        h = _Prepared()
        h.handle = ctypes.c_void_p(out.value)
        weakref.finalize(h, self.lib.tbk_prepared_destroy, h.handle)
        return h

    def workspace(self, nbytes):
        nbytes = int(nbytes)
        if nbytes <= 0:
            nbytes = 256
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
        return self._ws

    def to_host(self, t):
        """Device tensor -> numpy array backed by PINNED host memory from torch's caching host allocator: the
        copy runs at PCIe rate instead of through the driver's pageable staging (~2 GB/s for a 1 GB result), and a
        sweep that drops its previous result gets the same pinned block back without paying for the pinning."""
        if t.numel() * t.element_size() < (1 << 20):
            return t.cpu().numpy()
        h = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        self.sync()
        return h.numpy()

    def to_dev(self, arr, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype))
        return t.to(self.device, non_blocking=False)

    def model_handle(self, model):
        """Upload (once) the compiled plan of ``model``; cached on the plan."""
        plan = model._plan()
        h = getattr(plan, "_handle", None)
        if h is not None:
            return h, plan
        d = _lib.ModelDesc(plan.dim_k, plan.nsta, plan.nph, plan.nel, plan.nterm, plan.convention)
        for name in ("ph_R", "tau", "el_ptr", "el_row", "el_col", "t_ph", "t_amp", "pm_ptr", "pm_el", "pm_amp"):
            setattr(d, name, getattr(plan, name).ctypes.data)
        out = ctypes.c_void_p(0)
        _lib.check(self.lib.tbk_model_create(ctypes.byref(d), ctypes.byref(out)))
        handle = ctypes.c_void_p(out.value)
        plan._handle = handle
        weakref.finalize(plan, self.lib.tbk_model_destroy, handle)
        return handle, plan

    # ------------------------------------------------------- Hamiltonian / eigh
    def gen_ham(self, model, klist):
        """tb_model._gen_ham batched (pythtb.py:874-925) -> [nk,n,n] complex."""
        torch = self.torch
        handle, plan = self.model_handle(model)
        nk, n = klist.shape[0], plan.nsta
        kd = self.to_dev(klist, np.float64) if plan.dim_k > 0 else None
        ham = torch.empty((nk, n, n), dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.tbk_gen_ham(handle, _ptr(kd), nk, _ptr(ham), self.stream()))
        return ham.cpu().numpy()

    def eigh(self, ham, eig_vectors):
        """_sol_ham batched (pythtb.py:927-953): rows of evec are eigenvectors."""
        torch = self.torch
        ham = np.ascontiguousarray(ham, dtype=complex)
        batch, n = ham.shape[0], ham.shape[1]
        hd = self.to_dev(ham)
        ev = torch.empty((batch, n), dtype=torch.float64, device=self.device)
        vec = torch.empty((batch, n, n), dtype=torch.complex128, device=self.device) if eig_vectors else None
        wsb = self.lib.tbk_eigh_workspace(n, batch, int(eig_vectors))
        ws = self.workspace(wsb)
        _lib.check(self.lib.tbk_eigh_batched(_ptr(hd), n, batch, _ptr(ev), _ptr(vec), _ptr(ws), ws.numel(), self.stream()))
        return ev.cpu().numpy(), (vec.cpu().numpy() if eig_vectors else None)

    def solve_all_device(self, model, kd, nk, eig_vectors):
        """Fused assembly+eigh for a device k-list; returns device tensors in the
        reference layouts eval[band,k], evec[band,k,orb(,spin)] (pythtb.py:1040-1045)."""
        torch = self.torch
        handle, plan = self.model_handle(model)
        n = plan.nsta
        ev = torch.empty((n, nk), dtype=torch.float64, device=self.device)
        vec = None
        if eig_vectors:
            vec = torch.empty((n, nk, n), dtype=torch.complex128, device=self.device)
        ws = self.workspace(self.lib.tbk_solve_workspace(n, nk, int(eig_vectors)))
        _lib.check(self.lib.tbk_solve_k(handle, _ptr(kd), nk, _ptr(ev), nk, 1, _ptr(vec), nk * n, n,
                                        _ptr(ws), ws.numel(), self.stream()))
        return ev, vec

    CHUNK_K = 1 << 21          # k-points per chunk of a pipelined host-result sweep

    def solve_all_host(self, model, kd, nk, eig_vectors):
        """solve_all with HOST results for a large device k-list: the sweep is cut into chunks of CHUNK_K k-points whose
        kernels run on the engine's stream while the previous chunk's results cross PCIe on a copy stream (two device
        buffers, pinned destination in the reference layouts eval[band,k], evec[band,k,orb]).  A 256^3 eigenvalue sweep
        of an 8-band model is 1.07 GB of results: 20 ms of copy that used to follow 25 ms of kernels now hides behind them."""
        torch = self.torch
        handle, plan = self.model_handle(model)
        n = plan.nsta
        if nk < 2 * self.CHUNK_K:
            ev, vec = self.solve_all_device(model, kd, nk, eig_vectors)
            return self.to_host(ev), (self.to_host(vec) if vec is not None else None)
        C = self.CHUNK_K
        ev_h = torch.empty((n, nk), dtype=torch.float64, pin_memory=True)
        vec_h = torch.empty((n, nk, n), dtype=torch.complex128, pin_memory=True) if eig_vectors else None
        bufs = [(torch.empty((n, C), dtype=torch.float64, device=self.device),
                 torch.empty((n, C, n), dtype=torch.complex128, device=self.device) if eig_vectors else None) for _ in range(2)]
        ws = self.workspace(self.lib.tbk_solve_workspace(n, C, int(eig_vectors)))
        main = torch.cuda.current_stream(self.device)
        copier = getattr(self, "_copy_stream", None)
        if copier is None:
            copier = self._copy_stream = torch.cuda.Stream(device=self.device)
        freed = [None, None]
        dk = plan.dim_k
        for i, k0 in enumerate(range(0, nk, C)):
            cnt = min(C, nk - k0)
            ev_d, vec_d = bufs[i & 1]
            if freed[i & 1] is not None:
                main.wait_event(freed[i & 1])               # the copy out of this buffer (two chunks ago) has finished
            kptr = ctypes.c_void_p(kd.data_ptr() + k0 * dk * 8) if kd is not None else ctypes.c_void_p(0)
            _lib.check(self.lib.tbk_solve_k(handle, kptr, cnt, _ptr(ev_d), C, 1, _ptr(vec_d), C * n, n,
                                            _ptr(ws), ws.numel(), self.stream()))
            done = torch.cuda.Event()
            done.record(main)
            copier.wait_event(done)
            with torch.cuda.stream(copier):
                for b in range(n):                          # one contiguous run per band on either side
                    ev_h[b, k0:k0 + cnt].copy_(ev_d[b, :cnt], non_blocking=True)
                    if eig_vectors:
                        vec_h[b, k0:k0 + cnt].copy_(vec_d[b, :cnt], non_blocking=True)
                freed[i & 1] = torch.cuda.Event()
                freed[i & 1].record(copier)
        copier.synchronize()
        self.sync()
        return ev_h.numpy(), (vec_h.numpy() if eig_vectors else None)

    def solve_slice(self, model, store, dim_arr, fixed, free, kpts):
        """wf_array.solve_on_slice: eigenvectors of ``model`` at ``kpts[free..., dim_k]`` written straight
        into the slice ``store[fixed]`` of the device array through tbk_solve_k's output strides (one fused
        assemble + diagonalise launch per run of the last free axis); returns eval[free..., nsta] (host)."""
        torch = self.torch
        handle, plan = self.model_handle(model)
        n = plan.nsta
        wfs = store.dev(will_write=True)            # uploads the host mirror first if that is the newer copy
        stride = [int(x) for x in wfs.stride()]     # in complex elements
        base = sum(int(fixed[d]) * stride[d] for d in fixed)
        fshape = [int(store.shape[d]) for d in free]        # (a shard's store has its local extents)
        nk = fshape[-1]
        outer = fshape[:-1]
        kd = self.to_dev(kpts.reshape(-1, plan.dim_k), np.float64)
        ev = torch.empty((int(np.prod(fshape)), n), dtype=torch.float64, device=self.device)
        ws = self.workspace(self.lib.tbk_solve_workspace(n, nk, 1))
        run = 0
        for idx in np.ndindex(*outer):
            off = base + sum(int(i) * stride[d] for i, d in zip(idx, free[:-1]))
            _lib.check(self.lib.tbk_solve_k(
                handle, ctypes.c_void_p(kd.data_ptr() + run * nk * plan.dim_k * 8), nk,
                ctypes.c_void_p(ev.data_ptr() + run * nk * n * 8), 1, n,
                ctypes.c_void_p(wfs.data_ptr() + off * 16), stride[dim_arr], stride[free[-1]],
                _ptr(ws), ws.numel(), self.stream()))
            run += 1
        return ev.cpu().numpy().reshape(tuple(fshape) + (n,))

    def solve_all_mesh(self, model, mesh_size, eig_vectors, device_result=False):
        """solve_all on tb_model.k_uniform_mesh(mesh_size) with the k-points generated on the device
        (tbk_kmesh_uniform): nothing but the results crosses PCIe — and not even those with
        ``device_result`` (torch tensors in the reference layouts)."""
        torch = self.torch
        nd = len(mesh_size)
        nk = int(np.prod(mesh_size))
        kd = torch.empty((nk, nd), dtype=torch.float64, device=self.device)
        mesh = (ctypes.c_int32 * nd)(*[int(x) for x in mesh_size])
        _lib.check(self.lib.tbk_kmesh_uniform(mesh, nd, _ptr(kd), self.stream()))
        if not device_result:
            ev_h, vec_h = self.solve_all_host(model, kd, nk, eig_vectors)
            if not eig_vectors:
                return ev_h
            if model._nspin == 2:
                vec_h = vec_h.reshape(model._nsta, nk, model._norb, 2)
            return ev_h, vec_h
        ev, vec = self.solve_all_device(model, kd, nk, eig_vectors)
        if vec is not None and model._nspin == 2:
            vec = vec.reshape(model._nsta, nk, model._norb, 2)
        return (ev, vec) if eig_vectors else ev

    def solve_all(self, model, klist, eig_vectors):
        nk = klist.shape[0]
        if nk == 0:      # empty k-list: nothing to launch (reference returns empty arrays)
            ev_h = np.zeros((model._nsta, 0), dtype=float)
            if not eig_vectors:
                return ev_h
            tail = (model._norb,) if model._nspin == 1 else (model._norb, 2)
            return ev_h, np.zeros((model._nsta, 0) + tail, dtype=complex)
        kd = self.to_dev(klist, np.float64) if model._dim_k > 0 else None
        ev_h, vec_h = self.solve_all_host(model, kd, nk, eig_vectors)
        if not eig_vectors:
            return ev_h
        if model._nspin == 2:
            vec_h = vec_h.reshape(model._nsta, nk, model._norb, 2)
        return ev_h, vec_h

    # ------------------------------------------------------------ wf_array ops
    def new_store(self, shape, state_axis=None):
        return DeviceStore(self, shape, state_axis)

    def pbc_phases(self, orb, nspin, k_dirs, convention=1):
        """exp(-2 pi i tau_j[k_dir]) per state (pythtb.py:2729-2736), [len(k_dirs), nsta];
        Convention II eigenvectors are periodic in k (formalism tex:341-364): factor 1."""
        out = []
        for kd in k_dirs:
            ffac = np.exp(-2.0j * np.pi * np.asarray(orb)[:, kd])
            if convention == 2:
                ffac = np.ones_like(ffac)
            out.append(np.repeat(ffac, nspin))
        return np.array(out, dtype=complex)

    def _cached(self, owner, key, make):
        """Small device constants (pbc phases, occ lists, slice offsets) are
        uploaded once per owner object, not once per call."""
        cache = owner.__dict__.setdefault("_tbk_dev_cache", {})
        val = cache.get(key)
        if val is None:
            val = cache[key] = make()
        return val

    def solve_grid(self, model, store, mesh_arr, start_k, row0=0, nrows=None, wrap0=1, want_gaps=True,
                   host_result=False, reduce_ranks=None, defer_reduce=False):
        """wf_array.solve_on_grid + impose_pbc (pythtb.py:2421-2532) into ``store``
        (local rows [row0, row0+nrows] of mesh axis 0); returns the minimal gaps
        over the rows solved here as a device tensor — or, with ``host_result``, as a
        host array (the call then synchronises the stream) — or None.
        wrap0: 1 = write the periodic image of row 0 (single shard), 0 = leave the
        closing row to a halo exchange, 2 = compute the closing row in this launch.
        defer_reduce (sharded, device result only): the kernel only POSTS this rank's gaps to
        the peers; the returned tensor holds the minimum over the ranks once a later kernel has
        completed the reduction — a synchronous collective, a deferred flux_total at least one
        step later, or peer_flush() (include/tbk.h: tbk_peer_defer)."""
        torch = self.torch
        # ---- fast path: the same call as last time on this store (a parameter sweep, the bench loop):
        # reuse the marshalled arguments as long as every buffer they point to is still the live one
        fast_key = None
        if host_result or not want_gaps:
            fast_key = (id(model._plan()), tuple(int(x) for x in mesh_arr), tuple(float(x) for x in start_k), row0, nrows,
                        wrap0, want_gaps, host_result, reduce_ranks)
            hit = store.__dict__.get("_tbk_sg_fast")
            if hit is not None and hit[0] == fast_key and hit[1] is store._dev and hit[2] is self._ws:
                store.state = "device"                # every element is overwritten: a newer host mirror is irrelevant
                gaps_h = hit[5]
                _lib.check(self.lib.tbk_prepared_run(hit[3].handle, self.stream(), 1 if gaps_h is not None else 0))
                if gaps_h is not None:
                    if reduce_ranks is not None and not np.all(np.isfinite(gaps_h)):
                        raise _lib.TbkError("\n\nsolve_on_grid: a peer rank never delivered its gaps (fused reduction timed out)")
                    return gaps_h.copy()
                return None
        handle, plan = self.model_handle(model)
        nd, n = len(mesh_arr), plan.nsta
        if nrows is None:
            nrows = int(mesh_arr[0]) - 1
        wfs = store.dev(will_write=True, discard=True)      # the solve writes every element (the closing row of a
                                                            # halo-exchange shard is filled by halo_ring_shift right after)
        cache = plan.__dict__.setdefault("_tbk_dev_cache", {})
        phase = cache.get(("pbc", nd))
        if phase is None:
            phase = cache[("pbc", nd)] = self.to_dev(
                self.pbc_phases(model._orb, model._nspin, [model._per[d] for d in range(nd)], model._convention))
        gaps = gaps_h = None
        if n > 1 and want_gaps:
            if host_result:
                gaps, gaps_h = self.host_result(n - 1)
            else:
                gaps = torch.empty(n - 1, dtype=torch.float64, device=self.device)
        npts = nrows + 1
        for d in range(1, nd):
            npts *= int(mesh_arr[d]) - 1
        ws = self.workspace(self.lib.tbk_solve_workspace(n, npts, 1))
        start = (ctypes.c_double * nd)(*[float(x) for x in start_k])
        mesh = (ctypes.c_int32 * nd)(*[int(x) for x in mesh_arr])
        # a state-major store (DeviceStore.state_axis) is told to the kernels by its state stride
        sstride = int(wfs.stride(nd)) if store.state_axis is not None else 0
        args = (handle, start, mesh, nd, int(row0), int(nrows), int(wrap0), _ptr(wfs), _ptr(phase), _ptr(gaps),
                _ptr(ws), ws.numel(), sstride)
        reduced = reduce_ranks is None
        launched = False
        if gaps is not None and reduce_ranks is not None:
            # minimum over the ranks inside the kernel, through NVLink peer memory (csrc/tbk_peer.cuh)
            peer = self.peer_group(*reduce_ranks)
            if peer is not None:
                deferred = bool(defer_reduce and gaps_h is None)
                if deferred:
                    _lib.check(self.lib.tbk_peer_defer(peer, 1))
                rc = self.lib.tbk_solve_grid_x(*args, peer, self.stream())
                if deferred:
                    _lib.check(self.lib.tbk_peer_defer(peer, 0))
                if rc == 0:
                    reduced = launched = True
                    if deferred:
                        self._pending_keep.append(gaps)    # alive until the kernel that completes it has been enqueued
                    if fast_key is not None and gaps_h is not None:
                        store.__dict__["_tbk_sg_fast"] = (fast_key, store._dev, self._ws,
                                                          self._prepare(self.lib.tbk_solve_grid_prepare, args, peer),
                                                          gaps, gaps_h, (plan, phase), peer)
                elif rc != _lib.ERR_UNSUPPORTED:
                    _lib.check(rc)
        if not launched:
            _lib.check(self.lib.tbk_solve_grid_x(*args, None, self.stream()))
            if fast_key is not None and reduce_ranks is None and (gaps is None or gaps_h is not None):
                # keep plan / phase / start / mesh alive with the cached argument tuple
                store.__dict__["_tbk_sg_fast"] = (fast_key, store._dev, self._ws,
                                                  self._prepare(self.lib.tbk_solve_grid_prepare, args, None),
                                                  gaps, gaps_h, (plan, phase), None)
        if gaps is not None and not reduced:
            # not eligible for the fused reduction: NCCL all-reduce of the per-rank minima
            if gaps_h is not None:
                self.sync()
                return self.allreduce(gaps_h.copy(), "min")
            return self.allreduce(gaps, "min")
        if gaps_h is not None:
            self.sync()
            if reduce_ranks is not None and not np.all(np.isfinite(gaps_h)):
                raise _lib.TbkError("\n\nsolve_on_grid: a peer rank never delivered its gaps (fused reduction timed out)")
            return gaps_h.copy()
        return gaps

    @property
    def last_solve_kernel(self):
        return _lib.last_kernel(self.lib)

    # ---------------------------------------------------------- multi-GPU plumbing
    def peer_group(self, rank, nranks):
        """The NVLink peer-memory mailbox group used by the fused cross-rank reductions
        (csrc/tbk_peer.cuh): created once per process, IPC handles exchanged through
        torch.distributed.  Returns the opaque handle, or None when unavailable
        (more than 8 ranks, or PYTHTB_B200_PEER=0)."""
        if self._peer is not None:
            return self._peer if self._peer is not False else None
        self._peer = False
        if nranks > 8 or os.environ.get("PYTHTB_B200_PEER", "1") == "0":
            return None
        import torch.distributed as dist
        handle = ctypes.create_string_buffer(64)
        out = ctypes.c_void_p(0)
        _lib.check(self.lib.tbk_peer_create(int(rank), int(nranks), ctypes.byref(out), handle))
        all_handles = [None] * nranks
        dist.all_gather_object(all_handles, bytes(handle.raw))
        blob = ctypes.create_string_buffer(b"".join(all_handles), 64 * nranks)
        peer = ctypes.c_void_p(out.value)
        _lib.check(self.lib.tbk_peer_connect(peer, blob))
        dist.barrier()                      # every mailbox is mapped everywhere before the first collective
        self._peer = peer
        return peer

    def peer_flush(self):
        """Finish a deferred cross-rank reduction now (tbk_peer_flush); no-op when none is pending."""
        if self._peer:
            _lib.check(self.lib.tbk_peer_flush(self._peer, self.stream()))

    def peer_barrier(self):
        """Device-side barrier over the peer group, in stream order (tbk_peer_barrier): the kernels
        enqueued after it start together on every rank.  No-op without a peer group."""
        if self._peer:
            _lib.check(self.lib.tbk_peer_barrier(self._peer, self.stream()))

    def halo_ring_shift(self, store, dim_arr, phase, rank, nranks):
        """Close every rank's slab with the first row of the next rank: one
        grouped NCCL send/recv over NVLink (rank r sends its row 0 to r-1; rank 0's
        row goes to the last rank multiplied by the pbc phase, pythtb.py:2729)."""
        import torch.distributed as dist
        wfs = store.dev(will_write=True)
        row = wfs[0].contiguous()                 # (a copy only for a state-major store, whose rows are strided)
        nsta_arr = store.shape[dim_arr]
        n = int(np.prod(store.shape[dim_arr + 1:]))
        send = self.torch.empty_like(row)
        ph = self.to_dev(phase) if phase is not None else None
        _lib.check(self.lib.tbk_halo_pack(_ptr(row), _ptr(send), row.numel() // (nsta_arr * n), nsta_arr, n,
                                          _ptr(ph), self.stream()))
        last = wfs[wfs.shape[0] - 1]
        recv = last if last.is_contiguous() else self.torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, (rank - 1) % nranks), dist.P2POp(dist.irecv, recv, (rank + 1) % nranks)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if recv is not last:
            last.copy_(recv)

    def allreduce(self, x, op):
        """Sum / min of a small result vector over the ranks (NCCL)."""
        import torch.distributed as dist
        t = x if not isinstance(x, np.ndarray) else self.to_dev(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN)
        return t.cpu().numpy() if isinstance(x, np.ndarray) else t

    def allgather_rows(self, local, nranks, n0):
        """Gather per-rank row blocks (host arrays, rows along axis 0); returns the
        list of blocks in rank order.  Row counts follow from the sharding rule, so
        only one padded all-gather is needed."""
        import torch.distributed as dist
        torch = self.torch
        base, rem = divmod(n0 - 1, nranks)
        maxrows = base + 2
        tail = local.shape[1:]
        pad = np.zeros((maxrows,) + tail, dtype=np.float64)
        pad[:local.shape[0]] = local
        out = torch.empty((nranks, maxrows) + tail, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(out, self.to_dev(pad).reshape((1, maxrows) + tail))
        cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=self.device)
        cnts = torch.empty(nranks, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(cnts, cnt)
        out_h, cnts_h = out.cpu().numpy(), cnts.cpu().numpy()
        return [out_h[r, :cnts_h[r]] for r in range(nranks)]

    def impose_boundary(self, store, dim_arr, mesh_dir, phase):
        """impose_pbc / impose_loop on the device (pythtb.py:2674-2791)."""
        shape = store.shape
        wfs = store.dev(will_write=True)
        outer = int(np.prod(shape[:mesh_dir])) if mesh_dir > 0 else 1
        length = shape[mesh_dir]
        inner = int(np.prod(shape[mesh_dir + 1:dim_arr])) if mesh_dir + 1 < dim_arr else 1
        nsta_arr = shape[dim_arr]
        n = int(np.prod(shape[dim_arr + 1:]))
        if store.state_axis is not None:
            # state-major allocation [state][mesh...][orb]: every state plane is an array of single-state points
            outer, nsta_arr = outer * nsta_arr, 1
        ph = self.to_dev(phase) if phase is not None else None
        _lib.check(self.lib.tbk_impose_boundary(_ptr(wfs), outer, length, inner, nsta_arr, n, _ptr(ph), self.stream()))

    def _view(self, store, dim_arr, occ):
        wfs = store.dev(will_write=False)
        shape = store.shape
        key = ("view", shape, dim_arr, tuple(int(x) for x in occ))
        cache = store.__dict__.setdefault("_tbk_dev_cache", {})
        hit = cache.get(key)
        if hit is None:
            nsta_arr = shape[dim_arr]
            n = 1
            for x in shape[dim_arr + 1:]:
                n *= int(x)
            occ = np.asarray(occ, dtype=np.int64)
            if occ.size == 0:
                raise Exception("\n\nNo states selected.")
            occ = np.where(occ < 0, occ + nsta_arr, occ)
            if occ.min() < 0 or occ.max() >= nsta_arr:
                raise IndexError("index in occ out of bounds")
            occ_d = self.to_dev(occ.astype(np.int32))
            strides = [int(wfs.stride(d)) for d in range(dim_arr)]      # mesh-axis strides, complex elements
            sstride = int(wfs.stride(dim_arr))                          # n, or the state-major plane size
            hit = cache[key] = (n, nsta_arr, len(occ), occ_d, strides, sstride)
        n, nsta_arr, nocc, occ_d, strides, sstride = hit
        view = _lib.WfView(wfs.data_ptr(), n, nsta_arr, nocc, occ_d.data_ptr(), sstride)
        return view, strides, (wfs, occ_d)

    def _offsets(self, store, dim_arr, axes, strides):
        """Element offsets of every slice/string labelled by the mesh axes ``axes`` (device, cached)."""
        mesh = store.shape[:dim_arr]

        def make():
            oshape = tuple(mesh[d] for d in axes)
            offs = np.zeros(oshape if oshape else (1,), dtype=np.int64)
            for ax, d in enumerate(axes):
                sh = [1] * len(axes)
                sh[ax] = mesh[d]
                offs = offs + (np.arange(mesh[d], dtype=np.int64) * strides[d]).reshape(sh)
            return self.to_dev(np.ascontiguousarray(offs.reshape(-1)))
        return self._cached(store, ("offs", tuple(store.shape), tuple(axes)), make)

    def berry_strings(self, store, dim_arr, occ, dir, berry_evals):
        """_one_berry_loop for every string along ``dir`` (pythtb.py:2979-3029,
        3798-3838); raw phases in [-pi,pi), shape other_axes (+[nocc])."""
        torch = self.torch
        view, strides, keep = self._view(store, dim_arr, occ)
        mesh = store.shape[:dim_arr]
        other = [d for d in range(dim_arr) if d != dir]
        oshape = tuple(mesh[d] for d in other)
        offs_d = self._offsets(store, dim_arr, other, strides)
        nstr, npts, nocc = int(offs_d.numel()), mesh[dir], view.nocc
        out = torch.empty((nstr, nocc) if berry_evals else (nstr,), dtype=torch.float64, device=self.device)
        ws = self.workspace(self.lib.tbk_berry_workspace(nocc, view.n, nstr, npts, int(berry_evals)))
        _lib.check(self.lib.tbk_berry_strings(ctypes.byref(view), _ptr(offs_d), nstr, npts, strides[dir],
                                              int(berry_evals), _ptr(out), _ptr(ws), ws.numel(), self.stream()))
        res = out.cpu().numpy()
        if berry_evals and not np.all(np.isfinite(res)):
            raise Exception("\n\nberry_phase(berry_evals=True): the wave functions of a link are not finite (NaN / Inf); "
                            "the unitary polar factor of their overlap is undefined.")
        return res.reshape(oshape + ((nocc,) if berry_evals else ()))

    def wilson_phases_across_ranks(self, store, dim_arr, occ, dir, nranks):
        """berry_evals=True for strings that run along the sharded axis: every rank forms the ordered product
        of the polar link matrices of ITS links (tbk_wilson_products), the per-rank products are all-gathered in
        rank order over NCCL, and tbk_wilson_phases multiplies them and returns the sorted eigenphases
        (pythtb.py:3813-3838 with the product split by rank).  Shape other_axes + [nocc]."""
        import torch.distributed as dist
        torch = self.torch
        view, strides, keep = self._view(store, dim_arr, occ)
        mesh = store.shape[:dim_arr]
        other = [d for d in range(dim_arr) if d != dir]
        oshape = tuple(mesh[d] for d in other)
        offs_d = self._offsets(store, dim_arr, other, strides)
        nstr, npts, nocc = int(offs_d.numel()), mesh[dir], view.nocc
        prod = torch.empty((nstr, nocc, nocc), dtype=torch.complex128, device=self.device)
        ws = self.workspace(self.lib.tbk_berry_workspace(nocc, view.n, nstr, npts, 1))
        _lib.check(self.lib.tbk_wilson_products(ctypes.byref(view), _ptr(offs_d), nstr, npts, strides[dir], _ptr(prod),
                                                _ptr(ws), ws.numel(), self.stream()))
        return self.wilson_finish(prod, nranks).reshape(oshape + (nocc,))

    def wilson_finish(self, prod, nranks):
        """Per-rank ordered products prod[nstr, nocc, nocc] (device) -> sorted eigenphases [nstr, nocc] of
        prod_rank0 @ prod_rank1 @ ... (pythtb.py:3834-3838): one NCCL all-gather in rank order when nranks > 1,
        then tbk_wilson_phases."""
        torch = self.torch
        nstr, nocc = int(prod.shape[0]), int(prod.shape[1])
        if nranks > 1:
            import torch.distributed as dist
            allp = torch.empty((nranks, nstr, nocc, nocc), dtype=torch.complex128, device=self.device)
            dist.all_gather_into_tensor(torch.view_as_real(allp), torch.view_as_real(prod.contiguous()).unsqueeze(0))
            mats = allp.permute(1, 0, 2, 3).contiguous()      # [nstr][rank][nocc][nocc]: rank order = link order
        else:
            mats = prod.reshape(nstr, 1, nocc, nocc).clone()
        out = torch.empty((nstr, nocc), dtype=torch.float64, device=self.device)
        ws = self.workspace(self.lib.tbk_wilson_workspace(nocc, nstr, max(nranks, 1)))
        _lib.check(self.lib.tbk_wilson_phases(_ptr(mats), nstr, max(nranks, 1), nocc, _ptr(out), _ptr(ws), ws.numel(), self.stream()))
        res = out.cpu().numpy()
        if not np.all(np.isfinite(res)):
            raise Exception("\n\nberry_phase(berry_evals=True): the wave functions of a link are not finite (NaN / Inf); "
                            "the unitary polar factor of their overlap is undefined.")
        return res

    # ------------------------------------------------- streamed 1-D strings (BASELINE config 4)
    def stream_chunk(self, n):
        """Links per chunk of the streamed string: whole waves of the solver that takes matrices of this size
        (one 512-thread CTA per SM above n = 256, two 256-thread CTAs below), capped at ~1 GiB of eigenvectors
        (twice as long chunks were measured: same rate, twice the memory)."""
        if n > 256:
            c = 148 * 2
        elif n > 32:
            c = 148 * 4
        else:
            c = 1 << 16
        cap = max(8, (1 << 30) // (16 * n * n))
        return int(min(c, cap))

    def stream_links(self, model, npts, start_k, occ, l0, l1, berry_evals, want_gaps=False, chunk=None):
        """Links l0 .. l1-1 of the closed 1-D string k_g = start_k + g / (npts - 1), g = 0 .. npts-1 (the last point
        is the periodic image of the first, pythtb.py:2472-2486 + 2729) WITHOUT materialising the wave functions:
        chunks of <= `chunk` links go through assemble + diagonalise (tbk_solve_grid into a chunk buffer whose
        first row is the carried last point of the previous chunk) and straight into the link overlaps
        (pythtb.py:3813-3831).  Returns (result, gaps):
          berry_evals False: result = sum over the chunks of -arg det prod M (a float, NOT wrapped),
          berry_evals True : result = ordered product of the unitary link matrices, device tensor [1, nocc, nocc];
          gaps = minimal direct gaps over the points l0 .. l1 (host array) or None."""
        torch = self.torch
        handle, plan = self.model_handle(model)
        n = plan.nsta
        occ = np.asarray(occ, dtype=np.int64).reshape(-1)
        occ = np.where(occ < 0, occ + n, occ)
        if occ.size == 0:
            raise Exception("\n\nNo states selected.")
        if occ.min() < 0 or occ.max() >= n:
            raise IndexError("index in occ out of bounds")
        nocc = int(occ.size)
        l0, l1, npts = int(l0), int(l1), int(npts)
        if not (0 <= l0 < l1 <= npts - 1):
            raise Exception("\n\nstream_links: empty or out-of-range link interval")
        C = int(min(chunk or self.stream_chunk(n), l1 - l0))
        buf = torch.empty((C + 1, n, n), dtype=torch.complex128, device=self.device)
        cache = plan.__dict__.setdefault("_tbk_dev_cache", {})
        phase = cache.get(("pbc", 1))
        if phase is None:
            phase = cache[("pbc", 1)] = self.to_dev(self.pbc_phases(model._orb, model._nspin, [model._per[0]], model._convention))
        occ_d = self.to_dev(occ.astype(np.int32))
        off_d = self.to_dev(np.zeros(1, dtype=np.int64))
        nch = (l1 - l0 + C - 1) // C
        ws = self.workspace(max(self.lib.tbk_solve_workspace(n, C + 1, 1),
                                self.lib.tbk_berry_workspace(nocc, n, 1, C + 1, int(bool(berry_evals))),
                                self.lib.tbk_wilson_workspace(nocc, 1, nch)))
        gaps = gaps_run = None
        if want_gaps and n > 1:
            gaps = torch.empty(n - 1, dtype=torch.float64, device=self.device)
        if berry_evals:
            acc = torch.empty((nch, nocc, nocc), dtype=torch.complex128, device=self.device)
        else:
            acc = torch.empty(nch, dtype=torch.float64, device=self.device)
        start = (ctypes.c_double * 1)(float(np.asarray(start_k, dtype=float).reshape(-1)[0]))
        mesh = (ctypes.c_int32 * 1)(npts)
        blk = n * n                                   # complex elements per k-point
        lo, ci, prev = l0, 0, 0
        while lo < l1:
            c = min(C, l1 - lo)
            if ci == 0:                               # points lo .. lo+c
                dst, row0, nrows = buf.data_ptr(), lo, c
            else:                                     # carry the closing point of the previous chunk; points lo+1 .. lo+c
                buf[0].copy_(buf[prev])
                dst, row0, nrows = buf.data_ptr() + 16 * blk, lo + 1, c - 1
            _lib.check(self.lib.tbk_solve_grid(handle, start, mesh, 1, row0, nrows, 2, ctypes.c_void_p(dst), _ptr(phase),
                                               _ptr(gaps), _ptr(ws), ws.numel(), self.stream()))
            if gaps is not None:
                gaps_run = gaps.clone() if gaps_run is None else torch.minimum(gaps_run, gaps)
            view = _lib.WfView(buf.data_ptr(), n, n, nocc, occ_d.data_ptr(), 0)
            if berry_evals:
                _lib.check(self.lib.tbk_wilson_products(ctypes.byref(view), _ptr(off_d), 1, c + 1, blk, _ptr(acc[ci]),
                                                        _ptr(ws), ws.numel(), self.stream()))
            else:
                _lib.check(self.lib.tbk_berry_strings(ctypes.byref(view), _ptr(off_d), 1, c + 1, blk, 0,
                                                      ctypes.c_void_p(acc.data_ptr() + 8 * ci), _ptr(ws), ws.numel(), self.stream()))
            lo, ci, prev = lo + c, ci + 1, c
        gaps_h = gaps_run.cpu().numpy() if gaps_run is not None else None
        if not berry_evals:
            return float(acc.cpu().numpy().sum()), gaps_h
        prod = torch.empty((1, nocc, nocc), dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.tbk_wilson_chain(_ptr(acc), 1, nch, nocc, _ptr(prod), _ptr(ws), ws.numel(), self.stream()))
        return prod, gaps_h

    def stream_gaps(self, model, npts, start_k, l0, l1):
        """Minimal direct gaps over the points l0 .. l1-1 of the 1-D mesh (pythtb.py:2484, 2529-2530) from an
        eigenvalue-only pass; nothing but [nsta, chunk] eigenvalues is ever stored."""
        n = model._nsta
        if n <= 1:
            return None
        den = float(npts - 1)
        s0 = float(np.asarray(start_k, dtype=float).reshape(-1)[0])
        best = None
        step = 1 << 16
        for a in range(int(l0), int(l1), step):
            g = np.arange(a, min(int(l1), a + step), dtype=float)
            kd = self.to_dev((s0 + g / den).reshape(-1, 1), np.float64)           # pythtb.py:2477, same doubles as the kernels
            ev, _ = self.solve_all_device(model, kd, int(g.size), False)
            d = (ev[1:] - ev[:-1]).min(dim=1).values
            best = d if best is None else self.torch.minimum(best, d)
        return best.cpu().numpy()

    def flux(self, store, dim_arr, occ, dirs, individual):
        """_one_flux_plane on every 2-D slice spanned by ``dirs`` (pythtb.py:3133-3202).
        Returns plaquette phases [rest..., n0-1, n1-1] or their sums [rest...]."""
        tot, plq = self.flux_device(store, dim_arr, occ, dirs, want_total=not individual, want_plaq=individual)
        mesh = store.shape[:dim_arr]
        rest = [d for d in range(dim_arr) if d not in dirs]
        rshape = tuple(mesh[d] for d in rest)
        if individual:
            return plq.cpu().numpy().reshape(rshape + (mesh[dirs[0]] - 1, mesh[dirs[1]] - 1))
        return tot.cpu().numpy().reshape(rshape)

    def flux_total(self, store, dim_arr, occ, dirs, host_result=False, reduce_ranks=None, defer_reduce=False):
        """Sum of the plaquette phases of every local 2-D slice: a device tensor [nslice], or with
        ``host_result`` a host array (written by the kernel into pinned memory; the call synchronises).
        defer_reduce (sharded, device result): the sum over the ranks is only posted by this kernel (which
        completes the deferred reductions of EARLIER steps); the tensor holds the global sum once a later
        collective kernel or peer_flush() has completed it."""
        return self.flux_device(store, dim_arr, occ, dirs, want_total=True, want_plaq=False, host_result=host_result,
                                reduce_ranks=reduce_ranks, defer_reduce=defer_reduce)[0]

    def flux_device(self, store, dim_arr, occ, dirs, want_total=True, want_plaq=False, host_result=False,
                    reduce_ranks=None, defer_reduce=False):
        torch = self.torch
        fast_key = None
        if want_total and not want_plaq and host_result:
            fast_key = (tuple(int(x) for x in occ), tuple(dirs), tuple(store.shape), dim_arr, reduce_ranks)
            hit = store.__dict__.get("_tbk_fx_fast")
            if hit is not None and hit[0] == fast_key and hit[1] is store._dev and hit[2] is self._ws and \
                    store.state != "host":
                _lib.check(self.lib.tbk_prepared_run(hit[3].handle, self.stream(), 1))
                if reduce_ranks is not None and not np.all(np.isfinite(hit[4])):
                    raise _lib.TbkError("\n\nberry_flux: a peer rank never delivered its partial sum (fused reduction timed out)")
                return hit[4].copy(), None
        view, strides, keep = self._view(store, dim_arr, occ)
        mesh = store.shape[:dim_arr]
        rest = [d for d in range(dim_arr) if d not in dirs]
        offs_d = self._offsets(store, dim_arr, rest, strides)
        nslice = int(offs_d.numel())
        n0, n1 = mesh[dirs[0]], mesh[dirs[1]]
        plq = torch.empty((nslice, n0 - 1, n1 - 1), dtype=torch.float64, device=self.device) if want_plaq else None
        tot = tot_h = None
        if want_total:
            if host_result and nslice <= 4096:
                tot, tot_h = self.host_result(nslice)
            else:
                tot = torch.empty((nslice,), dtype=torch.float64, device=self.device)
        ws = self.workspace(self.lib.tbk_flux_workspace(view.nocc, view.n, nslice, n0, n1))
        args = (ctypes.byref(view), _ptr(offs_d), nslice, n0, strides[dirs[0]], n1, strides[dirs[1]], _ptr(plq),
                _ptr(tot), _ptr(ws), ws.numel())
        reduced = reduce_ranks is None
        launched = False
        if tot is not None and reduce_ranks is not None:
            # sum over the ranks inside the kernel, through NVLink peer memory (csrc/tbk_peer.cuh)
            peer = self.peer_group(*reduce_ranks)
            if peer is not None:
                deferred = bool(defer_reduce and tot_h is None)
                if deferred:
                    _lib.check(self.lib.tbk_peer_defer(peer, 1))
                rc = self.lib.tbk_flux_plane_x(*args, peer, self.stream())
                if deferred:
                    _lib.check(self.lib.tbk_peer_defer(peer, 0))
                if rc == 0:
                    reduced = launched = True
                    if deferred:
                        self._pending_keep.append(tot)
                    if fast_key is not None and tot_h is not None and store.state != "host":
                        store.__dict__["_tbk_fx_fast"] = (fast_key, store._dev, self._ws,
                                                          self._prepare(self.lib.tbk_flux_plane_prepare, args, peer),
                                                          tot_h, (view, keep, offs_d), peer)
                elif rc != _lib.ERR_UNSUPPORTED:
                    _lib.check(rc)
        if not launched:
            _lib.check(self.lib.tbk_flux_plane(*args, self.stream()))
            if fast_key is not None and reduce_ranks is None and tot_h is not None and store.state != "host":
                store.__dict__["_tbk_fx_fast"] = (fast_key, store._dev, self._ws,
                                                  self._prepare(self.lib.tbk_flux_plane_prepare, args, None),
                                                  tot_h, (view, keep, offs_d), None)
        if tot is not None and not reduced:
            # not eligible for the fused reduction: NCCL all-reduce of the per-rank sums
            if tot_h is not None:
                self.sync()
                return self.allreduce(tot_h.copy(), "sum"), plq
            tot = self.allreduce(tot, "sum")
            return (tot.cpu().numpy() if host_result else tot), plq
        if tot_h is not None:
            self.sync()
            if reduce_ranks is not None and not np.all(np.isfinite(tot_h)):
                raise _lib.TbkError("\n\nberry_flux: a peer rank never delivered its partial sum (fused reduction timed out)")
            return tot_h.copy(), plq
        if host_result and tot is not None:
            return tot.cpu().numpy(), plq
        return tot, plq

    # ------------------------------------------------------- position operator
    def _pos(self, model, dir):
        return np.repeat(np.asarray(model._orb, dtype=float)[:, dir], model._nspin)

    def position_matrix(self, model, evec, dir):
        """tb_model.position_matrix batched (pythtb.py:2034-2113): evec[batch,nocc,n]."""
        torch = self.torch
        evec = np.ascontiguousarray(evec, dtype=complex)
        batch, nocc, n = evec.shape
        ed = self.to_dev(evec)
        pos = self.to_dev(self._pos(model, dir), np.float64)
        x = torch.empty((batch, nocc, nocc), dtype=torch.complex128, device=self.device)
        _lib.check(self.lib.tbk_position_matrix(_ptr(ed), batch, nocc, n, _ptr(pos), _ptr(x), self.stream()))
        return x.cpu().numpy()

    def position_hwf_store(self, model, store, dim_arr, occ, dir, hwf_evec, out_store=None):
        """position_hwf at EVERY mesh point of a wf_array in one batched call, device to device
        (the per-k Python loop of examples/cubic_slab_hwf.py:63-79).  Returns the centres
        [mesh..., nocc] as a host array; with ``hwf_evec`` the hybrid Wannier functions in the orbital basis are
        written into ``out_store`` (shape mesh + (nocc, norb(,2))) without leaving the device."""
        torch = self.torch
        wfs = store.dev(will_write=False)
        mesh = store.shape[:dim_arr]
        nsta = store.shape[dim_arr]
        n = 1
        for x in store.shape[dim_arr + 1:]:
            n *= int(x)
        occ_t = torch.as_tensor(np.asarray(occ, dtype=np.int64), device=self.device)
        nocc = int(occ_t.numel())
        batch = 1
        for x in mesh:
            batch *= int(x)
        pos = self.to_dev(self._pos(model, dir), np.float64)
        hwfc = torch.empty((batch, nocc), dtype=torch.float64, device=self.device)
        hwf = hwf_log = None
        if hwf_evec:
            hwf_log = out_store.dev(will_write=True)
            # (a state-major output store is filled through a contiguous staging tensor)
            hwf = hwf_log.reshape(batch, nocc, n) if hwf_log.is_contiguous() else \
                torch.empty((batch, nocc, n), dtype=torch.complex128, device=self.device)
        flat = wfs.reshape(batch, nsta, n)
        # mesh points in chunks: the gathered occupied blocks and the position / eigenvector workspaces of a
        # chunk stay below ~2 GiB each (a [129, 129] mesh of the norb-499 slab would otherwise need 3 x 33 GB
        # of temporaries next to the 66 GB array)
        per = max(1, nocc * max(n, nocc) * 16)
        step = int(max(1, min(batch, (2 << 30) // per)))
        ws = self.workspace(self.lib.tbk_position_hwf_workspace(nocc, n, step))
        for a in range(0, batch, step):
            b = min(batch, a + step)
            ed = flat[a:b].index_select(1, occ_t).contiguous()                    # data movement only
            _lib.check(self.lib.tbk_position_hwf(_ptr(ed), b - a, nocc, n, _ptr(pos), _ptr(hwfc[a:b]),
                                                 _ptr(hwf[a:b]) if hwf is not None else ctypes.c_void_p(0), 1,
                                                 _ptr(ws), ws.numel(), self.stream()))
        if hwf_log is not None and not hwf_log.is_contiguous():
            hwf_log.copy_(hwf.reshape(hwf_log.shape))
        return hwfc.cpu().numpy().reshape(tuple(mesh) + (nocc,))

    def position_hwf(self, model, evec, dir, hwf_evec, orbital_basis):
        """tb_model.position_hwf batched (pythtb.py:2162-2279)."""
        torch = self.torch
        evec = np.ascontiguousarray(evec, dtype=complex)
        batch, nocc, n = evec.shape
        ed = self.to_dev(evec)
        pos = self.to_dev(self._pos(model, dir), np.float64)
        hwfc = torch.empty((batch, nocc), dtype=torch.float64, device=self.device)
        hwf = None
        if hwf_evec:
            hwf = torch.empty((batch, nocc, n if orbital_basis else nocc), dtype=torch.complex128, device=self.device)
        ws = self.workspace(self.lib.tbk_position_hwf_workspace(nocc, n, batch))
        _lib.check(self.lib.tbk_position_hwf(_ptr(ed), batch, nocc, n, _ptr(pos), _ptr(hwfc), _ptr(hwf),
                                             int(bool(orbital_basis)), _ptr(ws), ws.numel(), self.stream()))
        return hwfc.cpu().numpy(), (hwf.cpu().numpy() if hwf_evec else None)
