#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 k-mesh engine (BASELINE.json configs[1]).

A *step* is one pass of the hot path over one k-mesh:

    wf_array(model, [R*1024+1, 1025]).solve_on_grid([-1/2,-1/2])   (H build + eigh of every k-point,
                                                                   periodic images, min direct gaps)
    .berry_flux([0])                                               (one plaquette phase per k-point,
                                                                   summed -> Chern number)

on the Haldane model (delta=0, examples/haldane_bp.py:14-41 of the reference), R = number of GPUs
(weak scaling: every rank owns 1024 mesh rows of a (R*1024) x 1024 mesh).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload haldane|kane_mele]
                    [--extras all|none|kane_mele,3,4,5] [--budget-s S]

prints ONE JSON line (rank 0).  The top-level numbers are the Haldane headline; the same line carries
`workloads.kane_mele` (the other half of configs[1], same mesh, same measurement) and `configs.3/4/5`
(bench_extras.py: silicon 256^3, BN ribbon 1e5 k-points streamed, cubic slab [129, 129]) at the named shapes,
each with its own roofline and — at N = 1 — cpu_baseline; `fp64_peaks` are DFMA / DMMA microkernels timed in this
run (the FP64 roofline denominators).  Extras are skipped (and say so) when the time budget is used up.  `value` is k-points/s through build+eigh+flux with everything
resident in HBM, timed with CUDA events; `e2e` is the same step through the public PythTB-style
API (host arguments in, host numpy/float results out, every step); `roofline` is the dominant
kernel against the measured HBM peak; `cpu_baseline` is the numpy oracle (a vectorised port of the
reference's numpy/LAPACK path) on this box's host cores.  `--impl reference` times that CPU path
alone.  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import sys
import threading
import time

# the CPU legs run one process per host core: BLAS must not start a thread pool of its own inside every one of them
for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS_PER_RANK = 1024
COLS = 1024
START_K = [-0.5, -0.5]
METRIC = "k-points/sec (H build + eigh + Berry-flux plaquette), 1024x1024 mesh per GPU"
UNIT = "k-points/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the solve kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json, written from profiles/rNN/raw_*.csv); None if absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d[workload]["solve_kernel"]["dram_bytes"]
    except Exception:
        return None


def _build_model(mod, workload):
    from tests import models as M
    if workload == "haldane":
        return M.haldane(mod, delta=0.0), [0]
    if workload == "kane_mele":
        return M.kane_mele(mod, "odd"), [0, 1]
    raise SystemExit("unknown workload " + workload)


# ----------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle on the host cores
# ----------------------------------------------------------------------------------------------
def _ref_chunk(args):
    """Rows [r0, r1) of the mesh (plus the closing row r1) through the oracle: returns
    (sum of plaquette phases, min direct gaps).  Solving row r1 directly instead of copying it
    gives the same plaquette phases (they are gauge invariant per plaquette)."""
    workload, nrows_total, ncols, r0, r1 = args
    from oracle import pythtb_oracle as orc
    from tests import oracle_api
    model, occ = _build_model(oracle_api, workload)
    i = np.arange(r0, r1 + 1, dtype=float)
    j = np.arange(0, ncols + 1, dtype=float)
    k0 = START_K[0] + i / float(nrows_total)
    k1 = START_K[1] + j / float(ncols)
    kk = np.stack(np.meshgrid(k0, k1, indexing="ij"), axis=-1).reshape(-1, 2)
    ev, vec = orc.sol_ham(orc.gen_ham(model, kk), True)
    n = model._nsta
    wfs = vec.reshape(len(i), len(j), n, n)
    plaq = orc.one_flux_plane(wfs[:, :, occ])
    gaps = (ev[:, 1:] - ev[:, :-1]).reshape(len(i), len(j), n - 1)[:-1, :-1].min(axis=(0, 1))
    return float(plaq.sum()), gaps


def _chunks(workload, rows, procs):
    per = max(1, -(-rows // (2 * procs)))            # two chunks per worker: some load balance
    out, r = [], 0
    while r < rows:
        out.append((workload, ROWS_PER_RANK, COLS, r, min(rows, r + per)))
        r += per
    return out


def _cpu_rate(workload, rows, procs, min_seconds=8.0, max_passes=200):
    """k-points/s of the oracle on a rows x COLS slab of the mesh using `procs` processes: passes are
    repeated for at least `min_seconds`; returns (best rate, mean rate, passes, seconds, flux, min gaps)."""
    import multiprocessing as mp
    chunks = _chunks(workload, rows, procs)
    ctx = mp.get_context("fork")
    times = []
    flux = gaps = None
    with ctx.Pool(procs) as pool:
        pool.map(_ref_chunk, chunks)                  # warm the workers (imports, BLAS init)
        t_all = time.perf_counter()
        while len(times) < max_passes and (time.perf_counter() - t_all) < min_seconds:
            t0 = time.perf_counter()
            res = pool.map(_ref_chunk, chunks)
            times.append(time.perf_counter() - t0)
            flux = sum(x[0] for x in res)
            gaps = np.min(np.array([x[1] for x in res]), axis=0)
    n = rows * COLS
    return n / min(times), n * len(times) / sum(times), len(times), sum(times), flux, gaps


def run_reference(args):
    """The reference's CPU path (numpy oracle port: same numpy/LAPACK calls per k-point, vectorised
    over the mesh, all host cores through a process pool), same metric / workload as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    import multiprocessing as mp
    rows = ROWS_PER_RANK
    chunks = _chunks(args.workload, rows, procs)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        pool.map(_ref_chunk, chunks)
        pool.map(_ref_chunk, chunks)
        est = (time.perf_counter() - t0) / 2
        total_steps = args.steps + args.warmup
        if est * total_steps > 150.0:                 # keep the whole run inside a few minutes whatever K is
            rows = max(2 * procs, int(rows * 150.0 / (est * total_steps)))
            chunks = _chunks(args.workload, rows, procs)
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_chunk, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pool.map(_ref_chunk, chunks)
        dt = time.perf_counter() - t0
    value = rows * COLS * args.steps / dt
    sample = "%d rows x %d cols of the 1024x1024 mesh per step (numpy oracle, %d processes, os.cpu_count()=%d)" % (
        rows, COLS, procs, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": _config(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _workload_name(workload, nranks):
    return "%s %dx%d k-mesh: wf_array.solve_on_grid + berry_flux (configs[1])" % (
        {"haldane": "Haldane(delta=0)", "kane_mele": "Kane-Mele(odd)"}[workload], nranks * ROWS_PER_RANK, COLS)


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
SPIN_CYCLES = 1600000     # ~0.8 ms device-side spin before every timed step: the host is far ahead of the device when
                          # the timed launches are enqueued (8 Python ranks share the box's cores), so that the events
                          # time kernels — and the solve -> flux overlap — and not the launch path


def _config(workload, world, halo=None, reduction=None):
    """The `config` dict: the same keys from both arms (the reference arm states what does not apply)."""
    n = {"haldane": 2, "kane_mele": 4}[workload]
    return {"workload": _workload_name(workload, world), "arithmetic": "complex128 = pairs of f64", "norb": 2,
            "nspin": n // 2, "occ": {"haldane": [0], "kane_mele": [0, 1]}[workload], "mesh_per_gpu": [ROWS_PER_RANK, COLS],
            "l2": "256 MiB flush between steps (untimed); CUDA events per step, host kept ahead of the device by a device-side spin"
                  + ("; device-side barrier over the ranks before every timed step (untimed)" if world > 1 else ""),
            "parallelism": "mesh rows sliced over %d GPU(s)" % world, "halo": halo, "cross_rank_reduction": reduction}


def _fp64_peaks(eng):
    """DFMA and DMMA (mma.sync.m8n8k4.f64) peak of this GPU, measured now: TFLOP/s of a register-resident kernel
    that issues nothing but independent FMA chains / tensor-core FP64 MMAs (csrc/tbk_api.cu)."""
    import ctypes
    import torch
    from pythtb_b200 import _lib
    sink = torch.zeros(8, dtype=torch.float64, device=eng.device)
    out = {}
    for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops")):
        iters = 1 << 15
        flops = float(eng.lib.tbk_bench_fp64_flops(kind, iters))
        best = None
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(eng.lib.tbk_bench_fp64(kind, iters, ctypes.c_void_p(sink.data_ptr()), eng.stream()))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep >= 1:
                best = ms if best is None else min(best, ms)
        out[name] = flops / (best * 1e-3) / 1e12
    out["how"] = "148 x 4 CTAs x 256 threads, 16 independent DFMA chains per thread / 8 independent m8n8k4 DMMA per warp, 32768 iterations, best of 5 (CUDA events)"
    return out


def _oracle_plaquettes(model, occ, nrows_total, rows, cols):
    """Plaquette phases (pythtb.py:3840-3865) of the plaquettes (rows[i], cols[i]) of the (nrows_total+1) x (COLS+1) mesh
    through the oracle: the four corners are solved directly (a periodic image differs from the directly solved
    point by a gauge, which the plaquette phase does not see)."""
    from oracle import pythtb_oracle as orc
    n = model._nsta
    di = np.array([0, 1])
    k0 = START_K[0] + (rows[:, None] + di[None, :]) / float(nrows_total)          # [s, 2]
    k1 = START_K[1] + (cols[:, None] + di[None, :]) / float(COLS)
    kk = np.stack([np.repeat(k0[:, :, None], 2, axis=2), np.repeat(k1[:, None, :], 2, axis=1)], axis=-1)   # [s, 2, 2, 2]
    ev, vec = orc.sol_ham(orc.gen_ham(model, kk.reshape(-1, 2)), True)
    wfs = vec.reshape(len(rows), 2, 2, n, n)
    return np.array([orc.one_flux_plane(wfs[s][:, :, occ])[0, 0] for s in range(len(rows))])


def _oracle_gaps(model, nrows_total, row_lo, row_hi, stride):
    from oracle import pythtb_oracle as orc
    i = np.arange(row_lo, row_hi, stride, dtype=float)
    j = np.arange(0, COLS, stride, dtype=float)
    kk = np.stack(np.meshgrid(START_K[0] + i / float(nrows_total), START_K[1] + j / float(COLS), indexing="ij"), axis=-1).reshape(-1, 2)
    ev = orc.sol_ham(orc.gen_ham(model, kk), False)
    return (ev[:, 1:] - ev[:, :-1]).min(axis=0)


def measure_mesh(tb, eng, args, workload, world, rank, local, with_cpu, cpu_seconds=8.0):
    """The configs[1] measurement for one model: value / stages / e2e / roofline / check (/ cpu_baseline)."""
    import ctypes
    import torch
    import torch.distributed as dist
    from pythtb_b200 import _lib
    lib = eng.lib
    model, occ = _build_model(tb, workload)
    n = model._nsta
    mesh = [world * ROWS_PER_RANK + 1, COLS + 1]
    shard = (rank, world) if world > 1 else None
    w = tb.wf_array(model, mesh, shard=shard) if shard else tb.wf_array(model, mesh)
    kpts_per_step_rank = ROWS_PER_RANK * COLS
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)
    steps = args.steps
    pipelined = world > 1

    def flush_l2():
        torch.cuda._sleep(SPIN_CYCLES)
        _lib.check(lib.tbk_flush_l2(ctypes.c_void_p(flush.data_ptr()), flush.numel(), eng.stream()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step: the two engine calls, results stay on the device.  N > 1: both cross-rank
    # reductions (min of the gaps, sum of the flux) are POSTED by the producing kernels over NVLink and completed by
    # the flux kernel of the NEXT step (csrc/tbk_peer.cuh) — a sweep that reads its results at the end never waits
    # for a peer inside a step; the last step's reductions are flushed inside its own timed window.
    def step_device(last=False):
        gaps = w._solve_on_grid_device(START_K, defer_reduce=pipelined)
        flux = w._berry_flux_device(occ, defer_reduce=pipelined)
        if last and pipelined:
            eng.peer_flush()
        return gaps, flux

    def step_e2e():
        gaps = w.solve_on_grid(START_K)
        flux = w.berry_flux(occ)
        return gaps, flux

    for _ in range(max(3, args.warmup)):
        flush_l2()
        step_device()
    eng.peer_flush()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches0 = eng.launches
    step_device()
    launches = eng.launches - launches0           # kernels of libtbk_b200.so per step (counted inside the library)
    eng.peer_flush()
    barrier()
    for s in range(steps):
        flush_l2()
        eng.peer_barrier()        # N > 1: every timed step starts together on all ranks (device-side, untimed)
        ev0[s].record()
        out = step_device(last=(s == steps - 1))
        ev1[s].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    gaps_d, flux_d = out
    gaps_dev_host = gaps_d.cpu().numpy() if hasattr(gaps_d, "cpu") else np.asarray(gaps_d)
    flux_val = float(flux_d) if not hasattr(flux_d, "cpu") else float(flux_d.cpu().reshape(-1)[0])

    # ---- the two kernels alone (after an L2 flush each) for the roofline / stage split
    def alone(fn):
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for s in range(steps):
            flush_l2()
            e0[s].record()
            fn()
            e1[s].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / steps
    k_ms = alone(lambda: w._solve_on_grid_device(START_K, want_gaps=False))
    solve_kernel = w._last_solve_kernel()
    f_ms = alone(lambda: w._berry_flux_device(occ, local_only=True))

    # ---- `e2e`: K steps through the public API (each returns host values -> synchronous)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        gaps_h, flux_h = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # ---- the same step when the caller also wants the whole eigenvector array on the host (the reference keeps
    # `_wfs` in host memory): solve + flux + D2H of the local `_wfs` slab into its pinned mirror, every step
    nwf = max(3, min(20, steps))
    w._wfs                                            # allocate the pinned mirror outside the timed region
    barrier()
    t0 = time.perf_counter()
    for s in range(nwf):
        step_e2e()
        w._store.state = "device"
        w._wfs                                        # device -> pinned host copy of the whole slab
    torch.cuda.synchronize()
    e2e_wfs_s = (time.perf_counter() - t0) / nwf
    w._store.state = "device"
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- parity inside the bench: N > 1 has no full-mesh oracle beside it, so every rank checks 64 random
    # plaquettes of its slab and a sub-mesh gap bound against the oracle, and the fused in-kernel reductions
    # against NCCL reductions of the ranks' local results
    check = {}
    w._solve_on_grid_device(START_K, want_gaps=False)
    tot_loc, plq = eng.flux_device(w._store, 2, occ, [0, 1], want_total=True, want_plaq=True)
    rng = np.random.RandomState(7 + rank)
    nloc = ROWS_PER_RANK if world == 1 else w._shard.nrows
    row0 = 0 if world == 1 else w._shard.row0
    ri, ci = rng.randint(0, nloc, 64), rng.randint(0, COLS, 64)
    ri[:4], ci[:4] = [0, nloc - 1, 0, nloc - 1], [0, 0, COLS - 1, COLS - 1]       # slab corners: halo row / periodic images
    got = plq[0][torch.as_tensor(ri, device=eng.device), torch.as_tensor(ci, device=eng.device)].cpu().numpy()
    ref = _oracle_plaquettes(model, occ, world * ROWS_PER_RANK, row0 + ri, ci)
    pdev = float(np.max(np.abs((got - ref + np.pi) % (2 * np.pi) - np.pi)))
    sub = _oracle_gaps(model, world * ROWS_PER_RANK, row0, row0 + nloc, 16)
    loc_gaps = eng.solve_grid(model, w._store, w._mesh_arr, np.array(START_K), row0=row0, nrows=nloc,
                              wrap0=(2 if world > 1 else 1)).cpu().numpy()
    stats = torch.tensor([pdev] + list(-sub) + list(-loc_gaps) + [float(tot_loc.reshape(-1)[0])], dtype=torch.float64, device=eng.device)
    if world > 1:
        mx = stats[:1 + 2 * (n - 1)].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats[-1:].clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        stats = torch.cat([mx, sm])
    st = stats.cpu().numpy()
    sub_min, nccl_gaps, nccl_flux = -st[1:n], -st[n:2 * n - 1], float(st[-1])
    check["plaquettes_vs_oracle_max_dev"] = float(st[0])
    check["plaquettes_checked"] = 64 * world
    check["gaps_below_oracle_submesh"] = bool(np.all(gaps_dev_host <= sub_min + 1e-12) and np.all(sub_min - gaps_dev_host < 0.05))
    check["fused_gaps_equal_nccl"] = bool(np.array_equal(gaps_dev_host, nccl_gaps))
    check["fused_flux_vs_nccl"] = float(abs(flux_val - nccl_flux))

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, k_ms, f_ms, e2e_wfs_s], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, k_ms, f_ms, e2e_wfs_s = [float(x) for x in t.cpu()]
    if rank != 0:
        return None
    hbm_peak, peak_src = _peaks()
    chern = flux_val / (2 * np.pi)
    total_k = kpts_per_step_rank * world
    # algorithmic bytes of the solve kernel: every stored eigenvector block once (n*n complex128 per mesh point incl.
    # periodic images); k generated on device, H never leaves the SM.
    solve_bytes = (ROWS_PER_RANK + 1) * (COLS + 1) * n * n * 16
    flux_bytes = (ROWS_PER_RANK + 1) * (COLS + 1) * len(occ) * n * 16
    check.update({"chern": chern, "chern_is_integer": bool(abs(chern - round(chern)) < 1e-9),
                  "e2e_flux_equal": bool(abs(float(flux_h) - flux_val) < 1e-9),
                  "e2e_gaps_equal": bool(np.array_equal(np.asarray(gaps_h), gaps_dev_host))})
    res = {
        "value": total_k * steps / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / steps,
        "config": _config(workload, world, halo=(w._halo_mode() if world > 1 else None),
                          reduction=("posted in-kernel over NVLink peer memory, completed by the next step's flux kernel (last step: flushed inside its window)"
                                     if (world > 1 and eng._peer) else ("nccl all_reduce" if world > 1 else None))),
        "stages": {"solve_on_grid_ms": k_ms, "berry_flux_ms": f_ms,
                   "kpoints_per_s_solve": kpts_per_step_rank * world / (k_ms * 1e-3),
                   "plaquettes_per_s_flux": kpts_per_step_rank * world / (f_ms * 1e-3)},
        "check": check,
        "e2e": {"value": total_k * steps / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / steps,
                "h2d_bytes_per_step": 8 * len(START_K) + 4 * len(mesh), "d2h_bytes_per_step": 8 * (n - 1) + 8,
                "note": "public API wf_array.solve_on_grid + berry_flux; eigenvectors stay in HBM (lazy host mirror), "
                        "results (gaps, flux) are copied to the host every step (N > 1: two synchronous in-kernel reductions per step)"},
        "e2e_wfs_to_host": {"value": total_k / e2e_wfs_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_wfs_s,
                            "d2h_bytes_per_step": (ROWS_PER_RANK + 1) * (COLS + 1) * n * n * 16 + 8 * (n - 1) + 8,
                            "note": "as e2e, plus a device->pinned-host copy of the whole local _wfs slab every step "
                                    "(what a caller pays who, like the reference, wants the eigenvectors in host memory)"},
        "gpu_launches": launches * steps, "gpu_launches_per_step": launches,
        "roofline": {"bound": "hbm", "kernel": solve_kernel, "achieved": solve_bytes / (k_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": solve_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                     "traffic": _ncu_traffic(workload), "peak_source": peak_src, "algorithmic_bytes_per_launch": solve_bytes,
                     "flux_kernel": {"achieved": flux_bytes / (f_ms * 1e-3) / 1e9, "frac": flux_bytes / (f_ms * 1e-3) / 1e9 / hbm_peak,
                                     "algorithmic_bytes_per_launch": flux_bytes}},
        "clocks": sampler.summary(),
    }
    if with_cpu:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        cores = os.cpu_count() or 1
        procs = max(1, min(cores, 64))
        best, mean, passes, secs, cflux, cgaps = _cpu_rate(workload, ROWS_PER_RANK, procs, min_seconds=cpu_seconds)
        res["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": procs, "kind": "port", "mean_value": mean,
                               "sample": "full 1024x1024 mesh, %d passes in %.1f s (best pass reported), numpy "
                                         "oracle over %d processes, os.cpu_count()=%d" % (passes, secs, procs, cores),
                               "chern": cflux / (2 * np.pi)}
        # the full-mesh oracle is right here: total flux and minimal gaps of the whole 1024 x 1024 mesh
        res["check"]["flux_vs_full_mesh_oracle"] = float(abs((flux_val - cflux + np.pi) % (2 * np.pi) - np.pi))
        res["check"]["gaps_vs_full_mesh_oracle"] = float(np.max(np.abs(gaps_dev_host - cgaps)))
    del w, flush
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    t_start = time.perf_counter()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pythtb_b200 as tb
    from pythtb_b200 import _engine
    import bench_extras as X

    eng = _engine.get_engine()
    peaks = _fp64_peaks(eng)
    peaks["hbm_gbs"], peaks["hbm_source"] = _peaks()
    with_cpu = world == 1 and not args.no_cpu
    line = measure_mesh(tb, eng, args, args.workload, world, rank, local, with_cpu)
    extras = ["kane_mele", "3", "4", "5"] if args.extras == "all" else ([] if args.extras == "none" else args.extras.split(","))
    if args.workload != "haldane":
        extras = [e for e in extras if e != "kane_mele"]

    def left():
        # every rank must take the same decision: rank 0's clock
        t = torch.tensor([args.budget_s - (time.perf_counter() - t_start)], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.broadcast(t, 0)
        return float(t.cpu()[0])

    workloads, configs = {}, {}
    # (name, estimated seconds at N = 1, runner)
    plan = []
    if "kane_mele" in extras:
        plan.append(("kane_mele", 14.0, lambda: measure_mesh(tb, eng, args, "kane_mele", world, rank, local, with_cpu, cpu_seconds=4.0)))
    if "3" in extras:
        plan.append(("3", 16.0, lambda: X.config3(tb, eng, world, rank, peaks, with_cpu)))
    if "4" in extras:
        plan.append(("4", 16.0, lambda: X.config4(tb, eng, world, rank, peaks, with_cpu, ncell=100)))
    if "5" in extras:
        plan.append(("5", 40.0, lambda: X.config5(tb, eng, world, rank, peaks, with_cpu, budget_s=max(20.0, left() - 15.0))))
    if "4" in extras:
        plan.append(("4_norb400", 45.0, lambda: X.config4(tb, eng, world, rank, peaks, with_cpu, ncell=200)))
    for name, est, fn in plan:
        rem = left()
        if rem < est / max(1, world) + 5.0:
            rec = {"skipped": "time budget (%.0f s left of --budget-s %.0f, needs ~%.0f s)" % (rem, args.budget_s, est / max(1, world))}
        else:
            try:
                rec = fn()
            except Exception as e:                      # an extra must never cost the headline line
                import traceback
                rec = {"error": "%s: %s" % (type(e).__name__, str(e).strip()[:300]), "trace": traceback.format_exc()[-600:]}
                if world > 1:
                    raise
        (workloads if name == "kane_mele" else configs)[name] = rec
    if rank == 0:
        out = {"metric": METRIC, "value": line["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": line["ms_per_step"], "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        for key in ("config", "stages", "check", "e2e", "e2e_wfs_to_host", "gpu_launches", "gpu_launches_per_step",
                    "roofline", "clocks", "cpu_baseline"):
            if key in line:
                out[key] = line[key]
        out["fp64_peaks"] = peaks
        out["workloads"] = workloads
        out["configs"] = configs
        out["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="haldane", choices=["haldane", "kane_mele"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--extras", default="all", help="all | none | comma list of kane_mele,3,4,5")
    ap.add_argument("--budget-s", type=float, default=210.0, help="wall-clock budget of the whole run; extras that do not fit are skipped")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
