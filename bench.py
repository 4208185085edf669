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

prints ONE JSON line (rank 0).  `value` is k-points/s through build+eigh+flux with everything
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

import numpy as np

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
    repeated for at least `min_seconds`; returns (best rate, mean rate, passes, seconds, flux)."""
    import multiprocessing as mp
    chunks = _chunks(workload, rows, procs)
    ctx = mp.get_context("fork")
    times = []
    flux = None
    with ctx.Pool(procs) as pool:
        pool.map(_ref_chunk, chunks)                  # warm the workers (imports, BLAS init)
        t_all = time.perf_counter()
        while len(times) < max_passes and (time.perf_counter() - t_all) < min_seconds:
            t0 = time.perf_counter()
            res = pool.map(_ref_chunk, chunks)
            times.append(time.perf_counter() - t0)
            flux = sum(x[0] for x in res)
    n = rows * COLS
    return n / min(times), n * len(times) / sum(times), len(times), sum(times), flux


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
        "config": {"workload": _workload_name(args.workload, 1), "sample": sample},
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
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pythtb_b200 as tb
    from pythtb_b200 import _engine, _lib

    eng = _engine.get_engine()
    lib = eng.lib
    model, occ = _build_model(tb, args.workload)
    n = model._nsta
    mesh = [world * ROWS_PER_RANK + 1, COLS + 1]
    shard = (rank, world) if world > 1 else None
    w = tb.wf_array(model, mesh, shard=shard) if shard else tb.wf_array(model, mesh)
    kpts_per_step_rank = ROWS_PER_RANK * COLS
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)

    def flush_l2():
        # A ~150 us spin first, so that the host is always ahead of the device when the timed
        # launches are enqueued (a step is only tens of microseconds of device time: without it the
        # events would measure the Python launch path, not the kernels); then evict L2.
        torch.cuda._sleep(300000)
        _lib.check(lib.tbk_flush_l2(ctypes_ptr(flush), flush.numel(), eng.stream()))

    import ctypes

    def ctypes_ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step: the two engine calls, results stay on the device
    def step_device():
        # N > 1: the gap minimum over the ranks is posted by the solve kernel and completed inside the flux
        # kernel, whose own exchange carries both (one exposed NVLink round trip per step instead of two)
        gaps = w._solve_on_grid_device(START_K, defer_reduce=world > 1)
        flux = w._berry_flux_device(occ)
        return gaps, flux

    # ---- end-to-end step: public API, host arguments in, host results out
    def step_e2e():
        gaps = w.solve_on_grid(START_K)
        flux = w.berry_flux(occ)
        return gaps, flux

    for _ in range(max(3, args.warmup)):
        flush_l2()
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- `value`: K device-resident steps, CUDA events around every step (L2 flush in between, untimed)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = eng.launches
    step_device()
    launches = eng.launches - launches0           # kernels of libtbk_b200.so per step (counted inside the library)
    barrier()
    for s in range(args.steps):
        flush_l2()
        eng.peer_barrier()        # N > 1: every timed step starts together on all ranks (device-side, untimed)
        ev0[s].record()
        out = step_device()
        ev1[s].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    gaps_d, flux_d = out
    gaps_dev_host = gaps_d.cpu().numpy() if hasattr(gaps_d, "cpu") else np.asarray(gaps_d)
    flux_val = float(flux_d) if not hasattr(flux_d, "cpu") else float(flux_d.cpu().reshape(-1)[0])

    # ---- dominant kernel alone (solve_on_grid's fused assemble+eigh+pbc kernel) for the roofline
    kev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for s in range(args.steps):
        flush_l2()
        kev0[s].record()
        w._solve_on_grid_device(START_K, want_gaps=False)
        kev1[s].record()
    torch.cuda.synchronize()
    k_ms = sum(a.elapsed_time(b) for a, b in zip(kev0, kev1)) / args.steps
    fev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    fev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for s in range(args.steps):
        flush_l2()
        fev0[s].record()
        w._berry_flux_device(occ, local_only=True)
        fev1[s].record()
    torch.cuda.synchronize()
    f_ms = sum(a.elapsed_time(b) for a, b in zip(fev0, fev1)) / args.steps

    # ---- `e2e`: K steps through the public API (each returns host values -> synchronous)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        gaps_h, flux_h = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # ---- the same step when the caller also wants the whole eigenvector array on the host (the reference keeps
    # `_wfs` in host memory): solve + flux + D2H of the local `_wfs` slab into its pinned mirror, every step
    nwf = max(3, min(20, args.steps))
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

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, k_ms, f_ms, e2e_wfs_s], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, k_ms, f_ms, e2e_wfs_s = [float(x) for x in t.cpu()]

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        chern = flux_val / (2 * np.pi)
        total_k = kpts_per_step_rank * world
        # algorithmic bytes of the solve kernel: every stored eigenvector block once (n*n complex128
        # per mesh point incl. periodic images); k generated on device, H never leaves the SM.
        solve_bytes = (ROWS_PER_RANK + 1) * (COLS + 1) * n * n * 16
        flux_bytes = (ROWS_PER_RANK + 1) * (COLS + 1) * len(occ) * n * 16
        line = {
            "metric": METRIC, "value": total_k * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": _workload_name(args.workload, world), "arithmetic": "complex128 = pairs of f64", "norb": model._norb, "nspin": model._nspin,
                       "occ": occ, "mesh_per_gpu": [ROWS_PER_RANK, COLS], "l2": "256 MiB flush between steps (untimed); events per step, host kept ahead of the device"
                             + ("; device-side barrier over the ranks before every timed step (untimed)" if world > 1 else ""),
                       "parallelism": "mesh rows sliced over %d GPU(s)" % world,
                       "halo": (w._halo_mode() if world > 1 else None),
                       "cross_rank_reduction": ("in-kernel over NVLink peer memory" if (world > 1 and eng._peer) else
                                                ("nccl all_reduce" if world > 1 else None))},
            "stages": {"solve_on_grid_ms": k_ms, "berry_flux_ms": f_ms,
                       "kpoints_per_s_solve": kpts_per_step_rank * world / (k_ms * 1e-3),
                       "plaquettes_per_s_flux": kpts_per_step_rank * world / (f_ms * 1e-3)},
            "check": {"chern": chern, "chern_is_integer": bool(abs(chern - round(chern)) < 1e-9),
                      "e2e_flux_equal": bool(abs(float(flux_h) - flux_val) < 1e-9),
                      "e2e_gaps_equal": bool(np.array_equal(np.asarray(gaps_h), gaps_dev_host))},
            "e2e": {"value": total_k * args.steps / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "h2d_bytes_per_step": 8 * len(START_K) + 4 * len(mesh),
                    "d2h_bytes_per_step": 8 * (n - 1) + 8,
                    "note": "public API wf_array.solve_on_grid + berry_flux; eigenvectors stay in HBM "
                            "(lazy host mirror), results (gaps, flux) are copied to the host every step"},
            "e2e_wfs_to_host": {"value": total_k / e2e_wfs_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_wfs_s,
                                "d2h_bytes_per_step": (ROWS_PER_RANK + 1) * (COLS + 1) * n * n * 16 + 8 * (n - 1) + 8,
                                "note": "as e2e, plus a device->pinned-host copy of the whole local _wfs slab every step "
                                        "(what a caller pays who, like the reference, wants the eigenvectors in host memory)"},
            "gpu_launches": launches * args.steps,
            "gpu_launches_per_step": launches,
            "roofline": {"bound": "hbm", "kernel": w._last_solve_kernel(), "achieved": solve_bytes / (k_ms * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": solve_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                         "traffic": _ncu_traffic(args.workload), "peak_source": peak_src, "algorithmic_bytes_per_launch": solve_bytes,
                         "flux_kernel": {"achieved": flux_bytes / (f_ms * 1e-3) / 1e9,
                                         "frac": flux_bytes / (f_ms * 1e-3) / 1e9 / hbm_peak,
                                         "algorithmic_bytes_per_launch": flux_bytes}},
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu:
            os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, 64))
            best, mean, passes, secs, cflux = _cpu_rate(args.workload, ROWS_PER_RANK, procs)
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": procs, "kind": "port", "mean_value": mean,
                                    "sample": "full 1024x1024 mesh, %d passes in %.1f s (best pass reported), numpy "
                                              "oracle over %d processes, os.cpu_count()=%d" % (passes, secs, procs, cores),
                                    "chern": cflux / (2 * np.pi)}
        print(json.dumps(line))
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
