#!/usr/bin/env python
"""Per-CTA timeline of the two headline kernels (TBK_CTA_TRACE=1): where the launch ramp, the body and
the tail of the single wave go.  Usage: TBK_CTA_TRACE=1 python profiles/cta_trace.py [haldane|kane_mele]"""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TBK_CTA_TRACE"] = "1"
import torch
import pythtb_b200 as tb
from pythtb_b200 import _engine, _lib
from tests import models as M

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
wl = sys.argv[1] if len(sys.argv) > 1 else "haldane"
model, occ = (M.haldane(tb, delta=0.0), [0]) if wl == "haldane" else (M.kane_mele(tb, "odd"), [0, 1])
eng = _engine.get_engine()
w = tb.wf_array(model, [1024 * world + 1, 1025], shard=(rank, world)) if world > 1 else tb.wf_array(model, [1025, 1025])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)
CAP = 4096
buf = (ctypes.c_uint64 * (2 * CAP * 4))()


def grab(which=None):
    """Timeline of the last traced launch: the solve kernel's half of the buffer, the flux kernel's, or (which=None)
    whichever was launched last in the calls this script makes one at a time."""
    _lib.check(eng.lib.tbk_debug_cta_trace(buf, 2 * CAP, 1))
    a = np.ctypeslib.as_array(buf).reshape(2 * CAP, 4).astype(np.int64).copy()
    lo, hi = a[:CAP], a[CAP:]
    lo, hi = lo[lo[:, 2] > 0], hi[hi[:, 2] > 0]
    if which == "solve":
        return lo
    if which == "flux":
        return hi
    if which == "both":
        return lo, hi
    return hi if len(hi) else lo


def summarise(name, a):
    t0, t1 = a[:, 1].min(), a[:, 2].max()
    dur = a[:, 2] - a[:, 1]
    start = a[:, 1] - t0
    end = a[:, 2] - t0
    order = np.argsort(-end)
    per_sm = {}
    for sm, b, e, _ in a:
        s = per_sm.setdefault(int(sm), [b, e, 0]); s[0] = min(s[0], b); s[1] = max(s[1], e); s[2] += 1
    sm_end = np.array([v[1] - t0 for v in per_sm.values()])
    out = {"kernel": name, "ctas": int(len(a)), "sms": len(per_sm), "span_us": (t1 - t0) / 1e3,
           "cta_start_us": {"p50": float(np.median(start)) / 1e3, "max": float(start.max()) / 1e3},
           "cta_dur_us": {"min": float(dur.min()) / 1e3, "p50": float(np.median(dur)) / 1e3,
                          "p90": float(np.percentile(dur, 90)) / 1e3, "max": float(dur.max()) / 1e3},
           "cta_end_us": {"p10": float(np.percentile(end, 10)) / 1e3, "p50": float(np.median(end)) / 1e3,
                          "p90": float(np.percentile(end, 90)) / 1e3, "max": float(end.max()) / 1e3},
           "sm_end_us": {"min": float(sm_end.min()) / 1e3, "p50": float(np.median(sm_end)) / 1e3, "max": float(sm_end.max()) / 1e3},
           "ctas_per_sm": sorted(set(v[2] for v in per_sm.values())),
           "last_ctas(blockIdx,sm,start_us,dur_us)": [(int(a[i, 3]), int(a[i, 0]), round(start[i] / 1e3, 2), round(dur[i] / 1e3, 2)) for i in order[:12]],
           "first_ctas_to_end(blockIdx,sm,start_us,dur_us)": [(int(a[i, 3]), int(a[i, 0]), round(start[i] / 1e3, 2), round(dur[i] / 1e3, 2)) for i in order[-6:]]}
    return out


res = []
spans = []
for rep in range(12):
    eng.lib.tbk_flush_l2(ctypes.c_void_p(flush.data_ptr()), flush.numel(), eng.stream())
    eng.peer_barrier()
    w._solve_on_grid_device([-0.5, -0.5], defer_reduce=world > 1)
    torch.cuda.synchronize()
    a = grab()
    if world > 1:
        dist.barrier()
    eng.lib.tbk_flush_l2(ctypes.c_void_p(flush.data_ptr()), flush.numel(), eng.stream())
    eng.peer_barrier()
    w._berry_flux_device(occ)
    torch.cuda.synchronize()
    b = grab()
    if world > 1:
        dist.barrier()
    if rep >= 2:
        spans.append(((a[:, 2].max() - a[:, 1].min()) / 1e3, (b[:, 2].max() - b[:, 1].min()) / 1e3,
                      float(np.sort(b[:, 2])[-2] - b[:, 1].min()) / 1e3))
    if rep == 11:
        res = [summarise("mesh_small_kernel", a), summarise("flux_rows_kernel", b)]
# the flux kernel inside a step: launched right behind the grid solve (programmatic dependent launch, array still in L2)
instep = []
for rep in range(10):
    eng.lib.tbk_flush_l2(ctypes.c_void_p(flush.data_ptr()), flush.numel(), eng.stream())
    eng.peer_barrier()
    w._solve_on_grid_device([-0.5, -0.5], defer_reduce=world > 1)
    w._berry_flux_device(occ, defer_reduce=world > 1)
    torch.cuda.synchronize()
    a, b = grab("both")
    if world > 1:
        dist.barrier()
    if rep >= 2:
        t0 = a[:, 1].min()
        instep.append(((b[:, 2].max() - b[:, 1].min()) / 1e3, float(np.sort(b[:, 2])[-2] - b[:, 1].min()) / 1e3,
                       float(np.median(b[:, 2] - b[:, 1])) / 1e3,
                       # the step on one clock, relative to the first solve CTA: solve body end (all but the last CTA), solve end,
                       # first / median / last flux CTA start, flux body end, flux end
                       float(np.sort(a[:, 2])[-2] - t0) / 1e3, float(a[:, 2].max() - t0) / 1e3, float(b[:, 1].min() - t0) / 1e3,
                       float(np.median(b[:, 1]) - t0) / 1e3, float(b[:, 1].max() - t0) / 1e3, float(np.sort(b[:, 2])[-2] - t0) / 1e3,
                       float(b[:, 2].max() - t0) / 1e3))
med = np.round(np.median(np.array(instep), axis=0), 2)
print("rank %d world %d flux kernel inside a step, us (span, span without its last CTA, median CTA): %s" % (rank, world, med[:3]))
print("rank %d world %d step timeline us from the first solve CTA (solve body end, solve end, flux first/median/last CTA start, flux body end, flux end): %s" % (rank, world, med[3:]))
print("rank %d world %d spans us (mesh, flux, flux without its last CTA): %s" % (rank, world, np.round(np.median(np.array(spans), axis=0), 2)))
if rank == 0:
    print(json.dumps(res, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
