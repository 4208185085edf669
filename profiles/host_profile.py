# cProfile of the e2e step on the GPU box (single GPU, or sharded under torchrun: rank 0 prints)
import cProfile, pstats, sys, io, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
import pythtb_b200 as tb
from tests import models as M
m = M.haldane(tb, 0.0)
w = tb.wf_array(m, [1024 * world + 1, 1025], shard=(rank, world)) if world > 1 else tb.wf_array(m, [1025, 1025])
for _ in range(20):
    w.solve_on_grid([-0.5, -0.5]); w.berry_flux([0])
N = 2000
t0 = time.perf_counter(); ts = tf = 0.0
for _ in range(N):
    a = time.perf_counter(); w.solve_on_grid([-0.5, -0.5]); b = time.perf_counter(); w.berry_flux([0]); c = time.perf_counter()
    ts += b - a; tf += c - b
if rank == 0:
    print("world %d: solve_on_grid %.2f us, berry_flux %.2f us per call" % (world, 1e6 * ts / N, 1e6 * tf / N))
pr = cProfile.Profile(); pr.enable()
for _ in range(N):
    w.solve_on_grid([-0.5, -0.5]); w.berry_flux([0])
pr.disable()
if rank == 0:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:5000])
if world > 1:
    dist.barrier(); dist.destroy_process_group()
