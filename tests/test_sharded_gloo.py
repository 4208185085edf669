"""Multi-rank host logic on CPU: world_size 2 and 3 over gloo.  The per-rank
numerics come from the oracle engine (tests/oracle_api.py); what is under test
is the product's sharding code in pythtb_b200/wfarray.py — slab partition,
halo ring shift / recomputation, cross-rank reductions and gathers."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_wf_array_over_gloo(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
        assert "rank %d ok" % rank in out


def test_shard_partition_covers_mesh():
    from pythtb_b200.wfarray import _Shard
    for n0 in (5, 14, 1025, 8193):
        for world in (1, 2, 3, 4, 8):
            if n0 - 1 < world:
                continue
            rows = []
            for r in range(world):
                s = _Shard(r, world, n0)
                rows.extend(range(s.row0, s.row0 + s.nrows))
            assert rows == list(range(n0 - 1))
