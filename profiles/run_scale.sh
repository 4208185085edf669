#!/bin/bash
# weak-scaling pass on an 8-GPU box (gpurun --gpus 8): multi-rank parity, bench at N = 1, 2, 4, 8, CTA spans at N = 8
TAG=${1:-scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; tail -3 $OUT/pytest_multi.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > $OUT/bench_n1.json 2>$OUT/bench_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 200 --warmup 5 > $OUT/bench_n$n.json 2>$OUT/bench_n$n.err
done
TBK_CTA_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 profiles/cta_trace.py 2>&1 | grep spans > $OUT/spans_n8.log
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open("$OUT/bench_n%d.json"%n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable", e); continue
    if n==1: base=d["value"]; be=d["e2e"]["value"]
    print("N=%d value %.3f G/s ms/step %.4f eff %.3f | e2e %.3f G/s eff %.3f"%(n,d["value"]/1e9,d["ms_per_step"],d["value"]/(n*base),d["e2e"]["value"]/1e9,d["e2e"]["value"]/(n*be)), d["stages"]["solve_on_grid_ms"], d["stages"]["berry_flux_ms"], d["check"])
PY
cat $OUT/spans_n8.log
