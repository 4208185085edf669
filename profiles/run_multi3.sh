#!/bin/bash
# short multi-GPU pass: multi-rank parity worker + weak scaling of the headline at 1..N ranks
N=${1:-2}
TAG=${2:-multi3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log; tail -4 $OUT/pytest_multi.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --extras none > $OUT/bench_n1.json 2>$OUT/bench_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 100 --warmup 5 --extras none > $OUT/bench_n$n.json 2>$OUT/bench_n$n.err
    tail -2 $OUT/bench_n$n.err
  fi
done
python - <<PY
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open("$OUT/bench_n%d.json"%n).read().strip().splitlines()[-1])
    except Exception as e:
        continue
    if n==1: base=d["value"]
    print("N=%d value %.3f G/s ms/step %.4f e2e %.3f G/s eff %.3f"%(n,d["value"]/1e9,d["ms_per_step"],d["e2e"]["value"]/1e9,d["value"]/(n*base) if base else 0), d["stages"], d["check"])
PY
ls -la $OUT
