#!/bin/bash
# multi-GPU pass (gpurun --gpus N): the multi-rank parity worker, weak-scaling bench at 1..N ranks (headline only),
# one full bench line with the extras at N ranks, hetrd variant A/B on one GPU.
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_stream.py tests/test_slice_fill.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log; tail -6 $OUT/pytest_multi.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --extras none > $OUT/bench_n1.json 2>$OUT/bench_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 100 --warmup 5 --extras none > $OUT/bench_n$n.json 2>$OUT/bench_n$n.err
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
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_full_n$N.json 2>$OUT/bench_full_n$N.err; echo "full bench exit $?"; tail -5 $OUT/bench_full_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_full_n$N.json").read().strip().splitlines()[-1])
    print("wall %.0f s value %.3f G"%(d["bench_wall_s"], d["value"]/1e9))
    for k,v in list(d["workloads"].items())+list(d["configs"].items()):
        print(k, {kk: v.get(kk) for kk in ("value","skipped","error","stages","check")})
except Exception as e:
    print("no full line", e)
PY
for V in full sym; do
  TBK_HETRD=$V timeout 300 python profiles/prof_blocked.py > $OUT/prof_blocked_$V.json 2> $OUT/prof_blocked_$V.err
  python - <<PY
import json
d=json.load(open("$OUT/prof_blocked_$V.json"))
for k,v in d.items(): print("$V", k, "%.2f ms/matrix/CTA"%v["ms_per_matrix_per_cta"], v["share"], "fallbacks", v["fallbacks"])
PY
done
ls -la $OUT; du -sh gpurun_out
