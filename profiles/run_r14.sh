#!/bin/bash
# r14: full parity suite, smoke, bench (all workloads + configs) and the reference arm at HEAD; every step under a timeout
OUT=gpurun_out/${1:-r14}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"], "wfs_to_host ms", d["e2e_wfs_to_host"]["ms_per_step"], "wall", d.get("bench_wall_s"))
print("kane_mele", d["workloads"]["kane_mele"]["value"], d["workloads"]["kane_mele"]["stages"])
for k,v in d["configs"].items():
    print(k, v.get("value"), v.get("stages"), v["roofline"].get("frac"), v.get("check",{}).get("ok"))
PY
