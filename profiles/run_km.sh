#!/bin/bash
# Kane-Mele (n = 4) mesh kernel: parity of every n <= 4 path + bench at the three occupancy variants
OUT=gpurun_out/${1:-km1}
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "not config_scale and not large" > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for V in 0 1 2; do
  TBK_MESH_VARIANT4=$V timeout 200 python bench.py --workload kane_mele --no-cpu --extras none --steps 40 --warmup 5 > $OUT/bench_km_v$V.json 2> $OUT/bench_km_v$V.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_km_v$V.json").read().strip().splitlines()[-1])
    print("variant $V", "value %.3g"%d["value"], "us/step %.1f"%(d["ms_per_step"]*1e3), d["stages"], d["check"].get("plaquettes_vs_oracle_max_dev"), d["check"].get("chern"))
except Exception as e: print("variant $V failed", e)
PY
done
