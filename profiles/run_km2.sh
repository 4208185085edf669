#!/bin/bash
OUT=gpurun_out/${1:-km2}
mkdir -p $OUT
for L in default fcsmem; do
for V in 0 2; do
  if [ $L = default ]; then unset PYTHTB_B200_LIB; else export PYTHTB_B200_LIB=$PWD/profiles/ab/libtbk_$L.so; fi
  TBK_MESH_VARIANT4=$V timeout 150 python bench.py --workload kane_mele --no-cpu --extras none --steps 40 --warmup 5 > $OUT/bench_km_${L}_v$V.json 2> $OUT/bench_km_${L}_v$V.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_km_${L}_v$V.json").read().strip().splitlines()[-1])
    print("$L variant $V", "us/step %.1f"%(d["ms_per_step"]*1e3), "solve %.1f us"%(d["stages"]["solve_on_grid_ms"]*1e3), d["check"].get("plaquettes_vs_oracle_max_dev"))
except Exception as e: print("$L variant $V failed", e)
PY
done; done
