#!/bin/bash
# tuning sweep of mesh_small_kernel variants (resident CTAs per SM x rows per iteration)
OUT=gpurun_out/${1:-sweep}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for v in 0 1 2 3 4 5; do
  TBK_MESH_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > $OUT/bench_v$v.json 2>$OUT/bench_v$v.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_v$v.json"))
print("variant $v", d["stages"], d["value"], d["e2e"]["value"], d["check"])
PY
done
timeout 300 python bench.py --workload kane_mele --steps 50 --warmup 5 --no-cpu > $OUT/bench_km.json 2>$OUT/bench_km.err
python -c "import json;d=json.load(open('$OUT/bench_km.json'));print('km',d['stages'],d['check'])"
timeout 900 python profiles/bench_configs.py > $OUT/configs.json 2>$OUT/configs.err; cat $OUT/configs.json
