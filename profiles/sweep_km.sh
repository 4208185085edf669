#!/bin/bash
# Kane-Mele (n = 4) mesh kernel: resident-CTA variants of the direct small solver
OUT=gpurun_out/${1:-km}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for v in 0 1 2; do
  TBK_MESH_VARIANT4=$v timeout 300 python bench.py --workload kane_mele --steps 50 --warmup 5 --no-cpu > $OUT/bench_km_v$v.json 2>$OUT/bench_km_v$v.err
  python -c "
import json
d=json.load(open('$OUT/bench_km_v$v.json'))
print('variant4 $v', d['stages'], d['value'], d['check'])"
done
