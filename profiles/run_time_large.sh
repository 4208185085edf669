#!/bin/bash
OUT=gpurun_out/${1:-tl}
mkdir -p $OUT
timeout 400 python profiles/time_large.py > $OUT/time_large.json 2> $OUT/time_large.err; tail -2 $OUT/time_large.err
python - <<PY
import json
d=json.loads(open("$OUT/time_large.json").read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict): print(k, {kk: round(vv,2) for kk,vv in v.items() if kk.endswith("_ms") or kk.endswith("tflops")})
PY
timeout 600 python -m pytest tests/test_gpu_config_scale.py tests/test_gpu_parity.py -m gpu -q -x -k "wilson or position or slab or ribbon or config_scale" > $OUT/pytest_large.log 2>&1; tail -3 $OUT/pytest_large.log
