#!/bin/bash
# staged (compact-WY on DMMA) vs legacy back-transformation: parity tests + timings
OUT=gpurun_out/${1:-wy1}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_config_scale.py -m gpu -q -x > $OUT/pytest_config_scale.log 2>&1; tail -5 $OUT/pytest_config_scale.log
timeout 400 python profiles/time_large.py > $OUT/time_large_staged.json 2> $OUT/time_large_staged.err; tail -2 $OUT/time_large_staged.err
TBK_BACKTR=legacy timeout 400 python profiles/time_large.py > $OUT/time_large_legacy.json 2> $OUT/time_large_legacy.err; tail -2 $OUT/time_large_legacy.err
python - <<PY
import json
for v in ("staged","legacy"):
    try:
        d=json.loads(open("$OUT/time_large_%s.json"%v).read().strip().splitlines()[-1])
        for k,x in d.items():
            if isinstance(x,dict): print(v, k, {kk: round(vv,2) for kk,vv in x.items() if kk.endswith("_ms") or kk.endswith("per_s")})
    except Exception as e: print(v, "failed", e)
PY
