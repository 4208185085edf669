#!/bin/bash
# A/B of the front kernel's CTA shape (every step under a short timeout)
OUT=gpurun_out/${1:-shape1}
mkdir -p $OUT
timeout 60 python profiles/ring_smoke.py 2>&1 | tail -1
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 150 python profiles/time_large.py > $OUT/time_large_$tag.json 2> $OUT/time_large_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/time_large_$tag.json").read().strip().splitlines()[-1])
    for k,x in d.items():
        if isinstance(x,dict): print("$tag", k, {kk: round(vv,2) for kk,vv in x.items() if kk in ("solve_on_grid_ms","kpts_per_s","position_hwf_all_ms")})
except Exception as e: print("$tag", "failed", e)
PY
}
run default A=1
run s256_4 TBK_BLK_SHAPE=256,4
run s512_4 TBK_BLK_SHAPE=512,4
run minb3 PYTHTB_B200_LIB=$PWD/profiles/ab/libtbk_minb3.so
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "replayed" 2>&1 | tail -2
