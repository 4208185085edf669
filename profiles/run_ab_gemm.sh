#!/bin/bash
OUT=gpurun_out/${1:-ab_gemm}
mkdir -p $OUT
for V in default gemm_single gemm_pipe1; do
  if [ $V = default ]; then unset PYTHTB_B200_LIB; else export PYTHTB_B200_LIB=$PWD/profiles/ab/libtbk_$V.so; fi
  timeout 400 python profiles/time_large.py > $OUT/time_large_$V.json 2> $OUT/time_large_$V.err; tail -2 $OUT/time_large_$V.err
  python - <<PY
import json
d=json.loads(open("$OUT/time_large_$V.json").read().strip().splitlines()[-1])
for k,v in d.items():
    if isinstance(v,dict): print("$V", k, {kk: round(vv,2) for kk,vv in v.items() if kk.endswith("_ms") or kk.endswith("tflops")})
PY
done
unset PYTHTB_B200_LIB
timeout 600 python -m pytest tests/test_gpu_config_scale.py tests/test_gpu_parity.py -m gpu -q -x -k "wilson or position or slab or ribbon or config_scale" > $OUT/pytest_large.log 2>&1; tail -3 $OUT/pytest_large.log
