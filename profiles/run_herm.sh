#!/bin/bash
OUT=gpurun_out/${1:-herm1}
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "wilson or position or slab or hwf or config_scale or random_models_berry or case_matches" > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 150 python profiles/time_large.py > $OUT/time_large.json 2> $OUT/time_large.err
python - <<PY
import json
d=json.loads(open("$OUT/time_large.json").read().strip().splitlines()[-1])
for k,x in d.items():
    if isinstance(x,dict): print(k, {kk: round(vv,2) for kk,vv in x.items() if kk.endswith("_ms")})
PY
