#!/bin/bash
OUT=gpurun_out/${1:-slots2}
mkdir -p $OUT
timeout 400 python bench.py --no-cpu --extras 4,5 --steps 20 --warmup 3 > $OUT/bench_cfg45.json 2> $OUT/bench_cfg45.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_cfg45.json").read().strip().splitlines()[-1])
for k,c in d["configs"].items(): print(k, round(c["value"],1), c.get("stages"), c.get("peak_device_memory_gb"), c["check"])
PY
timeout 300 python -m pytest tests -m gpu -x -q -k "config_scale or stream or large" 2>&1 | tail -2
