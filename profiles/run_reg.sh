#!/bin/bash
OUT=gpurun_out/${1:-reg1}
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "not config_scale" > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 100 python profiles/split_cfg3.py 2>&1 | tail -1
timeout 300 python bench.py --no-cpu --extras 3 --steps 20 --warmup 3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_cfg3.json").read().strip().splitlines()[-1])
c=d["configs"]["3"]; print("config 3", c["value"], c["e2e"], c["roofline"]["frac"], c["check"], c["kernel"])
PY
