#!/bin/bash
OUT=gpurun_out/${1:-shape2}
mkdir -p $OUT
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu --extras 4 --steps 20 --warmup 3 > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$tag.json").read().strip().splitlines()[-1])
    for k,c in d["configs"].items(): print("$tag", k, round(c["value"],1), c["check"]["ok"])
except Exception as e: print("$tag failed", e)
PY
}
run default A=1
run s512_16 TBK_BLK_SHAPE_ALL=1 TBK_BLK_SHAPE=512,16
run s512_8 TBK_BLK_SHAPE_ALL=1 TBK_BLK_SHAPE=512,8
