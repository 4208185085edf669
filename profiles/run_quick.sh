#!/bin/bash
# quick GPU pass: parity tests, bench (haldane + kane_mele) with optional env knobs, optional ncu capture
# usage: bash profiles/run_quick.sh <tag> [ncu]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], {k:round(v,5) if v<1 else round(v/1e9,3) for k,v in d["stages"].items()}, "value G/s %.3f e2e G/s %.3f e2e_ms %.4f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]), d["check"], d["roofline"]["kernel"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
for w in haldane kane_mele; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu > $OUT/bench_$w.json 2>$OUT/bench_$w.err; show $OUT/bench_$w.json
  TBK_FLUX_RING=0 timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu > $OUT/bench_${w}_noring.json 2>$OUT/bench_${w}_noring.err; show $OUT/bench_${w}_noring.json
done
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_small|flux_r' -s 8 -c 4 -f -o $OUT/prof_haldane python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_small|flux_r' -s 8 -c 4 -f -o $OUT/prof_kane_mele python bench.py --workload kane_mele --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_km.log 2>&1
fi
