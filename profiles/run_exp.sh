#!/bin/bash
# experiment pass: parity, bench (haldane, kane_mele), CTA timeline, L2 fetch granularity knob
TAG=${1:-exp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], {k:round(v,5) if v<1 else round(v/1e9,3) for k,v in d["stages"].items()}, "value G/s %.3f e2e G/s %.3f e2e_ms %.4f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]), d["check"], "frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
for w in haldane kane_mele; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu > $OUT/bench_$w.json 2>$OUT/bench_$w.err; show $OUT/bench_$w.json
done
TBK_CTA_TRACE=1 timeout 300 python profiles/cta_trace.py haldane > $OUT/cta_trace_haldane.json 2>$OUT/cta_trace_haldane.err; cat $OUT/cta_trace_haldane.json | head -120
