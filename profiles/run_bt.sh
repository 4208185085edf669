#!/bin/bash
# back-transformation kernel: parity + per-kernel durations (ncu launch list of profiles/prof_stages.py)
OUT=gpurun_out/${1:-bt1}
mkdir -p $OUT
timeout 60 python profiles/ring_smoke.py 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_config_scale.py tests/test_gpu_parity.py -m gpu -q -x -k "config_scale or large_ribbon or blocked or slab or hwf" 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python profiles/prof_stages.py > $OUT/log.txt 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
for r in rows[1:]:
    if "backtransform" in r[ki]: print(r[ki][:40], r[gi], r[vi])
PY
