#!/bin/bash
OUT=gpurun_out/${1:-km_ncu}
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mesh_small' -s 6 -c 1 -f -o $OUT/prof_km python bench.py --workload kane_mele --no-cpu --extras none --steps 3 --warmup 3 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof_km.ncu-rep --page raw --csv > $OUT/raw_km.csv 2>/dev/null
ncu -i $OUT/prof_km.ncu-rep --page source --csv --print-source cuda,sass > $OUT/src_km.csv 2>/dev/null
python profiles/summarize_lines.py $OUT/src_km.csv "" 0.8 > $OUT/lines_km.txt 2>&1
rm -f $OUT/src_km.csv
tail -2 $OUT/ncu.log
