#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of the bench step, --set full captures of the two headline kernels
# (Haldane, Kane-Mele) and of the large-matrix families at n = 400, reduced on the box to CSV / per-line summaries.
TAG=${1:-r11}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --extras none > $OUT/bench_under_ncu.log 2>&1
for WL in haldane kane_mele; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_small|flux_rows' -s 8 -c 4 -f -o /tmp/prof_$WL python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --extras none > $OUT/ncu_full_$WL.log 2>&1
  ncu -i /tmp/prof_$WL.ncu-rep --page raw --csv > $OUT/raw_$WL.csv 2>/dev/null
  ncu -i /tmp/prof_$WL.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$WL.csv 2>/dev/null
  python profiles/summarize_lines.py /tmp/src_$WL.csv "" 1.0 > $OUT/lines_$WL.txt 2>/dev/null
  rm -f /tmp/prof_$WL.ncu-rep /tmp/src_$WL.csv
done
PROF_NCELL=200 PROF_NK=297 timeout 200 python profiles/prof_large.py > $OUT/prof_large_shapes.json 2> $OUT/prof_large.err
PROF_NCELL=200 PROF_NK=297 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'solve_blocked|link_matrix|position_matrix_dmma|hwf_to_orbital_dmma|string_product' \
  -c 8 -f -o /tmp/prof_large python profiles/prof_large.py > $OUT/ncu_large.log 2>&1
ncu -i /tmp/prof_large.ncu-rep --page raw --csv > $OUT/raw_large.csv 2>/dev/null
ncu -i /tmp/prof_large.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_large.csv 2>/dev/null
python profiles/summarize_lines.py /tmp/src_large.csv "" 1.5 > $OUT/lines_large.txt 2>/dev/null
rm -f /tmp/prof_large.ncu-rep /tmp/src_large.csv
tail -2 $OUT/ncu_large.log; ls -la $OUT; du -sh gpurun_out
