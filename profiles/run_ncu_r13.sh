#!/bin/bash
# r13: ncu --set full of the staged blocked eigensolver (n = 400 ribbon, 296 matrices) + raw copy bandwidth of the box
OUT=gpurun_out/${1:-r13}
mkdir -p $OUT
python profiles/d2h_peak.py > $OUT/d2h_peak.json 2>&1; cat $OUT/d2h_peak.json
PROF_WHICH=ribbon_n400 PROF_REPS=1 timeout 800 ncu --set full --clock-control none --import-source on \
  -k regex:'solve_blocked|blk_wy|blk_backtransform' -c 3 -f -o $OUT/prof_staged python profiles/prof_stages.py > $OUT/ncu_staged.log 2>&1
ncu -i $OUT/prof_staged.ncu-rep --page raw --csv > $OUT/raw_staged.csv 2>/dev/null
ncu -i $OUT/prof_staged.ncu-rep --page source --csv --print-source cuda > $OUT/src_staged.csv 2>/dev/null
tail -3 $OUT/ncu_staged.log; ls -la $OUT
