#!/bin/bash
# ncu --set full capture of the large-matrix kernel families (one launch each), under gpurun:
#   bash profiles/run_large_ncu.sh <tag>
TAG=${1:-r06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 100 python profiles/prof_large.py > $OUT/prof_large_shapes.json 2> $OUT/prof_large.err
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'solve_blocked|link_matrix|position_matrix_dmma|hwf_to_orbital_dmma|string_product|unitary_herm|unitary_rayleigh' \
  -c 24 -f -o $OUT/prof_large python profiles/prof_large.py > $OUT/ncu_large.log 2>&1
ncu -i $OUT/prof_large.ncu-rep --page raw --csv > $OUT/raw_large.csv 2>/dev/null
tail -3 $OUT/ncu_large.log; ls -la $OUT
