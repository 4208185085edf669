#!/bin/bash
# r06 pass (the headline kernels are unchanged since r05, whose launch list and full-set captures stay valid):
# parity tests, smoke, the bench line + reference arm at HEAD, and one ncu --set full capture of the
# large-matrix kernel families.  .ncu-rep files are reduced to CSV on the box and deleted (gpurun_out <= 64 MiB).
TAG=${1:-r06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench_haldane.json 2> $OUT/bench_haldane.err; cut -c1-400 $OUT/bench_haldane.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 100 python profiles/prof_large.py > $OUT/prof_large_shapes.json 2> $OUT/prof_large.err
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'solve_blocked|link_matrix|position_matrix_dmma|hwf_to_orbital_dmma|string_product|unitary_herm|unitary_rayleigh' \
  -c 24 -f -o /tmp/prof_large python profiles/prof_large.py > $OUT/ncu_large.log 2>&1
ncu -i /tmp/prof_large.ncu-rep --page raw --csv > $OUT/raw_large.csv 2>/dev/null
rm -f /tmp/prof_large.ncu-rep
tail -2 $OUT/ncu_large.log; ls -la $OUT; du -sh gpurun_out
