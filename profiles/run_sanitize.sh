#!/bin/bash
OUT=gpurun_out/${1:-r15}
mkdir -p $OUT
timeout 60 python profiles/sanitize_small.py 2>&1 | tail -1
timeout 100 python profiles/split_cfg3.py 2>&1 | tail -1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck.log python profiles/sanitize_small.py > $OUT/sanitizer_memcheck_run.log 2>&1; echo "memcheck exit $?"; tail -2 $OUT/sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/sanitizer_racecheck.log python profiles/sanitize_small.py > $OUT/sanitizer_racecheck_run.log 2>&1; echo "racecheck exit $?"; tail -2 $OUT/sanitizer_racecheck.log
TBK_REG_GEMM=0 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/sanitizer_racecheck_scalar.log python profiles/sanitize_small.py > /dev/null 2>&1; echo "racecheck (scalar sweep) exit $?"; tail -1 $OUT/sanitizer_racecheck_scalar.log
