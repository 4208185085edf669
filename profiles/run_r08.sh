#!/bin/bash
# r08 (round 2): parity suite with the state-major device layout + posted/completed reductions, smoke, the bench
# line with extras, the reference arm, and one compute-sanitizer pass (memcheck + racecheck) over the golden-size
# cases, the mesh kernels and the shard slabs (ticket / last-CTA / mailbox patterns).
TAG=${1:-r08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err; cut -c1-3000 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
if [ "$2" = "sanitize" ]; then
  SEL='case_matches or mesh_kernels or shard_slabs'
  timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/sanitizer_memcheck_pytest.log 2>&1; echo "memcheck exit $?" >> $OUT/sanitizer_memcheck_pytest.log
  tail -3 $OUT/sanitizer_memcheck_pytest.log; tail -3 $OUT/sanitizer_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/sanitizer_racecheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/sanitizer_racecheck_pytest.log 2>&1; echo "racecheck exit $?" >> $OUT/sanitizer_racecheck_pytest.log
  tail -3 $OUT/sanitizer_racecheck_pytest.log; tail -3 $OUT/sanitizer_racecheck.log
fi
ls -la $OUT; du -sh gpurun_out
