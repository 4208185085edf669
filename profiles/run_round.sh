#!/bin/bash
# One GPU-box pass: parity tests, bench (both workloads + reference arm), ncu launch list, ncu full capture.
# Usage (under gpurun): bash profiles/run_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_haldane.json 2> $OUT/bench_haldane.err; tail -c 3000 $OUT/bench_haldane.json
timeout 600 python bench.py --workload kane_mele > $OUT/bench_kane_mele.json 2> $OUT/bench_kane_mele.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_small|flux_rows' -s 8 -c 4 -f -o $OUT/prof_haldane python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mesh_small|flux_rows' -s 8 -c 4 -f -o $OUT/prof_kane_mele python bench.py --workload kane_mele --steps 3 --warmup 3 --no-cpu > $OUT/ncu_full_km.log 2>&1
ls -la $OUT
