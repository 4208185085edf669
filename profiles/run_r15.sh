#!/bin/bash
# r15: last full pass of round 2 — parity suite, smoke, bench line (every workload / config) + reference arm, launch list of the
# bench command, full-set capture of the register / tensor-pipe sweep kernel.  Every step under a timeout.
OUT=gpurun_out/${1:-r18}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --extras kane_mele,3 > $OUT/bench_under_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'solve_reg' -c 1 -f -o $OUT/prof_reg python profiles/split_cfg3.py > $OUT/ncu_reg.log 2>&1
ncu -i $OUT/prof_reg.ncu-rep --page raw --csv > $OUT/raw_reg.csv 2>/dev/null
ncu -i $OUT/prof_reg.ncu-rep --page source --csv --print-source cuda,sass > $OUT/src_reg.csv 2>/dev/null
python profiles/summarize_lines.py $OUT/src_reg.csv "" 1.5 > $OUT/lines_reg.txt 2>&1; rm -f $OUT/src_reg.csv
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"], "wfs_to_host ms", d["e2e_wfs_to_host"]["ms_per_step"], "wall", d.get("bench_wall_s"))
print("kane_mele", d["workloads"]["kane_mele"]["value"], d["workloads"]["kane_mele"]["stages"])
for k,v in d["configs"].items():
    print(k, v.get("value"), v.get("e2e",{}).get("value"), v.get("stages"), v["roofline"].get("frac"), v.get("check",{}).get("ok"))
PY
