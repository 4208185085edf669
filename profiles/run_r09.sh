#!/bin/bash
# r09: parity suite after the lower-triangle tridiagonalisation / lazily reduced models / deeper flux prefetch,
# A/B of the flux kernel's prefetch rotation, stage profile of the blocked eigensolver, full bench line.
TAG=${1:-r09}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log
for S in 4 6 8; do
  TBK_FLUX_SLOTS=$S timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu --extras none > $OUT/bench_slots$S.json 2> $OUT/bench_slots$S.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_slots$S.json").read().strip().splitlines()[-1])
print("slots $S: value %.3f G ms/step %.4f solve %.4f flux %.4f e2e %.4f"%(d["value"]/1e9,d["ms_per_step"],d["stages"]["solve_on_grid_ms"],d["stages"]["berry_flux_ms"],d["e2e"]["ms_per_step"]))
PY
done
timeout 400 python profiles/prof_blocked.py > $OUT/prof_blocked.json 2> $OUT/prof_blocked.err; python - <<PY
import json
d=json.load(open("$OUT/prof_blocked.json"))
for k,v in d.items(): print(k, "%.0f k/s"%v["kpts_per_s"], "%.2f ms/matrix/CTA"%v["ms_per_matrix_per_cta"], v["share"], "fallbacks", v["fallbacks"])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("wall %.0f s value %.3f G e2e %.3f G"%(d["bench_wall_s"], d["value"]/1e9, d["e2e"]["value"]/1e9))
for k,v in list(d["workloads"].items())+list(d["configs"].items()):
    print(k, {kk: v.get(kk) for kk in ("value","skipped","error","cpu_baseline")}, (v.get("roofline") or {}).get("frac"))
PY
ls -la $OUT; du -sh gpurun_out
