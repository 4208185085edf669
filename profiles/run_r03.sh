#!/bin/bash
# r03 pass: everything run_round.sh does plus the config-scale secondary benchmark
TAG=${1:-r03}
bash profiles/run_round.sh $TAG
timeout 900 python profiles/bench_configs.py > gpurun_out/$TAG/bench_configs.json 2> gpurun_out/$TAG/bench_configs.err; tail -c 2500 gpurun_out/$TAG/bench_configs.json
