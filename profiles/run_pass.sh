#!/bin/bash
# r04 pass: run_round.sh (parity, smoke, bench x2 + reference arm, ncu launch list + full captures) + CTA timeline + configs
TAG=${1:-r05}
bash profiles/run_round.sh $TAG
TBK_CTA_TRACE=1 timeout 300 python profiles/cta_trace.py haldane > gpurun_out/$TAG/cta_trace_haldane.json 2> gpurun_out/$TAG/cta_trace_haldane.err
grep "us (" gpurun_out/$TAG/cta_trace_haldane.json
timeout 900 python profiles/bench_configs.py > gpurun_out/$TAG/bench_configs.json 2> gpurun_out/$TAG/bench_configs.err; tail -c 600 gpurun_out/$TAG/bench_configs.json
