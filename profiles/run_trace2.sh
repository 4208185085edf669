#!/bin/bash
N=${1:-2}
OUT=gpurun_out/${2:-trace2}
mkdir -p $OUT
TBK_CTA_TRACE=1 timeout 200 python profiles/cta_trace.py 2>&1 | grep "^rank" > $OUT/spans_n1.log; cat $OUT/spans_n1.log
TBK_CTA_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 profiles/cta_trace.py 2>&1 | grep "^rank" > $OUT/spans_n$N.log; cat $OUT/spans_n$N.log
