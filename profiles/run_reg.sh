#!/bin/bash
OUT=gpurun_out/${1:-reg5}
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "not config_scale" > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 100 python profiles/split_cfg3_vec.py 2>&1 | tail -1
TBK_REG_EIGVALS=0 timeout 100 python profiles/split_cfg3_vec.py 2>&1 | tail -1
