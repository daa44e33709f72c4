#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s -k "config4 or training" > gpurun_out/tests7.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests7.log
grep -E "maxabs|passed|failed|exit|backward" gpurun_out/tests7.log
timeout 600 python tools/bench_train.py 2>&1 | tail -n 2 | tee gpurun_out/bench_train.log
