#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests14.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests14.log
tail -n 5 gpurun_out/tests14.log
timeout 300 python tools/tc_experiments.py --child 2>&1 | tee gpurun_out/exp14.txt
