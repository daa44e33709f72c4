#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s -k "training or drop_in or post_fusion" > gpurun_out/tests6.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests6.log
tail -n 25 gpurun_out/tests6.log
timeout 300 python tools/bench_postfusion.py 2>&1 | tail -n 1 | tee gpurun_out/bench_postfusion.log
