#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "post_fusion" -s > gpurun_out/tests5.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests5.log
tail -n 8 gpurun_out/tests5.log
timeout 300 python tools/bench_postfusion.py 2>&1 | tail -n 3 | tee gpurun_out/bench_postfusion.log
