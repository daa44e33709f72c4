#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/mma_rate 512 > gpurun_out/mma_rate.txt 2>&1; echo "exit $?" >> gpurun_out/mma_rate.txt
cat gpurun_out/mma_rate.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests11.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests11.log
tail -n 3 gpurun_out/tests11.log
