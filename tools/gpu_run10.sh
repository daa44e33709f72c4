#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/tests10.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests10.log
grep -E "impl 1 vs 2|sync window|passed|failed|exit|FAILED|Error" gpurun_out/tests10.log | head -20
