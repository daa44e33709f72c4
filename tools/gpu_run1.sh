#!/bin/bash
# first GPU contact: exact path first, tensor-core path in a separate process (a trapped kernel kills the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== fp32 / non-tc tests" > gpurun_out/run1.log
timeout 600 python -m pytest tests -q -m gpu -k "fp32 or audio or rows or composite or get_rays" -s --maxfail=50 >> gpurun_out/run1.log 2>&1
echo "exit $?" >> gpurun_out/run1.log
echo "=== tc tests" >> gpurun_out/run1.log
timeout 900 python -m pytest tests -q -m gpu -k "not fp32 and not audio and not rows and not composite and not get_rays" -s --maxfail=50 >> gpurun_out/run1.log 2>&1
echo "exit $?" >> gpurun_out/run1.log
echo "=== bench fp32 (tiny)" >> gpurun_out/run1.log
timeout 600 python bench.py --precision fp32 --frames 1 --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/run1.log 2>&1
echo "=== bench bf16x3" >> gpurun_out/run1.log
timeout 900 python bench.py --steps 5 --warmup 3 >> gpurun_out/run1.log 2>&1
echo "=== bench bf16x1" >> gpurun_out/run1.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision bf16x1 --no-cpu-baseline >> gpurun_out/run1.log 2>&1
echo "=== bench plain bf16x3" >> gpurun_out/run1.log
timeout 600 python bench.py --steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline >> gpurun_out/run1.log 2>&1
tail -c 6000 gpurun_out/run1.log
