#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/tests8.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests8.log
grep -E "fp16f8|passed|failed|exit|FAILED|Error" gpurun_out/tests8.log | head -40
for args in "--steps 5 --warmup 3 --no-cpu-baseline --precision fp16f8" "--steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline --precision fp16f8" "--steps 5 --warmup 3 --no-cpu-baseline"; do
  timeout 600 python bench.py $args 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['mode'],d['config']['precision'],'fps %.1f e2e %.1f kernel_ms %.2f frac %.3f issued %.0f TF clocks %s'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['roofline']['issued_mma_tflops'],d['clocks']))
    else: print(l.strip()[:200])
"
done
