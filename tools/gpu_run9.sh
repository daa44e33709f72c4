#!/bin/bash
mkdir -p gpurun_out
export S2L_TC_IMPL=2
timeout 600 python -m pytest tests -q -m gpu -x -k "plain_vs_golden or volumetric_vs_golden or ensemble4_vs_golden or ragged or independent" -s > gpurun_out/tests9.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests9.log
grep -E "maxabs|passed|failed|exit|FAILED|Error|timeout" gpurun_out/tests9.log | head -40
for prec in bf16x3 fp16f8 bf16x1; do
  timeout 300 python bench.py --steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline --no-extras --precision $prec 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('impl2', d['config']['precision'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], d['finite'])
    elif 'rror' in l or 'timeout' in l: print(l.strip()[:200])
"
done
