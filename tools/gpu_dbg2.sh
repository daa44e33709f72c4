#!/bin/bash
for v in "" loadlo; do
  if [ -z "$v" ]; then unset S2L_LIB_PATH; else export S2L_LIB_PATH=$PWD/speech2lip_b200/csrc/dbg_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline --precision bf16x1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('variant=$v', d['config']['precision'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'])
"
done
