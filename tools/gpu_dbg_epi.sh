#!/bin/bash
for v in "" 1 2 4 7; do
  if [ -z "$v" ]; then unset S2L_LIB_PATH; else export S2L_LIB_PATH=$PWD/speech2lip_b200/csrc/dbg_epi$v.so; fi
  for prec in bf16x3 bf16x1; do
    timeout 300 python bench.py --steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline --precision $prec 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('variant=$v', d['config']['precision'], 'kernel_ms %.3f'%d['roofline']['kernel_ms_per_launch'], 'finite', d['finite'])
"
  done
done
