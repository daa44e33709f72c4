#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests3.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests3.log
tail -n 3 gpurun_out/tests3.log
for args in "--steps 5 --warmup 3 --no-cpu-baseline" "--steps 5 --warmup 3 --precision bf16x1 --no-cpu-baseline" "--steps 5 --warmup 3 --mode plain --frames 64 --no-cpu-baseline"; do
  timeout 600 python bench.py $args >> gpurun_out/bench3.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/bench3.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['mode'],d['config']['precision'],'fps %.1f e2e %.1f kernel_ms %.2f frac %.3f clocks %s'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac'],d['clocks']))
    else: print(l.strip()[:300])
PY
