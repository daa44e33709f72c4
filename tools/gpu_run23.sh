#!/bin/bash
mkdir -p gpurun_out
S2L_TC_IMPL=2 timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests23.log 2>&1; echo "pytest(impl 2) exit $?" >> gpurun_out/tests23.log
tail -n 3 gpurun_out/tests23.log
for impl in 1 2; do
 for prec in fp16f8 bf16x3; do
  S2L_TC_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --precision $prec > gpurun_out/v23_${impl}_$prec.json 2>gpurun_out/v23_${impl}_$prec.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/v23_${impl}_$prec.json').read().strip().splitlines()[-1])
print('impl $impl $prec: %.1f fps  %.2f ms/step  clocks %s W %s'%(d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],d['clocks']['power_w_max']))
PY
 done
done
