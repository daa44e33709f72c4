#!/bin/bash
mkdir -p gpurun_out
S2L_TC_IMPL=2 timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests18.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests18.log
tail -n 12 gpurun_out/tests18.log
echo "== impl 1"; timeout 300 python tools/tc_experiments.py --child
echo "== impl 2"; S2L_TC_IMPL=2 timeout 300 python tools/tc_experiments.py --child
for impl in 1 2; do
  S2L_TC_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v18_$impl.json 2>gpurun_out/v18_$impl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/v18_$impl.json').read().strip().splitlines()[-1])
print('impl $impl: %.1f fps  %.2f ms/step  clocks %s W %s'%(d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],d['clocks']['power_w_max']))
PY
done
