#!/bin/bash
# round-2 validation: tests (all failures shown), smoke, both bench arms, profiles.  usage: tools/gpu_r2.sh <tag> [pytest -k expr]
tag=${1:-r2}
kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 1500 python -m pytest tests -q -m gpu -k "$kexpr" -s --tb=short > gpurun_out/tests_${tag}.log 2>&1
else
  timeout 1500 python -m pytest tests -q -m gpu --tb=short -s > gpurun_out/tests_${tag}.log 2>&1
fi
echo "pytest exit $?" >> gpurun_out/tests_${tag}.log
grep -E "passed|failed|error" gpurun_out/tests_${tag}.log | tail -n 5
grep -E "^FAILED|MODE-V SOAK|Error" gpurun_out/tests_${tag}.log | head -n 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2>gpurun_out/bench_${tag}.err
tail -c 600 gpurun_out/bench_${tag}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1])
    print('value %.1f fps  e2e %.1f  ms/step %.2f  frac %.3f  kernel_ms %.2f  launches %s clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms_per_launch'],d['gpu_launches'],d['clocks']))
    print('cpu', d.get('cpu_baseline'))
    for k,v in d.get('extras',{}).items(): print('  ',k, v)
except Exception as e:
    print('bench parse failed', e)
PY
