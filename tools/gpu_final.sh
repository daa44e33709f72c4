#!/bin/bash
# round-end style validation: tests, smoke, both bench arms, profiles
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests_${tag}.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests_${tag}.log
tail -n 3 gpurun_out/tests_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>gpurun_out/bench_ref_${tag}.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2>gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1])
print('value %.1f fps  e2e %.1f  ms/step %.2f  frac %.3f  issued %.0f TF  clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['issued_mma_tflops'],d['clocks']))
print('cpu', d.get('cpu_baseline'))
for k,v in d.get('extras',{}).items(): print('  ',k, '%.1f fps'%v['frames_per_s'])
r=json.loads(open('gpurun_out/bench_ref_${tag}.json').read().strip().splitlines()[-1]); print('reference arm %.4f fps on %d cores'%(r['value'], r['cpu_baseline']['cores']))
PY
bash tools/gpu_profile.sh ${tag} > /dev/null 2>&1
ls gpurun_out | grep ${tag}
