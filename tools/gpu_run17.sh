#!/bin/bash
# sustained (power-limited) V-mode throughput of debug variants
mkdir -p gpurun_out
for v in base halfload suspend noload noepi; do
  lib=$PWD/tools/dbg_$v.so; [ $v = base ] && lib=$PWD/speech2lip_b200/csrc/libs2l_b200.so
  S2L_LIB_PATH=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v17_$v.json 2>gpurun_out/v17_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/v17_$v.json').read().strip().splitlines()[-1])
print('$v: %.1f fps  %.2f ms/step  clocks %s W %s'%(d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],d['clocks']['power_w_max']))
PY
done
