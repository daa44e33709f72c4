#!/bin/bash
# Sustained (power-limited) throughput of debug builds of the library — what each part of the kernel costs in the
# regime the headline number is measured in.  Build the variants in the build container first:
#   python - <<PY
#   import sys; sys.path.insert(0, "speech2lip_b200/csrc"); import build
#   for n, d in (("noload", ["S2L_DBG_NOLOAD"]), ("noepi", ["S2L_DBG_NOEPI"]), ("noboth", ["S2L_DBG_NOLOAD", "S2L_DBG_NOEPI"])):
#       build.build(force=True, defines=d, out="tools/dbg_%s.so" % n)
#   PY
# then: gpurun -- 'bash tools/gpu_power_experiments.sh'      (S2L_TC_IMPL=1|2|3 selects the schedule)
# NOLOAD leaves the ring stages at their initial contents, so its MMAs multiply constant data: it bounds the cost of the
# weight stream from above (operand toggling in the tensor cores drops too).
mkdir -p gpurun_out
for v in base noload noepi noboth; do
  lib=$PWD/tools/dbg_$v.so; [ $v = base ] && lib=$PWD/speech2lip_b200/csrc/libs2l_b200.so
  [ -f $lib ] || continue
  S2L_LIB_PATH=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/power_$v.json 2>gpurun_out/power_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/power_$v.json').read().strip().splitlines()[-1])
print('$v: %.1f fps  %.2f ms/step  clocks %s MHz  %s W  schedule %s'%(d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],d['clocks']['power_w_max'],d['config']['tc_schedule']))
PY
done
