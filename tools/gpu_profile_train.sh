#!/bin/bash
# usage: tools/gpu_profile_train.sh <tag>  -> ncu --set full captures of the three training kernels at the config-5 geometry (24 renders of 80x120)
tag=${1:-rX}
mkdir -p gpurun_out
for k in dgrad_tc_kernel wgrad_tc_kernel "mlp_tc_kernel"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_train_${k}_${tag} -f python tools/bench_train_tc.py 80 120 > gpurun_out/ncu_train_${k}_${tag}.log 2>&1
done
ls -la gpurun_out | grep prof_train
