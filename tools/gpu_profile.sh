#!/bin/bash
# usage: tools/gpu_profile.sh <tag>   -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep (the BENCHED geometry)
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_${tag}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 1 -o gpurun_out/prof_${tag} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out | tail -n 8
