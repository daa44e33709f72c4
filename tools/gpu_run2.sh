#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests2.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke2.log
# launch list of the default bench command (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
# full capture of the tensor-core kernel on a 1-frame step
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 -o gpurun_out/prof_tc_r1 -f python bench.py --steps 1 --warmup 3 --frames 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/tests2.log gpurun_out/smoke2.log gpurun_out/ncu_full.log
ls -la gpurun_out
