#!/bin/bash
mkdir -p gpurun_out
for p in bf16x3 fp16f8 bf16x1; do
  S2L_TC_IMPL=2 S2L_LIB_PATH=$PWD/tools/dbg_tl.so timeout 200 python tools/tc_timeline.py $p > gpurun_out/tl2_$p.txt 2>&1
  echo "== tc2 $p"; grep "^#" gpurun_out/tl2_$p.txt; python tools/tl_analyze.py gpurun_out/tl2_$p.txt | tail -n 16
done
