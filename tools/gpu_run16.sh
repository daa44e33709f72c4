#!/bin/bash
mkdir -p gpurun_out
for v in tl tl_noboth; do
  for p in bf16x3 fp16f8 bf16x1; do
    S2L_LIB_PATH=$PWD/tools/dbg_$v.so timeout 200 python tools/tc_timeline.py $p > gpurun_out/${v}_$p.txt 2>&1
    echo "== $v $p"; grep "^#" gpurun_out/${v}_$p.txt; python tools/tl_analyze.py gpurun_out/${v}_$p.txt | tail -n 19
  done
done
