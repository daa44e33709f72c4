#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/tests15.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests15.log
tail -n 3 gpurun_out/tests15.log
timeout 600 python tools/tc_experiments.py speech2lip_b200/csrc/libs2l_b200.so tools/dbg_noload.so tools/dbg_noepi.so tools/dbg_noboth.so > gpurun_out/exp15.txt 2>&1
cat gpurun_out/exp15.txt
for v in tl tl_noboth; do
  for p in bf16x3 fp16f8; do
    S2L_LIB_PATH=$PWD/tools/dbg_$v.so timeout 200 python tools/tc_timeline.py $p > gpurun_out/${v}_$p.txt 2>&1
    echo "== $v $p"; python tools/tl_analyze.py gpurun_out/${v}_$p.txt | tail -n 24
  done
done
