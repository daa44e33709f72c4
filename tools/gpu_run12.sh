#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_experiments.py speech2lip_b200/csrc/libs2l_b200.so tools/dbg_noload.so tools/dbg_noepi.so tools/dbg_noboth.so > gpurun_out/exp12.txt 2>&1
cat gpurun_out/exp12.txt
