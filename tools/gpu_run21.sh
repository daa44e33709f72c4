#!/bin/bash
export S2L_TC_IMPL=2
timeout 600 python tools/tc_experiments.py speech2lip_b200/csrc/libs2l_b200.so tools/dbg_noload.so tools/dbg_noepi.so tools/dbg_noboth.so
