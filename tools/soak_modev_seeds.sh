#!/bin/bash
# Mode-V parity soak over several weight sets / audio seeds (each run = 64 frames of 256x256x64 in fp16f8 AND bf16x3 against the
# exact fp32 path on all rays and the oracle on 256 rays per frame).  usage: tools/soak_modev_seeds.sh <tag>
tag=${1:-rX}
mkdir -p gpurun_out
out=gpurun_out/modev_soak_seeds_${tag}.txt
: > $out
for spec in "kaiming 1 61" "kaiming 2 62" "kaiming 3 63" "kaiming 4 64" "default 0 65" "default 1 66"; do
  set -- $spec
  S2L_SOAK_KIND=$1 S2L_SOAK_SEED=$2 S2L_SOAK_AUDIO_SEED=$3 timeout 900 python -m pytest tests/test_gpu_modev.py -q -s -k soak --tb=line 2>&1 \
    | grep -E "MODE-V SOAK|passed|failed|Error" >> $out
done
cat $out
