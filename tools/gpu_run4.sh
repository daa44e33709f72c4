#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/run4.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 >> gpurun_out/run4.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 >> gpurun_out/run4.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> gpurun_out/run4.log 2>&1
cut -c1-700 gpurun_out/run4.log
