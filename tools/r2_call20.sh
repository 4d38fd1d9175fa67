#!/bin/bash
# round 2, GPU call 20 (2 GPUs): the bench at HEAD under torchrun
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench20_2gpu.json 2> gpurun_out/r2_bench20_2gpu.err; tail -1 gpurun_out/r2_bench20_2gpu.json | cut -c1-260; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench20_2gpu.err | tail -3
