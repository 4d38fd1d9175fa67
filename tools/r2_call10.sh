#!/bin/bash
# round 2, GPU call 10 (2 GPUs): partitioned == single GPU tests, the 2-GPU bench line (stacked slabs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== multi-GPU tests"
(timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXfs 2>&1 | tail -12) | tee gpurun_out/r2_tests10_multi.txt
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench10_2gpu.json 2> gpurun_out/r2_bench10_2gpu.err; tail -1 gpurun_out/r2_bench10_2gpu.json | cut -c1-300; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench10_2gpu.err | tail -5
echo "== bench 1 GPU (same box)"
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_bench10_1gpu.json 2> gpurun_out/r2_bench10_1gpu.err; tail -1 gpurun_out/r2_bench10_1gpu.json | cut -c1-300
