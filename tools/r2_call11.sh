#!/bin/bash
# round 2, GPU call 11 (8 GPUs): 4-rank parity tests, bench at 8 / 4 / 2 GPUs (clamped stack), 8 GPUs open stack
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== multi-GPU tests (4-rank cases)"
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXfs -k "4-" 2>&1 | tail -6) | tee gpurun_out/r2_tests11_multi.txt
for np in 8 4 2; do
  echo "== bench $np GPUs (stack, clamped)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 2953$np bench.py --gpus $np --steps 20 --warmup 5 > gpurun_out/r2_bench11_${np}gpu.json 2> gpurun_out/r2_bench11_${np}gpu.err; tail -1 gpurun_out/r2_bench11_${np}gpu.json | cut -c1-260; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench11_${np}gpu.err | tail -3
done
echo "== bench 8 GPUs (stack-open)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --weak stack-open --skip-e2e > gpurun_out/r2_bench11_8gpu_open.json 2> gpurun_out/r2_bench11_8gpu_open.err; tail -1 gpurun_out/r2_bench11_8gpu_open.json | cut -c1-260
echo "== bench 1 GPU"
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-e2e > gpurun_out/r2_bench11_1gpu.json 2> gpurun_out/r2_bench11_1gpu.err; tail -1 gpurun_out/r2_bench11_1gpu.json | cut -c1-260
