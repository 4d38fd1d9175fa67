#!/bin/bash
# round 2, GPU call 16 (2 GPUs): HEAD after the removal of the FGMRES / unroll-8 paths -- full suite incl. the 2-rank cases, smoke, bench at 1 and 2 GPUs
mkdir -p gpurun_out
echo "== gpu tests (all, 2 GPUs visible)"
(timeout 1200 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -6) | tee gpurun_out/r2_tests16.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench16_2gpu.json 2> gpurun_out/r2_bench16_2gpu.err; tail -1 gpurun_out/r2_bench16_2gpu.json | cut -c1-260; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench16_2gpu.err | tail -3
echo "== bench reference arm under torchrun (rank 0 prints, rank 1 exits)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 --cpu-size 10 2>/dev/null | tail -1 | cut -c1-200
echo "== bench 1 GPU"
timeout 900 python bench.py > gpurun_out/r2_bench16_1gpu.json 2> gpurun_out/r2_bench16_1gpu.err; tail -1 gpurun_out/r2_bench16_1gpu.json | cut -c1-260; tail -2 gpurun_out/r2_bench16_1gpu.err
