#!/bin/bash
# usage: multi.sh nproc shape pc [ENV=...]
np=$1; shape=$2; pc=$3; shift 3
echo "== np=$np $shape $pc $@"
env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$np --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_worker.py $shape $pc 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
