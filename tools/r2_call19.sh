#!/bin/bash
# round 2, GPU call 19: HEAD with lvpp_newton_begin_same_iterate -- full GPU suite, smoke, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -8) | tee gpurun_out/r2_tests19.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --skip-cpu > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; tail -1 gpurun_out/r2_bench19.json | cut -c1-200; tail -2 gpurun_out/r2_bench19.err
