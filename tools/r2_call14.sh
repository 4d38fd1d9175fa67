#!/bin/bash
# round 2, GPU call 14: the state of HEAD -- full GPU suite, smoke, both bench arms with the driver's arguments, n = 368 line
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -6) | tee gpurun_out/r2_tests14.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench14_ref.json 2> gpurun_out/r2_bench14_ref.err; tail -1 gpurun_out/r2_bench14_ref.json | cut -c1-300; tail -2 gpurun_out/r2_bench14_ref.err
echo "== bench, driver args"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench14_s20.json 2> gpurun_out/r2_bench14_s20.err; tail -1 gpurun_out/r2_bench14_s20.json | cut -c1-300; tail -2 gpurun_out/r2_bench14_s20.err
echo "== bench n = 368 with the psi bound"
LVPP_GMRES_RESTART=20 timeout 900 python bench.py --size 368 --steps 20 --warmup 5 --psi-cap 1 --psi-free 0 --skip-cpu > gpurun_out/r2_bench14_n368.json 2> gpurun_out/r2_bench14_n368.err; tail -1 gpurun_out/r2_bench14_n368.json | cut -c1-300; tail -2 gpurun_out/r2_bench14_n368.err
