#!/bin/bash
# round 2, GPU call 17: bench.py after the shared-config refactor (both arms, 3-D and 2-D)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; tail -1 gpurun_out/r2_bench17.json | cut -c1-400; tail -2 gpurun_out/r2_bench17.err
timeout 300 python bench.py --workload obstacle2d --steps 6 --warmup 3 --skip-cpu > gpurun_out/r2_bench17_2d.json 2> gpurun_out/r2_bench17_2d.err; tail -1 gpurun_out/r2_bench17_2d.json | cut -c1-400; tail -2 gpurun_out/r2_bench17_2d.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 --cpu-size 10 | cut -c1-400
