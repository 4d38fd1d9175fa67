#!/bin/bash
# round 2, GPU call 21: GMRES warm start from the previous Newton correction -- full GPU suite, bench with and without
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -8) | tee gpurun_out/r2_tests21.txt
timeout 600 python bench.py --skip-cpu --skip-aux > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; tail -1 gpurun_out/r2_bench21.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("warm: ms/step %.1f DOFs/s %.1fM krylov %d e2e %.1fM" % (d["ms_per_step"], d["value"]/1e6, d["krylov_iterations"], d["e2e"]["value"]/1e6), d["solve_in_progress"]["newton_steps"], d["solve_in_progress"]["krylov_iterations"], d["failures"])'; tail -2 gpurun_out/r2_bench21.err
LVPP_GMRES_WARM=0 timeout 600 python bench.py --skip-cpu --skip-aux --skip-e2e > gpurun_out/r2_bench21_cold.json 2> gpurun_out/r2_bench21_cold.err; tail -1 gpurun_out/r2_bench21_cold.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("cold: ms/step %.1f DOFs/s %.1fM krylov %d" % (d["ms_per_step"], d["value"]/1e6, d["krylov_iterations"]), d["solve_in_progress"]["krylov_iterations"])'
