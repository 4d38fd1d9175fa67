#!/bin/bash
# round 2, GPU call 9: tests + bench with the round's defaults (V(2,3), bf16 records, equilibrated norm, no fallback), profiles
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -6) | tee gpurun_out/r2_tests9.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"
timeout 900 python bench.py > gpurun_out/r2_bench9_default.json 2> gpurun_out/r2_bench9_default.err; tail -1 gpurun_out/r2_bench9_default.json | cut -c1-300; tail -3 gpurun_out/r2_bench9_default.err
echo "== bench, driver args"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench9_s20.json 2> gpurun_out/r2_bench9_s20.err; tail -1 gpurun_out/r2_bench9_s20.json | cut -c1-300; tail -3 gpurun_out/r2_bench9_s20.err
echo "== bench 2-D"
timeout 900 python bench.py --workload obstacle2d --steps 20 --warmup 5 > gpurun_out/r2_bench9_2d.json 2> gpurun_out/r2_bench9_2d.err; tail -1 gpurun_out/r2_bench9_2d.json | cut -c1-300; tail -3 gpurun_out/r2_bench9_2d.err
echo "== ncu full: k_packed2_op (first launches = fine level, eigenvalue estimate)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_packed2_op -c 3 -o gpurun_out/r2_packed2_n215 python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-aux > /dev/null 2> gpurun_out/r2_ncu_full.err; tail -2 gpurun_out/r2_ncu_full.err; ls -la gpurun_out/*.ncu-rep
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/r2_launches_n215.csv python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu --skip-aux > gpurun_out/r2_bench9_under_ncu.json 2> gpurun_out/r2_bench9_under_ncu.err
python tools/ncu_summary.py gpurun_out/r2_launches_n215.csv > gpurun_out/r2_launch_summary_n215.csv; head -16 gpurun_out/r2_launch_summary_n215.csv | cut -c1-120; rm -f gpurun_out/r2_launches_n215.csv
echo "== n = 368 (100.5 M rows), constant alpha"
LVPP_GMRES_RESTART=20 timeout 600 python tools/full_solve.py --verbose --tag n368 --size 368 --alpha-scheme constant --tol 1e-6 > gpurun_out/r2_full368.json 2> gpurun_out/r2_full368.err; grep -E "^outer" gpurun_out/r2_full368.err | cut -c1-140 | tail -16; tail -1 gpurun_out/r2_full368.json | cut -c1-600; tail -2 gpurun_out/r2_full368.err | cut -c1-300
