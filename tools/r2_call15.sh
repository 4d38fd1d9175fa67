#!/bin/bash
# round 2, GPU call 15: HEAD -- full GPU suite, mixed-form lines with the device-resident GMRES, experimental switches, bench
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -6) | tee gpurun_out/r2_tests15.txt
echo "== mixed-form engine"
for cfg in "gradient 200" "signorini 32" "multiphase 100" "signorini 48"; do
  set -- $cfg
  timeout 400 python bench.py --workload $1 --size $2 > gpurun_out/r2_forms2_$1_$2.json 2> gpurun_out/r2_forms2_$1_$2.err; echo "-- $1 $2: $(tail -1 gpurun_out/r2_forms2_$1_$2.json | cut -c1-120) ... $(tail -1 gpurun_out/r2_forms2_$1_$2.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("krylov", d["krylov_iterations"], "ms/step %.1f" % d["ms_per_step"])' 2>&1 | tail -1)"
done
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", sum(d["history"]["newton_steps"]), "krylov", sum(d["history"].get("krylov_iterations")), "solve_s %.2f" % d["solve_s"])' 2>&1 | tail -1)"
}
run sw_default --size 215 --alpha-scheme constant --tol 1e-6
LVPP_GMRES_FLEXIBLE=1 run sw_flexible --size 215 --alpha-scheme constant --tol 1e-6
LVPP_MG_UNROLL=8 run sw_unroll8 --size 215 --alpha-scheme constant --tol 1e-6
run sw_default2 --size 215 --alpha-scheme constant --tol 1e-6
echo "== bench (no flags)"
timeout 900 python bench.py > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; tail -1 gpurun_out/r2_bench15.json | cut -c1-300; tail -2 gpurun_out/r2_bench15.err
