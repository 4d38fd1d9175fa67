#!/bin/bash
# round 2, GPU call 12: the psi-increase bound (CI schedule at n = 215, whole solve at n = 368), mixed-form engine at larger sizes
mkdir -p gpurun_out
echo "== gpu tests (new)"
(timeout 600 python -m pytest tests/test_gpu_mg.py -m gpu -q -rxXf 2>&1 | tail -5) | tee gpurun_out/r2_tests12.txt
run() {  # tag, args...
  tag=$1; shift
  timeout 900 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.2f" % d["solve_s"], "inc", ["%.1e" % v for v in d["history"]["primal_increment"]][-3:])' 2>&1 | tail -1)"
}
run ci215_cap2 --size 215 --psi-cap 2
run ci215_cap4 --size 215 --psi-cap 4
run ci271_cap2 --size 271 --nz 272 --psi-cap 2
LVPP_GMRES_RESTART=20 run const368_cap2 --size 368 --alpha-scheme constant --tol 1e-6 --psi-cap 2
LVPP_GMRES_RESTART=20 run ci368_cap2 --size 368 --psi-cap 2
grep -E "^outer [23] " gpurun_out/r2_scan_ci215_cap2.err | cut -c1-130 | head -20
echo "== mixed-form engine"
for cfg in "gradient 200" "signorini 32" "multiphase 100"; do
  set -- $cfg
  timeout 400 python bench.py --workload $1 --size $2 > gpurun_out/r2_forms_$1_$2.json 2> gpurun_out/r2_forms_$1_$2.err; echo "-- $1 $2: $(tail -1 gpurun_out/r2_forms_$1_$2.json | cut -c1-500)"; tail -2 gpurun_out/r2_forms_$1_$2.err | cut -c1-200
done
