#!/bin/bash
# round 2, GPU call 13: psi bound with free rise below 0; forms tests (blocked NonlinearProblem)
mkdir -p gpurun_out
echo "== gpu tests (forms + mg)"
(timeout 600 python -m pytest tests/test_gpu_mg.py tests/test_gpu_forms.py -m gpu -q -rxXf 2>&1 | tail -5) | tee gpurun_out/r2_tests13.txt
run() {  # tag, args...
  tag=$1; shift
  timeout 900 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.2f" % d["solve_s"], "inc", ["%.1e" % v for v in d["history"]["primal_increment"]][-3:])' 2>&1 | tail -1)"
}
run ci215_free0_cap1 --size 215 --psi-cap 1 --psi-free 0
run ci215_free0_cap2 --size 215 --psi-cap 2 --psi-free 0
run ci215_freem2_cap2 --size 215 --psi-cap 2 --psi-free -2
run const215_free0_cap1 --size 215 --alpha-scheme constant --tol 1e-6 --psi-cap 1 --psi-free 0
LVPP_GMRES_RESTART=20 run const368_free0_cap1 --size 368 --alpha-scheme constant --tol 1e-6 --psi-cap 1 --psi-free 0
LVPP_GMRES_RESTART=20 run ci368_free0_cap1 --size 368 --psi-cap 1 --psi-free 0
