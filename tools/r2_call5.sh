#!/bin/bash
# round 2, GPU call 5 (repeats call 4, whose build carried an unstable norm-from-dot-products Gram-Schmidt variant, since reverted)
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -12) | tee gpurun_out/r2_tests5.txt
echo "== bench, driver args"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5_s20.json 2> gpurun_out/r2_bench5_s20.err; tail -1 gpurun_out/r2_bench5_s20.json | cut -c1-300; tail -3 gpurun_out/r2_bench5_s20.err
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(grep -c '^outer' gpurun_out/r2_scan_$tag.err) newton steps; $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.1f" % d["solve_s"], "inc", ["%.1e" % v for v in d["history"]["primal_increment"]])' 2>&1 | tail -1)"
}
run const_215_tol6 --size 215 --alpha-scheme constant --tol 1e-6
LVPP_MG_PACK=bf16 run const_215_bf16 --size 215 --alpha-scheme constant --tol 1e-6
LVPP_MG_CHEB=0 run const_215_plain --size 215 --alpha-scheme constant --tol 1e-6
LVPP_MG_CHEB=10 run const_215_cheb10 --size 215 --alpha-scheme constant --tol 1e-6
LVPP_GMRES_WEIGHT=off run const_215_euclid --size 215 --alpha-scheme constant --tol 1e-6
run const_240 --size 240 --alpha-scheme constant --tol 1e-6
run const_271 --size 271 --nz 272 --alpha-scheme constant --tol 1e-6
run const_320 --size 320 --alpha-scheme constant --tol 1e-6
echo "== solver variants on the overshoot state (n = 160, CI schedule, third proximal step)"
for cfg in "default:" "plain:LVPP_MG_CHEB=0" "fp64plain:LVPP_MG_FP32=0 LVPP_MG_CHEB=0" "over1:LVPP_MG_OVER=1.0 LVPP_MG_CHEB=0" "euclid:LVPP_GMRES_WEIGHT=off LVPP_MG_CHEB=0" "restart150:LVPP_GMRES_RESTART=150 LVPP_MG_CHEB=0"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs LVPP_MG_VERBOSE=1 timeout 200 python tools/diag_solve.py --size 160 --ksp-max-it 600 --max-newton 8 > gpurun_out/r2_diag160_$tag.json 2> gpurun_out/r2_diag160_$tag.err
  echo "-- $tag ($envs)"; grep -E "^outer 2|^== outer 2" gpurun_out/r2_diag160_$tag.err | cut -c1-230 | tail -12
done
