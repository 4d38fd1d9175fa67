#!/bin/bash
# round 2, GPU call 7: the state of psi in the late constant-alpha steps at n = 215 / 200 / 208; cycle shapes and coarse scaling at n = 215
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -6) | tee gpurun_out/r2_tests7.txt
for n in 215 200 208; do
  timeout 300 python tools/diag_solve.py --size $n --alpha-scheme constant --max-outer 9 --tol-exit 1e-6 > gpurun_out/r2_diagconst_$n.json 2> gpurun_out/r2_diagconst_$n.err
  echo "-- const alpha n=$n"; grep -E "^outer" gpurun_out/r2_diagconst_$n.err | cut -c1-30,60-75,215-290 | tail -24
done
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.2f" % d["solve_s"])' 2>&1 | tail -1)"
}
echo "== knobs at n = 215 (constant alpha, tol 1e-5)"
for cfg in "default:LVPP_DUMMY=1" "v13:LVPP_MG_NPRE=1 LVPP_MG_NPOST=3" "v14:LVPP_MG_NPRE=1 LVPP_MG_NPOST=4" "v23:LVPP_MG_NPRE=2 LVPP_MG_NPOST=3" "v33:LVPP_MG_NSMOOTH=3" "unroll8:LVPP_MG_UNROLL=8" "ks1.8:LVPP_MG_KSCALE=1.8 LVPP_MG_OVER=1.0" "ks1.4:LVPP_MG_KSCALE=1.4 LVPP_MG_OVER=1.3" "over1.5:LVPP_MG_OVER=1.5"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  export $envs
  run knob215_$tag --size 215 --alpha-scheme constant --tol 1e-5
  for e in $envs; do unset ${e%%=*}; done
done
