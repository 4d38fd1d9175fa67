#!/bin/bash
# round 2, GPU call 8: Chebyshev sweeps kept for the whole solve (no fallback to plain damping) x cycle shapes, n = 215
mkdir -p gpurun_out
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.2f" % d["solve_s"])' 2>&1 | tail -1)"
}
for cfg in "v22:LVPP_DUMMY=1" "v22adapt:LVPP_MG_CHEB_ADAPT=1" "v23:LVPP_MG_NPRE=2 LVPP_MG_NPOST=3" "v33:LVPP_MG_NSMOOTH=3" "v33u8:LVPP_MG_NSMOOTH=3 LVPP_MG_UNROLL=8" "v34:LVPP_MG_NPRE=3 LVPP_MG_NPOST=4" "v44:LVPP_MG_NSMOOTH=4" "v24:LVPP_MG_NPRE=2 LVPP_MG_NPOST=4" "v33c10:LVPP_MG_NSMOOTH=3 LVPP_MG_CHEB=10" "v33c4:LVPP_MG_NSMOOTH=3 LVPP_MG_CHEB=4" "v33o15:LVPP_MG_NSMOOTH=3 LVPP_MG_OVER=1.5" "v33o21:LVPP_MG_NSMOOTH=3 LVPP_MG_OVER=2.1" "v22fp32:LVPP_MG_PACK=fp32"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  export $envs
  run k215b_$tag --size 215 --alpha-scheme constant --tol 1e-6
  for e in $envs; do unset ${e%%=*}; done
done
