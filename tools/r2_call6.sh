#!/bin/bash
# round 2, GPU call 6: tests with the bf16 records as default; how the late-solve Krylov counts grow with the mesh; knobs
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -12) | tee gpurun_out/r2_tests6.txt
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.2f" % d["solve_s"])' 2>&1 | tail -1)"
}
for n in 48 64 96 128 160; do run late_$n --size $n --alpha-scheme constant --tol 1e-5; done
echo "== knobs at n = 128 (constant alpha, tol 1e-5)"
for cfg in "over1.4:LVPP_MG_OVER=1.4" "over2.2:LVPP_MG_OVER=2.2" "v33:LVPP_MG_NSMOOTH=3" "v13:LVPP_MG_NPRE=1 LVPP_MG_NPOST=3" "plain:LVPP_MG_CHEB=0" "cheb3:LVPP_MG_CHEB=3" "fp64:LVPP_MG_FP32=0" "restart100:LVPP_GMRES_RESTART=100" "margin1.3:LVPP_MG_MARGIN=1.3"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  export $envs
  run knob128_$tag --size 128 --alpha-scheme constant --tol 1e-5
  for e in $envs; do unset ${e%%=*}; done
done
echo "== bench, driver args (bf16 records)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6_s20.json 2> gpurun_out/r2_bench6_s20.err; tail -1 gpurun_out/r2_bench6_s20.json | cut -c1-300; tail -3 gpurun_out/r2_bench6_s20.err
echo "== bench 2-D (configs[0])"
timeout 900 python bench.py --workload obstacle2d --steps 20 --warmup 5 > gpurun_out/r2_bench6_2d.json 2> gpurun_out/r2_bench6_2d.err; tail -1 gpurun_out/r2_bench6_2d.json | cut -c1-400; tail -3 gpurun_out/r2_bench6_2d.err
