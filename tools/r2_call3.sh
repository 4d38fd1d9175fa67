#!/bin/bash
# round 2, GPU call 3: the device-resident GMRES (tests, bench), and which alpha schedules / sizes complete with the full Newton step
mkdir -p gpurun_out
echo "== gpu tests"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -15) | tee gpurun_out/r2_tests3.txt
timeout 300 python -m pytest "tests/test_gpu_mg.py::test_full_lvpp_solve_mg_matches_oracle" -q 2>&1 | grep -E "where False|assert|passed|failed" | cut -c1-1500
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2_bench3_default.json 2> gpurun_out/r2_bench3_default.err; tail -1 gpurun_out/r2_bench3_default.json | cut -c1-700; tail -3 gpurun_out/r2_bench3_default.err
echo "== bench driver args"
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/r2_bench3_s20.json 2> gpurun_out/r2_bench3_s20.err; tail -1 gpurun_out/r2_bench3_s20.json | cut -c1-900; tail -3 gpurun_out/r2_bench3_s20.err
run() {  # tag, args...
  tag=$1; shift
  timeout 400 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(grep -c '^outer' gpurun_out/r2_scan_$tag.err) newton steps; $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.1f" % d["solve_s"], "inc", ["%.1e" % v for v in d["history"]["primal_increment"]])' 2>&1 | tail -1)"
}
run const_215 --size 215 --alpha-scheme constant
run geom_215 --size 215 --alpha-scheme geometric
run none_128 --size 128
run none_160 --size 160
run none_96 --size 96
run jacobi_160 --size 160 --pc jacobi
