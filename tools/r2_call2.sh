#!/bin/bash
# round 2, GPU call 2: the linear solves are exact to 1e-15 in both row blocks (call 1) and the full Newton step still
# diverges at n = 215 in the third proximal step -> which meshes / SNES options complete the whole LVPP solve?
mkdir -p gpurun_out
run() {  # tag, args...
  tag=$1; shift
  timeout 300 python tools/full_solve.py --verbose --tag $tag "$@" > gpurun_out/r2_scan_$tag.json 2> gpurun_out/r2_scan_$tag.err
  echo "== $tag: $(grep -c '^outer' gpurun_out/r2_scan_$tag.err) newton steps; $(tail -1 gpurun_out/r2_scan_$tag.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("failure", d["failure"], "newton", d["history"]["newton_steps"], "krylov", d["history"].get("krylov_iterations"), "solve_s %.1f" % d["solve_s"], "inc", ["%.1e" % v for v in d["history"]["primal_increment"]])' 2>&1 | tail -1)"
}
for n in 200 208 212 214 216 218 220 224 240 256; do run none_$n --size $n; done
run bt_215 --size 215 --linesearch bt
run none_215_rtol8 --size 215 --snes-rtol 1e-8
run none_215_rtol10 --size 215 --snes-rtol 1e-10
run bt_215_rtol8 --size 215 --linesearch bt --snes-rtol 1e-8
run bt_216 --size 216 --linesearch bt
run none_271 --size 271 --nz 272
run bt_271 --size 271 --nz 272 --linesearch bt
grep -E "^outer" gpurun_out/r2_scan_bt_215.err | cut -c1-200 | tail -40
echo "== gpu tests (all)"
(timeout 900 python -m pytest tests -m gpu -q -rxXf 2>&1 | tail -40) | tee gpurun_out/r2_tests2.txt
