#!/bin/bash
# round 2, GPU call 1: which residual norm does the Krylov stopping test need at n = 215?  (tools/diag_solve.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for cfg in "euclid:" "weight:LVPP_GMRES_WEIGHT=auto" "weight_plain:LVPP_GMRES_WEIGHT=auto LVPP_MG_CHEB=0"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "== diag n=215 $tag ($envs)"
  env $envs LVPP_MG_VERBOSE=1 timeout 420 python tools/diag_solve.py --size 215 > gpurun_out/r2_diag215_$tag.json 2> gpurun_out/r2_diag215_$tag.err
  grep -E "^outer|^==|plain damping|retrying|Error|error" gpurun_out/r2_diag215_$tag.err | cut -c1-260 | tail -45
  tail -1 gpurun_out/r2_diag215_$tag.json | cut -c1-400
done
echo "== diag n=32 against host LU, both norms"
timeout 300 python tools/diag_solve.py --size 32 --lu > gpurun_out/r2_diag32_lu_euclid.json 2> gpurun_out/r2_diag32_lu_euclid.err
grep -E "^outer|^==" gpurun_out/r2_diag32_lu_euclid.err | cut -c1-330 | tail -30
LVPP_GMRES_WEIGHT=auto timeout 300 python tools/diag_solve.py --size 32 --lu > gpurun_out/r2_diag32_lu_weight.json 2> gpurun_out/r2_diag32_lu_weight.err
grep -E "^outer|^==" gpurun_out/r2_diag32_lu_weight.err | cut -c1-330 | tail -30
echo "== gpu tests"
(timeout 600 python -m pytest tests -m gpu -q -rxX -x --deselect "tests/test_gpu_mg.py::test_full_lvpp_solve_mg_matches_oracle" 2>&1 | tail -25) | tee gpurun_out/r2_tests1.txt
