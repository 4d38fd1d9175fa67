#!/bin/bash
# First GPU call of the next round (one B200; up to about 35 minutes when every full solve runs into its timeout --
# split it if the budget is tight: tests + bench + scans are about 12 minutes): everything that was changed after
# the last GPU minute of round 1, then the open question of DESIGN.md 7a.
#   gpurun --timeout 2400 -- tools/round2_first_run.sh
# Read first: gpurun_out/r2_tests.txt -- the tests of tests/test_zz_gpu_linesearch.py are non-strict xfail; every
# XPASS there can be made a plain test, every XFAIL is a defect of code written without a GPU.
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -15) | tee gpurun_out/r2_tests.txt
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -1 gpurun_out/r2_bench.json | cut -c1-300
echo "== experimental: one reduction per GMRES iteration"
tools/mg_scan.sh "LVPP_GMRES_FUSED_NORM=0" "LVPP_GMRES_FUSED_NORM=1" "LVPP_GMRES_WEIGHT=auto" "LVPP_GMRES_WEIGHT=auto LVPP_MG_CHEB=10" "LVPP_MG_CHEB=10" "LVPP_MG_CHEB=0"
echo "== experimental: bf16 pair records in the cycle (10 instead of 16 bytes per slot; tools/mg_precision.py: same iteration counts)"
tools/mg_scan.sh "LVPP_MG_PACK=bf16" "LVPP_MG_PACK=bf16 LVPP_MG_UNROLL=8" "LVPP_MG_PACK=bf16 LVPP_GMRES_WEIGHT=auto" "LVPP_GMRES_FLEXIBLE=1" "LVPP_GMRES_FLEXIBLE=1 LVPP_MG_PACK=bf16 LVPP_GMRES_WEIGHT=auto"
export LVPP_GMRES_WEIGHT=auto   # the equilibrated residual norm: 17-26 Krylov iterations through the whole 64^3 CPU emulation
for cfg in "--linesearch none" "--linesearch bt" "--linesearch none --snes-rtol 1e-9" "--linesearch bt --snes-rtol 1e-9"; do
  tag=$(echo "$cfg" | tr -d ' -' )
  echo "== full solve n=215 $cfg"
  LVPP_MG_VERBOSE=1 timeout 200 python tools/full_solve.py --size 215 --verbose $cfg > gpurun_out/r2_full215_$tag.json 2> gpurun_out/r2_full215_$tag.err
  grep -E "^outer|Chebyshev ratio|retrying|Error" gpurun_out/r2_full215_$tag.err | tail -40
  tail -1 gpurun_out/r2_full215_$tag.json | cut -c1-600
done
echo "== full solve n=215, failure-recovering alpha control (recovery.py)"
LVPP_MG_VERBOSE=1 timeout 300 python tools/full_solve.py --size 215 --verbose --alpha-scheme adaptive > gpurun_out/r2_full215_adaptive.json 2> gpurun_out/r2_full215_adaptive.err
grep -E "^outer|retrying|Error" gpurun_out/r2_full215_adaptive.err | tail -40
tail -1 gpurun_out/r2_full215_adaptive.json | cut -c1-800
