#!/bin/bash
run() {
  echo "== FP32=$1"
  LVPP_MG_FP32=$1 timeout 200 python bench.py --steps 4 --warmup 3 --skip-cpu --skip-aux --skip-e2e 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('ms_per_step', round(d['ms_per_step'],1), 'krylov', d['krylov_iterations'], 'launches', d['gpu_launches'], 'hist', d['outer_history'])
except Exception as e:
    print('FAILED', e)
"
}
run 1
run 0
timeout 600 python -m pytest tests/test_gpu_mg.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
