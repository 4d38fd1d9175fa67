#!/bin/bash
# usage: scan.sh "VAR=val VAR=val" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 4 --warmup 3 --skip-cpu --skip-aux --skip-e2e 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('ms_per_step', round(d['ms_per_step'],1), 'krylov', d['krylov_iterations'], 'launches', d['gpu_launches'], 'hist', d['outer_history']['newton_steps'])
except Exception as e:
    print('FAILED', e)
"
done
