#!/bin/bash
# ncu evidence for round 2 (one B200, about 6 minutes; never a bench number): launch list of one bench run and a full
# capture of the cycle's fine-level sweep, default build and the bf16 pair records.
#   gpurun --timeout 900 -- tools/round2_ncu.sh
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-aux --skip-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_n215.csv $BENCH > gpurun_out/r02_launches.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_launches_n215.csv > gpurun_out/r02_launch_summary_n215.csv 2>/dev/null || true
for mode in fp32 bf16; do
  if [ $mode = bf16 ]; then export LVPP_MG_PACK=bf16; kern=k_packed2_op; else unset LVPP_MG_PACK; kern=k_packed_op; fi
  ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip 40 -c 2 -f -o gpurun_out/r02_${kern}_n215 $BENCH > gpurun_out/r02_${kern}.log 2>&1
  ncu -i gpurun_out/r02_${kern}_n215.ncu-rep --page raw --csv > gpurun_out/r02_${kern}_n215_ncu_raw.csv 2>/dev/null
  python - "$kern" <<'PY'
import csv, json, sys
kern = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r02_{kern}_n215_ncu_raw.csv")))
hdr = rows[0]
want = {"dram__bytes_read.sum": None, "dram__bytes_write.sum": None, "gpu__time_duration.sum": None}
idx = {k: hdr.index(k) for k in want if k in hdr}
units = rows[1]
out = []
for r in rows[2:]:
    out.append({k: (r[i], units[i]) for k, i in idx.items()})
print(kern, json.dumps(out))
PY
done
