#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2i}
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
echo "== main lib"; python scripts/variant_probe.py 2>&1 | tail -1
timeout 300 python scripts/bench_inflate.py 2>&1 | tail -5
ZB_CASE="reference" ZB_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_inf_$TAG.csv python scripts/bench_inflate.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_inf_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
print("inflate launches:", [(r[4].split('(')[0], round(int(r[-1].replace(',',''))/1e6,3)) for r in rows][8:16])
PY
