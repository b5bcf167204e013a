#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2f}
echo "== main lib"; python scripts/variant_probe.py 2>&1 | tail -1
ZB_CASE="reference" ZB_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_inf_$TAG.csv python scripts/bench_inflate.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_inf_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
print("inflate launches:", [(r[4].split('(')[0], round(int(r[-1].replace(',',''))/1e6,3)) for r in rows])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_l9_$TAG.csv python scripts/one_deflate.py 1 9 > gpurun_out/ncu_l9_$TAG.log 2>&1; tail -2 gpurun_out/ncu_l9_$TAG.log | cut -c1-300
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_l9_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
print("L9 launches:", [(r[4].split('(')[0], round(int(r[-1].replace(',',''))/1e6,3)) for r in rows])
PY
python scripts/variant_probe.py 9 2>&1 | tail -1
