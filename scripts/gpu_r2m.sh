#!/bin/bash
mkdir -p gpurun_out
TAG=r2m
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
for lv in 9 7 8; do echo "== L$lv $(timeout 200 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-200)"; done
python scripts/bench_inflate.py 2>&1 | tail -4
ZB_INFLATE_BLOCKWISE=1 ZB_CASE=reference python scripts/bench_inflate.py 2>&1 | tail -1
ZB_CASE="reference" ZB_REPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_inf_$TAG.csv python scripts/bench_inflate.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_inf_$TAG.csv')) if len(r)>10 and r[0].isdigit()]
print("inflate launches:", [(r[4].split('(')[0], round(int(r[-1].replace(',',''))/1e6,3)) for r in rows][9:18])
PY
