#!/bin/bash
# inflate loop: timing, per-kernel launch list, inflate tests
mkdir -p gpurun_out
timeout 300 python scripts/bench_inflate.py 2>&1 | tail -8
ZB_CASE="L6 zlib" ZB_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_inf.csv python scripts/bench_inflate.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_inf.csv')) if len(r)>10 and r[0].isdigit()]
print({r[4].split('(')[0]: int(r[-1])/1e6 for r in rows[:8]})
PY
timeout 500 python -m pytest tests -q -m gpu -x --timeout 300 -k "inflate or uncompress" 2>&1 | tail -3
