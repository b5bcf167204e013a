#!/bin/bash
# quick validation of a kernel change: parity tests, level-6 probe with phase profile, per-iteration debug, warm launch list
mkdir -p gpurun_out
TAG=${1:-r2n}
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
echo "== L6 $(timeout 200 python scripts/variant_probe.py 6 2>&1 | tail -1 | cut -c1-600)"
ZB_DEBUG=1 timeout 200 python scripts/one_deflate.py 1 2>&1 | grep "^iter" | cut -c1-160
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1
python scripts/summarize_profile.py $TAG > /dev/null 2>&1
python - <<PY
import json
d=json.load(open('profiles/${TAG}_summary.json'))['launch_list']
print('warm total', d['total_ms'])
for k,v in d['kernels'].items(): print(' ',k, v['launches'], v['total_ms'], v['per_launch_ms'][:13])
PY
