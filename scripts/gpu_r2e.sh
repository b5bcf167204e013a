#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2e}
bash scripts/gpu_sweep.sh $TAG
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -k "inflate or uncompress or stream or golden" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 300 python scripts/bench_inflate.py 2>&1 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list_$TAG.log 2>&1; tail -1 gpurun_out/ncu_list_$TAG.log
python scripts/summarize_profile.py $TAG > /dev/null 2>&1; python - <<PY
import json
d=json.load(open('profiles/${TAG}_summary.json'))
for k,v in d['launch_list']['kernels'].items(): print(k, v['launches'], v['total_ms'], v['per_launch_ms'])
PY
