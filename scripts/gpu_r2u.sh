#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2u}
# a hang must not eat the budget: one deflate first, with a short limit
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
bash scripts/gpu_sweep.sh $TAG 6 | cut -c1-420
CNT=zlib_rs_b200/variants/libz_b200_cnt.so; [ -f $CNT ] && export ZB_LIB_PATH=$PWD/$CNT; ZB_DEBUG=1 timeout 200 python scripts/one_deflate.py 1 2>&1 | grep "^iter\|walks" | cut -c1-160; unset ZB_LIB_PATH
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1
python scripts/summarize_profile.py $TAG > /dev/null 2>&1
python - <<PY
import json
d=json.load(open('profiles/${TAG}_summary.json'))['launch_list']
print('warm total', d['total_ms'])
for k,v in d['kernels'].items(): print(' ',k, v['launches'], v['total_ms'], v['per_launch_ms'][:13])
PY
