#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
echo "== main lib"; python scripts/variant_probe.py 2>&1 | tail -1
for cfg in "4096 1024" "4096 512" "2048 512" "2048 1024" "8192 512"; do set -- $cfg; echo "== msub $1 threads $2: $(ZB_MSUB=$1 ZB_MTHREADS=$2 timeout 120 python scripts/variant_probe.py 2>&1 | tail -1)"; done
bash scripts/gpu_sweep.sh $TAG
