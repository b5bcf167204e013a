#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
bash scripts/gpu_sweep.sh $TAG 6 | cut -c1-420
