#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-t}
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
