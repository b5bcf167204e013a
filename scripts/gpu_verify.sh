#!/bin/bash
# Lean round-end style check: gpu test-suite, one bench line, ncu launch list of one level-6 deflate.
mkdir -p gpurun_out
TAG=${1:-v}
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log
