#!/bin/bash
# New-kernel check: memcheck on small cases, the gpu test-suite, one bench line.
mkdir -p gpurun_out
TAG=${1:-v2}
echo "== sanitizer"; timeout 200 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitizer_$TAG.log 2>&1; echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|ok|crc" gpurun_out/sanitizer_$TAG.log | head -12
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log
