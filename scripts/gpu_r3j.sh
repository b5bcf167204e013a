#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3k.log 2>&1 || { echo "SMOKE L9 FAILED"; tail -3 gpurun_out/smoke_r3k.log; exit 1; }
tail -2 gpurun_out/smoke_r3k.log | cut -c1-160
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -q -m gpu --timeout 120 -x -k "level9 or slow or config4 or dictionary or levels" > gpurun_out/pytest_gpu_r3k.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3k.log | cut -c1-300
for lv in 9; do echo "== L$lv $(timeout 100 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-200)"; done
