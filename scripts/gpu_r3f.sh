#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3f.log 2>&1 || { echo "SMOKE L9 FAILED"; tail -3 gpurun_out/smoke_r3f.log; exit 1; }
tail -2 gpurun_out/smoke_r3f.log | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -x -k "level9 or slow or config4 or dictionary" > gpurun_out/pytest_gpu_r3f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3f.log | cut -c1-300
for lv in 9 8 7; do echo "== L$lv $(timeout 100 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-260)"; done
