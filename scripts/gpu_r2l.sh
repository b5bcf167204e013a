#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r2l.log | cut -c1-300
for lv in 9 7 8; do echo "== L$lv $(timeout 200 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-200)"; done
