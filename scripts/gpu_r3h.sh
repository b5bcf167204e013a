#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3h.log 2>&1 || { echo "SMOKE L9 FAILED"; tail -3 gpurun_out/smoke_r3h.log; exit 1; }
for ss in 8192 12288 16384 24576; do echo "== L9 sub $ss $(ZB_SLOW_SUB=$ss timeout 100 python scripts/variant_probe.py 9 2>&1 | tail -1 | cut -c1-120)"; done
for ss in 4096 6144; do echo "== L8 sub $ss $(ZB_SLOW_SUB=$ss timeout 100 python scripts/variant_probe.py 8 2>&1 | tail -1 | cut -c1-120)"; done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -x -k "level9 or slow or config4" > gpurun_out/pytest_gpu_r3h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3h.log | cut -c1-300
