#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3g.log 2>&1 || { echo "SMOKE L9 FAILED"; exit 1; }
for lv in 9 8 7; do bash scripts/gpu_sweep.sh r3g_L$lv $lv | cut -c1-200; done
