#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3l.log 2>&1 || { echo "SMOKE L9 FAILED"; exit 1; }
bash scripts/gpu_sweep.sh r3l 9 | cut -c1-140
