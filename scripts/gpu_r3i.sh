#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 9 > gpurun_out/smoke_r3i.log 2>&1 || { echo "SMOKE L9 FAILED"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_slow -s 0 -c 1 -f -o gpurun_out/prof_k_slow_L9_r3i python scripts/one_deflate.py 1 9 > gpurun_out/ncu_full_L9_r3i.log 2>&1; echo "L9 ncu rc=$?"
