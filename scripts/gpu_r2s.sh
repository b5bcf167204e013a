#!/bin/bash
# ncu --set full captures (first launch) of several kernels of one level-6 deflate
mkdir -p gpurun_out
TAG=${1:-r2s}
for k in k_match k_nxt k_skip k_path_mark; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 0 -c 1 -f -o gpurun_out/prof_${k}_$TAG python scripts/one_deflate.py 1 6 > gpurun_out/ncu_full_${k}_$TAG.log 2>&1; echo "$k ncu rc=$?"
done
ls -la gpurun_out/*_$TAG.ncu-rep
