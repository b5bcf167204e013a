#!/bin/bash
# ncu --set full captures of single launches of one level-6 deflate: "kernel:skip" pairs
mkdir -p gpurun_out
TAG=${1:-r2s}; shift
for spec in "$@"; do
  k=${spec%%:*}; sk=${spec#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$k -s $sk -c 1 -f -o gpurun_out/prof_${k}_s${sk}_$TAG python scripts/one_deflate.py 1 6 > gpurun_out/ncu_full_${k}_$TAG.log 2>&1; echo "$k skip $sk ncu rc=$?"
done
ls -la gpurun_out/*_$TAG.ncu-rep
