#!/bin/bash
# one ncu --set full capture (with source) of a kernel: TAG KERNEL_REGEX SKIP [level]
mkdir -p gpurun_out
TAG=${1:-n}; KRN=${2:-k_match}; SKIP=${3:-0}; LEVEL=${4:-6}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRN -s $SKIP -c 1 -f -o gpurun_out/prof_${KRN}_$TAG python scripts/one_deflate.py 1 $LEVEL > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/prof_${KRN}_$TAG.ncu-rep
