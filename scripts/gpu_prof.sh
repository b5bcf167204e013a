#!/bin/bash
# quick loop + one full ncu capture of a kernel (regex $2, launch skip $3)
TAG=${1:-p}; KRN=${2:-k_match}; SKIP=${3:-0}
bash scripts/gpu_quick.sh $TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRN -s $SKIP -c 1 -f -o gpurun_out/prof_${KRN}_$TAG python scripts/one_deflate.py > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
