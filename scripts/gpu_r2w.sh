#!/bin/bash
TAG=${1:-r2w}
bash scripts/gpu_r2u.sh $TAG
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$TAG.json
