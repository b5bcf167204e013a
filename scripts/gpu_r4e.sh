#!/bin/bash
# last call of the round: whole GPU suite (no -x), the k_match capture that pins profiles/k_match_traffic.json to this build, the bench line
mkdir -p gpurun_out
TAG=r4e
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 300 python -m pytest tests -q -m gpu --timeout 100 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-400
timeout 120 ncu --set full --clock-control none --import-source on -k regex:^k_match -s 0 -c 1 -f -o gpurun_out/prof_k_match_$TAG python scripts/one_deflate.py 1 6 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$TAG.json
