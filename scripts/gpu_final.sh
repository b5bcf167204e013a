#!/bin/bash
# the round's final measurement batch: tests, bench (both arms), launch list, one full capture of k_match, the other levels, inflate
mkdir -p gpurun_out
TAG=${1:-r3z}
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -3 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 420 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference > gpurun_out/ref_$TAG.json 2> gpurun_out/ref_$TAG.err; echo "reference rc=$?"; cut -c1-600 gpurun_out/ref_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_match -s 0 -c 1 -f -o gpurun_out/prof_k_match_$TAG python scripts/one_deflate.py 1 6 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
for lv in 7 8 9; do echo "== L$lv $(timeout 100 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-260)"; done
timeout 120 python scripts/bench_inflate.py 2>&1 | tail -4 | cut -c1-300
timeout 100 python scripts/e2e_probe.py
