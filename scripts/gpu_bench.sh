#!/bin/bash
# pytest (gpu), bench both arms, ncu launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
TAG=${1:-r01}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_list.log
echo "== ncu full k_match"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/prof_k_match_$TAG python scripts/one_deflate.py > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
