#!/bin/bash
# full check: gpu test-suite, 8 GiB checksum bench, bench line, ncu capture of the crc kernel
mkdir -p gpurun_out
TAG=${1:-all}
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$TAG.log
echo "== checksum 8 GiB"; timeout 300 python scripts/bench_checksum.py 8 > gpurun_out/checksum_$TAG.json 2> gpurun_out/checksum_$TAG.err; echo "rc=$?"; cat gpurun_out/checksum_$TAG.json; tail -3 gpurun_out/checksum_$TAG.err
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== ncu full k_crc_partial"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_crc_partial -s 3 -c 1 -f -o gpurun_out/prof_k_crc_$TAG python scripts/bench_checksum.py 1 > gpurun_out/ncu_crc_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_crc_$TAG.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
