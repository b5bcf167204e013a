#!/bin/bash
# crc32 kernel check: checksum / gzip / inflate tests, the 8 GiB checksum bench, one ncu --set full capture of k_crc_partial (1 GiB).
mkdir -p gpurun_out
TAG=${1:-crc}
echo "== pytest"; timeout 600 python -m pytest tests -q -m gpu --timeout 300 -k "checksum or inflate or golden or uncompress or stitch or flush or strategies" > gpurun_out/pytest_crc_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_crc_$TAG.log
echo "== bench 8 GiB"; timeout 300 python scripts/bench_checksum.py 8 > gpurun_out/checksum_$TAG.json 2> gpurun_out/checksum_$TAG.err; echo "rc=$?"; cat gpurun_out/checksum_$TAG.json; tail -3 gpurun_out/checksum_$TAG.err
echo "== ncu full k_crc_partial"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_crc_partial -s 3 -c 1 -f -o gpurun_out/prof_k_crc_$TAG python scripts/bench_checksum.py 1 > gpurun_out/ncu_crc_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_crc_$TAG.log
