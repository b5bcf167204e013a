#!/bin/bash
mkdir -p gpurun_out
for so in zlib_rs_b200/variants/libz_b200_*.so; do
  name=$(basename $so .so); name=${name#libz_b200_}
  for lv in 9 7 8; do echo "== $name L$lv $(ZB_LIB_PATH=$PWD/$so timeout 200 python scripts/variant_probe.py $lv 2>&1 | tail -1 | cut -c1-160)"; done
done 2>&1 | tee gpurun_out/sweep_slow.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/prof_k_match_r2k python scripts/one_deflate.py > gpurun_out/ncu_full_r2k.log 2>&1; tail -1 gpurun_out/ncu_full_r2k.log
