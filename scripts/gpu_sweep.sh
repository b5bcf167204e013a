#!/bin/bash
# probe every variant under zlib_rs_b200/variants (scripts/variant_probe.py): golden check + timings
mkdir -p gpurun_out
for so in zlib_rs_b200/variants/libz_b200_*.so; do
  name=$(basename $so .so); name=${name#libz_b200_}
  mt=""; [[ $name =~ _mt([0-9]+) ]] && mt=${BASH_REMATCH[1]}   # a variant named *_mt768 runs its dense passes with 768 threads per CTA
  if [ -n "$mt" ]; then export ZB_MTHREADS=$mt; else unset ZB_MTHREADS; fi
  ms=""; [[ $name =~ _ms([0-9]+) ]] && ms=${BASH_REMATCH[1]}   # *_ms16384: dense passes with 16384 positions per CTA
  if [ -n "$ms" ]; then export ZB_MSUB=$ms; else unset ZB_MSUB; fi
  fs=""; [[ $name =~ _fs([0-9]+) ]] && fs=${BASH_REMATCH[1]}   # *_fs16384: first pass with 16384 positions per CTA
  if [ -n "$fs" ]; then export ZB_MSUB1=$fs; else unset ZB_MSUB1; fi
  echo "== $name $(ZB_LIB_PATH=$PWD/$so timeout 60 python scripts/variant_probe.py ${2:-6} 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/sweep_${1:-s}.log
