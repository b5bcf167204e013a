#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_r4c.log 2>&1 || { echo SMOKE FAILED; exit 1; }
echo "== default $(timeout 60 python scripts/variant_probe.py 6 2>&1 | tail -1 | cut -c1-150)"
for ms in 1024 2048 8192; do echo "== later passes sub $ms $(ZB_MSUB1=4096 ZB_MSUB=$ms timeout 60 python scripts/variant_probe.py 6 2>&1 | tail -1 | cut -c1-150)"; done
for mt in 512; do echo "== later passes threads $mt $(ZB_MSUB1=4096 ZB_MTHREADS=$mt timeout 60 python scripts/variant_probe.py 6 2>&1 | tail -1 | cut -c1-150)"; done
