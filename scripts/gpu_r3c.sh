#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_r3c.log 2>&1 || { echo "SMOKE FAILED"; exit 1; }
timeout 100 python scripts/e2e_probe.py
ZB_UPLOAD_CHUNKED=0 timeout 100 python scripts/e2e_probe.py
timeout 100 python scripts/e2e_probe.py
ZB_UPLOAD_CHUNKED=0 timeout 100 python scripts/e2e_probe.py
