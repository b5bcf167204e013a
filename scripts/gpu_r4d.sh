#!/bin/bash
mkdir -p gpurun_out
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_r4d.log 2>&1 || { echo SMOKE FAILED; exit 1; }
timeout 200 python scripts/gpu_periodic.py 2>&1 | tee gpurun_out/periodic_r4d.log | tail -16
timeout 100 python - <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import zlib_rs_b200 as Z
from corpus import calgary_mix
e = Z.Engine(0); d = calgary_mix()
p = e.alloc(len(d)); e.to_device(p, d)
cap = Z.lib().zb_deflate_bound(len(d)) + 64; q = e.alloc(cap)
for lv in (6,):
    best = 1e9
    for _ in range(4):
        _, r = e.deflate(p, n=len(d), level=lv, src_on_device=True, dst=q, dst_cap=cap, dst_on_device=True)
        best = min(best, r.gpu_ms)
    print("calgary-mix 64 MiB level", lv, "best_ms", round(best, 2), "GiB/s", round(len(d)/best/1e-3/2**30, 3), "iters", r.iterations, "launches", r.gpu_launches)
PY
