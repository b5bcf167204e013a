#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2d}
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
echo "== main lib"; python scripts/variant_probe.py 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1]); r=j["roofline"]
    print("value GiB/s", round(j["value"],4), "ms/step", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],4), "iters", r["iterations"], "launches", j["gpu_launches"])
    print("phases", r["phases_ms"]); print("cpu", j.get("cpu_baseline"))
    print("other", json.dumps(j.get("other_configs"))[:3000])
except Exception as e:
    print("bench parse failed", e)
PY
