#!/bin/bash
# round-2 loop: gpu tests, bench line, per-iteration debug counters, ncu launch list
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_$TAG.json")); r=j["roofline"]
    print("value GiB/s", round(j["value"],4), "ms/step", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],4), "iters", r["iterations"], "launches", j["gpu_launches"])
    print("phases", r["phases_ms"])
    print("other", json.dumps(j.get("other_configs"))[:1500])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
ZB_DEBUG=1 timeout 300 python scripts/one_deflate.py 2 > gpurun_out/debug_$TAG.log 2>&1; tail -30 gpurun_out/debug_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_deflate.py > gpurun_out/ncu_list_$TAG.log 2>&1; tail -1 gpurun_out/ncu_list_$TAG.log
python scripts/summarize_profile.py $TAG 2>/dev/null | head -40
