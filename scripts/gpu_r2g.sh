#!/bin/bash
# 2-GPU check of bench.py's N>1 path (weak-scaling headline + chunk-sharded single streams) and both arms' JSON lines
mkdir -p gpurun_out
TAG=${1:-r2g}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_$TAG.json 2> gpurun_out/bench2_$TAG.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench2_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/bench2_$TAG.json") if l.startswith("{")][-1])
    print("N=2 value", round(j["value"],4), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],4), "launches", j["gpu_launches"])
    print(json.dumps(j["other_configs"])[:2500])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-600
timeout 300 python scripts/bench_inflate.py 2>&1 | grep "reference file"
