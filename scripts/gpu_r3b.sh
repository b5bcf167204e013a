#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r3b}
timeout 90 python scripts/one_deflate.py 1 > gpurun_out/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/smoke_$TAG.log; exit 1; }
timeout 60 python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import zlib, zlib_rs_b200 as Z
from corpus import silesia_tar
d = silesia_tar()
out = Z.compress2(d, 6)
assert zlib.decompress(out) == d
import hashlib; print('compress2 host path ok', len(out), hashlib.sha256(out).hexdigest()[:8])
" || { echo "HOST PATH FAILED"; exit 1; }
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
print('roofline', d.get('roofline')); print('clocks', d.get('clocks'))
oc=d.get('other_configs',{})
for k,v in oc.items(): print(' ', k, json.dumps(v)[:300])
PY
