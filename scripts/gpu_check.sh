#!/bin/bash
# First-contact GPU run: smoke, sanitizer on a small case, the gpu test-suite, a phase profile.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sanitizer (small)"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python - > gpurun_out/sanitizer.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'tests')
import zlib, zlib_rs_b200 as Z
from corpus import synthetic_mix
d = synthetic_mix(150000, seed=4)
out = Z.compress2(d, 6)
print("deflate ok", len(out), zlib.decompress(out) == d)
print("inflate ok", Z.uncompress(out, len(d)) == d)
print("crc", Z.crc32(d) == zlib.crc32(d), "adler", Z.adler32(d) == zlib.adler32(d))
PY
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|deflate ok|inflate ok|crc" gpurun_out/sanitizer.log | head -20
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== profile"; timeout 300 python - > gpurun_out/profile.log 2>&1 <<'PY'
import sys, time; sys.path.insert(0,'tests')
import zlib_rs_b200 as Z
from corpus import silesia_tar
e = Z.Engine(0); d = silesia_tar()
for rep in range(3):
    e.set_profile(rep == 2)
    t = time.time(); out, res = e.deflate(d, level=6); dt = time.time() - t
    print("rep", rep, "out", len(out), "gpu_ms", res.gpu_ms, "wall_ms", dt * 1e3, "iters", res.iterations, "launches", res.gpu_launches, "syms", res.n_symbols, "blocks", res.n_blocks)
print(e.get_profile())
import zlib
t=time.time(); rc, o, r = e.inflate(out, len(d)); print("inflate rc", rc, "ok", o == d, "gpu_ms", r.gpu_ms)
PY
cat gpurun_out/profile.log | tail -8
