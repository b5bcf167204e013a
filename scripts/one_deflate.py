"""One device-resident level-6 deflate of silesia-small.tar (for ncu launch lists / captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlib_rs_b200 as Z
from corpus import silesia_tar
e = Z.Engine(0)
d = silesia_tar()
p = e.alloc(len(d)); e.to_device(p, d)
cap = Z.lib().zb_deflate_bound(len(d)) + 64
q = e.alloc(cap)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
e.set_profile(True)
for _ in range(reps):
    _, r = e.deflate(p, n=len(d), level=level, src_on_device=True, dst=q, dst_cap=cap, dst_on_device=True)
print("level", level, "out", r.out_bytes, "gpu_ms", r.gpu_ms, "iters", r.iterations, "launches", r.gpu_launches)
print("phases", e.get_profile())
