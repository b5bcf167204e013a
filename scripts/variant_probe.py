"""One kernel variant (ZB_LIB_PATH): level-6 deflate of silesia-small.tar, output checked against the golden hash, phase profile."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlib_rs_b200 as Z
from corpus import silesia_tar
GOLD = "939f96c8"
e = Z.Engine(0)
d = silesia_tar()
level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
out, r = e.deflate(d, level=level)
ok = hashlib.sha256(out).hexdigest().startswith(GOLD) if level == 6 else None
p = e.alloc(len(d)); e.to_device(p, d)
cap = Z.lib().zb_deflate_bound(len(d)) + 64
q = e.alloc(cap)
e.set_profile(True)
for _ in range(2):
    _, r = e.deflate(p, n=len(d), level=level, src_on_device=True, dst=q, dst_cap=cap, dst_on_device=True)
prof = {k: round(v["ms"], 3) for k, v in e.get_profile().items() if v["ms"] > 0}
e.set_profile(False)
best = 1e9
for _ in range(5):
    _, r = e.deflate(p, n=len(d), level=level, src_on_device=True, dst=q, dst_cap=cap, dst_on_device=True)
    best = min(best, r.gpu_ms)
print("golden_ok", ok, "out", r.out_bytes, "best_ms", round(best, 3), "iters", r.iterations, prof)
