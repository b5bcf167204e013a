"""Exactness and cost of the hole fixed point on its worst input class (periodic data with mutations)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlib_rs_b200 as Z
import oracle_lib as O
from corpus import periodic_mutated
e = Z.Engine(0)
bad = 0
for (n, period, nmut, seed) in ((200003, 222, 30, 1), (200003, 74, 20, 2), (300000, 500, 10, 3), (150000, 37, 40, 4), (400000, 1000, 25, 5)):
    d = periodic_mutated(n, period, nmut, seed)
    for level in (3, 5, 6):
        out, r = e.deflate(d, level=level)
        ok = out == O.compress(d, level)[1]
        bad += not ok
        print("n", n, "period", period, "level", level, "exact", ok, "iters", r.iterations, "gpu_ms", round(r.gpu_ms, 2), flush=True)
print("bad", bad)
