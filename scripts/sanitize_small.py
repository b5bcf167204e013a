"""compute-sanitizer target: small deflate (levels 1, 2, 4, 6, 9, memLevel 1/9), inflate (parallel + serial) and checksum calls."""
import sys, os, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlib_rs_b200 as Z
from corpus import synthetic_mix, silesia_member
e = Z.Engine(0)
cases = [synthetic_mix(150000, 4), bytes(200000) + synthetic_mix(70000, 5), silesia_member(1)[:200000], silesia_member(9)[:140000], b"abc" * 30000]
for d in cases:
    for level in (6, 9, 4, 1, 2):
        out, r = e.deflate(d, level=level)
        assert zlib.decompress(out) == d, (len(d), level)
    for mem in (1, 9):
        out, r = e.deflate(d, level=6, mem_level=mem)
        assert zlib.decompress(out) == d, (len(d), mem)
    print("deflate ok", len(d), flush=True)
big = silesia_member(7)[:700000]
comp = zlib.compress(big, 6)
print("inflate input", len(comp))
rc, got, r = e.inflate(comp, len(big))
assert rc == 0 and got == big
rc, got, r = e.inflate(zlib.compress(cases[0], 9), len(cases[0]))
assert rc == 0 and got == cases[0]
print("inflate ok", r.gpu_launches)
print("crc", Z.crc32(big) == zlib.crc32(big), "adler", Z.adler32(big) == zlib.adler32(big))
