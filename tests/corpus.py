"""Corpus helpers shared by tests and bench.py (no code under test is used here)."""
import hashlib
import os
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAR_GZ = os.path.join(ROOT, "data", "silesia-small.tar.gz")
TAR_SHA256 = "b781b5ef664889c9e8690da7a3988b2d65b1574ba9c5a11c5c1dc46bb0c2831d"
MEMBERS = ["ooffice", "nci", "mozilla", "dickens", "webster", "samba", "osdb", "x-ray", "sao", "xml", "mr", "reymont"]
MEMBER_LEN = 1310690
_cache = {}


def silesia_gz():
    return open(TAR_GZ, "rb").read()


def silesia_tar():
    """silesia-small.tar (15,736,320 B), inflated with stock zlib from the committed .tar.gz."""
    if "tar" not in _cache:
        d = zlib.decompress(silesia_gz())
        assert hashlib.sha256(d).hexdigest() == TAR_SHA256
        _cache["tar"] = d
    return _cache["tar"]


def silesia_member(k):
    """k-th member (1,310,690 B); data offsets 1024 + k*1,311,232."""
    off = 1024 + k * 1311232
    return silesia_tar()[off: off + MEMBER_LEN]


def xorshift_bytes(n, seed=0x9E3779B97F4A7C15):
    """Deterministic pseudo-random bytes (xorshift64*)."""
    import numpy as np
    out = np.empty((n + 7) // 8, dtype=np.uint64)
    x = np.uint64(seed)
    # vectorised splitmix64 over the index keeps this fast for MiB sizes
    idx = np.arange(out.size, dtype=np.uint64) + x
    z = idx * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return z.tobytes()[:n]


def synthetic_mix(n, seed=1):
    """Mixed-compressibility buffer: text-like, repetitive, zero runs and noise slices."""
    import numpy as np
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=rng.integers(2, 9), dtype=np.uint8)) for _ in range(512)]
    parts = []
    total = 0
    while total < n:
        kind = rng.integers(0, 10)
        ln = int(rng.integers(200, 6000))
        if kind < 5:
            w = b" ".join(words[i] for i in rng.integers(0, 512, size=ln // 5))
        elif kind < 6:
            w = bytes(ln)
        elif kind < 7:
            w = bytes([int(rng.integers(0, 256))]) * ln
        elif kind < 8:
            pat = bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8))
            w = pat * (ln // len(pat) + 1)
        else:
            w = bytes(rng.integers(0, 256, size=ln, dtype=np.uint8))
        parts.append(w)
        total += len(w)
    return b"".join(parts)[:n]


CALGARY_MIX_BYTES = 64 << 20
CALGARY_MIX_SHA256 = "fc0e027bacb9bf3be3c23479b3ea674087602f372086ed3b9c3e180488c1f011"


def calgary_extra():
    """The four non-Silesia sources SURVEY.md 8(d)4 names (lcet10.txt, paper-100k.pdf, fireworks.jpg, gix-blame-readme.bin of the
    reference's test-data), committed as data/calgary_extra.tar.gz."""
    if "extra" not in _cache:
        import tarfile
        with tarfile.open(os.path.join(ROOT, "data", "calgary_extra.tar.gz")) as tf:
            _cache["extra"] = [tf.extractfile(nm).read() for nm in ("lcet10.txt", "paper-100k.pdf", "fireworks.jpg", "gix-blame-readme.bin")]
    return _cache["extra"]


def calgary_mix(n=CALGARY_MIX_BYTES):
    """BASELINE config 4's "64 MiB synthetic Calgary-mix" exactly as SURVEY.md 8(d)4 defines it (the Calgary corpus itself is not
    available offline): 64 KiB slices drawn round-robin from {lcet10.txt, paper-100k.pdf, fireworks.jpg, gix-blame-readme.bin, the 12
    Silesia members} at xorshift64*-chosen offsets (seed 0x9E3779B97F4A7C15), every 16th slice replaced by seeded random bytes and
    every 32nd by zeros."""
    key = ("calgary", n)
    if key in _cache:
        return _cache[key]
    SL = 65536
    x = 0x9E3779B97F4A7C15
    mask = (1 << 64) - 1
    members = calgary_extra() + [silesia_member(k) for k in range(12)]
    parts = []
    i = 0
    while i * SL < n:
        x ^= x >> 12
        x ^= (x << 25) & mask
        x ^= x >> 27
        r = (x * 0x2545F4914F6CDD1D) & mask
        if i % 32 == 31:
            parts.append(bytes(SL))
        elif i % 16 == 15:
            parts.append(xorshift_bytes(SL, seed=r))
        else:
            m = members[i % len(members)]
            off = r % (len(m) - SL)
            parts.append(m[off: off + SL])
        i += 1
    out = b"".join(parts)[:n]
    if n == CALGARY_MIX_BYTES:
        assert hashlib.sha256(out).hexdigest() == CALGARY_MIX_SHA256
    _cache[key] = out
    return out


def periodic_mutated(n, period, nmut, seed):
    """A period of random bytes repeated to n bytes with nmut single-byte mutations: long matches whose sources sit in the holes of
    earlier long matches -- the input class on which the hole fixed point needs the most iterations (found by scripts/fuzz_hostmodel.py)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    per = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
    b = bytearray((per * (n // period + 1))[:n])
    for _ in range(nmut):
        b[int(rng.integers(0, n))] = int(rng.integers(0, 256))
    return bytes(b)
