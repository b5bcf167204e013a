"""Sharding one stream over ranks: contiguous ranges, raw deflate segments closed by the Z_SYNC_FLUSH
marker (zlib-rs/src/deflate.rs:2733-2738, test split_deflate :4149-4221), one all-gather of the segments, and the
combine algebra for the check value (zlib-rs/src/adler32.rs:58-87, crc32/combine.rs:3-15).  Host-side plumbing only:
the segment compressor is passed in (the GPU engine in production, the oracle in the CPU tests)."""
import struct


def plan_shards(n, world):
    """Contiguous byte ranges, one per rank; the last rank takes the remainder."""
    base = n // world
    out = []
    for r in range(world):
        lo = r * base
        hi = n if r == world - 1 else lo + base
        out.append((lo, hi))
    return out


def adler32_combine(a1, a2, len2):
    BASE = 65521
    rem = len2 % BASE
    s1 = a1 & 0xFFFF
    s2 = (rem * s1) % BASE
    s1 += (a2 & 0xFFFF) + BASE - 1
    s2 += ((a1 >> 16) & 0xFFFF) + ((a2 >> 16) & 0xFFFF) + BASE - rem
    if s1 >= BASE:
        s1 -= BASE
    if s1 >= BASE:
        s1 -= BASE
    if s2 >= (BASE << 1):
        s2 -= BASE << 1
    if s2 >= BASE:
        s2 -= BASE
    return s1 | (s2 << 16)


def zlib_header(level):
    lf = 0 if level < 2 else 1 if level < 6 else 2 if level == 6 else 3
    h = ((8 + (7 << 4)) << 8) | (lf << 6)
    h += 31 - (h % 31)
    return struct.pack(">H", h)


def stitch_zlib(segments, adlers, lengths, level=6):
    """segments[i]: raw deflate of shard i (all but the last end with 00 00 ff ff and have BFINAL clear)."""
    a = 1
    for ad, ln in zip(adlers, lengths):
        a = adler32_combine(a, ad, ln)
    return zlib_header(level) + b"".join(segments) + struct.pack(">I", a)


def compress_sharded(data, rank, world, compress_segment, adler32, all_gather_object, level=6):
    """Every rank compresses its range; all ranks end up with the whole zlib stream.
    compress_segment(bytes, last) -> raw deflate bytes; adler32(bytes) -> int; all_gather_object(obj) -> list."""
    lo, hi = plan_shards(len(data), world)[rank]
    part = data[lo:hi]
    seg = compress_segment(part, rank == world - 1)
    gathered = all_gather_object((seg, adler32(part), hi - lo))
    return stitch_zlib([g[0] for g in gathered], [g[1] for g in gathered], [g[2] for g in gathered], level)
