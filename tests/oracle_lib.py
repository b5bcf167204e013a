"""ctypes binding of the CPU oracle (oracle/_build/libzoracle.so) -- TESTS ONLY.

The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may load it.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libzoracle.so")


class ZoStream(ctypes.Structure):
    _fields_ = [("next_in", ctypes.c_void_p), ("avail_in", ctypes.c_uint32), ("total_in", ctypes.c_uint64),
                ("next_out", ctypes.c_void_p), ("avail_out", ctypes.c_uint32), ("total_out", ctypes.c_uint64),
                ("msg", ctypes.c_char_p), ("state", ctypes.c_void_p), ("data_type", ctypes.c_int),
                ("adler", ctypes.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ROOT, "oracle", f) for f in ("zo_deflate.c", "zo_inflate.c", "zo_checksum.c", "zoracle.h")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(_SO)
        sz, u32, u64, p = ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p
        L.zo_adler32.argtypes, L.zo_adler32.restype = [u32, ctypes.c_char_p, sz], u32
        L.zo_crc32.argtypes, L.zo_crc32.restype = [u32, ctypes.c_char_p, sz], u32
        L.zo_adler32_combine.argtypes, L.zo_adler32_combine.restype = [u32, u32, u64], u32
        L.zo_crc32_combine.argtypes, L.zo_crc32_combine.restype = [u32, u32, u64], u32
        L.zo_crc32_combine_gen.argtypes, L.zo_crc32_combine_gen.restype = [u64], u32
        L.zo_crc32_combine_op.argtypes, L.zo_crc32_combine_op.restype = [u32, u32, u32], u32
        L.zo_hash_standard.argtypes, L.zo_hash_standard.restype = [u32], u32
        L.zo_hash_roll.argtypes, L.zo_hash_roll.restype = [u32, u32], u32
        L.zo_compress_ex.argtypes = [ctypes.c_char_p, ctypes.POINTER(sz), ctypes.c_char_p, sz] + [ctypes.c_int] * 5
        L.zo_compress_bound.argtypes, L.zo_compress_bound.restype = [sz], sz
        L.zo_uncompress.argtypes = [ctypes.c_char_p, ctypes.POINTER(sz), ctypes.c_char_p, sz]
        for f in ("zo_deflate_init",):
            getattr(L, f).argtypes = [ctypes.POINTER(ZoStream)] + [ctypes.c_int] * 4
        for f in ("zo_deflate", "zo_inflate", "zo_inflate_init"):
            getattr(L, f).argtypes = [ctypes.POINTER(ZoStream), ctypes.c_int]
        L.zo_deflate_set_dictionary.argtypes = [ctypes.POINTER(ZoStream), ctypes.c_char_p, sz]
        for f in ("zo_deflate_end", "zo_inflate_end"):
            getattr(L, f).argtypes = [ctypes.POINTER(ZoStream)]
        _lib = L
    return _lib


def adler32(data, start=1):
    return lib().zo_adler32(start, bytes(data), len(data))


def crc32(data, start=0):
    return lib().zo_crc32(start, bytes(data), len(data))


def compress(data, level=6, window_bits=15, mem_level=8, strategy=0, flush=4):
    """One-shot deflate::compress_with_flush restatement. Returns (rc, bytes)."""
    data = bytes(data)
    n = ctypes.c_size_t(len(data) + (len(data) >> 3) + 1024)
    buf = ctypes.create_string_buffer(n.value)
    rc = lib().zo_compress_ex(buf, ctypes.byref(n), data, len(data), level, window_bits, mem_level, strategy, flush)
    return rc, buf.raw[: n.value]


def compress_dict(data, dictionary, level=6, window_bits=15, mem_level=8, strategy=0):
    """deflateInit2 + deflateSetDictionary (deflate.rs:498-564) + one deflate(Z_FINISH). Returns (rc, bytes, dictid/adler)."""
    L = lib()
    s = ZoStream()
    assert L.zo_deflate_init(ctypes.byref(s), level, window_bits, mem_level, strategy) == 0
    rc = L.zo_deflate_set_dictionary(ctypes.byref(s), bytes(dictionary), len(dictionary))
    if rc != 0:
        L.zo_deflate_end(ctypes.byref(s))
        return rc, b"", 0
    dictid = s.adler
    data = bytes(data)
    src = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
    cap = len(data) + (len(data) >> 3) + 1024
    obuf = ctypes.create_string_buffer(cap)
    s.next_in, s.avail_in = ctypes.addressof(src), len(data)
    s.next_out, s.avail_out = ctypes.addressof(obuf), cap
    rc = L.zo_deflate(ctypes.byref(s), 4)
    out = obuf.raw[: cap - s.avail_out]
    L.zo_deflate_end(ctypes.byref(s))
    return (0 if rc == 1 else rc), out, dictid


def uncompress(data, out_cap):
    data = bytes(data)
    n = ctypes.c_size_t(out_cap)
    buf = ctypes.create_string_buffer(max(out_cap, 1))
    rc = lib().zo_uncompress(buf, ctypes.byref(n), data, len(data))
    return rc, buf.raw[: n.value]


def deflate_stream(data, level=6, window_bits=15, mem_level=8, strategy=0, in_chunk=1 << 30, out_chunk=1 << 20):
    """Streaming deflate through zo_deflate with bounded avail_in/avail_out (zpipe.c shape)."""
    L = lib()
    s = ZoStream()
    assert L.zo_deflate_init(ctypes.byref(s), level, window_bits, mem_level, strategy) == 0
    data = bytes(data)
    src = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
    out = bytearray()
    obuf = ctypes.create_string_buffer(out_chunk)
    pos = 0
    while True:
        n = min(in_chunk, len(data) - pos)
        s.next_in = ctypes.addressof(src) + pos
        s.avail_in = n
        pos += n
        flush = 4 if pos >= len(data) else 0
        while True:
            s.next_out = ctypes.addressof(obuf)
            s.avail_out = out_chunk
            rc = L.zo_deflate(ctypes.byref(s), flush)
            assert rc in (0, 1, -5), rc
            out += obuf.raw[: out_chunk - s.avail_out]
            if s.avail_out != 0:
                break
        if flush == 4:
            assert rc == 1
            break
    adler = s.adler
    L.zo_deflate_end(ctypes.byref(s))
    return bytes(out), adler


def inflate_stream(data, window_bits=15, in_chunk=1 << 30, out_chunk=1 << 20, flush=0):
    """Streaming inflate through zo_inflate. Returns (rc, bytes, msg, adler)."""
    L = lib()
    s = ZoStream()
    rc = L.zo_inflate_init(ctypes.byref(s), window_bits)
    assert rc == 0, rc
    data = bytes(data)
    src = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
    out = bytearray()
    obuf = ctypes.create_string_buffer(out_chunk)
    pos = 0
    rc = 0
    while True:
        n = min(in_chunk, len(data) - pos)
        s.next_in = ctypes.addressof(src) + pos
        s.avail_in = n
        while True:
            s.next_out = ctypes.addressof(obuf)
            s.avail_out = out_chunk
            rc = L.zo_inflate(ctypes.byref(s), flush)
            out += obuf.raw[: out_chunk - s.avail_out]
            if rc == -5 and s.avail_in == 0 and pos + n < len(data):
                rc = 0  # Z_BUF_ERROR is not fatal: no progress was possible without more input
                break
            if rc != 0 or s.avail_out != 0:
                break
        pos += n - s.avail_in
        if rc != 0 or pos >= len(data):
            break
    msg = s.msg.decode() if s.msg else None
    adler = s.adler
    L.zo_inflate_end(ctypes.byref(s))
    return rc, bytes(out), msg, adler


class ZoGzHeader(ctypes.Structure):
    _fields_ = [("text", ctypes.c_int), ("time", ctypes.c_ulong), ("xflags", ctypes.c_int), ("os", ctypes.c_int),
                ("extra", ctypes.c_char_p), ("extra_len", ctypes.c_uint), ("extra_max", ctypes.c_uint),
                ("name", ctypes.c_char_p), ("name_max", ctypes.c_uint), ("comment", ctypes.c_char_p), ("comm_max", ctypes.c_uint),
                ("hcrc", ctypes.c_int), ("done", ctypes.c_int)]


def gzip_with_header(data, level=6, text=0, time=0, os=0, extra=None, name=None, comment=None, hcrc=0, mem_level=8, strategy=0):
    """One-shot gzip stream with a gz_header (deflateSetHeader, zlib-rs/src/deflate.rs:2574-2697) from the oracle."""
    L = lib()
    h = ZoGzHeader(text, time, 0, os, extra, len(extra) if extra else 0, 0, name, 0, comment, 0, hcrc, 0)
    s = ZoStream()
    assert L.zo_deflate_init(ctypes.byref(s), level, 31, mem_level, strategy) == 0
    L.zo_deflate_set_header.argtypes = [ctypes.POINTER(ZoStream), ctypes.POINTER(ZoGzHeader)]
    assert L.zo_deflate_set_header(ctypes.byref(s), ctypes.byref(h)) == 0
    src = ctypes.create_string_buffer(bytes(data), max(len(data), 1))
    cap = len(data) + len(data) // 8 + 1024 + (len(extra) if extra else 0) + (len(name) if name else 0) + (len(comment) if comment else 0)
    out = ctypes.create_string_buffer(cap)
    s.next_in, s.avail_in = ctypes.addressof(src), len(data)
    s.next_out, s.avail_out = ctypes.addressof(out), cap
    assert L.zo_deflate(ctypes.byref(s), 4) == 1
    n = cap - s.avail_out
    L.zo_deflate_end(ctypes.byref(s))
    return out.raw[:n]
