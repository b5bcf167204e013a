"""CPU tests of the GPU engine's algorithm: tests/hostmodel compiles the device functions of
zlib_rs_b200/csrc (zb_core.h, zb_huff.h) for the host and runs them in the kernels' phase order.
Checked against the oracle: the parse (symbol by symbol) and the final bytes."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from corpus import periodic_mutated, silesia_member, synthetic_mix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_H = None


def H():
    global _H
    if _H is None:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostmodel")], stdout=subprocess.DEVNULL)
        _H = ctypes.CDLL(os.path.join(ROOT, "tests", "hostmodel", "_build", "libhostmodel.so"))
    return _H


def _syms(fn, data, level, extra=False):
    n = len(data)
    out = np.zeros((n + 16) * 2, dtype=np.uint32)
    ns, it = ctypes.c_uint32(0), ctypes.c_uint32(0)
    args = [data, n, level, out.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(ns)]
    if extra:
        args.append(ctypes.byref(it))
    assert getattr(H(), fn)(*args) == 0
    return out[: ns.value * 2].copy()


def _deflate(data, level):
    cap = len(data) + len(data) // 8 + 1024
    buf = ctypes.create_string_buffer(cap)
    n, it, dt = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_int(0)
    rc = H().hm_deflate(data, len(data), level, buf, cap, ctypes.byref(n), ctypes.byref(it), ctypes.byref(dt))
    assert rc == 0, rc
    return buf.raw[: n.value]


CASES = [("empty", b""), ("one", b"a"), ("three", b"abc"), ("mix262", synthetic_mix(262, 1)), ("mix5000", synthetic_mix(5000, 2)),
         ("mix65536", synthetic_mix(65536, 3)), ("mix131072", synthetic_mix(131072, 131072)), ("zeros", bytes(200000)),
         ("rand", np.random.default_rng(7).integers(0, 256, 70000, dtype=np.uint8).tobytes()),
         ("lit16383", np.random.default_rng(2).integers(0, 256, 16383, dtype=np.uint8).tobytes()),
         ("nci", silesia_member(1)[:150000]), ("xml", silesia_member(9)[:150000]), ("mozilla", silesia_member(2)[:150000])]


@pytest.mark.parametrize("name,data", CASES, ids=[c[0] for c in CASES])
def test_parse_matches_reference_parser(name, data):
    """serial_medium (absolute-coordinate restatement) and the parallel pipeline both reproduce every
    symbol the reference's deflate_medium tallies."""
    for level in (6,):
        o = _syms("hm_oracle_trace", data, level)
        s = _syms("hm_parse_serial", data, level)
        p = _syms("hm_parse_parallel", data, level, True)
        assert len(o) == len(s) and (o == s).all()
        assert len(o) == len(p) and (o == p).all()


@pytest.mark.parametrize("name,data", CASES, ids=[c[0] for c in CASES])
def test_bytes_match_reference(name, data):
    assert _deflate(data, 6) == O.compress(data, 6)[1]


@pytest.mark.parametrize("level", [3, 4, 5])
def test_other_medium_levels(level):
    for k in (5, 9):
        data = silesia_member(k)[:120000]
        assert _deflate(data, level) == O.compress(data, level)[1]


def test_window_base_schedule():
    """wbase(): slides at the first loop-top beyond base+65274 (zlib-rs/src/deflate.rs:1787)."""
    data = synthetic_mix(300000, 11)
    o = _syms("hm_oracle_trace", data, 6)
    p = _syms("hm_parse_parallel", data, 6, True)
    assert (o == p).all()


@pytest.mark.parametrize("level", [7, 8, 9])
def test_slow_levels_parse_matches_reference_parser(level):
    """slow_step() (zb_slow.h) from fresh loop-top to fresh loop-top reproduces every symbol deflate_slow tallies."""
    for name, data in CASES + [("dickens", silesia_member(3)[:400000]), ("mix1M", synthetic_mix(1 << 20, 21))]:
        o = _syms("hm_oracle_trace", data, level)
        s = _syms("hm_parse_slow", data, level)
        assert len(o) == len(s) and (o == s).all(), name


def test_inflate_scout_finds_exactly_the_dynamic_blocks():
    """parse_dynamic_header (zb_inflate_core.h) run on EVERY bit position of a stream flags the true dynamic block starts and
    nothing else: the block-parallel inflate chains blocks through these candidates."""
    import zlib
    for data, level in ((silesia_member(3)[:300000], 6), (silesia_member(3)[:300000], 1), (synthetic_mix(400000, 17), 9)):
        comp = zlib.compress(data, level)
        cap = 4096
        starts = (ctypes.c_uint64 * cap)()
        nd, outlen = ctypes.c_uint32(0), ctypes.c_uint64(0)
        rc = H().hm_inflate_walk(comp, len(comp), ctypes.c_uint64(16), starts, cap, ctypes.byref(nd), ctypes.byref(outlen))
        assert rc == 0 and outlen.value == len(data) and nd.value >= 1
        cands = (ctypes.c_uint64 * 65536)()
        nc = ctypes.c_uint32(0)
        assert H().hm_inflate_scout(comp, len(comp), cands, 65536, ctypes.byref(nc)) == 0
        assert set(starts[: nd.value]) <= set(cands[: nc.value])
        assert nc.value - nd.value <= 2  # false candidates are possible in principle, they never reach the chain


def _deflate_low(data, level, fixed=0, mem_level=8):
    cap = len(data) + len(data) // 8 + 1024
    buf = ctypes.create_string_buffer(cap)
    n, dt = ctypes.c_uint32(0), ctypes.c_int(0)
    rc = H().hm_deflate_low(data, len(data), level, fixed, mem_level, buf, cap, ctypes.byref(n), ctypes.byref(dt))
    assert rc == 0, rc
    return buf.raw[: n.value]


@pytest.mark.parametrize("level", [1, 2])
def test_low_levels_serial_restatement(level):
    """zb_serial.h (what k_serial_low runs on one warp) with scalar Ops: every symbol deflate_quick / deflate_fast emits,
    and the final bytes (quick: one static block encoded in pieces; fast: a block per full sym_buf)."""
    for name, data in CASES + [("dickens", silesia_member(3)[:400000]), ("mix1M", synthetic_mix(1 << 20, 21)),
                               ("sao", silesia_member(8)[:300000]), ("slide", (b"x" * 65274 + synthetic_mix(1000, 1)) * 3)]:
        o = _syms("hm_oracle_trace", data, level)
        s = _syms("hm_parse_low", data, level)
        assert len(o) == len(s) and (o == s).all(), name
        assert _deflate_low(data, level) == O.compress(data, level)[1], name
    d = silesia_member(5)[:300000]
    for mem in (1, 4, 9):
        assert _deflate_low(d, level, 0, mem) == O.compress(d, level, 15, mem, 0)[1]
    assert _deflate_low(d, level, 1) == O.compress(d, level, 15, 8, 4)[1]  # Z_FIXED


def test_small_window_serial_path():
    """serial_medium with the DynWin policy (what k_tail runs for windowBits 9..14 when the input slides the window and fits the serial
    path): symbols and bytes equal the oracle's, and the reference's two small-window vectors are reproduced."""
    import json
    def dfl(data, level, wbits, mem=8):
        cap = len(data) + len(data) // 4 + 1024
        buf = ctypes.create_string_buffer(cap)
        n, dt = ctypes.c_uint32(0), ctypes.c_int(0)
        assert H().hm_deflate_small_window(data, len(data), level, wbits, mem, buf, cap, ctypes.byref(n), ctypes.byref(dt)) == 0
        return buf.raw[: n.value]
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
    for v in kat["vectors"]:
        if v["name"] in ("hash_calc_difference", "longest_match_difference"):
            assert dfl(bytes.fromhex(v["input_hex"]), v["level"], max(v["window_bits"], 9), v["mem_level"]) == bytes.fromhex(v["expected_hex"])
    for wb in (9, 11, 14):
        for src in (synthetic_mix(31000, wb), silesia_member(9)[:30000], silesia_member(2)[:31000]):
            for n in (len(src), 700):
                for level in (3, 4, 5, 6):
                    d = src[:n]
                    assert dfl(d, level, wb) == O.compress(d, level, wb)[1], (wb, n, level)


def test_small_windows_serial_levels_and_huffman_only():
    """windowBits 9..14: zb_serial.h with the window size as a parameter (levels 1, 2) and the window-base rule of
    Z_HUFFMAN_ONLY (k_literal_syms / k_block_hist: base moves at 2w + k*w, one more slide at the end of the input)."""
    def run(fn, *args):
        d = args[0]
        cap = len(d) + len(d) // 4 + 2048
        buf = ctypes.create_string_buffer(cap)
        n, dt = ctypes.c_uint32(0), ctypes.c_int(0)
        assert getattr(H(), fn)(d, len(d), *args[1:], buf, cap, ctypes.byref(n), ctypes.byref(dt)) == 0
        return buf.raw[: n.value]
    rng = np.random.default_rng(5)
    srcs = [synthetic_mix(200000, 9), silesia_member(9)[:200000], rng.integers(0, 256, 100000, dtype=np.uint8).tobytes()]
    for wb in (9, 12, 14, 15):
        w = 1 << wb
        for src in srcs:
            for n in (len(src), 3 * w - 100, 2 * w, 1100, 700):
                d = src[: min(n, len(src))]
                for level in (1, 2):
                    assert run("hm_deflate_low_w", d, level, wb, 8) == O.compress(d, level, wb)[1], (wb, n, level)
                for mem in (1, 8, 9):
                    assert run("hm_deflate_huff", d, wb, mem) == O.compress(d, 6, wb, mem, 2)[1], (wb, n, mem)


def test_fuzz_smoke():
    """A short, seeded run of scripts/fuzz_hostmodel.py: structured random inputs through every host-model path against the oracle."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_hostmodel", os.path.join(ROOT, "scripts", "fuzz_hostmodel.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    cases, bad = m.run(60, 7, max_cases=150)
    assert cases == 150 and bad == 0


def test_lazy_levels_and_rle_with_small_windows():
    """Levels 7..9 (slow_step with the window size in SlowParams) and Z_RLE (the step of k_rle) against the oracle's symbol trace for
    windowBits 9..14: window schedule, match range, the reach of the level-9 tables, the look-ahead at the fill boundaries."""
    rng = np.random.default_rng(5)
    runs = bytearray()
    while len(runs) < 90000:
        runs += bytes([int(rng.integers(0, 4))]) * int(rng.integers(1, 700))
    srcs = {"mix": synthetic_mix(70000, 9), "m9": silesia_member(9)[:70000], "m1": silesia_member(1)[:60000], "runs": bytes(runs)}
    for name, data in srcs.items():
        n = len(data)
        a = np.zeros((n + 16) * 2, dtype=np.uint32); b = np.zeros((n + 16) * 2, dtype=np.uint32)
        na = ctypes.c_uint32(); nb = ctypes.c_uint32()
        for wbits in (9, 10, 12, 14):
            for level in (7, 8, 9):
                assert H().hm_parse_slow_w(data, n, level, wbits, a.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(na)) == 0
                assert H().hm_oracle_trace_w(data, n, level, wbits, 8, b.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(nb)) == 0
                assert na.value == nb.value and (a[: na.value * 2] == b[: nb.value * 2]).all(), (name, wbits, level)
            assert H().hm_parse_rle_w(data, n, wbits, a.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(na)) == 0
            assert H().hm_oracle_trace_ws(data, n, 6, wbits, 8, 3, b.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(nb)) == 0
            assert na.value == nb.value and (a[: na.value * 2] == b[: nb.value * 2]).all(), (name, wbits, "rle")


def test_lazy_formulation_is_exact():
    """The next design: M and the macro steps evaluated on demand along one walker per 256-position chunk and along
    the true path, the hole fixed point on top -- the same symbols as the oracle, with M evaluated at a fraction of the positions."""
    cases = [(silesia_member(1)[:300000], 6), (silesia_member(9)[:250000], 6), (silesia_member(10)[:300000], 5), (silesia_member(0)[:200000], 3),
             (periodic_mutated(150000, 37, 40, 4), 6)]
    for data, level in cases:
        n = len(data)
        a = np.zeros((n + 16) * 2, dtype=np.uint32); b = np.zeros((n + 16) * 2, dtype=np.uint32)
        na, nb, it = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        ev = (ctypes.c_uint64 * 64)()
        assert H().hm_parse_lazy(data, n, level, 256, a.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(na), ev, 64, ctypes.byref(it)) == 0
        assert H().hm_oracle_trace(data, n, level, b.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(nb)) == 0
        assert na.value == nb.value and (a[: na.value * 2] == b[: nb.value * 2]).all(), level
        assert ev[0] < 0.8 * n  # and far fewer than one evaluation per position


def test_parallel_formulation_with_small_windows():
    """windowBits 9..14 at levels 3..6 on inputs that slide the window many times: the phases of the GPU pipeline (links capped at
    the window's match range, M for every position, macro steps with the DynWin window schedule, path, hole fixed point, serial
    tail) give every symbol the oracle's deflate_medium tallies."""
    cases = [("xml", silesia_member(9)[:120000]), ("mix", synthetic_mix(150000, 5)), ("nci", silesia_member(1)[:90000])]
    for name, data in cases:
        n = len(data)
        for wbits, level in ((9, 6), (10, 4), (12, 3), (12, 6), (14, 5), (14, 6)):
            a = np.zeros((n + 16) * 2, dtype=np.uint32)
            b = np.zeros((n + 16) * 2, dtype=np.uint32)
            na, nb, it = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
            assert H().hm_parse_parallel_w(data, n, level, wbits, a.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(na), ctypes.byref(it)) == 0
            assert H().hm_oracle_trace_w(data, n, level, wbits, 8, b.ctypes.data_as(ctypes.c_void_p), n + 16, ctypes.byref(nb)) == 0
            assert na.value == nb.value and (a[: na.value * 2] == b[: nb.value * 2]).all(), (name, wbits, level)
