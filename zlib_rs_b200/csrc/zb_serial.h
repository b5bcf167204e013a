// zb_serial.h -- levels 1 and 2 (deflate_quick, deflate_fast): exact restatement, one warp per stream.
//
// These two parsers insert into the hash table only what they emit (quick: loop-tops only, quick.rs:104-141; fast:
// the interior of matches up to max_insert_length = 4 and the last position of longer ones, fast.rs:62-77), so
// "which earlier positions are candidates" is the parse itself: the fixed-point formulation of the other levels
//  degenerates to the serial order.  They are therefore run as what they are -- a serial loop over
// the reference's own head/prev tables -- with the tables in shared memory and the warp doing the wide parts
// (match compare, slide_hash) cooperatively.  The GPU gets its throughput at these levels from concurrent
// streams (one warp and 192 KiB of shared memory each), not from one stream.
//
// Everything is `__host__ __device__`; tests/hostmodel runs the same code with the scalar Ops below against the
// oracle.  Coordinates: B = absolute position of window index 0, F = absolute end of the filled window,
// p = absolute strstart; window index = absolute - B (deflate.rs:1776-1861 keeps these implicit).
#pragma once
#include "zb_core.h"

namespace zb {

struct SerialLowParams {
    uint32_t level;                       // 1 = deflate_quick, 2 = deflate_fast
    uint32_t good, lazy, nice, chain;     // deflate/algorithm/mod.rs:69-82 rows 1..2
    uint32_t block_syms;                  // lit_bufsize - 1 (sym_buf.rs:23): symbols per block of deflate_fast
};
ZB_HD SerialLowParams serial_low_params(int level, uint32_t block_syms)
{
    if (level == 1) return {1, 0, 0, 0, 0, block_syms};
    return {2, 4, 4, 8, 4, block_syms};
}

// Scalar reference operations (host model; the CUDA kernel substitutes warp-cooperative versions).
struct ScalarOps {
    // slide_hash (deflate/slide_hash.rs:11-47): saturating subtraction of the window size
    static ZB_HD void slide(uint16_t *t, uint32_t n)
    {
        for (uint32_t i = 0; i < n; i++) t[i] = t[i] >= kWSize ? (uint16_t)(t[i] - kWSize) : 0;
    }
    // compare256 (deflate/compare256.rs:4-39) on absolute positions
    template <class D>
    static ZB_HD uint32_t compare256(const D &d, uint32_t a, uint32_t b)
    {
        uint32_t n = 0;
        while (n < 256 && d.byte(a + n) == d.byte(b + n)) n++;
        return n;
    }
};

// D: byte(y) (stale-window semantics behind the input, see GAcc), word(y) = little-endian 32 bits at y.
// OPS: slide / compare256.  EMIT(Sym).  FLUSH(block_index, B): a full sym_buf was flushed (deflate_fast only).
template <class D, class OPS>
struct SerialLow {
    const D &d;
    uint16_t *head;   // 65536 entries (window indices)
    uint16_t *prev;   // 32768 entries; nullptr at level 1 (deflate_quick never reads it)
    uint32_t N;
    SerialLowParams sp;
    uint32_t B = 0, F = 0, p = 0;

    ZB_HD SerialLow(const D &d_, uint16_t *h, uint16_t *pv, uint32_t n, const SerialLowParams &s) : d(d_), head(h), prev(pv), N(n), sp(s) {}

    // StandardHashCalc::quick_insert_value (hash_calc.rs:48-59); str is a window index
    ZB_HD uint32_t insert_value(uint32_t str, uint32_t val)
    {
        const uint32_t hm = hash_u32(val);
        const uint32_t hd = head[hm];
        if (hd != (str & 0xffffu)) {
            if (prev) prev[str & (kWSize - 1)] = (uint16_t)hd;
            head[hm] = (uint16_t)str;
        }
        return hd;
    }
    ZB_HD uint32_t insert_at(uint32_t str) { return insert_value(str, d.word(B + str)); }

    // fill_window (deflate.rs:1776-1861) for a one-shot call: the whole input is available, state.insert is 0
    ZB_HD void fill_window()
    {
        for (;;) {
            const uint32_t sw = p - B;
            uint32_t more = 2 * kWSize - (F - p) - sw;
            if (sw >= kWSize + kMaxDist) {
                B += kWSize;
                OPS::slide(head, 65536u);
                if (prev) OPS::slide(prev, kWSize);
                more += kWSize;
            }
            if (F >= N) break; // avail_in == 0
            const uint32_t n = N - F < more ? N - F : more;
            F += n;
            if (F - p >= 3) {
                // :1836-1838 -- re-seeds the hash with the string before strstart (max_chain_length <= 1024 here)
                const uint32_t str = p - B;
                if (str >= 1) insert_at(str - 1);
            }
            if (!(F - p < kMinLookahead && F < N)) break;
        }
    }

    // longest_match, non-SLOW (longest_match.rs:15-350) with prev_length == 0.  cur: first candidate (window index).
    // Returns the length; `start` receives the match start (window index) when a longer match than 2 was found.
    ZB_HD uint32_t longest_match(uint32_t cur, uint32_t &start)
    {
        const uint32_t sw = p - B, lookahead = F - p;
        uint32_t best = 2, chain = sp.chain;
        if (best >= sp.good) chain >>= 2;
        const uint32_t limit = sw > kMaxDist ? sw - kMaxDist : 0;
        const bool early_exit = sp.level < 5;
        for (;;) {
            if (cur >= sw) break;
            uint32_t len = 0;
            bool found = true;
            for (;;) {
                const uint32_t c = B + cur;
                if (best < 8) {
                    // :198-224 -- the first 8 bytes decide lengths 3..7 directly
                    const uint32_t x0 = d.word(p) ^ d.word(c);
                    uint32_t c8;
                    if (x0) c8 = ctz_bytes(x0);
                    else { const uint32_t x1 = d.word(p + 4) ^ d.word(c + 4); c8 = x1 ? 4 + ctz_bytes(x1) : 8; }
                    if (c8 == 8) break;
                    if (c8 > best) { len = c8; break; }
                } else {
                    // :226-234 -- 8 bytes ending at best_len and the first 8 bytes
                    const uint32_t off = best - 7;
                    if (d.word(p + off) == d.word(c + off) && d.word(p + off + 4) == d.word(c + off + 4) && d.word(p) == d.word(c) &&
                        d.word(p + 4) == d.word(c + 4))
                        break;
                }
                // next in chain or return (:136-160)
                if (--chain == 0) { found = false; break; }
                cur = prev[cur & (kWSize - 1)];
                if (cur <= limit) { found = false; break; }
            }
            if (!found) return best;
            if (len == 0) len = OPS::compare256(d, p + 2, B + cur + 2) + 2;
            if (len > best) {
                start = cur;
                if (len >= lookahead) return lookahead;
                best = len;
                if (best >= sp.nice) return best;
            } else if (early_exit) {
                break;
            }
            if (--chain == 0) return best;
            cur = prev[cur & (kWSize - 1)];
            if (cur <= limit) return best;
        }
        return best;
    }

    static ZB_HD uint32_t ctz_bytes(uint32_t x)
    {
        uint32_t n = 0;
        while (!(x & 0xffu)) { x >>= 8; n++; }
        return n;
    }

    // deflate_quick (quick.rs:12-169), one call with Z_FINISH and ample output: one static block.
    template <class E>
    ZB_HD uint32_t run_quick(E &&emit)
    {
        for (;;) {
            uint32_t lookahead = F - p;
            if (lookahead < kMinLookahead) {
                fill_window();
                lookahead = F - p;
                if (lookahead == 0) break;
            }
            uint32_t lc;
            if (lookahead >= 4) {
                const uint32_t sw = p - B;
                const uint32_t val = d.word(p);
                const uint32_t hh = insert_value(sw, val);
                if (hh < sw && sw - hh <= kMaxDist) {
                    const uint32_t c = B + hh;
                    if (val == d.word(c)) {
                        uint32_t len = OPS::compare256(d, p + 2, c + 2) + 2;
                        if (len >= 4) {
                            if (len > lookahead) len = lookahead;
                            if (len > kMaxMatch) len = kMaxMatch;
                            emit(Sym{(uint16_t)(sw - hh), (uint16_t)(len - 3), p});
                            p += len;
                            continue;
                        }
                    }
                }
                lc = val & 0xffu;
            } else {
                lc = d.byte(p);
            }
            emit(Sym{0, (uint16_t)lc, p});
            p++;
        }
        return B;
    }

    // deflate_fast (fast.rs:12-118), one call with Z_FINISH and ample output.  Returns the final window base.
    template <class E, class FL>
    ZB_HD uint32_t run_fast(E &&emit, FL &&flush)
    {
        uint32_t nsym = 0, nblk = 0;
        for (;;) {
            uint32_t lookahead = F - p;
            if (lookahead < kMinLookahead) {
                fill_window();
                lookahead = F - p;
                if (lookahead == 0) break;
            }
            bool matched = false;
            uint32_t lc = 0;
            if (lookahead >= 4) {
                const uint32_t sw = p - B;
                const uint32_t val = d.word(p);
                const uint32_t hh = insert_value(sw, val);
                if (hh < sw && sw - hh <= kMaxDist && hh != 0) {
                    uint32_t start = 0;
                    uint32_t len = longest_match(hh, start);
                    if (len >= 4) {
                        emit(Sym{(uint16_t)(sw - start), (uint16_t)(len - 3), p});
                        lookahead -= len;
                        if (len <= sp.lazy && lookahead >= 4) {
                            // insert_string(strstart + 1, len - 1) (hash_calc.rs:61-82)
                            const uint32_t str = sw + 1, count = len - 1;
                            const uint32_t avail = 2 * kWSize - str;
                            const uint32_t n = avail < count + 3 ? avail : count + 3;
                            for (uint32_t i = 0; i + 4 <= n; i++) insert_at(str + i);
                            p += len;
                        } else {
                            p += len;
                            insert_at(p - B - 1); // quick_insert_string(strstart + 2 - STD_MIN_MATCH)
                        }
                        matched = true;
                    }
                }
                lc = val & 0xffu;
            } else {
                lc = d.byte(p);
            }
            if (!matched) {
                emit(Sym{0, (uint16_t)lc, p});
                p++;
            }
            if (++nsym == sp.block_syms) { // tally_* reported a full sym_buf: flush_block!(stream, false)
                flush(nblk++, B);
                nsym = 0;
            }
        }
        return B;
    }
};

} // namespace zb
