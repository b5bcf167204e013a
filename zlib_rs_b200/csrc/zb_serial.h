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
    uint32_t wsize;                       // 1 << windowBits (512 .. 32768)
};
ZB_HD SerialLowParams serial_low_params(int level, uint32_t block_syms, uint32_t wsize = kWSize)
{
    if (level == 1) return {1, 0, 0, 0, 0, block_syms, wsize};
    return {2, 4, 4, 8, 4, block_syms, wsize};
}

// Scalar reference operations (host model; the CUDA kernel substitutes warp-cooperative versions).
struct ScalarOps {
    // slide_hash (deflate/slide_hash.rs:11-47): saturating subtraction of the window size
    static ZB_HD void slide(uint16_t *t, uint32_t n, uint32_t w)
    {
        for (uint32_t i = 0; i < n; i++) t[i] = t[i] >= w ? (uint16_t)(t[i] - w) : 0;
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

// ---------------------------------------------------------------------------------------------
// Lane-parallel code that is shared by the device (32 real lanes, one slot per variable) and the host model (a loop
// over 32 lanes, 32 slots per variable).  W supplies: kSlots, first()/end() of the lane loop, slot(l), leader(),
// ballot, match_any, bcast, sync.  Uniform values (the same in every lane) are plain variables.
// ---------------------------------------------------------------------------------------------
#define ZB_FOR_LANES(W_, l) for (uint32_t l = W_::first(), zb_e_ = W_::end(); l < zb_e_; l++)
template <class W, class T>
struct LaneVar {
    T v[W::kSlots];
    ZB_HD T &operator[](uint32_t l) { return v[W::slot(l)]; }
    ZB_HD const T &operator[](uint32_t l) const { return v[W::slot(l)]; }
};
struct HostWarp { // scalar emulation
    static constexpr uint32_t kSlots = 32;
    static ZB_HD uint32_t first() { return 0; }
    static ZB_HD uint32_t end() { return 32; }
    static ZB_HD uint32_t slot(uint32_t l) { return l; }
    static ZB_HD bool leader() { return true; }
    static ZB_HD uint32_t ballot(const LaneVar<HostWarp, uint32_t> &p)
    {
        uint32_t m = 0;
        for (uint32_t l = 0; l < 32; l++) m |= (p.v[l] ? 1u : 0u) << l;
        return m;
    }
    static ZB_HD void match_any(const LaneVar<HostWarp, uint32_t> &key, LaneVar<HostWarp, uint32_t> &out)
    {
        for (uint32_t l = 0; l < 32; l++) {
            uint32_t m = 0;
            for (uint32_t k = 0; k < 32; k++) m |= (key.v[k] == key.v[l] ? 1u : 0u) << k;
            out.v[l] = m;
        }
    }
    static ZB_HD uint32_t bcast(const LaneVar<HostWarp, uint32_t> &x, uint32_t src) { return x.v[src]; }
    static ZB_HD void sync() {}
};
ZB_HD uint32_t bit_range(uint32_t a, uint32_t b) // bits a .. b-1, 0 <= a <= b <= 32
{
    const uint32_t hi = b >= 32 ? 0xffffffffu : (1u << b) - 1u;
    const uint32_t lo = a >= 32 ? 0xffffffffu : (1u << a) - 1u;
    return hi & ~lo;
}
ZB_HD uint32_t top_bit(uint32_t m) // index of the highest set bit, m != 0
{
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)m);
#endif
    uint32_t i = 31;
    while (!((m >> i) & 1u)) i--;
    return i;
}
ZB_HD uint32_t low_bit(uint32_t m) // index of the lowest set bit, m != 0
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)m) - 1u;
#endif
    uint32_t i = 0;
    while (!((m >> i) & 1u)) i++;
    return i;
}

// ---------------------------------------------------------------------------------------------
// RingAcc: the part of the input the parser can touch -- the 32 KiB behind strstart and a few hundred bytes ahead -- kept in a ring
// of R bytes (shared memory on the device), refilled a few KiB at a time as strstart advances.  A single warp cannot hide the
// latency of dependent global loads, and with 192 KiB of tables resident the L1 left over is too small for the window.
// R is a multiple of 16; 16 bytes behind the ring mirror its first bytes so that unaligned word reads never wrap.  Everything outside
// [lo, hi) -- and the stale-window bytes behind the end of the input -- is read from the input itself.
// CP: copy16(dst, src, n_units, first_unit) cooperative copy, sync().
// ---------------------------------------------------------------------------------------------
#if defined(ZB_RING_STATS)
static unsigned long long zb_ring_misses = 0;
#endif
constexpr uint32_t kRingBack = kWSize + 16;  // bytes kept behind strstart
constexpr uint32_t kRingAhead = 352;         // bytes guaranteed ahead of strstart after ensure()
struct ScalarCopy {
    static ZB_HD uint32_t first() { return 0; }
    static ZB_HD uint32_t stride() { return 1; }
    static ZB_HD void sync() {}
};
template <uint32_t R, class CP>
struct RingAcc {
    uint8_t *ring;     // R + 16 bytes
    const uint8_t *in; // N bytes, zero padded by at least 16
    uint32_t N;
    uint32_t W;              // window size (what the reference's window buffer holds behind the input depends on it)
    uint32_t lo = 0, hi = 0; // the ring holds absolute [lo, hi); multiples of 16
    ZB_HD RingAcc(uint8_t *r, const uint8_t *i, uint32_t n, uint32_t w = kWSize) : ring(r), in(i), N(n), W(w) {}
    ZB_HD uint32_t raw(uint32_t y) const // y < N
    {
#if defined(ZB_RING_STATS)
        if (!(y - lo < hi - lo)) zb_ring_misses++;
#endif
        return (y - lo < hi - lo) ? ring[y % R] : in[y];
    }
    ZB_HD uint32_t byte(uint32_t y) const
    {
        while (y >= N) { // what the reference's window buffer still holds behind the end of the input (cf. GAcc)
            if (y < 2 * W) return 0;
            y -= W;
        }
        return raw(y);
    }
    ZB_HD uint32_t word(uint32_t y) const
    {
        if (y + 4 <= N && y - lo < hi - lo && y + 4 <= hi) {
            const uint32_t i = y % R;
#if defined(__CUDA_ARCH__)
            const uint32_t *w = reinterpret_cast<const uint32_t *>(ring + (i & ~3u));
            return __funnelshift_r(w[0], w[1], (i & 3u) * 8u);
#else
            return ring[i] | (ring[i + 1] << 8) | (ring[i + 2] << 16) | ((uint32_t)ring[i + 3] << 24);
#endif
        }
        return byte(y) | (byte(y + 1) << 8) | (byte(y + 2) << 16) | (byte(y + 3) << 24);
    }
    // make [p - kRingBack, p + kRingAhead) resident (clipped to the padded input); uniform call
    ZB_HD void ensure(uint32_t p)
    {
        const uint32_t npad = (N + 15u) & ~15u;
        const uint32_t need = p + kRingAhead < npad ? p + kRingAhead : npad;
        if (hi >= need) return;
        uint32_t nh = ((p > kRingBack ? p - kRingBack : 0u) + R) & ~15u;
        if (nh > npad) nh = npad;
        CP::sync();
        const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(ring)) & 15u) == 0;
        for (uint32_t u = hi / 16 + CP::first(); u < nh / 16; u += CP::stride()) {
            const uint32_t di = (u * 16u) % R;
            if (vec) {
                struct alignas(16) U16 { uint32_t w[4]; };
                const U16 v = *reinterpret_cast<const U16 *>(in + (size_t)u * 16);
                *reinterpret_cast<U16 *>(ring + di) = v;
                if (di == 0) *reinterpret_cast<U16 *>(ring + R) = v;
            } else {
                for (uint32_t k = 0; k < 16; k++) {
                    const uint8_t v = in[(size_t)u * 16 + k];
                    ring[di + k] = v;
                    if (di == 0) ring[R + k] = v;
                }
            }
        }
        hi = nh;
        lo = nh > R ? nh - R : 0;
        CP::sync();
    }
};

// D: byte(y) (stale-window semantics behind the input, see GAcc), word(y) = little-endian 32 bits at y, ensure(p) before every step.
// OPS: slide / compare256.  EMIT(Sym).  FLUSH(block_index, B): a full sym_buf was flushed (deflate_fast only).
template <class D, class OPS>
struct SerialLow {
    D &d;
    uint16_t *head;   // 65536 entries (window indices)
    uint16_t *prev;   // 32768 entries; nullptr at level 1 (deflate_quick never reads it)
    uint32_t N;
    SerialLowParams sp;
    uint32_t B = 0, F = 0, p = 0;
    uint32_t WS, MD; // window size, max_dist = W - MIN_LOOKAHEAD (deflate.rs:1423)

    ZB_HD SerialLow(D &d_, uint16_t *h, uint16_t *pv, uint32_t n, const SerialLowParams &s)
        : d(d_), head(h), prev(pv), N(n), sp(s), WS(s.wsize), MD(s.wsize - kMinLookahead) {}

    // StandardHashCalc::quick_insert_value (hash_calc.rs:48-59); str is a window index
    ZB_HD uint32_t insert_value(uint32_t str, uint32_t val)
    {
        const uint32_t hm = hash_u32(val);
        const uint32_t hd = head[hm];
        if (hd != (str & 0xffffu)) {
            if (prev) prev[str & (WS - 1)] = (uint16_t)hd;
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
            uint32_t more = 2 * WS - (F - p) - sw;
            if (sw >= WS + MD) {
                B += WS;
                OPS::slide(head, 65536u, WS);
                if (prev) OPS::slide(prev, WS, WS);
                more += WS;
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
        const uint32_t limit = sw > MD ? sw - MD : 0;
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
                cur = prev[cur & (WS - 1)];
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
            cur = prev[cur & (WS - 1)];
            if (cur <= limit) return best;
        }
        return best;
    }

    static ZB_HD uint32_t ctz_bytes(uint32_t x) // index of the lowest non-zero byte, x != 0
    {
#if defined(__CUDA_ARCH__)
        return ((uint32_t)__ffs((int)x) - 1u) >> 3;
#endif
        uint32_t n = 0;
        while (!(x & 0xffu)) { x >>= 8; n++; }
        return n;
    }

    // 32 loop-tops of deflate_quick at once.  Every lane takes one position and assumes that all positions before it in the step
    // are loop-tops (literals); the first lane whose table candidate matches ends the run of literals with its match, and the
    // lanes behind the match resume as loop-tops of a further round.  "The table" for a lane is the stored head plus the lanes of
    // this step that are loop-tops so far (same-hash lanes found with match_any), so the stored table is written once, at the
    // end.  Requires that no position of the step can trigger fill_window (lookahead >= MIN_LOOKAHEAD + 32).
    template <class W, class EA>
    ZB_HD bool quick_wide_step(EA &&emit_at, uint32_t &nsym)
    {
        if (F - p < kMinLookahead + 32u) return false;
        const uint32_t sw0 = p - B;
        LaneVar<W, uint32_t> val, hsh, peers, cand, eq;
        ZB_FOR_LANES(W, l) { val[l] = d.word(p + l); hsh[l] = hash_u32(val[l]); }
        W::match_any(hsh, peers);
        uint32_t inserted = 0; // lanes of this step that are loop-tops (hence in the table)
        uint32_t s = 0;        // first lane that is not parsed yet
        while (s < 32) {
            ZB_FOR_LANES(W, l) {
                uint32_t e = 0, hh = 0;
                if (l >= s) {
                    const uint32_t lower = peers[l] & (inserted | bit_range(s, l));
                    hh = lower ? sw0 + top_bit(lower) : head[hsh[l]];
                    const uint32_t sw = sw0 + l;
                    // quick.rs:113-119: distance in range and the first four bytes equal => a match of at least 4
                    if (hh < sw && sw - hh <= MD && d.word(B + hh) == val[l]) e = 1;
                }
                cand[l] = hh;
                eq[l] = e;
            }
            const uint32_t m = W::ballot(eq);
            const uint32_t f = m ? low_bit(m) : 32u;
            ZB_FOR_LANES(W, l) { if (l >= s && l < f) emit_at(nsym + (l - s), Sym{0, (uint16_t)(val[l] & 0xffu), p + l}); }
            nsym += f - s;
            inserted |= bit_range(s, f < 32 ? f + 1 : 32);
            if (f == 32) { s = 32; break; }
            const uint32_t hh = W::bcast(cand, f);
            uint32_t len = OPS::compare256(d, p + f + 2, B + hh + 2) + 2;
            if (len > kMaxMatch) len = kMaxMatch; // lookahead > 258 here
            ZB_FOR_LANES(W, l) { if (l == f) emit_at(nsym, Sym{(uint16_t)(sw0 + f - hh), (uint16_t)(len - 3), p + f}); }
            nsym++;
            s = f + len;
        }
        // the last loop-top of every hash value is what the table keeps (prev is never read at level 1)
        ZB_FOR_LANES(W, l) {
            if ((inserted >> l) & 1u) {
                const uint32_t mine = peers[l] & inserted;
                if ((mine >> l) == 1u) head[hsh[l]] = (uint16_t)(sw0 + l);
            }
        }
        W::sync();
        p += s;
        return true;
    }

    // deflate_quick (quick.rs:12-169), one call with Z_FINISH and ample output: one static block.  emit_at(index, Sym).
    // Returns the final window base; nsym receives the number of symbols.
    template <class W, class EA>
    ZB_HD uint32_t run_quick(EA &&emit_at, uint32_t &nsym)
    {
        nsym = 0;
        for (;;) {
            d.ensure(p);
            uint32_t lookahead = F - p;
            if (lookahead < kMinLookahead) {
                fill_window();
                lookahead = F - p;
                if (lookahead == 0) break;
            }
            if (quick_wide_step<W>(emit_at, nsym)) continue;
            uint32_t lc;
            if (lookahead >= 4) {
                const uint32_t sw = p - B;
                const uint32_t val = d.word(p);
                const uint32_t hh = insert_value(sw, val);
                if (hh < sw && sw - hh <= MD) {
                    const uint32_t c = B + hh;
                    if (val == d.word(c)) {
                        uint32_t len = OPS::compare256(d, p + 2, c + 2) + 2;
                        if (len >= 4) {
                            if (len > lookahead) len = lookahead;
                            if (len > kMaxMatch) len = kMaxMatch;
                            if (W::leader()) emit_at(nsym, Sym{(uint16_t)(sw - hh), (uint16_t)(len - 3), p});
                            nsym++;
                            p += len;
                            continue;
                        }
                    }
                }
                lc = val & 0xffu;
            } else {
                lc = d.byte(p);
            }
            if (W::leader()) emit_at(nsym, Sym{0, (uint16_t)lc, p});
            nsym++;
            p++;
        }
        return B;
    }

    // 32 positions of deflate_fast at once, same idea as quick_wide_step.  Differences: a lane walks up to max_chain candidates
    // (same-hash lanes of this step that are in the table, newest first, then the stored head/prev chain) with the 8-byte
    // pre-check of longest_match (:198-224; nice_match = 8 ends the walk at the first candidate that passes it); a match of
    // length 4 puts its interior into the table, a longer one only its last position (fast.rs:62-77); prev links are written with
    // the head at commit time.  Positions behind lane 31 that a match inserts are inserted by the scalar code after the commit.
    template <class W, class EA, class FL>
    ZB_HD bool fast_wide_step(EA &&emit_at, FL &&flush, uint32_t &nsym, uint32_t &fill, uint32_t &nblk)
    {
        if (F - p < kMinLookahead + 32u || sp.block_syms < 64u) return false;
        const uint32_t sw0 = p - B;
        LaneVar<W, uint32_t> val, hsh, peers, mstart, mlen, ism;
        ZB_FOR_LANES(W, l) { val[l] = d.word(p + l); hsh[l] = hash_u32(val[l]); }
        W::match_any(hsh, peers);
        uint32_t inserted = 0; // lanes of this step that are in the table (all below s)
        uint32_t s = 0;        // first lane that is not parsed yet
        uint32_t post_lo = 0, post_hi = 0; // window indices [post_lo, post_hi) behind the step that a match inserts
        while (s < 32) {
            ZB_FOR_LANES(W, l) {
                uint32_t is_match = 0, best = 2, start = 0;
                if (l >= s) {
                    const uint32_t sw = sw0 + l;
                    const uint32_t limit = sw > MD ? sw - MD : 0;
                    uint32_t pm = peers[l] & (inserted | bit_range(s, l));
                    bool in_table = false;
                    uint32_t cur, chain = sp.chain;
                    if (pm) { const uint32_t j = top_bit(pm); pm &= ~(1u << j); cur = sw0 + j; }
                    else { cur = head[hsh[l]]; in_table = true; }
                    // fast.rs:45: dist in range, hash_head != 0
                    if (cur < sw && sw - cur <= MD && cur != 0) {
                        for (;;) {
                            const uint32_t c = B + cur;
                            const uint32_t x0 = val[l] ^ d.word(c);
                            uint32_t c8;
                            if (x0) c8 = ctz_bytes(x0);
                            else { const uint32_t x1 = d.word(p + l + 4) ^ d.word(c + 4); c8 = x1 ? 4 + ctz_bytes(x1) : 8; }
                            if (c8 == 8) { start = cur; best = 8; break; } // full length is computed for the winning lane only
                            if (c8 > best) { start = cur; best = c8; }
                            if (--chain == 0) break;
                            if (pm) { const uint32_t j = top_bit(pm); pm &= ~(1u << j); cur = sw0 + j; }
                            else if (!in_table) { cur = head[hsh[l]]; in_table = true; }
                            else cur = prev[cur & (WS - 1)];
                            if (cur <= limit) break;
                        }
                        is_match = best >= 4 ? 1u : 0u;
                    }
                }
                ism[l] = is_match;
                mstart[l] = start;
                mlen[l] = best;
            }
            const uint32_t m = W::ballot(ism);
            const uint32_t f = m ? low_bit(m) : 32u;
            ZB_FOR_LANES(W, l) { if (l >= s && l < f) emit_at(nsym + (l - s), Sym{0, (uint16_t)(val[l] & 0xffu), p + l}); }
            uint32_t added = f - s;
            inserted |= bit_range(s, f < 32 ? f + 1 : 32);
            if (f == 32) { s = 32; }
            else {
                const uint32_t st = W::bcast(mstart, f);
                uint32_t len = W::bcast(mlen, f);
                if (len == 8) {
                    len = OPS::compare256(d, p + f + 2, B + st + 2) + 2;
                    if (len > kMaxMatch) len = kMaxMatch; // lookahead > 258 here
                }
                ZB_FOR_LANES(W, l) { if (l == f) emit_at(nsym + added, Sym{(uint16_t)(sw0 + f - st), (uint16_t)(len - 3), p + f}); }
                added++;
                // table: a match of max_insert_length (4) inserts its interior, a longer one only its last position
                const uint32_t i0 = len <= sp.lazy ? f + 1 : f + len - 1, i1 = f + len; // lanes [i0, i1)
                if (i0 < 32) inserted |= bit_range(i0, i1 < 32 ? i1 : 32);
                if (i1 > 32) { post_lo = sw0 + (i0 > 32 ? i0 : 32); post_hi = sw0 + i1; }
                s = f + len;
            }
            nsym += added;
            fill += added;
            if (fill >= sp.block_syms) { // a full sym_buf inside the step: the base cannot change before the next fill_window
                if (W::leader()) flush(nblk, B);
                nblk++;
                fill -= sp.block_syms;
            }
        }
        // commit: every inserted lane links to the entry that was the head when it was inserted; the newest lane of a hash is the head
        LaneVar<W, uint32_t> pred;
        ZB_FOR_LANES(W, l) {
            uint32_t pv = 0;
            if ((inserted >> l) & 1u) {
                const uint32_t lower = peers[l] & inserted & bit_range(0, l);
                pv = lower ? sw0 + top_bit(lower) : head[hsh[l]];
            }
            pred[l] = pv;
        }
        W::sync();
        ZB_FOR_LANES(W, l) {
            if ((inserted >> l) & 1u) {
                prev[(sw0 + l) & (WS - 1)] = (uint16_t)pred[l];
                const uint32_t mine = peers[l] & inserted;
                if ((mine >> l) == 1u) head[hsh[l]] = (uint16_t)(sw0 + l);
            }
        }
        W::sync();
        for (uint32_t str = post_lo; str < post_hi; str++) insert_at(str);
        p += s;
        return true;
    }

    // deflate_fast (fast.rs:12-118), one call with Z_FINISH and ample output.  emit_at(index, Sym); flush(block, base) at every full
    // sym_buf.  Returns the final window base; nsym receives the number of symbols.
    template <class W, class EA, class FL>
    ZB_HD uint32_t run_fast(EA &&emit_at, FL &&flush, uint32_t &nsym)
    {
        uint32_t fill = 0, nblk = 0;
        nsym = 0;
        for (;;) {
            d.ensure(p);
            uint32_t lookahead = F - p;
            if (lookahead < kMinLookahead) {
                fill_window();
                lookahead = F - p;
                if (lookahead == 0) break;
            }
            if (fast_wide_step<W>(emit_at, flush, nsym, fill, nblk)) continue;
            bool matched = false;
            uint32_t lc = 0;
            if (lookahead >= 4) {
                const uint32_t sw = p - B;
                const uint32_t val = d.word(p);
                const uint32_t hh = insert_value(sw, val);
                if (hh < sw && sw - hh <= MD && hh != 0) {
                    uint32_t start = 0;
                    uint32_t len = longest_match(hh, start);
                    if (len >= 4) {
                        if (W::leader()) emit_at(nsym, Sym{(uint16_t)(sw - start), (uint16_t)(len - 3), p});
                        lookahead -= len;
                        if (len <= sp.lazy && lookahead >= 4) {
                            // insert_string(strstart + 1, len - 1) (hash_calc.rs:61-82)
                            const uint32_t str = sw + 1, count = len - 1;
                            const uint32_t avail = 2 * WS - str;
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
                if (W::leader()) emit_at(nsym, Sym{0, (uint16_t)lc, p});
                p++;
            }
            nsym++;
            if (++fill == sp.block_syms) { // tally_* reported a full sym_buf: flush_block!(stream, false)
                if (W::leader()) flush(nblk, B);
                nblk++;
                fill = 0;
            }
        }
        return B;
    }
};

} // namespace zb
