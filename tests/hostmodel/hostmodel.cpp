// hostmodel.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles zlib_rs_b200/csrc/zb_core.h (the device functions of the CUDA engine) for the host and
// drives them in the same phase order as the kernels, so the parallel algorithm can be checked against
// the oracle on a machine without a GPU.  Never linked into the shipped library.
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <utility>
#include <string.h>
#include <vector>
#include "../../zlib_rs_b200/csrc/zb_core.h"
extern "C" {
#include "../../oracle/zoracle.h"
}
using namespace zb;

struct HostAcc {
    const uint8_t *data; uint32_t N; const uint16_t *L; const uint32_t *holes; const uint32_t *M;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 65536) return 0; y -= 32768; }
        return data[y];
    }
    uint32_t link(uint32_t y) const { return y + 4 <= N ? L[y] : 0; }
    bool inserted(uint32_t y) const { return !((holes[y >> 5] >> (y & 31)) & 1u); }
    Match mlook(uint32_t x) const { uint32_t v = M[x]; return Match{v >> 16, x - (v & 0xffff)}; }
};

static void build_links(const uint8_t *d, uint32_t N, std::vector<uint16_t> &L)
{
    L.assign(N + 8, 0);
    std::vector<int64_t> head(65536, -1);
    for (uint32_t x = 0; x + 4 <= N; x++) {
        uint32_t v = d[x] | (d[x + 1] << 8) | (d[x + 2] << 16) | ((uint32_t)d[x + 3] << 24);
        uint32_t h = hash_u32(v);
        if (head[h] >= 0 && x - head[h] <= kMaxDist) L[x] = (uint16_t)(x - head[h]);
        head[h] = x;
    }
}

struct SymOut { uint32_t pos; uint16_t dist; uint16_t lc; };

extern "C" int hm_parse_serial(const uint8_t *data, uint32_t N, int level, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), ins((N >> 5) + 2, 0);
    HostAcc a{data, N, L.data(), holes.data(), nullptr};
    LevelParams lp = level_params(level);
    uint32_t n = 0;
    serial_medium(a, N, 0, ins.data(), (uint32_t)ins.size(), lp, [&](Sym s, uint32_t) {
        if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc};
        n++;
    });
    *nsyms = n;
    return 0;
}

// Phase-by-phase model of the GPU pipeline.
extern "C" int hm_parse_parallel(const uint8_t *data, uint32_t N, int level, SymOut *out, uint32_t cap, uint32_t *nsyms,
                                 uint32_t *iters_out)
{
    LevelParams lp = level_params(level);
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), newholes((N >> 5) + 2, 0);
    std::vector<uint32_t> M(N + 1024, 0), nxt(N + 1, 0);
    HostAcc a{data, N, L.data(), holes.data(), M.data()};
    uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    uint32_t iters = 0;
    std::vector<uint32_t> path;
    uint32_t tail_entry = 0;
    bool first = true;
    for (;;) {
        iters++;
        // K2: M for every position (hole-aware)
        for (uint32_t x = 0; x < N; x++) {
            if (!first) { /* model recomputes everything; the GPU restricts this to affected tiles */ }
            Match m = (x + kMSafe <= N) ? lm_walk(a, x, 0xffffffffu, lp) : Match{0, 0};
            M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
        }
        first = false;
        // P1: nxt for every canonical position below the tail
        for (uint32_t p = 0; p < tail_start; p++) {
            uint32_t ns, nlong = 0, lpos = 0, llen = 0;
            uint32_t np = macro_step(a, p, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy) { nlong++; lpos = s.pos; llen = s.lc + 3u; }
            }, &ns);
            nxt[p] = np;
            // what k_nxt / k_holes rely on (zb_kernels.cu): a long match is the last symbol of its macro step, alone, and 257 or 258
            // bytes long at levels 5/6 (the step is the match at levels 3/4)
            if (nlong && (nlong != 1 || lpos + llen != np || (lp.early_exit ? lpos != p : (llen != 257u && llen != 258u)))) return -2;
        }
        // P2: path from 0
        path.clear();
        uint32_t p = 0;
        tail_entry = 0;
        while (p < tail_start) {
            if (nxt[p] >= tail_start) break; // p is the tail entry (re-simulated serially)
            path.push_back(p);
            p = nxt[p];
        }
        tail_entry = p;
        // holes from the path's long matches
        std::fill(newholes.begin(), newholes.end(), 0);
        for (uint32_t q : path) {
            uint32_t ns;
            macro_step(a, q, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy)
                    for (uint32_t y = s.pos + 1; y + 1 < s.pos + s.lc + 3; y++) newholes[y >> 5] |= 1u << (y & 31);
            }, &ns);
        }
        if (newholes == holes) break;
        holes = newholes;
        a.holes = holes.data();
        if (iters > N / 257u + 64u) return -1;
    }
    // P3: symbols
    uint32_t n = 0;
    for (uint32_t q : path) {
        uint32_t ns;
        macro_step(a, q, lp, tail_start, [&](Sym s) { if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc}; n++; }, &ns);
    }
    // tail
    std::vector<uint32_t> ins(64 + (N - tail_entry) / 32 + 2, 0);
    serial_medium(a, N, tail_entry, ins.data(), (uint32_t)ins.size(), lp, [&](Sym s, uint32_t) {
        if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc};
        n++;
    });
    *nsyms = n;
    *iters_out = iters;
    return 0;
}


// The parallel formulation with a window smaller than 32 KiB (windowBits 9..14): same phases as hm_parse_parallel, links capped at the
// window's match range, the window schedule of DynWin everywhere.
struct HostAccW {
    const uint8_t *data; uint32_t N; const uint16_t *L; const uint32_t *holes; const uint32_t *M; uint32_t w;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 2 * w) return 0; y -= w; }
        return data[y];
    }
    uint32_t link(uint32_t y) const { return y + 4 <= N ? L[y] : 0; }
    bool inserted(uint32_t y) const { return !((holes[y >> 5] >> (y & 31)) & 1u); }
    Match mlook(uint32_t x) const { uint32_t v = M[x]; return Match{v >> 16, x - (v & 0xffff)}; }
};

extern "C" int hm_parse_parallel_w(const uint8_t *data, uint32_t N, int level, int wbits, SymOut *out, uint32_t cap, uint32_t *nsyms,
                                   uint32_t *iters_out)
{
    const DynWin wn{1u << wbits};
    LevelParams lp = level_params(level);
    std::vector<uint16_t> L(N + 8, 0);
    {
        std::vector<int64_t> head(65536, -1);
        for (uint32_t x = 0; x + 4 <= N; x++) {
            uint32_t v = data[x] | (data[x + 1] << 8) | (data[x + 2] << 16) | ((uint32_t)data[x + 3] << 24);
            uint32_t h = hash_u32(v);
            if (head[h] >= 0 && x - head[h] <= wn.maxdist()) L[x] = (uint16_t)(x - head[h]);
            head[h] = x;
        }
    }
    std::vector<uint32_t> holes((N >> 5) + 2, 0), newholes((N >> 5) + 2, 0);
    std::vector<uint32_t> M(N + 1024, 0), nxt(N + 1, 0);
    HostAccW a{data, N, L.data(), holes.data(), M.data(), wn.w};
    uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    uint32_t iters = 0;
    std::vector<uint32_t> path;
    uint32_t tail_entry = 0;
    for (;;) {
        iters++;
        for (uint32_t x = 0; x < N; x++) {
            Match m = (x + kMSafe <= N) ? lm_walk(a, x, 0xffffffffu, lp, wn) : Match{0, 0};
            M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
        }
        for (uint32_t p = 0; p < tail_start; p++) {
            uint32_t ns;
            nxt[p] = macro_step(a, p, lp, tail_start, [](Sym) {}, &ns, wn);
        }
        path.clear();
        uint32_t p = 0;
        while (p < tail_start) {
            if (nxt[p] >= tail_start) break;
            path.push_back(p);
            p = nxt[p];
        }
        tail_entry = p;
        std::fill(newholes.begin(), newholes.end(), 0);
        for (uint32_t q : path) {
            uint32_t ns;
            macro_step(a, q, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy)
                    for (uint32_t y = s.pos + 1; y + 1 < s.pos + s.lc + 3; y++) newholes[y >> 5] |= 1u << (y & 31);
            }, &ns, wn);
        }
        if (newholes == holes) break;
        holes = newholes;
        a.holes = holes.data();
        if (iters > N / 257u + 64u) return -1;
    }
    uint32_t n = 0;
    for (uint32_t q : path) {
        uint32_t ns;
        macro_step(a, q, lp, tail_start, [&](Sym s) { if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc}; n++; }, &ns, wn);
    }
    std::vector<uint32_t> ins(64 + (N - tail_entry) / 32 + 2, 0);
    serial_medium(a, N, tail_entry, ins.data(), (uint32_t)ins.size(), lp, [&](Sym s, uint32_t) {
        if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc};
        n++;
    }, wn);
    *nsyms = n;
    *iters_out = iters;
    return 0;
}

// Oracle trace (reference parser's tallied symbols)
struct TraceCtx { SymOut *out; uint32_t cap, n; };
static void trace_cb(void *ctx, uint64_t pos, unsigned dist, unsigned lc_or_len)
{
    TraceCtx *t = (TraceCtx *)ctx;
    if (t->n < t->cap) t->out[t->n] = SymOut{(uint32_t)pos, (uint16_t)dist, (uint16_t)(dist ? lc_or_len - 3 : lc_or_len)};
    t->n++;
}
extern "C" int hm_oracle_trace(const uint8_t *data, uint32_t N, int level, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    zo_stream s;
    memset(&s, 0, sizeof s);
    if (zo_deflate_init(&s, level, 15, 8, 0) != 0) return -1;
    TraceCtx t{out, cap, 0};
    zo_deflate_set_trace(&s, trace_cb, &t);
    std::vector<uint8_t> dst(zo_compress_bound(N) + 64);
    s.next_in = data; s.avail_in = N; s.next_out = dst.data(); s.avail_out = (uint32_t)dst.size();
    int rc = zo_deflate(&s, ZO_FINISH);
    zo_deflate_end(&s);
    *nsyms = t.n;
    return rc == ZO_STREAM_END ? 0 : -2;
}

// ------------------------------------------------------------------------------------------
// Encode model: blocks -> trees -> bit stream, phase by phase as the kernels do it.
// ------------------------------------------------------------------------------------------
#include "../../zlib_rs_b200/csrc/zb_huff.h"

static void put_bits(std::vector<uint8_t> &out, uint64_t bitpos, uint64_t val, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        uint64_t p = bitpos + i;
        if ((p >> 3) >= out.size()) out.resize((p >> 3) + 1, 0);
        out[p >> 3] |= (uint8_t)(((val >> i) & 1) << (p & 7));
    }
}

// syms -> blocks -> trees -> bit stream.  blockB(b): window base when block b is flushed.  quick: one static block cut into pieces.
template <class BF>
static int encode_stream(const uint8_t *data, uint32_t N, int level, const std::vector<Sym> &syms, BF &&blockB, bool quick, bool fixed, uint32_t block_syms,
                         uint8_t *dst, uint32_t cap, uint32_t *out_len, int *data_type_out, uint32_t cinfo = 7)
{
    // ---- blocks ----
    HuffTables T;
    init_tables(T);
    uint32_t nsyms = (uint32_t)syms.size();
    uint32_t nblocks = nsyms / block_syms + 1;
    std::vector<BlockDesc> blocks(nblocks);
    TreeScratch scratch;
    uint32_t in_pos = 0;
    int data_type = 2;
    for (uint32_t b = 0; b < nblocks; b++) {
        BlockDesc &bd = blocks[b];
        bd.sym_begin = b * block_syms;
        bd.sym_count = (b + 1 < nblocks) ? block_syms : nsyms - bd.sym_begin;
        bd.last = b + 1 == nblocks;
        bd.in_start = in_pos;
        uint32_t end = bd.last ? N : (syms[bd.sym_begin + bd.sym_count - 1].pos +
                                      (syms[bd.sym_begin + bd.sym_count - 1].dist ? syms[bd.sym_begin + bd.sym_count - 1].lc + 3u : 1u));
        bd.in_len = end - in_pos;
        in_pos = end;
        uint32_t Bflush = blockB(b, bd);
        uint32_t lfreq[kLCodes] = {0}, dfreq[kDCodes] = {0};
        for (uint32_t i = 0; i < bd.sym_count; i++) {
            const Sym &s = syms[bd.sym_begin + i];
            if (s.dist == 0) lfreq[s.lc]++;
            else { lfreq[257 + T.length_code[s.lc]]++; dfreq[d_code(T, s.dist - 1u)]++; }
        }
        if (quick) build_quick_piece(T, bd, lfreq, dfreq, b == 0, b + 1 == nblocks, true);
        else build_block(T, scratch, bd, lfreq, dfreq, bd.in_start >= Bflush, fixed);
        if (data_type == 2 && bd.sym_count) data_type = (int)bd.data_type;
    }
    // ---- scan + pack ----
    std::vector<uint8_t> out(2, 0);
    uint32_t lf = (level < 2 || fixed) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3; // deflate.rs:1591-1601
    uint32_t h = ((8 + (cinfo << 4)) << 8) | (lf << 6);
    h += 31 - (h % 31);
    out[0] = (uint8_t)(h >> 8); out[1] = (uint8_t)h;
    uint64_t bit = 16;
    for (uint32_t b = 0; b < nblocks; b++) {
        BlockDesc &bd = blocks[b];
        bd.bit_base = bit;
        if (bd.type == 0) {
            put_bits(out, bit, bd.hdr[0], 3);
            uint64_t p = (bit + 3 + 7) & ~7ull;
            uint16_t sl = (uint16_t)bd.in_len;
            put_bits(out, p, sl, 16); put_bits(out, p + 16, (uint16_t)~sl, 16);
            for (uint32_t i = 0; i < sl; i++) put_bits(out, p + 32 + 8ull * i, data[bd.in_start + i], 8);
        } else {
            for (uint32_t i = 0; i < bd.hdr_bits; i++) put_bits(out, bit + i, (bd.hdr[i >> 3] >> (i & 7)) & 1, 1);
            uint64_t q = bit + bd.hdr_bits;
            for (uint32_t i = 0; i < bd.sym_count; i++) {
                const Sym &s = syms[bd.sym_begin + i];
                uint64_t v; uint32_t n = sym_bits(T, bd, s.dist, s.lc, v);
                put_bits(out, q, v, n); q += n;
            }
            if (!bd.no_eob) { put_bits(out, q, bd.lcode[kEndBlock], bd.llen[kEndBlock]); q += bd.llen[kEndBlock]; }
            if (q != bit + bd.hdr_bits + bd.body_bits) return -7;
        }
        bit = block_end_bit(bd, bit);
    }
    uint64_t bytes = (bit + 7) >> 3;
    out.resize(bytes, 0);
    uint32_t ad = zo_adler32(1, data, N);
    out.push_back((uint8_t)(ad >> 24)); out.push_back((uint8_t)(ad >> 16)); out.push_back((uint8_t)(ad >> 8)); out.push_back((uint8_t)ad);
    if (out.size() > cap) return -5;
    memcpy(dst, out.data(), out.size());
    *out_len = (uint32_t)out.size();
    *data_type_out = data_type;
    return 0;
}

extern "C" int hm_deflate(const uint8_t *data, uint32_t N, int level, uint8_t *dst, uint32_t cap, uint32_t *out_len,
                          uint32_t *iters_out, int *data_type_out)
{
    LevelParams lp = level_params(level);
    // ---- parse (same as hm_parse_parallel, but keeping per-symbol window bases for the tail) ----
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), newholes((N >> 5) + 2, 0);
    std::vector<uint32_t> M(N + 1024, 0), nxt(N + 1, 0);
    HostAcc a{data, N, L.data(), holes.data(), M.data()};
    uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    uint32_t iters = 0, tail_entry = 0;
    std::vector<uint32_t> path;
    for (;;) {
        iters++;
        for (uint32_t x = 0; x < N; x++) {
            Match m = (x + kMSafe <= N) ? lm_walk(a, x, 0xffffffffu, lp) : Match{0, 0};
            M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
        }
        for (uint32_t p = 0; p < tail_start; p++) { uint32_t ns; nxt[p] = macro_step(a, p, lp, tail_start, [](Sym) {}, &ns); }
        path.clear();
        uint32_t p = 0;
        while (p < tail_start && nxt[p] < tail_start) { path.push_back(p); p = nxt[p]; }
        tail_entry = p;
        std::fill(newholes.begin(), newholes.end(), 0);
        for (uint32_t q : path) {
            uint32_t ns;
            macro_step(a, q, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy)
                    for (uint32_t y = s.pos + 1; y + 1 < s.pos + s.lc + 3; y++) newholes[y >> 5] |= 1u << (y & 31);
            }, &ns);
        }
        if (newholes == holes) break;
        holes = newholes;
        a.holes = holes.data();
        if (iters > N / 257u + 64u) return -1;
    }
    std::vector<Sym> syms;
    std::vector<uint32_t> symB;
    for (uint32_t q : path) {
        uint32_t ns;
        macro_step(a, q, lp, tail_start, [&](Sym s) { syms.push_back(s); symB.push_back(wbase(s.pos)); }, &ns);
    }
    std::vector<uint32_t> ins(64 + (N - tail_entry) / 32 + 2, 0);
    uint32_t finalB = serial_medium(a, N, tail_entry, ins.data(), (uint32_t)ins.size(), lp,
                                    [&](Sym s, uint32_t B) { syms.push_back(s); symB.push_back(B); });
    uint32_t it_dummy = iters;
    (void)it_dummy;
    *iters_out = iters;
    return encode_stream(data, N, level, syms, [&](uint32_t, const BlockDesc &bd) { return bd.last ? finalB : symB[bd.sym_begin + bd.sym_count - 1]; },
                         false, false, kBlockSyms, dst, cap, out_len, data_type_out);
}

// ------------------------------------------------------------------------------------------
// Block-parallel inflate: header scout (zb_inflate_core.h) against a serial walk of the stream.
// ------------------------------------------------------------------------------------------
#include "../../zlib_rs_b200/csrc/zb_inflate_core.h"

namespace {
struct Canon { uint16_t cnt[16], first[16], offs[16], sorted[320]; };
void canon_build(Canon &c, const uint16_t *lens, uint32_t n)
{
    for (int i = 0; i < 16; i++) c.cnt[i] = 0;
    for (uint32_t i = 0; i < n; i++) c.cnt[lens[i]]++;
    c.cnt[0] = 0;
    uint32_t code = 0, o = 0;
    uint16_t nx[16];
    for (int len = 1; len <= 15; len++) { c.first[len] = (uint16_t)code; c.offs[len] = (uint16_t)o; nx[len] = (uint16_t)o; code = (code + c.cnt[len]) << 1; o += c.cnt[len]; }
    for (uint32_t s = 0; s < n; s++) if (lens[s]) c.sorted[nx[lens[s]]++] = (uint16_t)s;
}
int canon_decode(const Canon &c, const BitSrc &s, uint64_t &pos)
{
    uint32_t bits = s.peek32(pos), code = 0;
    for (int len = 1; len <= 15; len++) {
        code = (code << 1) | (bits & 1u);
        bits >>= 1;
        if (c.cnt[len] && code >= c.first[len] && code - c.first[len] < c.cnt[len]) { pos += len; return c.sorted[c.offs[len] + code - c.first[len]]; }
    }
    return -1;
}
}

// Serial walk: returns the bit positions of all dynamic block headers, -1 on a stream this simple walker cannot follow.
extern "C" int hm_inflate_walk(const uint8_t *src, uint32_t n, uint64_t start_bit, uint64_t *dyn_starts, uint32_t cap, uint32_t *ndyn,
                               uint64_t *out_len)
{
    BitSrc s{src, n};
    uint64_t pos = start_bit, out = 0;
    uint32_t k = 0;
    std::vector<uint16_t> lens(320);
    for (;;) {
        const uint32_t w = s.peek32(pos);
        const uint32_t last = w & 1, type = (w >> 1) & 3;
        if (type == 0) {
            uint64_t p = (pos + 3 + 7) & ~7ull;
            const uint32_t v = s.peek32(p);
            out += v & 0xffff;
            pos = p + 32 + 8ull * (v & 0xffff);
        } else if (type == 2) {
            DynHeader h;
            if (!parse_dynamic_header(s, pos, h, lens.data())) return -2;
            if (k < cap) dyn_starts[k] = pos;
            k++;
            Canon L, D;
            canon_build(L, lens.data(), h.hlit);
            canon_build(D, lens.data() + h.hlit, h.hdist);
            pos = h.body_bit;
            for (;;) {
                int sym = canon_decode(L, s, pos);
                if (sym < 0) return -3;
                if (sym < 256) { out++; continue; }
                if (sym == 256) break;
                const uint32_t c = sym - 257;
                uint32_t len = len_base(c) + (s.peek32(pos) & ((1u << len_extra(c)) - 1));
                pos += len_extra(c);
                int ds = canon_decode(D, s, pos);
                if (ds < 0) return -4;
                pos += dist_extra(ds);
                out += len;
            }
        } else return -1;
        if (last) break;
    }
    *ndyn = k;
    *out_len = out;
    return 0;
}

extern "C" int hm_inflate_scout(const uint8_t *src, uint32_t n, uint64_t *cands, uint32_t cap, uint32_t *ncand)
{
    BitSrc s{src, n};
    std::vector<uint16_t> lens(320);
    uint32_t k = 0;
    for (uint64_t b = 0; b + 40 < (uint64_t)n * 8; b++) {
        DynHeader h;
        if (parse_dynamic_header(s, b, h, lens.data())) { if (k < cap) cands[k] = b; k++; }
    }
    *ncand = k;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Level 7..9 (deflate_slow) model: static chains + slow_step() from fresh loop-top to fresh loop-top.
// ------------------------------------------------------------------------------------------
#include "../../zlib_rs_b200/csrc/zb_slow.h"

struct SlowAcc {
    const uint8_t *data; uint32_t N; const uint16_t *L; uint32_t need;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 65536) return 0; y -= 32768; }
        return data[y];
    }
    uint32_t link(uint32_t y) const { return y + need <= N ? L[y] : 0; }
};

static void build_links_roll(const uint8_t *d, uint32_t N, std::vector<uint16_t> &L)
{
    L.assign(N + 8, 0);
    std::vector<int64_t> head(32768, -1);
    for (uint32_t x = 0; x + 3 <= N; x++) {
        uint32_t h = hash_roll3(d[x], d[x + 1], d[x + 2]);
        if (head[h] >= 0 && x - head[h] <= kLinkCapSlow) L[x] = (uint16_t)(x - head[h]);
        head[h] = x;
    }
}

extern "C" int hm_parse_slow(const uint8_t *data, uint32_t N, int level, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    SlowParams sp = slow_params(level);
    std::vector<uint16_t> L;
    if (sp.slow) build_links_roll(data, N, L); else build_links(data, N, L);
    SlowAcc a{data, N, L.data(), sp.slow ? 3u : 4u};
    uint32_t n = 0, p = 0;
    while (p < N) {
        SlowStep s = slow_step(a, p, N, sp);
        for (uint32_t i = 0; i < s.nlit; i++) { if (n < cap) out[n] = SymOut{p + i, 0, data[p + i]}; n++; }
        if (s.len) { if (n < cap) out[n] = SymOut{p + s.nlit, (uint16_t)s.dist, (uint16_t)(s.len - 3)}; n++; }
        if (s.next <= p) return -3;
        p = s.next;
    }
    *nsyms = n;
    return 0;
}

// The lazy levels with a window smaller than 32 KiB (windowBits 9..14): the same steps with the window size in SlowParams
// (window schedule, match range, reach of the level-9 tables), links capped accordingly.
struct SlowAccW {
    const uint8_t *data; uint32_t N; const uint16_t *L; uint32_t need; uint32_t w;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 2 * w) return 0; y -= w; }
        return data[y];
    }
    uint32_t link(uint32_t y) const { return y + need <= N ? L[y] : 0; }
};

extern "C" int hm_parse_slow_w(const uint8_t *data, uint32_t N, int level, int wbits, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    SlowParams sp = slow_params(level);
    sp.wsize = 1u << wbits;
    std::vector<uint16_t> L(N + 8, 0);
    if (sp.slow) {
        std::vector<int64_t> head(32768, -1);
        for (uint32_t x = 0; x + 3 <= N; x++) {
            uint32_t h = hash_roll3(data[x], data[x + 1], data[x + 2]);
            if (head[h] >= 0 && x - head[h] <= sp.wsize - 1) L[x] = (uint16_t)(x - head[h]);
            head[h] = x;
        }
    } else {
        std::vector<int64_t> head(65536, -1);
        for (uint32_t x = 0; x + 4 <= N; x++) {
            uint32_t v = data[x] | (data[x + 1] << 8) | (data[x + 2] << 16) | ((uint32_t)data[x + 3] << 24);
            uint32_t h = hash_u32(v);
            if (head[h] >= 0 && x - head[h] <= sp.maxdist()) L[x] = (uint16_t)(x - head[h]);
            head[h] = x;
        }
    }
    SlowAccW a{data, N, L.data(), sp.slow ? 3u : 4u, sp.wsize};
    uint32_t n = 0, p = 0;
    while (p < N) {
        SlowStep s = slow_step(a, p, N, sp);
        for (uint32_t i = 0; i < s.nlit; i++) { if (n < cap) out[n] = SymOut{p + i, 0, data[p + i]}; n++; }
        if (s.len) { if (n < cap) out[n] = SymOut{p + s.nlit, (uint16_t)s.dist, (uint16_t)(s.len - 3)}; n++; }
        if (s.next <= p) return -3;
        p = s.next;
    }
    *nsyms = n;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Levels 1 and 2 (zb_serial.h): the warp-serial restatement of deflate_quick / deflate_fast with scalar Ops.
// ------------------------------------------------------------------------------------------
#include "../../zlib_rs_b200/csrc/zb_serial.h"

struct LowAcc {
    const uint8_t *data; uint32_t N;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 65536) return 0; y -= 32768; }
        return data[y];
    }
    uint32_t word(uint32_t y) const { return byte(y) | (byte(y + 1) << 8) | (byte(y + 2) << 16) | (byte(y + 3) << 24); }
};

template <uint32_t R>
static uint32_t run_low_ring(const uint8_t *data, uint32_t N, int level, uint32_t block_syms, std::vector<Sym> &syms, std::vector<uint32_t> &blockB,
                             uint32_t wsize)
{
    std::vector<uint16_t> head(65536, 0), prev(32768, 0);
    std::vector<uint8_t> padded(N + 64, 0), ring(R + 16, 0xAA);
    if (N) memcpy(padded.data(), data, N);
    RingAcc<R, ScalarCopy> a(ring.data(), padded.data(), N, wsize);
    SerialLow<RingAcc<R, ScalarCopy>, ScalarOps> m(a, head.data(), level == 2 ? prev.data() : nullptr, N, serial_low_params(level, block_syms, wsize));
    uint32_t n = 0, fb;
    auto emit_at = [&](uint32_t i, Sym s) { if (syms.size() <= i) syms.resize(i + 1); syms[i] = s; };
    if (level == 1) fb = m.template run_quick<HostWarp>(emit_at, n);
    else fb = m.template run_fast<HostWarp>(emit_at, [&](uint32_t b, uint32_t B) { if (blockB.size() <= b) blockB.resize(b + 1); blockB[b] = B; }, n);
    syms.resize(n);
    return fb;
}

// the ring sizes of k_serial_low: 64 KiB next to the 128 KiB head table at level 1, 35824 B next to head + prev at level 2
static uint32_t run_low(const uint8_t *data, uint32_t N, int level, uint32_t block_syms, std::vector<Sym> &syms, std::vector<uint32_t> &blockB,
                        uint32_t wsize = kWSize)
{
    return level == 1 ? run_low_ring<65536>(data, N, level, block_syms, syms, blockB, wsize) : run_low_ring<35824>(data, N, level, block_syms, syms, blockB, wsize);
}

extern "C" int hm_parse_low(const uint8_t *data, uint32_t N, int level, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    std::vector<Sym> syms;
    std::vector<uint32_t> blockB;
    run_low(data, N, level, kBlockSyms, syms, blockB);
    for (size_t i = 0; i < syms.size() && i < cap; i++) out[i] = SymOut{syms[i].pos, syms[i].dist, syms[i].lc};
    *nsyms = (uint32_t)syms.size();
    return 0;
}

extern "C" int hm_deflate_low(const uint8_t *data, uint32_t N, int level, int fixed, int mem_level, uint8_t *dst, uint32_t cap,
                              uint32_t *out_len, int *data_type_out)
{
    const uint32_t block_syms = (1u << (mem_level + 6)) - 1;
    std::vector<Sym> syms;
    std::vector<uint32_t> blockB;
    const uint32_t finalB = run_low(data, N, level, block_syms, syms, blockB);
    return encode_stream(data, N, level, syms, [&](uint32_t b, const BlockDesc &bd) { return (bd.last || b >= blockB.size()) ? finalB : blockB[b]; }, level == 1,
                         fixed != 0, level == 1 ? kBlockSyms : block_syms, dst, cap, out_len, data_type_out);
}


// ------------------------------------------------------------------------------------------
// Windows smaller than 32 KiB: serial_medium with the DynWin policy over the whole input (what k_tail runs for inputs that fit it).
// ------------------------------------------------------------------------------------------
struct WinAcc {
    const uint8_t *data; uint32_t N, w; const uint16_t *L; const uint32_t *holes;
    uint32_t byte(uint32_t y) const {
        while (y >= N) { if (y < 2 * w) return 0; y -= w; }
        return data[y];
    }
    uint32_t link(uint32_t y) const { return y + 4 <= N ? L[y] : 0; }
    bool inserted(uint32_t y) const { return !((holes[y >> 5] >> (y & 31)) & 1u); }
};

static uint32_t run_small_window(const uint8_t *data, uint32_t N, int level, int wbits, std::vector<Sym> &syms, std::vector<uint32_t> &symB)
{
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), ins((N >> 5) + 2, 0);
    WinAcc a{data, N, 1u << wbits, L.data(), holes.data()};
    return serial_medium(a, N, 0, ins.data(), (uint32_t)ins.size(), level_params(level),
                         [&](Sym s, uint32_t B) { syms.push_back(s); symB.push_back(B); }, DynWin{1u << wbits});
}

extern "C" int hm_oracle_trace_w(const uint8_t *data, uint32_t N, int level, int wbits, int mem_level, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    zo_stream s;
    memset(&s, 0, sizeof s);
    if (zo_deflate_init(&s, level, wbits, mem_level, 0) != 0) return -1;
    TraceCtx t{out, cap, 0};
    zo_deflate_set_trace(&s, trace_cb, &t);
    std::vector<uint8_t> dst(zo_compress_bound(N) + N / 4 + 1024);
    s.next_in = data; s.avail_in = N; s.next_out = dst.data(); s.avail_out = (uint32_t)dst.size();
    int rc = zo_deflate(&s, ZO_FINISH);
    zo_deflate_end(&s);
    *nsyms = t.n;
    return rc == ZO_STREAM_END ? 0 : -2;
}

// the oracle's symbol trace with a strategy (Z_RLE = 3, Z_FILTERED = 1)
extern "C" int hm_oracle_trace_ws(const uint8_t *data, uint32_t N, int level, int wbits, int mem_level, int strategy, SymOut *out, uint32_t cap,
                                  uint32_t *nsyms)
{
    zo_stream s;
    memset(&s, 0, sizeof s);
    if (zo_deflate_init(&s, level, wbits, mem_level, strategy) != 0) return -1;
    TraceCtx t{out, cap, 0};
    zo_deflate_set_trace(&s, trace_cb, &t);
    std::vector<uint8_t> dst(zo_compress_bound(N) + N / 4 + 1024);
    s.next_in = data; s.avail_in = N; s.next_out = dst.data(); s.avail_out = (uint32_t)dst.size();
    int rc = zo_deflate(&s, ZO_FINISH);
    zo_deflate_end(&s);
    *nsyms = t.n;
    return rc == ZO_STREAM_END ? 0 : -2;
}

// Z_RLE with any window: the step of k_rle (zb_slow.cu) -- the window only enters through the look-ahead at a loop-top
extern "C" int hm_parse_rle_w(const uint8_t *d, uint32_t N, int wbits, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    const uint32_t w = 1u << wbits;
    uint32_t n = 0, p = 0;
    while (p < N) {
        const uint32_t B = base_at(p, N, w), la = lookahead_at(p, B, N, w);
        uint32_t len = 0;
        if (la >= 3 && p > 0 && p + 1 < N && d[p - 1] == d[p] && d[p] == d[p + 1]) {
            const uint32_t c = d[p - 1];
            uint32_t k = 0;
            while (k < 256 && p + 2 + k < N && d[p + 2 + k] == c) k++;
            len = k + 2;
            if (len > la) len = la;
            if (len > kMaxMatch) len = kMaxMatch;
            if (len < 3) len = 0;
        }
        if (len) { if (n < cap) out[n] = SymOut{p, 1, (uint16_t)(len - 3)}; n++; p += len; }
        else { if (n < cap) out[n] = SymOut{p, 0, d[p]}; n++; p++; }
    }
    *nsyms = n;
    return 0;
}

extern "C" int hm_parse_small_window(const uint8_t *data, uint32_t N, int level, int wbits, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    std::vector<Sym> syms;
    std::vector<uint32_t> symB;
    run_small_window(data, N, level, wbits, syms, symB);
    for (size_t i = 0; i < syms.size() && i < cap; i++) out[i] = SymOut{syms[i].pos, syms[i].dist, syms[i].lc};
    *nsyms = (uint32_t)syms.size();
    return 0;
}

extern "C" int hm_deflate_small_window(const uint8_t *data, uint32_t N, int level, int wbits, int mem_level, uint8_t *dst, uint32_t cap,
                                       uint32_t *out_len, int *data_type_out)
{
    const uint32_t block_syms = (1u << (mem_level + 6)) - 1;
    std::vector<Sym> syms;
    std::vector<uint32_t> symB;
    const uint32_t finalB = run_small_window(data, N, level, wbits, syms, symB);
    return encode_stream(data, N, level, syms, [&](uint32_t, const BlockDesc &bd) { return bd.last ? finalB : symB[bd.sym_begin + bd.sym_count - 1]; },
                         false, false, block_syms, dst, cap, out_len, data_type_out, (uint32_t)(wbits - 8));
}

// Z_HUFFMAN_ONLY (deflate/algorithm/huff.rs:9-46) with any window size: every byte a literal; the window base only matters for the
// stored-block rule.  deflate_huff refills when lookahead == 0, so the base moves when strstart reaches 2w + k*w.
static uint32_t huff_base(uint32_t q, uint32_t w) { return q < 2 * w ? 0 : w * (1 + (q - 2 * w) / w); }
static uint32_t huff_final_base(uint32_t N, uint32_t w)
{
    // the last fill_window call (lookahead == 0 at strstart == N) still slides when strstart >= w + max_dist (deflate.rs:1787)
    uint32_t B = N == 0 ? 0 : huff_base(N - 1, w);
    if (N - B >= 2 * w - kMinLookahead) B += w;
    return B;
}
extern "C" int hm_deflate_huff(const uint8_t *data, uint32_t N, int wbits, int mem_level, uint8_t *dst, uint32_t cap, uint32_t *out_len, int *dt)
{
    const uint32_t w = 1u << wbits, block_syms = (1u << (mem_level + 6)) - 1;
    std::vector<Sym> syms(N);
    for (uint32_t p = 0; p < N; p++) syms[p] = Sym{0, data[p], p};
    return encode_stream(data, N, 1 /* header level flags 0: strategy >= Z_HUFFMAN_ONLY */, syms, [&](uint32_t, const BlockDesc &bd) {
        return bd.last ? huff_final_base(N, w) : huff_base(syms[bd.sym_begin + bd.sym_count - 1].pos, w); }, false, false,
        block_syms, dst, cap, out_len, dt, (uint32_t)(wbits - 8));
}


extern "C" int hm_parse_low_w(const uint8_t *data, uint32_t N, int level, int wbits, SymOut *out, uint32_t cap, uint32_t *nsyms)
{
    std::vector<Sym> syms;
    std::vector<uint32_t> blockB;
    run_low(data, N, level, kBlockSyms, syms, blockB, 1u << wbits);
    for (size_t i = 0; i < syms.size() && i < cap; i++) out[i] = SymOut{syms[i].pos, syms[i].dist, syms[i].lc};
    *nsyms = (uint32_t)syms.size();
    return 0;
}

extern "C" int hm_deflate_low_w(const uint8_t *data, uint32_t N, int level, int wbits, int mem_level, uint8_t *dst, uint32_t cap,
                                uint32_t *out_len, int *data_type_out)
{
    const uint32_t block_syms = (1u << (mem_level + 6)) - 1;
    std::vector<Sym> syms;
    std::vector<uint32_t> blockB;
    const uint32_t finalB = run_low(data, N, level, level == 1 ? kBlockSyms : block_syms, syms, blockB, 1u << wbits);
    return encode_stream(data, N, level, syms, [&](uint32_t b, const BlockDesc &bd) { return (bd.last || b >= blockB.size()) ? finalB : blockB[b]; }, level == 1,
                         false, level == 1 ? kBlockSyms : block_syms, dst, cap, out_len, data_type_out, (uint32_t)(wbits - 8));
}

// ------------------------------------------------------------------------------------------
// Experiment (round-2 planning): how many fixed-point iterations does the level-6 parse need from a given initial
// hole set, and how many 32 KiB tiles change per iteration?  mode 0: no holes (what the engine does); mode 1: holes guessed from the
// data alone -- a greedy walk that takes the nearest chain candidate and treats matches >= 257 as long.
// stats[2*i] = changed hole words, stats[2*i+1] = dirty 32 KiB tiles of iteration i.
// ------------------------------------------------------------------------------------------
extern "C" int hm_iter_experiment(const uint8_t *data, uint32_t N, int level, int mode, uint32_t *stats, uint32_t cap, uint32_t *iters_out)
{
    LevelParams lp = level_params(level);
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), newholes((N >> 5) + 2, 0);
    std::vector<uint32_t> M(N + 1024, 0), nxt(N + 1, 0);
    uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    if (mode == 1) {
        uint32_t p = 0;
        while (p + 262 < tail_start) {
            uint32_t d = L[p], len = 0;
            if (d) { while (len < 258 && data[p + len] == data[p - d + len]) len++; }
            if (len >= 257) {
                for (uint32_t y = p + 1; y + 1 < p + len; y++) holes[y >> 5] |= 1u << (y & 31);
                p += len;
            } else p += len >= 4 ? len : 1;
        }
    }
    HostAcc a{data, N, L.data(), holes.data(), M.data()};
    uint32_t iters = 0;
    std::vector<uint32_t> path;
    for (;;) {
        for (uint32_t x = 0; x < N; x++) {
            Match m = (x + kMSafe <= N) ? lm_walk(a, x, 0xffffffffu, lp) : Match{0, 0};
            M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
        }
        for (uint32_t p = 0; p < tail_start; p++) { uint32_t ns; nxt[p] = macro_step(a, p, lp, tail_start, [](Sym) {}, &ns); }
        path.clear();
        uint32_t p = 0;
        while (p < tail_start && nxt[p] < tail_start) { path.push_back(p); p = nxt[p]; }
        std::fill(newholes.begin(), newholes.end(), 0);
        for (uint32_t q : path) {
            uint32_t ns;
            macro_step(a, q, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy)
                    for (uint32_t y = s.pos + 1; y + 1 < s.pos + s.lc + 3; y++) newholes[y >> 5] |= 1u << (y & 31);
            }, &ns);
        }
        uint32_t changed = 0, tiles = 0, last_tile = 0xffffffffu;
        for (uint32_t w = 0; w < holes.size(); w++)
            if (holes[w] != newholes[w]) { changed++; const uint32_t t = w / 1024; if (t != last_tile) { tiles++; last_tile = t; } }
        if (2 * iters + 1 < cap) { stats[2 * iters] = changed; stats[2 * iters + 1] = tiles; }
        iters++;
        if (!changed) break;
        holes = newholes;
        a.holes = holes.data();
        if (iters > N / 257u + 64u) return -1;
    }
    *iters_out = iters;
    return 0;
}

// Experiment ("what the next factor needs"): how much of M would a LAZY evaluation need?  One walker per chunk of C
// positions follows nxt[] from the chunk's first position to its end; the true path enters a chunk somewhere else and is followed
// until it meets the chunk walker's trail.  stats: [0] positions on the true path, [1] positions visited by the chunk walkers,
// [2] true-path nodes not on a walker's trail (the second round's work), [3] the longest such run, [4] chunks whose entry needed > 32 steps.
extern "C" int hm_lazy_experiment(const uint8_t *data, uint32_t N, int level, uint32_t C, uint64_t *stats)
{
    LevelParams lp = level_params(level);
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), M(N + 1024, 0), nxt(N + 1, 0);
    HostAcc a{data, N, L.data(), holes.data(), M.data()};
    const uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    for (uint32_t x = 0; x < N; x++) {
        Match m = (x + kMSafe <= N) ? lm_walk(a, x, 0xffffffffu, lp) : Match{0, 0};
        M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
    }
    for (uint32_t p = 0; p < tail_start; p++) { uint32_t ns; nxt[p] = macro_step(a, p, lp, tail_start, [](Sym) {}, &ns); }
    std::vector<uint8_t> trail(N + 1, 0);
    uint64_t visited = 0;
    for (uint32_t c0 = 0; c0 < tail_start; c0 += C) {
        uint32_t p = c0;
        while (p < tail_start && p < c0 + C) { if (!trail[p]) { trail[p] = 1; visited++; } p = nxt[p]; }
    }
    uint64_t on_path = 0, off = 0, longest = 0, run = 0, slow_chunks = 0;
    uint32_t p = 0, cur_chunk = 0xffffffffu;
    while (p < tail_start) {
        on_path++;
        if (p / C != cur_chunk) { cur_chunk = p / C; if (run > 32) slow_chunks++; run = 0; }
        if (!trail[p]) { off++; run++; if (run > longest) longest = run; } else run = 0;
        p = nxt[p];
    }
    stats[0] = on_path; stats[1] = visited; stats[2] = off; stats[3] = longest; stats[4] = slow_chunks;
    return 0;
}

// The lazy formulation as a whole: M and the macro steps are evaluated on demand along chunk walkers and along the
// true path; the hole fixed point runs on top of it exactly as in hm_parse_parallel.  Output: the symbols (to be compared with the
// oracle's trace) and, per iteration, how many positions had their M evaluated.  evals[i] = M evaluations of iteration i.
struct LazyAcc {
    const uint8_t *data; uint32_t N; const uint16_t *L; const uint32_t *holes; uint32_t *M; uint8_t *have; const LevelParams *lp; uint64_t *count;
    uint32_t byte(uint32_t y) const { while (y >= N) { if (y < 65536) return 0; y -= 32768; } return data[y]; }
    uint32_t link(uint32_t y) const { return y + 4 <= N ? L[y] : 0; }
    bool inserted(uint32_t y) const { return !((holes[y >> 5] >> (y & 31)) & 1u); }
    Match mlook(uint32_t x) const
    {
        if (!have[x]) {
            Match m = (x + kMSafe <= N) ? lm_walk(*this, x, 0xffffffffu, *lp) : Match{0, 0};
            M[x] = m.len ? ((m.len << 16) | (x - m.start)) : 0;
            have[x] = 1;
            (*count)++;
        }
        const uint32_t v = M[x];
        return Match{v >> 16, x - (v & 0xffff)};
    }
};

extern "C" int hm_parse_lazy(const uint8_t *data, uint32_t N, int level, uint32_t C, SymOut *out, uint32_t cap, uint32_t *nsyms, uint64_t *evals,
                             uint32_t evals_cap, uint32_t *iters_out)
{
    LevelParams lp = level_params(level);
    std::vector<uint16_t> L;
    build_links(data, N, L);
    std::vector<uint32_t> holes((N >> 5) + 2, 0), newholes((N >> 5) + 2, 0), M(N + 1024, 0), nxt(N + 1, 0);
    std::vector<uint8_t> have(N + 1024, 0), hnxt(N + 1, 0);
    uint64_t count = 0;
    LazyAcc a{data, N, L.data(), holes.data(), M.data(), have.data(), &lp, &count};
    const uint32_t tail_start = N > 2 * kTailZone ? N - kTailZone : 0;
    auto step = [&](uint32_t p) -> uint32_t {
        if (!hnxt[p]) { uint32_t ns; nxt[p] = macro_step(a, p, lp, tail_start, [](Sym) {}, &ns); hnxt[p] = 1; }
        return nxt[p];
    };
    std::vector<uint32_t> path;
    uint32_t iters = 0, tail_entry = 0;
    for (;;) {
        std::fill(have.begin(), have.end(), 0);
        std::fill(hnxt.begin(), hnxt.end(), 0);
        count = 0;
        // round 1: one walker per chunk (on the GPU: one lane each, all chunks at once)
        for (uint32_t c0 = 0; c0 < tail_start; c0 += C) {
            uint32_t p = c0;
            while (p < tail_start && p < c0 + C) p = step(p);
        }
        // round 2: the true path; where it is off the trails it is evaluated on demand (<= 20 steps per chunk boundary on the corpus)
        path.clear();
        uint32_t p = 0;
        while (p < tail_start) {
            const uint32_t np = step(p);
            if (np >= tail_start) break;
            path.push_back(p);
            p = np;
        }
        tail_entry = p;
        if (iters < evals_cap) evals[iters] = count;
        iters++;
        std::fill(newholes.begin(), newholes.end(), 0);
        for (uint32_t q : path) {
            uint32_t ns;
            macro_step(a, q, lp, tail_start, [&](Sym s) {
                if (s.dist && (uint32_t)s.lc + 3 > 16 * lp.lazy)
                    for (uint32_t y = s.pos + 1; y + 1 < s.pos + s.lc + 3; y++) newholes[y >> 5] |= 1u << (y & 31);
            }, &ns);
        }
        if (newholes == holes) break;
        holes = newholes;
        a.holes = holes.data();
        if (iters > N / 257u + 64u) return -1;
    }
    uint32_t n = 0;
    for (uint32_t q : path) {
        uint32_t ns;
        macro_step(a, q, lp, tail_start, [&](Sym s) { if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc}; n++; }, &ns);
    }
    std::vector<uint32_t> ins(64 + (N - tail_entry) / 32 + 2, 0);
    serial_medium(a, N, tail_entry, ins.data(), (uint32_t)ins.size(), lp, [&](Sym s, uint32_t) {
        if (n < cap) out[n] = SymOut{s.pos, s.dist, s.lc};
        n++;
    });
    *nsyms = n;
    *iters_out = iters;
    return 0;
}
