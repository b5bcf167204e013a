// zb_inflate_core.h -- device functions of the block-parallel inflate (host-testable like zb_core.h).
//
// A foreign deflate stream carries no index, so its blocks cannot be located without decoding -- except that
// a dynamic block header is highly self-checking: BTYPE, HLIT/HDIST ranges, a COMPLETE code-length code and
// complete literal/length and distance codes (the conditions of inflate_table, zlib-rs/src/inflate/inftrees.rs:42-245,
// and of the CodeLens state, zlib-rs/src/inflate.rs:1660-1777).  The engine therefore
//   1. tests EVERY bit position for a valid dynamic header (scout), in parallel;
//   2. decodes every candidate block without producing output, to learn its end bit and its output length;
//   3. walks the chain first block -> end -> next candidate ... (stored blocks are followed directly);
//   4. decodes all chained blocks in parallel into 16-bit symbols (a match that reaches back before its own
//      block leaves a marker: an index into the 32 KiB window in front of the block);
//   5. resolves the markers block after block.
// False candidates never reach the chain; a stream the chain cannot follow (fixed-code blocks, damage) is handed
// to the serial decoder kernel, which also produces the reference error messages.
#pragma once
#include <stdint.h>
#include "zb_core.h"

namespace zb {

struct BitSrc {
    const uint8_t *p;
    uint64_t nbytes;
    // 32 bits starting at bit position b (zero beyond the end)
    ZB_HD uint32_t peek32(uint64_t b) const
    {
        const uint64_t by = b >> 3;
        uint64_t v = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const uint64_t a = by + i;
            v |= (uint64_t)(a < nbytes ? p[a] : 0) << (8 * i);
        }
        return (uint32_t)(v >> (b & 7));
    }
};

ZB_HD uint32_t cl_order(uint32_t i)
{
    // 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15
    return i < 3 ? 16 + i : (i & 1) ? (i == 3 ? 0 : 8 - (i - 3) / 2) : 8 + (i - 4) / 2;
}

struct DynHeader {
    uint32_t hlit, hdist, hclen;  // counts (257.., 1.., 4..)
    uint64_t body_bit;            // first bit after the code-length section
    uint32_t bfinal;
};

// Kraft test of a set of code lengths: returns 0 complete, >0 incomplete, <0 over-subscribed.
ZB_HD int kraft_left(const uint16_t *count, int maxbits)
{
    int left = 1;
    for (int len = 1; len <= maxbits; len++) {
        left = (left << 1) - (int)count[len];
        if (left < 0) return -1;
    }
    return left;
}

// Stage A+B of the scout: is there a structurally valid dynamic block header at bit b?
// `lens` (>= 320 entries) receives the literal/length + distance code lengths on success.
ZB_HDN inline bool parse_dynamic_header(const BitSrc &s, uint64_t b, DynHeader &h, uint16_t *lens)
{
    const uint64_t nbits = s.nbytes * 8;
    if (b + 17 + 12 > nbits) return false;
    const uint32_t w = s.peek32(b);
    if (((w >> 1) & 3u) != 2u) return false;
    h.bfinal = w & 1u;
    h.hlit = ((w >> 3) & 31u) + 257u;
    h.hdist = ((w >> 8) & 31u) + 1u;
    h.hclen = ((w >> 13) & 15u) + 4u;
    if (h.hlit > 286 || h.hdist > 30) return false; // "too many length or distance symbols"
    uint64_t pos = b + 17;
    uint16_t cnt[16];
    uint8_t cl[19];
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < h.hclen; i++) {
        const uint32_t v = (s.peek32(pos) & 7u);
        pos += 3;
        cl[cl_order(i)] = (uint8_t)v;
        cnt[v]++;
    }
    cnt[0] = 0;
    // code-length code must be complete (inflate_table CODES: an incomplete set is invalid), and not empty
    {
        int any = 0;
        for (int i = 1; i <= 7; i++) any += cnt[i];
        if (!any) return false;
        if (kraft_left(cnt, 7) != 0) return false;
    }
    if (pos > nbits) return false;
    // canonical decode of the code-length code, bit by bit
    uint16_t first[9], offs[9], sorted[19];
    {
        uint32_t code = 0, o = 0;
        for (int len = 1; len <= 7; len++) { first[len] = (uint16_t)code; offs[len] = (uint16_t)o; code = (code + cnt[len]) << 1; o += cnt[len]; }
        uint16_t nx[9];
        for (int len = 1; len <= 7; len++) nx[len] = offs[len];
        for (uint32_t sym = 0; sym < 19; sym++) if (cl[sym]) sorted[nx[cl[sym]]++] = (uint16_t)sym;
    }
    const uint32_t total = h.hlit + h.hdist;
    uint32_t have = 0;
    while (have < total) {
        if (pos + 7 > nbits + 64) return false;
        uint32_t bits = s.peek32(pos);
        uint32_t code = 0, sym = 0xffff;
        for (int len = 1; len <= 7; len++) {
            code = (code << 1) | (bits & 1u);
            bits >>= 1;
            const uint32_t idx = code - first[len];
            if (cnt[len] && code >= first[len] && idx < cnt[len]) { sym = sorted[offs[len] + idx]; pos += len; break; }
        }
        if (sym == 0xffff) return false;
        if (sym < 16) { lens[have++] = (uint16_t)sym; continue; }
        uint32_t rep, val = 0;
        const uint32_t eb = s.peek32(pos);
        if (sym == 16) {
            if (have == 0) return false; // "invalid bit length repeat"
            val = lens[have - 1];
            rep = 3 + (eb & 3u); pos += 2;
        } else if (sym == 17) { rep = 3 + (eb & 7u); pos += 3; }
        else { rep = 11 + (eb & 127u); pos += 7; }
        if (have + rep > total) return false;
        while (rep--) lens[have++] = (uint16_t)val;
    }
    if (pos > nbits) return false;
    if (lens[256] == 0) return false; // "invalid code -- missing end-of-block"
    // literal/length and distance codes: not over-subscribed; incomplete only as a single 1-bit code
    for (int t = 0; t < 2; t++) {
        const uint16_t *L = t == 0 ? lens : lens + h.hlit;
        const uint32_t n = t == 0 ? h.hlit : h.hdist;
        uint16_t c[16];
        for (int i = 0; i < 16; i++) c[i] = 0;
        uint32_t mx = 0;
        for (uint32_t i = 0; i < n; i++) { c[L[i]]++; if (L[i] > mx) mx = L[i]; }
        c[0] = 0;
        if (mx == 0) { if (t == 0) return false; continue; } // no distance codes at all is accepted by inflate_table
        const int left = kraft_left(c, 15);
        if (left < 0) return false;
        if (left > 0 && mx != 1) return false;
    }
    h.body_bit = pos;
    return true;
}

// deflate symbol tables (RFC 1951 3.2.5), as in inflate/inftrees.rs LBASE/LEXT/DBASE/DEXT
ZB_HD uint32_t len_base(uint32_t c) { return c < 8 ? 3 + c : c == 28 ? 258 : 3 + ((4 + (c & 3)) << ((c - 4) >> 2)) - 0; }
ZB_HD uint32_t len_extra(uint32_t c) { return (c < 8 || c == 28) ? 0 : (c - 4) >> 2; }
ZB_HD uint32_t dist_base(uint32_t c) { return c < 4 ? 1 + c : 1 + ((2 + (c & 1)) << ((c - 2) >> 1)); }
ZB_HD uint32_t dist_extra(uint32_t c) { return c < 4 ? 0 : (c - 2) >> 1; }

} // namespace zb
