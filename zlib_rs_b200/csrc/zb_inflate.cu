// zb_inflate.cu -- GPU inflate (sm_100a), round-1 version: one warp per stream.
//
// Reference: zlib-rs/src/inflate.rs (Mode machine :896-1839, inflate_fast_help :1880-2158),
// inflate/inftrees.rs:42-245 (two-level decode tables), inflate/writer.rs (match copy),
// inflate/window.rs.  Lane 0 runs the bit reader and the table-driven literal/length/distance decode
// out of shared memory; the 32 lanes cooperate on input refill (128-bit loads into a shared-memory ring),
// long match copies, and 128-bit flushes of the 64 KiB output ring to HBM.  The check value (adler32 /
// crc32) is computed afterwards by the checksum kernels over the device-resident output.
// A single foreign stream is inherently serial in its Huffman decode; parallel speculative decode is the
// next-round item.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "zb_engine_internal.h"

namespace zb {

constexpr uint32_t kEnoughLens = 1332, kEnoughDists = 592;
constexpr uint32_t kInRing = 8192, kOutRing = 65536;

struct ICode { uint8_t op, bits; uint16_t val; };

enum InfErr {
    IE_OK = 0, IE_HEADER_CHECK, IE_METHOD, IE_WINDOW, IE_BLOCK_TYPE, IE_STORED_LEN, IE_TOO_MANY, IE_CODE_LENS, IE_REPEAT,
    IE_NO_EOB, IE_LITLEN_SET, IE_DIST_SET, IE_LITLEN_CODE, IE_DIST_CODE, IE_TOO_FAR, IE_TRUNCATED, IE_OUTPUT_FULL, IE_GZ_FLAGS,
    IE_HCRC, IE_NEED_DICT
};

struct InfState { // device result block
    uint64_t out_bytes, in_bytes;
    uint32_t err;
    uint32_t trailer_check; // value stored in the stream
    uint32_t trailer_len;   // gzip ISIZE
    uint32_t kind;          // 0 raw, 1 zlib, 2 gzip
};

__device__ static const uint16_t d_lbase[31] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0};
__device__ static const uint8_t d_lext[31] = {16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 18, 18, 18, 18, 19, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21, 16, 77, 202};
__device__ static const uint16_t d_dbase[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
__device__ static const uint8_t d_dext[32] = {16, 16, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22, 23, 23, 24, 24, 25, 25, 26, 26, 27, 27, 28, 28, 29, 29, 64, 64};

// inflate/inftrees.rs:42-245.  type: 0 codes, 1 lens, 2 dists.  Returns 0 ok, else failure.
__device__ static int inflate_table(int type, const uint16_t *lens, uint32_t codes, ICode *table, uint32_t bits, uint16_t *work,
                                    uint32_t *root_out)
{
    uint16_t count[16], offs[16];
    uint32_t min = 15, max = 0, len, root, curr, drop;
    for (len = 0; len < 16; len++) count[len] = 0;
    for (uint32_t i = 0; i < codes; i++)
        if (lens[i]) { count[lens[i]]++; if (lens[i] > max) max = lens[i]; if (lens[i] < min) min = lens[i]; }
    if (max == 0) {
        ICode c = {64, 1, 0};
        table[0] = table[1] = c;
        *root_out = 1;
        return 0;
    }
    root = bits < min ? min : bits > max ? max : bits;
    int left = 1;
    for (len = 1; len <= 15; len++) { left = (left << 1) - count[len]; if (left < 0) return -1; }
    if (left > 0 && (type == 0 || max != 1)) return -1;
    offs[0] = offs[1] = 0;
    for (len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (uint32_t sym = 0; sym < codes; sym++) if (lens[sym]) work[offs[lens[sym]]++] = (uint16_t)sym;
    const uint32_t match = type == 0 ? 20 : type == 1 ? 257 : 0;
    uint32_t used = 1u << root;
    if ((type == 1 && used > kEnoughLens) || (type == 2 && used > kEnoughDists)) return 1;
    uint32_t huff = 0, next = 0, low = 0xffffffffu, mask = used - 1, sym = 0, rhuff = 0;
    len = min; curr = root; drop = 0;
    for (;;) {
        ICode here;
        here.bits = (uint8_t)(len - drop);
        const uint32_t w = work[sym];
        if (w >= match) {
            here.op = type == 1 ? d_lext[w - match] : d_dext[w - match];
            here.val = type == 1 ? d_lbase[w - match] : d_dbase[w - match];
        } else if (w + 1 < match) { here.op = 0; here.val = (uint16_t)w; }
        else { here.op = 96; here.val = 0; }
        const uint32_t incr = 1u << (len - drop);
        uint32_t fill = 1u << curr;
        const uint32_t mn = fill;
        do { fill -= incr; table[next + (huff >> drop) + fill] = here; } while (fill != 0);
        rhuff += 0x80000000u >> (len - 1);
        huff = __brev(rhuff);
        sym++;
        if (--count[len] == 0) {
            if (len == max) break;
            len = lens[work[sym]];
        }
        if (len > root && (huff & mask) != low) {
            if (drop == 0) drop = root;
            next += mn;
            curr = len - drop;
            int l2 = 1 << curr;
            while (curr + drop < max) {
                l2 -= count[curr + drop];
                if (l2 <= 0) break;
                curr++;
                l2 <<= 1;
            }
            used += 1u << curr;
            if ((type == 1 && used > kEnoughLens) || (type == 2 && used > kEnoughDists)) return 1;
            low = huff & mask;
            table[low].op = (uint8_t)curr;
            table[low].bits = (uint8_t)root;
            table[low].val = (uint16_t)next;
        }
    }
    if (huff != 0) { ICode h = {64, (uint8_t)(len - drop), 0}; table[next + huff] = h; }
    *root_out = root;
    return 0;
}

struct InfShared {
    ICode lencode[kEnoughLens], distcode[kEnoughDists], lenfix[512], distfix[32];
    uint16_t lens[320], work[288];
    uint8_t in[kInRing];
    uint8_t out[kOutRing];
};

enum Cmd { C_NONE = 0, C_REFILL, C_FLUSH, C_COPY, C_DONE };

__global__ void __launch_bounds__(32) k_inflate(const uint8_t *__restrict__ src, uint64_t n, uint8_t *__restrict__ dst, uint64_t cap,
                                                int window_bits, InfState *res)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InfShared &S = *reinterpret_cast<InfShared *>(smem_raw);
    const uint32_t lane = threadIdx.x;
    // shared between lanes through shuffles from lane 0
    uint64_t ifill = 0;   // input bytes loaded into the ring so far (absolute)
    uint64_t oflush = 0;  // output bytes flushed to HBM so far
    // lane-0 state
    uint64_t hold = 0, ipos = 0, opos = 0;
    uint32_t bits = 0;
    uint64_t consumed_bits = 0; // bits taken from the stream (zero padding beyond n is detected with this)
    uint32_t err = IE_OK;
    int last = 0, mode = 0; // 0 header, 1 block header, 2 stored, 3 codes, 4 trailer, 5 done
    uint32_t lenbits = 9, distbits = 5;
    const ICode *lencode = S.lenfix, *distcode = S.distfix;
    uint32_t stored_left = 0;
    uint32_t kind = window_bits < 0 ? 0 : 1; // refined by the header
    uint32_t tr_check = 0, tr_len = 0;
    uint32_t copy_len = 0, copy_dist = 0;
    uint32_t gz_fl = 0, gz_xl = 0;

    if (lane == 0) {
        uint32_t root;
        uint32_t sym = 0;
        while (sym < 144) S.lens[sym++] = 8;
        while (sym < 256) S.lens[sym++] = 9;
        while (sym < 280) S.lens[sym++] = 7;
        while (sym < 288) S.lens[sym++] = 8;
        inflate_table(1, S.lens, 288, S.lenfix, 9, S.work, &root);
        for (sym = 0; sym < 32; sym++) S.lens[sym] = 5;
        inflate_table(2, S.lens, 32, S.distfix, 5, S.work, &root);
    }
    __syncwarp();

#define NEED(nb) do { while (bits < (nb)) { hold |= (uint64_t)((ipos < n) ? S.in[ipos & (kInRing - 1)] : 0) << bits; ipos++; bits += 8; } } while (0)
#define BITS(nb) ((uint32_t)(hold & ((1ull << (nb)) - 1)))
#define DROP(nb) do { const uint32_t nb_ = (nb); hold >>= nb_; bits -= nb_; consumed_bits += nb_; } while (0)
#define FAIL(code) do { err = (code); mode = 5; } while (0)

    for (;;) {
        uint32_t cmd = C_NONE;
        if (lane == 0) {
            // run until a cooperative action is needed
            while (cmd == C_NONE) {
                if (mode == 5) { cmd = C_DONE; break; }
                if (ifill < n && ifill - ipos < 1024) { cmd = C_REFILL; break; }
                if (opos - oflush >= kOutRing / 2 + 2048) { cmd = C_FLUSH; break; }
                if (consumed_bits > 8 * n) { FAIL(IE_TRUNCATED); continue; }
                if (mode == 0) {
                    if (window_bits < 0) { mode = 1; continue; }
                    NEED(16);
                    const uint32_t h = BITS(16);
                    if ((window_bits > 15) && h == 0x8b1f) { // gzip (inflate.rs:934-946)
                        kind = 2;
                        DROP(16);
                        NEED(16);
                        const uint32_t fl = BITS(16);
                        DROP(16);
                        if ((fl & 0xff) != 8) { FAIL(IE_METHOD); continue; }
                        if (fl & 0xe000) { FAIL(IE_GZ_FLAGS); continue; }
                        NEED(32); DROP(32); // mtime
                        NEED(16); DROP(16); // xfl, os
                        gz_fl = fl;
                        if (fl & 0x0400) { NEED(16); gz_xl = BITS(16); DROP(16); }
                        mode = 6;
                        continue;
                    }
                    if (window_bits > 15 && window_bits < 32) { FAIL(IE_HEADER_CHECK); continue; } // gzip only
                    kind = 1;
                    if ((((h & 0xff) << 8) + (h >> 8)) % 31) { FAIL(IE_HEADER_CHECK); continue; }
                    if ((h & 0xf) != 8) { FAIL(IE_METHOD); continue; }
                    const uint32_t wb = ((h >> 4) & 0xf) + 8;
                    const uint32_t want = (uint32_t)(window_bits & 15);
                    if (wb > 15 || (want != 0 && wb > want)) { FAIL(IE_WINDOW); continue; }
                    if (h & 0x2000) { FAIL(IE_NEED_DICT); continue; }
                    DROP(16);
                    mode = 1;
                    continue;
                }
                if (mode >= 6) { // gzip optional fields, at most 512 bytes per round so the input ring never starves
                    uint32_t k = 0;
                    if (mode == 6) {
                        while ((gz_fl & 0x0400) && gz_xl && k < 512) { NEED(8); DROP(8); gz_xl--; k++; }
                        if (!(gz_fl & 0x0400) || gz_xl == 0) mode = 7;
                    } else if (mode == 7 || mode == 8) {
                        const uint32_t flag = mode == 7 ? 0x0800u : 0x1000u;
                        bool end = !(gz_fl & flag);
                        while (!end && k < 512) { NEED(8); const uint32_t c = BITS(8); DROP(8); k++; if (!c) end = true; }
                        if (end) mode = mode + 1;
                    } else { // 9: header crc (not verified, as with inflateValidate(0)); then the first block
                        if (gz_fl & 0x0200) { NEED(16); DROP(16); }
                        mode = 1;
                    }
                    continue;
                }
                if (mode == 1) {
                    if (last) { DROP(bits & 7); mode = 4; continue; }
                    NEED(3);
                    last = (int)BITS(1);
                    const uint32_t type = (BITS(3) >> 1);
                    DROP(3);
                    if (type == 0) {
                        DROP(bits & 7);
                        NEED(32);
                        const uint32_t v = BITS(32);
                        DROP(32);
                        if ((v & 0xffff) != ((v >> 16) ^ 0xffff)) { FAIL(IE_STORED_LEN); continue; }
                        stored_left = v & 0xffff;
                        mode = 2;
                    } else if (type == 1) {
                        lencode = S.lenfix; distcode = S.distfix; lenbits = 9; distbits = 5;
                        mode = 3;
                    } else if (type == 2) {
                        NEED(14);
                        const uint32_t nlen = BITS(5) + 257; DROP(5);
                        const uint32_t ndist = BITS(5) + 1; DROP(5);
                        const uint32_t ncode = BITS(4) + 4; DROP(4);
                        if (nlen > 286 || ndist > 30) { FAIL(IE_TOO_MANY); continue; }
                        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                        uint32_t have = 0;
                        while (have < ncode) { NEED(3); S.lens[order[have++]] = (uint16_t)BITS(3); DROP(3); }
                        while (have < 19) S.lens[order[have++]] = 0;
                        uint32_t root;
                        if (inflate_table(0, S.lens, 19, S.lencode, 7, S.work, &root)) { FAIL(IE_CODE_LENS); continue; }
                        lenbits = root;
                        have = 0;
                        bool bad = false;
                        while (have < nlen + ndist) {
                            NEED(lenbits + 7);
                            ICode here = S.lencode[BITS(lenbits)];
                            if (here.val < 16) { DROP(here.bits); S.lens[have++] = here.val; continue; }
                            uint32_t len = 0, copy;
                            DROP(here.bits);
                            if (here.val == 16) {
                                if (have == 0) { FAIL(IE_REPEAT); bad = true; break; }
                                len = S.lens[have - 1];
                                copy = 3 + BITS(2); DROP(2);
                            } else if (here.val == 17) { copy = 3 + BITS(3); DROP(3); }
                            else { copy = 11 + BITS(7); DROP(7); }
                            if (have + copy > nlen + ndist) { FAIL(IE_REPEAT); bad = true; break; }
                            while (copy--) S.lens[have++] = (uint16_t)len;
                            if (consumed_bits > 8 * n) { FAIL(IE_TRUNCATED); bad = true; break; }
                        }
                        if (bad) continue;
                        if (S.lens[256] == 0) { FAIL(IE_NO_EOB); continue; }
                        if (inflate_table(1, S.lens, nlen, S.lencode, 10, S.work, &root)) { FAIL(IE_LITLEN_SET); continue; }
                        lenbits = root;
                        if (inflate_table(2, S.lens + nlen, ndist, S.distcode, 9, S.work, &root)) { FAIL(IE_DIST_SET); continue; }
                        distbits = root;
                        lencode = S.lencode; distcode = S.distcode;
                        mode = 3;
                    } else { FAIL(IE_BLOCK_TYPE); }
                    continue;
                }
                if (mode == 2) { // stored bytes (bit buffer is byte aligned here)
                    uint32_t k = 0;
                    while (stored_left && k < 4096) {
                        if (ifill < n && ifill - ipos < 16) break;
                        if (opos >= cap) { FAIL(IE_OUTPUT_FULL); break; }
                        NEED(8);
                        S.out[opos & (kOutRing - 1)] = (uint8_t)BITS(8);
                        DROP(8);
                        opos++; stored_left--; k++;
                        if (consumed_bits > 8 * n) { FAIL(IE_TRUNCATED); break; }
                    }
                    if (mode == 2 && stored_left == 0) mode = 1;
                    continue; // re-check refill/flush conditions
                }
                if (mode == 3) { // literal/length/distance loop (inflate.rs:1918-2158)
                    uint32_t budget = 512;
                    while (budget--) {
                        if (ifill < n && ifill - ipos < 64) break;
                        if (opos - oflush >= kOutRing - 4096) break;
                        NEED(48);
                        ICode here = lencode[BITS(lenbits)];
                        if (here.op && (here.op & 0xf0) == 0) {
                            const ICode l = here;
                            here = lencode[l.val + (BITS(l.bits + l.op) >> l.bits)];
                            DROP(l.bits);
                        }
                        DROP(here.bits);
                        if (here.op == 0) {
                            if (opos >= cap) { FAIL(IE_OUTPUT_FULL); break; }
                            S.out[opos & (kOutRing - 1)] = (uint8_t)here.val;
                            opos++;
                            continue;
                        }
                        if (here.op & 32) { mode = 1; break; }
                        if (here.op & 64) { FAIL(IE_LITLEN_CODE); break; }
                        uint32_t len = here.val;
                        uint32_t ex = here.op & 15;
                        if (ex) { len += BITS(ex); DROP(ex); }
                        here = distcode[BITS(distbits)];
                        if ((here.op & 0xf0) == 0) {
                            const ICode l = here;
                            here = distcode[l.val + (BITS(l.bits + l.op) >> l.bits)];
                            DROP(l.bits);
                        }
                        DROP(here.bits);
                        if (here.op & 64) { FAIL(IE_DIST_CODE); break; }
                        uint32_t dist = here.val;
                        ex = here.op & 15;
                        if (ex) { dist += BITS(ex); DROP(ex); }
                        if (dist > opos) { FAIL(IE_TOO_FAR); break; }
                        if (opos + len > cap) { FAIL(IE_OUTPUT_FULL); break; }
                        if (consumed_bits > 8 * n) { FAIL(IE_TRUNCATED); break; }
                        if (len >= 24) { copy_len = len; copy_dist = dist; cmd = C_COPY; break; }
                        for (uint32_t j = 0; j < len; j++) S.out[(opos + j) & (kOutRing - 1)] = S.out[(opos + j - dist) & (kOutRing - 1)];
                        opos += len;
                    }
                    continue;
                }
                if (mode == 4) { // trailer (inflate.rs:1398-1430, 1779-1795)
                    if (kind == 1) {
                        NEED(32);
                        const uint32_t v = BITS(32);
                        DROP(32);
                        tr_check = __byte_perm(v, 0, 0x0123);
                    } else if (kind == 2) {
                        NEED(32); tr_check = BITS(32); DROP(32);
                        NEED(32); tr_len = BITS(32); DROP(32);
                    }
                    if (consumed_bits > 8 * n) { FAIL(IE_TRUNCATED); continue; }
                    mode = 5;
                    continue;
                }
            }
        }
        cmd = __shfl_sync(0xffffffffu, cmd, 0);
        if (cmd == C_REFILL) {
            const uint64_t ipos0 = __shfl_sync(0xffffffffu, ipos, 0);
            // fill up to the ring size, never overwriting unread bytes
            uint64_t room = kInRing - (ifill - ipos0);
            uint64_t cnt = n - ifill < room ? n - ifill : room;
            for (uint64_t i = lane; i < cnt; i += 32) S.in[(ifill + i) & (kInRing - 1)] = src[ifill + i];
            ifill += cnt;
            __syncwarp();
        } else if (cmd == C_FLUSH || cmd == C_DONE) {
            const uint64_t opos0 = __shfl_sync(0xffffffffu, opos, 0);
            const uint64_t upto = cmd == C_DONE ? opos0 : oflush + kOutRing / 2; // the trigger guarantees opos0 >= upto
            const uint64_t end = upto > cap ? cap : upto;
            for (uint64_t i = oflush + lane; i < end; i += 32) dst[i] = S.out[i & (kOutRing - 1)];
            if (end > oflush) oflush = end;
            __syncwarp();
            if (cmd == C_DONE) break;
        } else if (cmd == C_COPY) {
            const uint32_t len = __shfl_sync(0xffffffffu, copy_len, 0);
            const uint32_t dist = __shfl_sync(0xffffffffu, copy_dist, 0);
            const uint64_t o = __shfl_sync(0xffffffffu, opos, 0);
            if (dist >= 32) {
                for (uint32_t b = 0; b < len; b += 32) {
                    const uint32_t j = b + lane;
                    uint8_t v = 0;
                    if (j < len) v = S.out[(o + j - dist) & (kOutRing - 1)];
                    __syncwarp();
                    if (j < len) S.out[(o + j) & (kOutRing - 1)] = v;
                    __syncwarp();
                }
            } else {
                for (uint32_t b = 0; b < len; b += 32) {
                    const uint32_t j = b + lane;
                    if (j < len) S.out[(o + j) & (kOutRing - 1)] = S.out[(o - dist + (j % dist)) & (kOutRing - 1)];
                }
                __syncwarp();
            }
            if (lane == 0) opos += len;
        }
    }
    if (lane == 0) {
        res->out_bytes = opos;
        res->in_bytes = (consumed_bits + 7) >> 3;
        res->err = err;
        res->trailer_check = tr_check;
        res->trailer_len = tr_len;
        res->kind = kind;
    }
#undef NEED
#undef BITS
#undef DROP
#undef FAIL
}

static const char *inf_msg(uint32_t e)
{
    switch (e) {
    case IE_HEADER_CHECK: return "incorrect header check";
    case IE_METHOD: return "unknown compression method";
    case IE_WINDOW: return "invalid window size";
    case IE_BLOCK_TYPE: return "invalid block type";
    case IE_STORED_LEN: return "invalid stored block lengths";
    case IE_TOO_MANY: return "too many length or distance symbols";
    case IE_CODE_LENS: return "invalid code lengths set";
    case IE_REPEAT: return "invalid bit length repeat";
    case IE_NO_EOB: return "invalid code -- missing end-of-block";
    case IE_LITLEN_SET: return "invalid literal/lengths set";
    case IE_DIST_SET: return "invalid distances set";
    case IE_LITLEN_CODE: return "invalid literal/length code";
    case IE_DIST_CODE: return "invalid distance code";
    case IE_TOO_FAR: return "invalid distance too far back";
    case IE_GZ_FLAGS: return "unknown header flags set";
    case IE_HCRC: return "header crc mismatch";
    case IE_NEED_DICT: return "need dictionary";
    case IE_TRUNCATED: return "unexpected end of input";
    default: return "";
    }
}

int Engine::inflate_init()
{
    cudaError_t e = cudaFuncSetAttribute(k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InfShared));
    if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "k_inflate attr: %s", cudaGetErrorString(e)); return ZB_E_CUDA; }
    if (cudaMalloc(&d_inf_state, sizeof(InfState)) != cudaSuccess) return ZB_E_MEM;
    if (cudaMallocHost(&h_inf_state, sizeof(InfState)) != cudaSuccess) return ZB_E_MEM;
    return ZB_OK;
}

#define CKI(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s: %s", #call, cudaGetErrorString(e_)); return ZB_E_CUDA; } \
    } while (0)

int Engine::inflate(const void *src, size_t n, bool src_dev, void *dst, size_t dst_cap, bool dst_dev, int window_bits,
                    zb_inflate_result *res)
{
    if (!res || (!src && n) || (!dst && dst_cap)) return ZB_E_PARAM;
    memset(res, 0, sizeof *res);
    if (window_bits < 0) { if (window_bits < -15 || window_bits > -8) return ZB_E_PARAM; }
    else if (window_bits != 0 && ((window_bits & 15) < 8) ) return ZB_E_PARAM;
    if (window_bits > 47) return ZB_E_PARAM;
    CKI(cudaSetDevice(device));
    int rc;
    void *p;
    const uint8_t *d_src = static_cast<const uint8_t *>(src);
    CKI(cudaEventRecord(ev0, st));
    if (!src_dev) {
        if ((rc = reserve(19 /*S_INF0*/, n + 64, &p)) != ZB_OK) return rc;
        if (n) CKI(cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, st));
        d_src = static_cast<const uint8_t *>(p);
    }
    uint8_t *d_dst = static_cast<uint8_t *>(dst);
    if (!dst_dev) {
        if ((rc = reserve(20 /*S_INF1*/, dst_cap + 64, &p)) != ZB_OK) return rc;
        d_dst = static_cast<uint8_t *>(p);
    }
    InfState *dis = static_cast<InfState *>(d_inf_state), *his = static_cast<InfState *>(h_inf_state);
    k_inflate<<<1, 32, sizeof(InfShared), st>>>(d_src, n, d_dst, dst_cap, window_bits, dis);
    launches = 1;
    CKI(cudaMemcpyAsync(his, dis, sizeof(InfState), cudaMemcpyDeviceToHost, st));
    CKI(cudaStreamSynchronize(st));
    CKI(cudaGetLastError());
    res->out_bytes = his->out_bytes;
    res->in_bytes = his->in_bytes;
    int status = ZB_OK;
    if (his->err == IE_OUTPUT_FULL) status = ZB_E_BUF;
    else if (his->err != IE_OK) { status = ZB_E_DATA; snprintf(res->msg, sizeof res->msg, "%s", inf_msg(his->err)); }
    // check value of what was produced (inflate.rs:1398-1430 verifies it against the trailer)
    uint32_t check = his->kind == 2 ? 0 : 1;
    if (his->kind != 0 && his->out_bytes) {
        void *d_ck;
        const size_t ck_bytes = ((size_t)his->out_bytes / 16384 + 16) * 8;
        if ((rc = reserve(18 /*S_CK*/, ck_bytes, &d_ck)) != ZB_OK) return rc;
        if (his->kind == 2) CKI(launch_crc32(d_dst, his->out_bytes, 0, d_ck, ck_bytes, d_check, st));
        else CKI(launch_adler32(d_dst, his->out_bytes, 1, d_ck, ck_bytes, d_check, st));
        launches += 2;
        CKI(cudaMemcpyAsync(h_info, d_check, 4, cudaMemcpyDeviceToHost, st));
        CKI(cudaStreamSynchronize(st));
        check = *reinterpret_cast<uint32_t *>(h_info);
    }
    res->check = check;
    if (status == ZB_OK && his->kind != 0) {
        if (check != his->trailer_check) { status = ZB_E_DATA; snprintf(res->msg, sizeof res->msg, "incorrect data check"); }
        else if (his->kind == 2 && (uint32_t)his->out_bytes != his->trailer_len) { status = ZB_E_DATA; snprintf(res->msg, sizeof res->msg, "incorrect length check"); }
    }
    if (!dst_dev && his->out_bytes) CKI(cudaMemcpyAsync(dst, d_dst, his->out_bytes, cudaMemcpyDeviceToHost, st));
    CKI(cudaEventRecord(ev1, st));
    CKI(cudaStreamSynchronize(st));
    CKI(cudaEventElapsedTime(&res->gpu_ms, ev0, ev1));
    res->status = status;
    res->gpu_launches = launches;
    return status;
}

} // namespace zb
