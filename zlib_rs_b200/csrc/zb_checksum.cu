// zb_checksum.cu -- adler32 / crc32 kernels (sm_100a).
//
// Replaces zlib-rs/src/adler32.rs:19-47 (+ SIMD variants) and zlib-rs/src/crc32.rs:19-29 (braid /
// pclmulqdq folding).  Both checksums are chunk-parallel:
//   adler32 : a range contributes (A = sum d_i, B = sum (n-i) d_i, n); concatenation is
//             A = A1+A2, B = B1 + n2*A1 + B2 (mod 65521)       [the algebra of adler32_combine, adler32.rs:58-87]
//   crc32   : the raw CRC (zero register, no final xor) of a concatenation is raw(A)*x^(8|B|) mod P ^ raw(B)
//             [crc32_combine_op, crc32/combine.rs:3-61]; leading zero bytes do not change a raw CRC, so all
//             pieces are aligned to the END of the buffer and every shift is by a full piece length.
// One pass over the data with 128-bit loads; the partials are reduced by one CTA with a log-depth tree.
#include <cuda_runtime.h>
#include <stdint.h>

namespace zb {

constexpr uint32_t kAdlerBase = 65521u;
constexpr uint32_t kCrcPoly = 0xedb88320u;

// ------------------------------------------------------------------------------------------------
// adler32
// ------------------------------------------------------------------------------------------------
struct AdlerPartial { uint32_t a, b; };

__global__ void __launch_bounds__(256) k_adler_partial(const uint8_t *__restrict__ buf, uint64_t len, uint32_t chunk,
                                                        AdlerPartial *__restrict__ part)
{
    const uint64_t c0 = (uint64_t)blockIdx.x * chunk;
    const uint32_t n = (uint32_t)min((uint64_t)chunk, len - c0);
    const uint8_t *p = buf + c0;
    unsigned long long a = 0, b = 0;
    const uint32_t tid = threadIdx.x;
    const bool aligned = (((uintptr_t)p) & 15) == 0;
    const uint32_t n16 = aligned ? n / 16 : 0;
    for (uint32_t g = tid; g < n16; g += 256) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + g);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t s = 0, wt = 0; // wt = sum j * d_j, j = 0..15  (byte sums with DP4A)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t sk = __dp4a(w[k], 0x01010101u, 0u);
            s += sk;
            wt += (4 * k) * sk + __dp4a(w[k], 0x03020100u, 0u);
        }
        a += s;
        b += (unsigned long long)(n - g * 16) * s - wt;
    }
    for (uint32_t i = n16 * 16 + tid; i < n; i += 256) {
        const uint32_t d = p[i];
        a += d;
        b += (unsigned long long)(n - i) * d;
    }
    __shared__ unsigned long long sa[8], sb[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((tid & 31) == 0) { sa[tid >> 5] = a; sb[tid >> 5] = b; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 8; i++) { a += sa[i]; b += sb[i]; }
        part[blockIdx.x] = AdlerPartial{(uint32_t)(a % kAdlerBase), (uint32_t)(b % kAdlerBase)};
    }
}

// (A, B, n) triples combine associatively; one CTA reduces all chunk partials.
__global__ void __launch_bounds__(1024) k_adler_final(const AdlerPartial *part, uint32_t nchunks, uint64_t len, uint32_t chunk, uint32_t start,
                                                       uint32_t *out)
{
    __shared__ unsigned long long sA[1024], sB[1024], sN[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nchunks + 1023) / 1024;
    unsigned long long A = 0, B = 0, n = 0;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t c = tid * per + k;
        if (c >= nchunks) break;
        const uint64_t c0 = (uint64_t)c * chunk;
        const unsigned long long cn = min((uint64_t)chunk, len - c0);
        B = (B + (cn % kAdlerBase) * A + part[c].b) % kAdlerBase;
        A = (A + part[c].a) % kAdlerBase;
        n += cn;
    }
    sA[tid] = A; sB[tid] = B; sN[tid] = n;
    __syncthreads();
    for (uint32_t s = 1; s < 1024; s <<= 1) {
        if ((tid & (2 * s - 1)) == 0) {
            const unsigned long long A2 = sA[tid + s], B2 = sB[tid + s], n2 = sN[tid + s];
            sB[tid] = (sB[tid] + (n2 % kAdlerBase) * sA[tid] + B2) % kAdlerBase;
            sA[tid] = (sA[tid] + A2) % kAdlerBase;
            sN[tid] += n2;
        }
        __syncthreads();
    }
    if (tid == 0) {
        // prepend the running value: s1 = start.lo + A, s2 = start.hi + n*start.lo + B
        const unsigned long long s1 = start & 0xffffu, s2 = (start >> 16) & 0xffffu;
        const unsigned long long r2 = (s2 + (len % kAdlerBase) * s1 + sB[0]) % kAdlerBase;
        const unsigned long long r1 = (s1 + sA[0]) % kAdlerBase;
        *out = (uint32_t)(r1 | (r2 << 16));
    }
}

// ------------------------------------------------------------------------------------------------
// crc32
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b)
{
    // a(x)*b(x) mod p(x), reflected (crc32/combine.rs:27-47); a != 0
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}

__device__ __forceinline__ uint32_t x2nmodp(uint64_t n, uint32_t k)
{
    // x^(n * 2^k) mod p(x)
    uint32_t p = 1u << 31;
    uint32_t sq = 1u << 30; // x^1
    for (uint32_t i = 0; i < k; i++) sq = multmodp(sq, sq);
    while (n) {
        if (n & 1) p = multmodp(sq, p);
        n >>= 1;
        if (n) sq = multmodp(sq, sq);
    }
    return p;
}

// k_crc_partial: byte-sliced table CRC is bound by shared-memory lookups (one per input byte).  Two things decide its speed:
//  * bank conflicts -- a 256-entry table hit by 32 lanes at random indices serialises ~3.5-fold.  Here every lane owns a private copy
//    of the four slice tables in its own bank (entry (k, idx) of lane t lives at word (k*256 + idx)*32 + t: 128 KiB per CTA), so
//    every lookup of a warp is one conflict-free wavefront;
//  * issue slots -- per 32-bit word: 4 lookups, 4 address computations (shift + and + add), two 3-input xors.
// One persistent CTA of 1024 threads per SM builds the tables once and loops over chunks; thread t computes the raw CRC of a
// contiguous segment with 128-bit loads.  Chunks and segments are aligned to E = the last 16-byte boundary of the buffer (the
// first ones are the short ones), so segment k of a chunk always has exactly (T-1-k) full segments behind it, every segment
// start is 16-byte aligned, and every shift in the combine tree is by a full piece length.  The < 16 bytes behind E are
// appended by k_crc_final.
constexpr uint32_t kCrcThreads = 1024;
constexpr uint32_t kCrcTabBytes = 4 * 256 * 32 * 4 + 32768; // + slack to align the tables to 32 KiB in the shared window

template <uint32_t KOFF, uint32_t SHIFT>
__device__ __forceinline__ uint32_t crc_lds(uint32_t tb, uint32_t x)
{
    // table word of byte ((x >> SHIFT*8) & 0xff) of slice KOFF/32768.  tb = 32 KiB-aligned table base | lane*4, so the entry offset
    // (idx * 128, bits 7..14) is OR-ed in: one shift + one LOP3 per lookup; the slice offset rides in the LDS immediate
    uint32_t v;
    const uint32_t sh = SHIFT == 0 ? x << 7 : SHIFT == 1 ? x >> 1 : SHIFT == 2 ? x >> 9 : x >> 17;
    const uint32_t a = (sh & 0x7f80u) | tb;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(KOFF));
    return v;
}
// one 32-bit step of the sliced CRC: x = data word ^ register; returns T3[b0] ^ T2[b1] ^ T1[b2] ^ T0[b3]
__device__ __forceinline__ uint32_t crc_step(uint32_t tb, uint32_t x)
{
    return crc_lds<3u * 32768u, 0>(tb, x) ^ crc_lds<2u * 32768u, 1>(tb, x) ^ crc_lds<32768u, 2>(tb, x) ^ crc_lds<0u, 3>(tb, x);
}

__global__ void __launch_bounds__(kCrcThreads) k_crc_partial(const uint8_t *__restrict__ buf, uint64_t main_len, uint32_t seg, uint32_t nchunks,
                                                              uint32_t *__restrict__ part)
{
    extern __shared__ __align__(16) uint8_t crc_smem[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(crc_smem);
    const uint32_t tab_base = (smem_base + 32767u) & ~32767u;
    uint32_t *tab = reinterpret_cast<uint32_t *>(crc_smem + (tab_base - smem_base)); // [4][256][32], 32 KiB aligned
    __shared__ uint32_t t0[256];
    __shared__ uint32_t segcrc[kCrcThreads];
    __shared__ uint32_t shiftop[10];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    if (tid < 256) {
        uint32_t c = tid;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        t0[tid] = c;
    }
    if (tid >= 256 && tid < 266) shiftop[tid - 256] = x2nmodp((uint64_t)seg << (tid - 256), 3); // x^(8*seg*2^level)
    __syncthreads();
    for (uint32_t i = tid; i < 256 * 32; i += kCrcThreads) {
        const uint32_t idx = i >> 5;
        uint32_t c = t0[idx];
        tab[i] = c; // T0
        for (uint32_t t = 1; t < 4; t++) { c = t0[c & 0xff] ^ (c >> 8); tab[t * 8192 + i] = c; }
    }
    __syncthreads();
    const uint32_t tb = tab_base | (lane * 4u);
    const uint64_t chunk = (uint64_t)seg * kCrcThreads;
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        // chunk j covers [E - (nchunks-j)*C, E - (nchunks-j-1)*C) clipped at 0
        const uint64_t cend = main_len - (uint64_t)(nchunks - 1 - ch) * chunk;
        const uint64_t cbeg = cend > chunk ? cend - chunk : 0;
        // segment tid covers [cend - (T-tid)*S, cend - (T-tid-1)*S) clipped at cbeg
        const uint64_t back = (uint64_t)(kCrcThreads - 1 - tid) * seg;
        uint32_t crc = 0;
        if (cend >= back && cend - back > cbeg) {
            const uint64_t send = cend - back;
            const uint64_t sbeg = (send - cbeg > seg) ? send - seg : cbeg;
            const uint8_t *p = buf + sbeg;
            const uint32_t slen = (uint32_t)(send - sbeg);
            uint32_t i = 0;
            // only the first segment of the buffer can start unaligned
            while (i < slen && (((uintptr_t)(p + i)) & 15)) { crc = t0[(crc ^ p[i]) & 0xff] ^ (crc >> 8); i++; }
#pragma unroll 2
            for (; i + 16 <= slen; i += 16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + i));
                crc = crc_step(tb, v.x ^ crc);
                crc = crc_step(tb, v.y ^ crc);
                crc = crc_step(tb, v.z ^ crc);
                crc = crc_step(tb, v.w ^ crc);
            }
            for (; i < slen; i++) crc = t0[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
        }
        segcrc[tid] = crc;
        __syncthreads();
        // tree: at level l, segment i absorbs segment i + 2^l, which is 2^l full segments long
        for (uint32_t l = 0; (1u << l) < kCrcThreads; l++) {
            const uint32_t s2 = 1u << l;
            if ((tid & (2 * s2 - 1)) == 0) {
                const uint32_t left = segcrc[tid];
                segcrc[tid] = (left ? multmodp(shiftop[l], left) : 0) ^ segcrc[tid + s2];
            }
            __syncthreads();
        }
        if (tid == 0) part[ch] = segcrc[0];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_crc_final(const uint32_t *part, uint32_t nchunks, uint64_t chunk_bytes, const uint8_t *tail, uint32_t tail_len,
                                                     uint64_t len, uint32_t start, uint32_t *out)
{
    __shared__ uint32_t sc[1024];
    __shared__ uint32_t ops[12];
    const uint32_t tid = threadIdx.x;
    // thread t reduces `per` consecutive chunk partials (aligned to the end: the leading threads may have fewer)
    const uint32_t per = (nchunks + 1023) / 1024;
    if (tid == 0) ops[0] = x2nmodp(chunk_bytes, 3);
    if (tid >= 1 && tid <= 10) ops[tid] = x2nmodp((chunk_bytes * per) << (tid - 1), 3);
    __syncthreads();
    uint32_t acc = 0;
    {
        // group g = tid covers chunks [nchunks - (1024-g)*per, nchunks - (1023-g)*per) clipped at 0
        const int64_t gend = (int64_t)nchunks - (int64_t)(1023 - tid) * per;
        const int64_t gbeg = gend - per;
        for (int64_t c = gbeg < 0 ? 0 : gbeg; c < gend; c++) acc = (acc ? multmodp(ops[0], acc) : 0) ^ part[c];
    }
    sc[tid] = acc;
    __syncthreads();
    for (uint32_t l = 0; l < 10; l++) {
        const uint32_t s = 1u << l;
        if ((tid & (2 * s - 1)) == 0) {
            const uint32_t left = sc[tid];
            sc[tid] = (left ? multmodp(ops[l + 1], left) : 0) ^ sc[tid + s];
        }
        __syncthreads();
    }
    if (tid == 0) {
        // register after the data starting from r0 = ~start:  raw(M) ^ r0 * x^(8 len); then the final xor
        uint32_t raw = sc[0];
        for (uint32_t i = 0; i < tail_len; i++) { // the bytes behind the last 16-byte boundary continue the raw register
            raw ^= tail[i];
            for (int k = 0; k < 8; k++) raw = (raw & 1) ? (raw >> 1) ^ kCrcPoly : raw >> 1;
        }
        const uint32_t r0 = ~start;
        const uint32_t sh = r0 ? multmodp(x2nmodp(len, 3), r0) : 0;
        *out = ~(raw ^ sh);
    }
}

// host-side launchers -------------------------------------------------------------------------
cudaError_t launch_adler32(const uint8_t *d_buf, uint64_t len, uint32_t start, void *d_scratch, size_t scratch_bytes, uint32_t *d_out,
                           cudaStream_t st)
{
    uint64_t chunk = 16384, nchunks = 0;
    if (len) {
        chunk = (len + 148 * 16 - 1) / (148 * 16);
        chunk = (chunk + 4095) & ~4095ull;
        if (chunk < 16384) chunk = 16384;
        if (chunk > (1u << 20)) chunk = 1u << 20;
        nchunks = (len + chunk - 1) / chunk;
        while (nchunks * sizeof(AdlerPartial) > scratch_bytes) { chunk *= 2; nchunks = (len + chunk - 1) / chunk; }
        k_adler_partial<<<(uint32_t)nchunks, 256, 0, st>>>(d_buf, len, (uint32_t)chunk, static_cast<AdlerPartial *>(d_scratch));
    }
    k_adler_final<<<1, 1024, 0, st>>>(static_cast<AdlerPartial *>(d_scratch), (uint32_t)nchunks, len, (uint32_t)chunk, start, d_out);
    return cudaGetLastError();
}

cudaError_t launch_crc32(const uint8_t *d_buf, uint64_t len, uint32_t start, void *d_scratch, size_t scratch_bytes, uint32_t *d_out,
                         cudaStream_t st)
{
    static int sm_count[64]; // per device; the opt-in for 128 KiB of dynamic shared memory is per device too
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (sm_count[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaFuncSetAttribute(k_crc_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrcTabBytes);
        if (e != cudaSuccess) return e;
        sm_count[dev] = v > 0 ? v : 148;
    }
    const int n_sm = sm_count[dev];
    // main part: up to the last 16-byte boundary; the rest (< 16 bytes) is appended by k_crc_final
    uint32_t tail_len = (uint32_t)((reinterpret_cast<uintptr_t>(d_buf) + len) & 15u);
    if (tail_len > len) tail_len = (uint32_t)len;
    const uint64_t main_len = len - tail_len;
    // bytes per thread: enough segments to fill every SM once, at most 4 KiB (the combine tree costs ~2500 cycles per chunk)
    uint64_t seg = (main_len + (uint64_t)n_sm * kCrcThreads - 1) / ((uint64_t)n_sm * kCrcThreads);
    seg = (seg + 15) & ~15ull;
    if (seg < 256) seg = 256;
    if (seg > 4096) seg = 4096;
    const uint64_t chunk = seg * kCrcThreads;
    const uint64_t nchunks = (main_len + chunk - 1) / chunk;
    if (nchunks * 4 > scratch_bytes) return cudaErrorInvalidValue;
    uint32_t *part = static_cast<uint32_t *>(d_scratch);
    if (nchunks) {
        const uint32_t grid = (uint32_t)(nchunks < (uint64_t)n_sm ? nchunks : (uint64_t)n_sm);
        k_crc_partial<<<grid, kCrcThreads, kCrcTabBytes, st>>>(d_buf, main_len, (uint32_t)seg, (uint32_t)nchunks, part);
    }
    k_crc_final<<<1, 1024, 0, st>>>(part, (uint32_t)nchunks, chunk, d_buf + main_len, tail_len, len, start, d_out);
    return cudaGetLastError();
}

} // namespace zb
