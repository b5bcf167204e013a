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

constexpr uint32_t kCrcSeg = 2048; // bytes per thread
constexpr uint32_t kCrcThreads = 128;
constexpr uint32_t kCrcChunk = kCrcSeg * kCrcThreads; // 256 KiB per CTA

// Raw CRC of each chunk.  Chunks and segments are aligned to the END of the buffer (the first ones are the short
// ones), so segment k of a chunk always has exactly (kCrcThreads-1-k) full segments behind it.
__global__ void __launch_bounds__(kCrcThreads) k_crc_partial(const uint8_t *__restrict__ buf, uint64_t len, uint32_t nchunks,
                                                              uint32_t *__restrict__ part)
{
    __shared__ uint32_t tab[8][256];
    __shared__ uint32_t segcrc[kCrcThreads];
    __shared__ uint32_t shiftop[8];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 256; i += kCrcThreads) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        tab[0][i] = c;
    }
    if (tid < 7) shiftop[tid] = x2nmodp((uint64_t)kCrcSeg << tid, 3); // x^(8*seg*2^level)
    __syncthreads();
    for (uint32_t i = tid; i < 256; i += kCrcThreads) {
        uint32_t c = tab[0][i];
        for (int t = 1; t < 8; t++) { c = tab[0][c & 0xff] ^ (c >> 8); tab[t][i] = c; }
    }
    __syncthreads();
    // chunk j covers [len - (nchunks-j)*C, len - (nchunks-j-1)*C) clipped at 0
    const uint64_t cend = len - (uint64_t)(nchunks - 1 - blockIdx.x) * kCrcChunk;
    const uint64_t cbeg = cend > kCrcChunk ? cend - kCrcChunk : 0;
    // segment tid covers [cend - (T-tid)*S, cend - (T-tid-1)*S) clipped at cbeg
    const uint64_t send = cend - (uint64_t)(kCrcThreads - 1 - tid) * kCrcSeg;
    uint32_t crc = 0;
    if (send > cbeg && cend >= (uint64_t)(kCrcThreads - 1 - tid) * kCrcSeg) {
        const uint64_t sbeg = (send - cbeg > kCrcSeg) ? send - kCrcSeg : cbeg;
        const uint8_t *p = buf + sbeg;
        const uint32_t slen = (uint32_t)(send - sbeg);
        uint32_t i = 0;
        while (i < slen && (((uintptr_t)(p + i)) & 15)) { crc = tab[0][(crc ^ p[i]) & 0xff] ^ (crc >> 8); i++; }
        for (; i + 16 <= slen; i += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + i));
            uint32_t lo = v.x ^ crc, hi = v.y;
            crc = tab[7][lo & 0xff] ^ tab[6][(lo >> 8) & 0xff] ^ tab[5][(lo >> 16) & 0xff] ^ tab[4][lo >> 24] ^
                  tab[3][hi & 0xff] ^ tab[2][(hi >> 8) & 0xff] ^ tab[1][(hi >> 16) & 0xff] ^ tab[0][hi >> 24];
            lo = v.z ^ crc; hi = v.w;
            crc = tab[7][lo & 0xff] ^ tab[6][(lo >> 8) & 0xff] ^ tab[5][(lo >> 16) & 0xff] ^ tab[4][lo >> 24] ^
                  tab[3][hi & 0xff] ^ tab[2][(hi >> 8) & 0xff] ^ tab[1][(hi >> 16) & 0xff] ^ tab[0][hi >> 24];
        }
        for (; i < slen; i++) crc = tab[0][(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    }
    segcrc[tid] = crc;
    __syncthreads();
    // tree: at level l, segment i absorbs segment i + 2^l, which is 2^l full segments long
    for (uint32_t l = 0; (1u << l) < kCrcThreads; l++) {
        const uint32_t s = 1u << l;
        if ((tid & (2 * s - 1)) == 0) {
            const uint32_t left = segcrc[tid];
            segcrc[tid] = (left ? multmodp(shiftop[l], left) : 0) ^ segcrc[tid + s];
        }
        __syncthreads();
    }
    if (tid == 0) part[blockIdx.x] = segcrc[0];
}

__global__ void __launch_bounds__(1024) k_crc_final(const uint32_t *part, uint32_t nchunks, uint64_t len, uint32_t start, uint32_t *out)
{
    __shared__ uint32_t sc[1024];
    __shared__ uint32_t ops[12];
    const uint32_t tid = threadIdx.x;
    // thread t reduces `per` consecutive chunk partials (aligned to the end: the leading threads may have fewer)
    const uint32_t per = (nchunks + 1023) / 1024;
    if (tid == 0) ops[0] = x2nmodp(kCrcChunk, 3);
    if (tid >= 1 && tid <= 10) ops[tid] = x2nmodp(((uint64_t)kCrcChunk * per) << (tid - 1), 3);
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
        const uint32_t r0 = ~start;
        const uint32_t sh = r0 ? multmodp(x2nmodp(len, 3), r0) : 0;
        *out = ~(sc[0] ^ sh);
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
    const uint64_t nchunks = (len + kCrcChunk - 1) / kCrcChunk;
    if (nchunks * 4 > scratch_bytes) return cudaErrorInvalidValue;
    uint32_t *part = static_cast<uint32_t *>(d_scratch);
    if (nchunks) k_crc_partial<<<(uint32_t)nchunks, kCrcThreads, 0, st>>>(d_buf, len, (uint32_t)nchunks, part);
    k_crc_final<<<1, 1024, 0, st>>>(part, (uint32_t)nchunks, len, start, d_out);
    return cudaGetLastError();
}

} // namespace zb
