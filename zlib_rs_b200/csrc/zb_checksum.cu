// zb_checksum.cu -- adler32 / crc32 kernels (sm_100a).
//
// Replaces zlib-rs/src/adler32.rs:19-47 (+ SIMD variants) and zlib-rs/src/crc32.rs:19-29 (braid /
// pclmulqdq folding).  Both checksums are linear in a way that makes them chunk-parallel:
//   adler32 : a chunk contributes A = sum(d_i), B = sum((n-i) d_i); chunks chain with
//             s2 += n*s1 + B, s1 += A (mod 65521)   [same algebra as adler32_combine, adler32.rs:58-87]
//   crc32   : raw CRC (zero init, no final xor) of a concatenation is raw(A)*x^(8|B|) mod P ^ raw(B)
//             [crc32_combine_op, crc32/combine.rs:3-61]
// Memory-bound: 128-bit loads, one pass over the data.
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
        uint32_t s = 0, wt = 0; // wt = sum j * d_j, j = 0..15
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b0 = w[k] & 0xff, b1 = (w[k] >> 8) & 0xff, b2 = (w[k] >> 16) & 0xff, b3 = w[k] >> 24;
            s += b0 + b1 + b2 + b3;
            wt += (4 * k) * b0 + (4 * k + 1) * b1 + (4 * k + 2) * b2 + (4 * k + 3) * b3;
        }
        a += s;
        b += (unsigned long long)(n - g * 16) * s - wt;
    }
    for (uint32_t i = n16 * 16 + tid; i < n; i += 256) {
        const uint32_t d = p[i];
        a += d;
        b += (unsigned long long)(n - i) * d;
    }
    // block reduce
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

__global__ void k_adler_final(const AdlerPartial *part, uint32_t nchunks, uint64_t len, uint32_t chunk, uint32_t start, uint32_t *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long s1 = start & 0xffffu, s2 = (start >> 16) & 0xffffu;
    for (uint32_t c = 0; c < nchunks; c++) {
        const uint64_t c0 = (uint64_t)c * chunk;
        const uint32_t n = (uint32_t)min((uint64_t)chunk, len - c0);
        s2 = (s2 + (unsigned long long)(n % kAdlerBase) * s1 + part[c].b) % kAdlerBase;
        s1 = (s1 + part[c].a) % kAdlerBase;
    }
    *out = (uint32_t)(s1 | (s2 << 16));
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

constexpr uint32_t kCrcSeg = 1024; // bytes per thread
constexpr uint32_t kCrcThreads = 128;
constexpr uint32_t kCrcChunk = kCrcSeg * kCrcThreads;

// Raw CRC (zero init, no final xor) of each 128 KiB chunk; thread = contiguous 1 KiB, slice-by-8.
__global__ void __launch_bounds__(kCrcThreads) k_crc_partial(const uint8_t *__restrict__ buf, uint64_t len, uint32_t *__restrict__ part)
{
    __shared__ uint32_t tab[8][256];
    __shared__ uint32_t segcrc[kCrcThreads];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 256; i += kCrcThreads) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ kCrcPoly : c >> 1;
        tab[0][i] = c;
    }
    __syncthreads();
    for (uint32_t i = tid; i < 256; i += kCrcThreads) {
        uint32_t c = tab[0][i];
        for (int t = 1; t < 8; t++) { c = tab[0][c & 0xff] ^ (c >> 8); tab[t][i] = c; }
    }
    __syncthreads();
    const uint64_t c0 = (uint64_t)blockIdx.x * kCrcChunk;
    const uint64_t clen = min((uint64_t)kCrcChunk, len - c0);
    const uint64_t s0 = (uint64_t)tid * kCrcSeg;
    uint32_t crc = 0;
    uint32_t slen = 0;
    if (s0 < clen) {
        slen = (uint32_t)min((uint64_t)kCrcSeg, clen - s0);
        const uint8_t *p = buf + c0 + s0;
        uint32_t i = 0;
        if ((((uintptr_t)p) & 15) == 0) {
            for (; i + 16 <= slen; i += 16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + i));
                uint32_t lo = v.x ^ crc, hi = v.y;
                crc = tab[7][lo & 0xff] ^ tab[6][(lo >> 8) & 0xff] ^ tab[5][(lo >> 16) & 0xff] ^ tab[4][lo >> 24] ^
                      tab[3][hi & 0xff] ^ tab[2][(hi >> 8) & 0xff] ^ tab[1][(hi >> 16) & 0xff] ^ tab[0][hi >> 24];
                lo = v.z ^ crc; hi = v.w;
                crc = tab[7][lo & 0xff] ^ tab[6][(lo >> 8) & 0xff] ^ tab[5][(lo >> 16) & 0xff] ^ tab[4][lo >> 24] ^
                      tab[3][hi & 0xff] ^ tab[2][(hi >> 8) & 0xff] ^ tab[1][(hi >> 16) & 0xff] ^ tab[0][hi >> 24];
            }
        }
        for (; i < slen; i++) crc = tab[0][(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    }
    segcrc[tid] = crc;
    __syncthreads();
    // combine the segments left to right (all full segments have length kCrcSeg except possibly the last)
    if (tid == 0) {
        const uint32_t nseg = (uint32_t)((clen + kCrcSeg - 1) / kCrcSeg);
        const uint32_t shift_full = x2nmodp(kCrcSeg, 3);
        uint32_t acc = 0;
        for (uint32_t s = 0; s < nseg; s++) {
            const uint32_t sl = (uint32_t)min((uint64_t)kCrcSeg, clen - (uint64_t)s * kCrcSeg);
            const uint32_t sh = sl == kCrcSeg ? shift_full : x2nmodp(sl, 3);
            acc = (acc ? multmodp(sh, acc) : 0) ^ segcrc[s];
        }
        part[blockIdx.x] = acc;
    }
}

__global__ void k_crc_final(const uint32_t *part, uint32_t nchunks, uint64_t len, uint32_t start, uint32_t *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // crc(M) with initial register r0 = ~start:  raw(M) ^ r0 * x^(8|M|), then final xor
    const uint32_t shift_full = x2nmodp(kCrcChunk, 3);
    uint32_t acc = ~start; // the initial register behaves like a prefix whose raw CRC is r0
    for (uint32_t c = 0; c < nchunks; c++) {
        const uint64_t cl = min((uint64_t)kCrcChunk, len - (uint64_t)c * kCrcChunk);
        const uint32_t sh = cl == kCrcChunk ? shift_full : x2nmodp(cl, 3);
        acc = (acc ? multmodp(sh, acc) : 0) ^ part[c];
    }
    *out = ~acc;
}

// host-side launchers -------------------------------------------------------------------------
cudaError_t launch_adler32(const uint8_t *d_buf, uint64_t len, uint32_t start, void *d_scratch, size_t scratch_bytes, uint32_t *d_out,
                           cudaStream_t st)
{
    if (len == 0) {
        k_adler_final<<<1, 32, 0, st>>>(nullptr, 0, 0, 1, start, d_out);
        return cudaGetLastError();
    }
    uint64_t chunk = (len + 148 * 8 - 1) / (148 * 8);
    chunk = (chunk + 4095) & ~4095ull;
    if (chunk < 16384) chunk = 16384;
    if (chunk > (1u << 20)) chunk = 1u << 20;
    uint64_t nchunks = (len + chunk - 1) / chunk;
    while (nchunks * sizeof(AdlerPartial) > scratch_bytes) { chunk *= 2; nchunks = (len + chunk - 1) / chunk; }
    AdlerPartial *part = static_cast<AdlerPartial *>(d_scratch);
    k_adler_partial<<<(uint32_t)nchunks, 256, 0, st>>>(d_buf, len, (uint32_t)chunk, part);
    k_adler_final<<<1, 32, 0, st>>>(part, (uint32_t)nchunks, len, (uint32_t)chunk, start, d_out);
    return cudaGetLastError();
}

cudaError_t launch_crc32(const uint8_t *d_buf, uint64_t len, uint32_t start, void *d_scratch, size_t scratch_bytes, uint32_t *d_out,
                         cudaStream_t st)
{
    const uint64_t nchunks = (len + kCrcChunk - 1) / kCrcChunk;
    if (nchunks * 4 > scratch_bytes) return cudaErrorInvalidValue;
    uint32_t *part = static_cast<uint32_t *>(d_scratch);
    if (nchunks) k_crc_partial<<<(uint32_t)nchunks, kCrcThreads, 0, st>>>(d_buf, len, part);
    k_crc_final<<<1, 32, 0, st>>>(part, (uint32_t)nchunks, len, start, d_out);
    return cudaGetLastError();
}

} // namespace zb
