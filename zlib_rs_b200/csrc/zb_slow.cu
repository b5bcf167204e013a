// zb_slow.cu -- kernels of the level 7..9 path (deflate_slow, lazy matching; sm_100a).
//
//   k_links_roll  L[x]   : previous position with the same rolling 3-byte hash (level 9; levels 7/8 use k_links)
//   k_slow        nxt[p] : macro step of the lazy parser from every position taken as a fresh loop-top
//                 M[p]   : its symbols (literal count, match length, distance)
//   k_path_*             : shared with the level-6 path
//   k_emit_slow / k_tail_slow : symbols of the path nodes
// All chains are static (every position is inserted), so there is no fixed-point iteration here.
#include "zb_kernels.cuh"
#include "zb_slow.h"

namespace zb {

constexpr uint32_t kRollWarm = 32768;

__global__ void __launch_bounds__(1024) k_links_roll(JobBufs jb)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *head = reinterpret_cast<uint32_t *>(smem);
    uint8_t *sd = smem + 32768 * 4;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t N = jb.N;
    const uint32_t ts = blockIdx.x * kLinkTile;
    const uint32_t te = min(ts + kLinkTile, N);
    const uint32_t ws = ts > kRollWarm ? ts - kRollWarm : 0;
    for (uint32_t i = tid; i < 32768; i += 1024) head[i] = 0;
    {
        const uint32_t n16 = (te + 16 - ws + 15) / 16; // the input buffer is zero padded
        const uint4 *src = reinterpret_cast<const uint4 *>(jb.in + ws);
        uint4 *dst = reinterpret_cast<uint4 *>(sd);
        for (uint32_t i = tid; i < n16; i += 1024) dst[i] = src[i];
    }
    __syncthreads();
    for (uint32_t base = ws; base < te; base += 32) {
        const uint32_t x = base + lane;
        const bool valid = x < te && x + 3 <= N;
        uint32_t key = 0;
        if (valid) key = hash_roll3(sd[x - ws], sd[x - ws + 1], sd[x - ws + 2]);
        const bool mine = valid && (key & 31u) == warp;
        const uint32_t m = __ballot_sync(0xffffffffu, mine);
        if (m == 0) continue;
        uint32_t pred_rel = 0, peers_ins = 0;
        if (mine) {
            peers_ins = __match_any_sync(m, key);
            const uint32_t lower = peers_ins & ((1u << lane) - 1u);
            if (lower) pred_rel = (base + (31 - __clz(lower))) - ws + 1;
            else pred_rel = head[key];
        }
        __syncwarp();
        if (mine) {
            const uint32_t rel = x - ws + 1;
            if (x >= ts) {
                const uint32_t d = pred_rel ? rel - pred_rel : 0;
                jb.L[x] = (uint16_t)((d && d <= kLinkCapSlow) ? d : 0);
            }
            if ((peers_ins >> lane) == 1u) head[key] = rel;
        }
        __syncwarp();
    }
    for (uint32_t x = max(ts, N >= 2 ? N - 2 : 0) + tid; x < te; x += 1024) jb.L[x] = 0;
}

// shared-memory window of k_slow: data and links of [ws, ws + span)
struct SlowSAcc {
    const uint8_t *sdata;
    const uint16_t *sL;
    uint32_t ws, N, need;
    __device__ __forceinline__ uint32_t byte(uint32_t y) const
    {
        // bytes beyond the input are what the reference's window buffer still holds there
        while (y >= N) {
            if (y < 2 * kWSize) return 0;
            y -= kWSize;
        }
        return sdata[y - ws];
    }
    __device__ __forceinline__ uint32_t link(uint32_t y) const { return y + need <= N ? sL[y - ws] : 0; }
};

__device__ __forceinline__ uint32_t pack_step(const SlowStep &s)
{
    return (s.nlit << 24) | (s.len ? ((s.len - 3u) << 16) | 0x8000u | (s.dist - 1u) : 0u);
}

__global__ void __launch_bounds__(1024) k_slow(JobBufs jb)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_next;
    const uint32_t sub = kSlowSub;
    const uint32_t ts = blockIdx.x * sub;
    const uint32_t N = jb.N;
    if (ts >= N) return;
    const uint32_t te = min(ts + sub, N);
    const uint32_t ws = ts >= kWSize ? ts - kWSize : 0;
    const uint32_t span = te + kSlowAhead - ws; // <= kWSize + sub + kSlowAhead
    uint8_t *sdata = smem;
    uint16_t *sL = reinterpret_cast<uint16_t *>(smem + kWSize + kSlowSub + kSlowAhead);
    const uint32_t tid = threadIdx.x;
    if (tid == 0) s_next = ts;
    {
        // the input and L allocations are padded with kPad (>= kSlowAhead) zero entries
        const uint32_t n16 = (span + 15) / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(jb.in + ws);
        uint4 *dst = reinterpret_cast<uint4 *>(sdata);
        for (uint32_t i = tid; i < n16; i += 1024) dst[i] = src[i];
        const uint32_t nl = (span + 7) / 8;
        const uint4 *ls = reinterpret_cast<const uint4 *>(jb.L + ws);
        uint4 *ld = reinterpret_cast<uint4 *>(sL);
        for (uint32_t i = tid; i < nl; i += 1024) ld[i] = ls[i];
    }
    __syncthreads();
    const SlowSAcc a{sdata, sL, ws, N, jb.sp.slow ? 3u : 4u};
    const SlowParams sp = jb.sp;
    for (;;) {
        const uint32_t p = atomicAdd(&s_next, 1u);
        if (p >= te) break;
        const SlowStep s = slow_step(a, p, N, sp);
        const uint32_t delta = s.next - p, ns = s.nlit + (s.len ? 1u : 0u);
        if (delta > 0xffffu || ns > 0xffu || delta == 0) atomicOr(&jb.info->error, 1u);
        jb.nxt[p] = (delta & 0xffffu) | ((ns & 0xffu) << 16) | (s.next >= N ? kNxtTail : 0u);
        jb.M[p] = pack_step(s);
    }
}

__device__ __forceinline__ uint32_t emit_step(const JobBufs &jb, uint32_t p, Sym *out)
{
    const uint32_t v = jb.M[p];
    const uint32_t nlit = v >> 24;
    uint32_t k = 0;
    for (uint32_t i = 0; i < nlit; i++) out[k++] = Sym{0, jb.in[p + i], p + i};
    if (v & 0x8000u) out[k++] = Sym{(uint16_t)((v & 0x7fffu) + 1u), (uint16_t)((v >> 16) & 0xffu), p + nlit};
    return k;
}

__global__ void __launch_bounds__(256) k_emit_slow(JobBufs jb)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= jb.N) return;
    const uint32_t idx = jb.symidx[p];
    if (!idx) return;
    emit_step(jb, p, jb.syms + jb.tile_symbase[p / kPathTile] + idx - 1);
}

// the last node of the path (its step reaches the end of the input) and the job totals
__global__ void k_tail_slow(JobBufs jb)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t n = jb.info->n_mid_syms;
    if (jb.N > 0) n += emit_step(jb, jb.info->tail_entry, jb.syms + n);
    jb.info->n_syms = n;
    jb.info->final_base = base_at(jb.N, jb.N);
    jb.info->n_blocks = n / kBlockSyms + 1;
}

} // namespace zb
