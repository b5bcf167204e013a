// zb_engine_internal.h -- Engine: one CUDA device + stream + grow-only buffers (host side).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/zb_engine.h"
#include "zb_kernels.cuh"

namespace zb {

extern thread_local char g_err[256];
size_t deflate_bound(size_t n);

struct Engine {
    static constexpr int kSlots = 40;
    struct Buf { void *p = nullptr; size_t cap = 0; };
    int device = -1;
    cudaStream_t st = nullptr, st2 = nullptr; // st2: the serial tail runs beside k_emit
    cudaEvent_t evf = nullptr, evt = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    static constexpr int kUpChunks = 8;
    cudaEvent_t evk = nullptr;              // ... and the checksum behind the last chunk, also on st2
    cudaEvent_t evc[kUpChunks] = {nullptr}; // chunks of a host input arriving on st2 while the link pass already runs on st
    Buf bufs[kSlots];
    void *h_stage = nullptr;
    size_t h_stage_cap = 0;
    JobInfo *h_info = nullptr, *d_info = nullptr;
    uint32_t *d_check = nullptr;
    void *d_inf_state = nullptr, *h_inf_state = nullptr; // inflate result block
    uint32_t launches = 0;
    uint8_t h_prime = 0; // staging byte of ZB_FLAG_PRIME (must outlive the async copy)
    // optional per-phase device timing (zb_engine_set_profile): 0 links, 1 match, 2 nxt, 3 path, 4 emit+holes,
    // 5 tail, 6 blocks(hist+trees+scan), 7 encode, 8 checksum, 9 h2d, 10 d2h
    static constexpr int kPhases = 12;
    bool profile = false;
    float phase_ms[kPhases] = {0};
    uint32_t phase_launches[kPhases] = {0};
    cudaEvent_t pev0 = nullptr, pev1 = nullptr;
    void pbegin();
    void pend(int phase, uint32_t nlaunch);

    int init(int dev);
    int inflate_init();
    ~Engine();
    int reserve(int slot, size_t bytes, void **out);
    int stage(size_t bytes);
    int deflate(const void *src, size_t n, bool src_dev, void *dst, size_t dst_cap, bool dst_dev, int level, int strategy,
                int window_bits, uint32_t flags, zb_deflate_result *res, const void *dict = nullptr, size_t dict_len = 0);
    int inflate(const void *src, size_t n, bool src_dev, void *dst, size_t dst_cap, bool dst_dev, int window_bits,
                zb_inflate_result *res, uint32_t flags = 0);
    int inflate_blocks(const void *src, size_t n, uint64_t start_bit, const void *dict, size_t dict_len, void *dst, size_t dst_cap,
                       int check_kind, uint32_t check_start, zb_inflate_seg *out);
    int checksum(bool crc, uint32_t start, const void *buf, size_t len, bool on_dev, uint32_t *out, float *ms);
};

} // namespace zb
