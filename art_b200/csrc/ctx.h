// Internal context of libart_hotpath.so -- not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/art_hotpath.h"

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct art_hp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // the stream work is queued on (own or caller's)
    cudaStream_t copy_stream = nullptr;  // second stream for copy/compute overlap
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long launches = 0;
    std::string err;
    // device scratch, grown on demand and kept across calls
    DevBuf d_raw, d_out[3], d_scratch;
    // pinned staging (two halves for double buffering)
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_bytes = 0;

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define ART_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return (ctx)->fail(ART_HP_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                               cudaGetErrorString(e_), __FILE__, __LINE__);                   \
    } while (0)

static inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// grow-only device buffer
int art_reserve(art_hp_ctx* ctx, DevBuf& b, size_t bytes);

// kernels (device-resident planes, pitch in floats); each returns an art_hp_status
int art_rcd_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                float* R, float* G, float* B, size_t op);
int art_border_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, int bord, const float* raw, size_t rp,
                   float* R, float* G, float* B, size_t op);
int art_amaze_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                  float* R, float* G, float* B, size_t op, double initialGain, int border);
