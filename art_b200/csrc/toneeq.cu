// ImProcFunctions::toneEqualizer (reference rtengine/iptoneequalizer.cc L343-371) and tone_eq() (L68-338), STAGE_1 of
// ImProcFunctions::process (improcfun.cc L584), without the colour-map preview branch (lcms2, PREVIEW pipeline only):
//   frame * gain (gain = 1 / 65535 * 2^-pivot), Y = LIM(rgbLuminance, 1e-5, 32)
//   regularization > 0: guidedFilterLog(10, Y, 5 / scale + 0.5, 0.014) -- the log image guides itself (guidedfilter.cc L243-269)
//   regularization > 1: Y2 = Y, Y posterised to 1/5 EV, guidedFilter(Y2, Y, Y, 350 / scale, 0.004) and, for reg = 5 - min(regularization, 4) > 1,
//                       guidedFilter(Y2, Y, Y, radius (reg - 1), 0.00004)
//   RGB *= correction(Y): the sum of twelve 2-EV gaussian bands, from a 65536-entry table for Y <= 1, evaluated directly above it (a whole
//   4-pixel SSE2 group through sleef's VECTOR xlogf / xexpf when one of its pixels is above 1; per pixel in the scalar row tail); frame / gain
// The band factors, their normalisation and the table are built on the device with the same sleef steps the reference's host code takes.
// Bit-identical to the reference.
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>

namespace {

__device__ __forceinline__ float maxr(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float minr(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float lim_f(float v, float lo, float hi) { return maxr(lo, minr(v, hi)); }
__device__ __forceinline__ float vmaxf_(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float vminf_(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }
__device__ __forceinline__ float lut_s(const float* __restrict__ data, int size, float index)
{   // LUT.h L437-459, clip below and above
    const int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) return data[size - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
__device__ __forceinline__ float lut_v(const float* __restrict__ data, int size, float index)
{   // LUT.h L349-377
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}
__device__ __forceinline__ float log2_(float x) { return sleef::xlogf_scalar(x) / sleef::xlogf_scalar(2.f); }
__device__ __forceinline__ float exp2_(float x) { return sleef::pow_F_scalar(2.f, x); }
__device__ __forceinline__ float gauss_(float b, float x) { return sleef::xexpf_scalar(-((x - b) * (x - b)) / 4.0f); }

__constant__ float c_centers[12] = {-16.0f, -14.0f, -12.0f, -10.0f, -8.0f, -6.0f, -4.0f, -2.0f, 0.0f, 2.0f, 4.0f, 6.0f};

struct Bands { int v[5]; };

// consts[0..11] = factors (L93-110), consts[12] = w_sum (L156-159)
__global__ void k_teq_consts(Bands b, float* __restrict__ consts)
{
    const int band_of[12] = {0, 0, 0, 0, 0, 1, 2, 3, 4, 4, 4, 4};
    const float lo_of[12] = {2.f, 2.f, 2.f, 2.f, 2.f, 2.f, 2.5f, 3.f, 3.f, 3.f, 3.f, 3.f};
    const float hi_of[12] = {3.f, 3.f, 3.f, 3.f, 3.f, 3.f, 2.5f, 2.f, 2.f, 2.f, 2.f, 2.f};
    float w_sum = 0.f;
    for (int i = 0; i < 12; ++i) {
        const int v = b.v[band_of[i]];
        const float f = v < 0 ? lo_of[i] : hi_of[i];
        consts[i] = exp2_((float)v / 100.f * f);
        w_sum += gauss_(c_centers[i], 0.f);
    }
    consts[12] = w_sum;
}

__device__ __forceinline__ float process_pixel(float y, const float* __restrict__ k)
{   // L164-178
    const float luma = lim_f(log2_(maxr(y, 0.f)), -14.f, 4.f);
    float correction = 0.0f;
#pragma unroll
    for (int c = 0; c < 12; ++c) correction += gauss_(c_centers[c], luma) * k[c];
    return correction / k[12];
}
__device__ __forceinline__ float vprocess_pixel(float y, const float* __restrict__ k)
{   // L243-254
    const float luma = vminf_(vmaxf_(sleef::xlogf_vector(vmaxf_(y, 0.f)) / sleef::xlogf_scalar(2.f), -14.f), 4.f);
    float correction = 0.f;
#pragma unroll
    for (int c = 0; c < 12; ++c) { const float d = luma - c_centers[c]; correction += sleef::xexpf_vector(-(d * d) / 4.f) * k[c]; }
    return correction / k[12];
}

__global__ void __launch_bounds__(256) k_teq_lut(float* __restrict__ lut, const float* __restrict__ consts)
{
    __shared__ float k[13];
    if (threadIdx.x < 13) k[threadIdx.x] = consts[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536) lut[i] = process_pixel((float)i / 65535.f, k);
}

// frame * gain; Y (L114-120); tolog: the first loop of guidedFilterLog
__global__ void __launch_bounds__(256) k_teq_prep(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H,
                                                  float* __restrict__ Y, size_t yp, float gain, float w0, float w1, float w2, int tolog)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x;
        const float R = r[i] * gain, G = g[i] * gain, B = b[i] * gain;
        r[i] = R; g[i] = G; b[i] = B;
        float l = lim_f(R * w0 + G * w1 + B * w2, 1e-5f, 32.f);
        if (tolog) l = sleef::xlin2log_scalar(maxr(l, 0.f), 10.f);
        Y[(size_t)y * yp + x] = l;
    }
}

// unlog: the last loop of guidedFilterLog; poster: L131-140 (Y2 = Y, Y = 2^(round(5 l) / 5))
__global__ void __launch_bounds__(256) k_teq_mid(float* __restrict__ Y, float* __restrict__ Y2, size_t yp, int W, int H, int unlog, int poster)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * yp + x;
        float v = Y[i];
        if (unlog) v = sleef::xlog2lin_scalar(maxr(v, 0.f), 10.f);
        if (poster) {
            Y2[i] = v;
            const float l = lim_f(log2_(maxr(v, 1e-9f)), -16.0f, 6.0f);
            v = exp2_(roundf(l * 5.f) / 5.f);
        }
        Y[i] = v;
    }
}

// L311-337, then Imagefloat::multiply(1.f / gain)
__global__ void __launch_bounds__(128) k_teq_apply(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H,
                                                   const float* __restrict__ Y, size_t yp, const float* __restrict__ lut, const float* __restrict__ consts, float back, int aligned)
{
    __shared__ float k[13];
    if (threadIdx.x < 13) k[threadIdx.x] = consts[threadIdx.x];
    __syncthreads();
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const float* cy = Y + (size_t)y * yp;
        const size_t row = (size_t)y * ip;
        if (x0 + 4 <= W) {
            float c[4], corr[4];
            if (aligned) { const float4 v = *reinterpret_cast<const float4*>(cy + x0); c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w; }
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = cy[x0 + j];
            }
            const bool any = c[0] > 1.f || c[1] > 1.f || c[2] > 1.f || c[3] > 1.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) corr[j] = any ? vprocess_pixel(c[j], k) : lut_v(lut, 65536, c[j] * 65535.f);
            if (aligned) {      // 16-byte aligned planes: one 128-bit access per plane and group
                float* pl[3] = {r, g, b};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    float4* p4 = reinterpret_cast<float4*>(pl[q] + row + x0);
                    float4 v = *p4;
                    v.x = v.x * corr[0] * back; v.y = v.y * corr[1] * back; v.z = v.z * corr[2] * back; v.w = v.w * corr[3] * back;
                    *p4 = v;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const size_t i = row + x0 + j;
                    r[i] = r[i] * corr[j] * back; g[i] = g[i] * corr[j] * back; b[i] = b[i] * corr[j] * back;
                }
            }
        } else {
            for (int x = x0; x < W; ++x) {
                const float cY = cy[x];
                const float corr = cY > 1.f ? process_pixel(cY, k) : lut_s(lut, 65536, cY * 65535.f);
                const size_t i = row + x;
                r[i] = r[i] * corr * back; g[i] = g[i] * corr * back; b[i] = b[i] * corr * back;
            }
        }
    }
}

}  // namespace

int art_tone_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_toneeq_params* p)
{
    if (!p->ws) return ctx->fail(ART_HP_ERR_INVALID, "the tone equalizer needs the working-space matrix");
    if (!(p->scale > 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "scale must be positive");
    if (p->regularization < 0) return ctx->fail(ART_HP_ERR_INVALID, "regularization %d", p->regularization);
    cudaStream_t st = ctx->stream;
    // iptoneequalizer.cc L351: the product is formed in double (pivot is a double) and rounded to float once
    const float gain = (float)(1.f / 65535.f * std::pow(2.f, -p->pivot));
    const float back = 1.f / gain;
    const int detail = p->regularization > 0 ? 5 : 0;
    int radius = (int)((float)detail / p->scale + 0.5f);
    float epsilon = 0.01f + 0.002f * (float)std::max(detail - 3, 0);
    int radius2 = 0, reg = 0;
    if (p->regularization > 1) {
        radius2 = (int)(350.f / p->scale);
        reg = 5 - std::min(p->regularization, 4);
    }
    const size_t yp = round_up((size_t)W, 32), n = yp * (size_t)H;
    void* blk = nullptr;
    int rc = art_pool_alloc(ctx, (2 * n + 65536 + 64) * sizeof(float), &blk);
    if (rc) return rc;
    float* Y = (float*)blk;
    float* Y2 = Y + n;
    float* lut = Y2 + n;
    float* consts = lut + 65536;
    Bands bands;
    for (int i = 0; i < 5; ++i) bands.v[i] = p->bands[i];
    const dim3 grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_teq_prep");
    k_teq_consts<<<1, 1, 0, st>>>(bands, consts);
    k_teq_lut<<<256, 256, 0, st>>>(lut, consts);
    k_teq_prep<<<grid, 256, 0, st>>>(r, g, b, ip, W, H, Y, yp, gain, (float)p->ws[3], (float)p->ws[4], (float)p->ws[5], radius > 0);
    art_prof_end(ctx);
    ctx->launches += 3;
    if (radius > 0) rc = art_guided_dev(ctx, Y, yp, Y, yp, Y, yp, W, H, radius, epsilon, 0);
    if (!rc && (radius > 0 || p->regularization > 1)) {
        art_prof_begin(ctx, "k_teq_mid");
        k_teq_mid<<<grid, 256, 0, st>>>(Y, Y2, yp, W, H, radius > 0, p->regularization > 1);
        art_prof_end(ctx);
        ctx->launches++;
    }
    if (!rc && p->regularization > 1) {
        epsilon = 0.004f;
        rc = art_guided_dev(ctx, Y2, yp, Y, yp, Y, yp, W, H, radius2, epsilon, 0);
        if (!rc && reg > 1) rc = art_guided_dev(ctx, Y2, yp, Y, yp, Y, yp, W, H, radius2 * (reg - 1), epsilon / 100, 0);
    }
    if (!rc) {
        const int aligned = ip % 4 == 0 && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
        art_prof_begin(ctx, "k_teq_apply");
        k_teq_apply<<<dim3(((W + 3) / 4 + 127) / 128, std::min(H, 148 * 8)), 128, 0, st>>>(r, g, b, ip, W, H, Y, yp, lut, consts, back, aligned);
        art_prof_end(ctx);
        ctx->launches++;
    }
    art_pool_free(ctx, blk);
    if (rc) return rc;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
