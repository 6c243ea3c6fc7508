// denoise::denoiseGuidedSmoothing (reference rtengine/ipsmoothing.cc L875-897), the chroma smoothing ImProcFunctions::denoise runs
// after RGB_denoise when smoothingEnabled and guidedChromaRadius != 0 (rtengine/ipdenoise.cc L1171-1172):
//   Imagefloat::normalizeFloatTo1                       (x *= 1.f / 65535.f)
//   guided_smoothing(R, G, B, ws, iws, Channel::C, guidedChromaRadius, 0.001f, scale)    (ipsmoothing.cc L334-409)
//       r = max(int(round(radius / scale)), 0); guide = xlin2log(max(rgbLuminance, 0), 10)
//       guidedFilterLog(guide, 10, chan, r, eps) for R, G, B  (guidedfilter.cc L243-263: xlin2log, guidedFilter with automatic
//       subsampling, xlog2lin)
//       keep the INPUT luminance, take the filtered chroma scaled by Y_in / Y_filtered (Color::rgb2yuv / yuv2rgb, color.h L783-796)
//   Imagefloat::normalizeFloatTo65535                   (x *= 65535.f)
// Three kernels around guided.cu's filter: k_gs_prep (scale, input luminance, log guide, log channels -- one pass over the three
// planes instead of the reference's five), the guided filter per channel, k_gs_final (back from log, chroma transfer, scale).
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>

namespace {

__device__ __forceinline__ float max0(float v) { return v < 0.f ? 0.f : v; }     // rtengine::max(v, 0.f)

__global__ void __launch_bounds__(256) k_gs_prep(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H,
                                                 float* __restrict__ iY, float* __restrict__ guide, size_t sp, float w0, float w1, float w2, int filter)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float down = 1.f / 65535.f;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x;
        const float R = r[i] * down, G = g[i] * down, B = b[i] * down;
        if (!filter) { r[i] = R * 65535.f; g[i] = G * 65535.f; b[i] = B * 65535.f; continue; }      // rounded radius 0: the two scalings alone
        const float l = R * w0 + G * w1 + B * w2;
        const size_t o = (size_t)y * sp + x;
        iY[o] = l;
        guide[o] = sleef::xlin2log_scalar(max0(l), 10.f);
        r[i] = sleef::xlin2log_scalar(max0(R), 10.f);
        g[i] = sleef::xlin2log_scalar(max0(G), 10.f);
        b[i] = sleef::xlin2log_scalar(max0(B), 10.f);
    }
}

__global__ void __launch_bounds__(256) k_gs_final(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H,
                                                  const float* __restrict__ iYp, size_t sp, float w0, float w1, float w2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x;
        const float R = sleef::xlog2lin_scalar(max0(r[i]), 10.f), G = sleef::xlog2lin_scalar(max0(g[i]), 10.f), B = sleef::xlog2lin_scalar(max0(b[i]), 10.f);
        const float iY = iYp[(size_t)y * sp + x];
        float oY = R * w0 + G * w1 + B * w2;
        float ou = oY - B, ov = R - oY;
        const float bump = oY > 1e-5f ? iY / oY : 1.f;
        ou *= bump; ov *= bump; oY = iY;
        const float bb = oY - ou, rr = ov + oY;
        const float gg = (oY - rr * w0 - bb * w2) / w1;
        r[i] = rr * 65535.f; g[i] = gg * 65535.f; b[i] = bb * 65535.f;
    }
}

}  // namespace

int art_guided_smoothing_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const double* ws9, int guidedChromaRadius, double scale)
{
    if (guidedChromaRadius == 0) return ART_HP_OK;                  // ipsmoothing.cc L877-879
    if (!ws9) return ctx->fail(ART_HP_ERR_INVALID, "guided smoothing needs the working-space matrix");
    if (!(scale > 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "scale must be positive");
    cudaStream_t st = ctx->stream;
    const int rad = std::max((int)std::round(guidedChromaRadius / scale), 0);
    const float w0 = (float)ws9[3], w1 = (float)ws9[4], w2 = (float)ws9[5];
    const size_t sp = round_up((size_t)W, 32), n = sp * (size_t)H;
    float* iY = nullptr;
    if (rad > 0) {
        void* blk = nullptr;
        int rc = art_pool_alloc(ctx, 2 * n * sizeof(float), &blk);
        if (rc) return rc;
        iY = (float*)blk;
    }
    float* guide = iY ? iY + n : nullptr;
    const dim3 grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_gs_prep");
    k_gs_prep<<<grid, 256, 0, st>>>(r, g, b, ip, W, H, iY, guide, sp, w0, w1, w2, rad > 0);
    art_prof_end(ctx);
    ctx->launches++;
    int rc = ART_HP_OK;
    if (rad > 0) {
        float* ch[3] = {r, g, b};
        for (int c = 0; c < 3 && !rc; ++c) rc = art_guided_dev(ctx, guide, sp, ch[c], ip, ch[c], ip, W, H, rad, 0.001f, 0);
        if (!rc) {
            art_prof_begin(ctx, "k_gs_final");
            k_gs_final<<<grid, 256, 0, st>>>(r, g, b, ip, W, H, iY, sp, w0, w1, w2);
            art_prof_end(ctx);
            ctx->launches++;
        }
        art_pool_free(ctx, iY);
    }
    if (rc) return rc;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
