// The two pixel loops of ImProcFunctions::blackAndWhite (reference rtengine/ipbw.cc L283-312, L343-362), the last step of
// ImProcFunctions::process STAGE_3: bw = (bwr r + bwg g + bwb b) kcorec with optional per-channel gamma tables, then the colour cast --
// Imagefloat::setMode(YUV) (imagefloat.cc L700-725), u += ulut[Y], v += vlut[Y], and the setMode(RGB) of the next stage (L779-803).
// computeBWMixerConstants (L50-217) and the five 65536-entry tables stay host code of the reference and arrive as parameters.
// One kernel, one thread per SSE2 group of the reference's row loops (vector LUT rule) or per scalar row tail; 12 B read + 12 B written
// per pixel where the reference makes two to four passes.  Bit-identical to the reference.
#include "ctx.h"

namespace {

__device__ __forceinline__ float vmaxf_(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float vminf_(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }
__device__ __forceinline__ float lut_s(const float* __restrict__ data, float index)
{   // LUT.h L437-459, LUTf(65536): clips below and above
    const int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > 65534.f) return data[65535];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
__device__ __forceinline__ float lut_v(const float* __restrict__ data, float index)
{   // LUT.h L349-377
    const int idx = (int)vclampf_(index, 0.f, 65534.f);
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, 65535.f) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}

struct BwArgs {
    float *r, *g, *b; size_t ip; int W, H;
    float bwr, bwg, bwb, kcorec;
    const float *gr, *gg, *gb, *ul, *vl;
    float w0, w1, w2;
    int aligned;
};

template <bool VEC>
__device__ __forceinline__ void bw_pixel(const BwArgs& a, float& r, float& g, float& b)
{
    if (a.gr) {
        r = VEC ? lut_v(a.gr, r) : lut_s(a.gr, r);
        g = VEC ? lut_v(a.gg, g) : lut_s(a.gg, g);
        b = VEC ? lut_v(a.gb, b) : lut_s(a.gb, b);
    }
    const float bw = ((a.bwr * r + a.bwg * g + a.bwb * b) * a.kcorec);
    r = g = b = bw;
    if (a.ul) {
        const float Y = r * a.w0 + g * a.w1 + b * a.w2;
        float u = Y - b, v = r - Y;
        u += VEC ? lut_v(a.ul, Y) : lut_s(a.ul, Y);
        v += VEC ? lut_v(a.vl, Y) : lut_s(a.vl, Y);
        b = Y - u; r = v + Y;
        g = (Y - r * a.w0 - b * a.w2) / a.w1;
    }
}

__global__ void __launch_bounds__(128) k_bw(const BwArgs a)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= a.W) return;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t row = (size_t)y * a.ip;
        if (x0 + 4 <= a.W && a.aligned) {        // 16-byte aligned planes: one 128-bit access per plane and group
            float4 vr = *reinterpret_cast<const float4*>(a.r + row + x0), vg = *reinterpret_cast<const float4*>(a.g + row + x0),
                   vb = *reinterpret_cast<const float4*>(a.b + row + x0);
            bw_pixel<true>(a, vr.x, vg.x, vb.x); bw_pixel<true>(a, vr.y, vg.y, vb.y); bw_pixel<true>(a, vr.z, vg.z, vb.z); bw_pixel<true>(a, vr.w, vg.w, vb.w);
            *reinterpret_cast<float4*>(a.r + row + x0) = vr; *reinterpret_cast<float4*>(a.g + row + x0) = vg; *reinterpret_cast<float4*>(a.b + row + x0) = vb;
        } else if (x0 + 4 <= a.W) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const size_t i = row + x0 + k;
                float r = a.r[i], g = a.g[i], b = a.b[i];
                bw_pixel<true>(a, r, g, b);
                a.r[i] = r; a.g[i] = g; a.b[i] = b;
            }
        } else {
            for (int x = x0; x < a.W; ++x) {
                const size_t i = row + x;
                float r = a.r[i], g = a.g[i], b = a.b[i];
                bw_pixel<false>(a, r, g, b);
                a.r[i] = r; a.g[i] = g; a.b[i] = b;
            }
        }
    }
}

// proPhotoBlue (reference rtengine/improcfun.cc L312-357): what ImProcFunctions::process runs after toneEqualizer when the working profile is ProPhoto
// (L585-587).  A pixel with r == 0 or g == 0 and no negative channel goes through Color::rgb2hsv (fp64, color.cc L586-622), loses 1 % of its saturation and
// comes back through Color::hsv2rgb (L654-694); the reference's SSE2 group test only skips groups without such a pixel.
__global__ void __launch_bounds__(256) k_prophoto_blue(float* __restrict__ R, float* __restrict__ G, float* __restrict__ B, size_t ip, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x;
        const float r = R[i], g = G[i], b = B[i];
        const float mn0 = g < r ? g : r, mn = b < mn0 ? b : mn0;
        if (!((r == 0.0f || g == 0.0f) && mn >= 0.f)) continue;
        const double var_R = r / 65535.0, var_G = g / 65535.0, var_B = b / 65535.0;
        const double m0 = var_G < var_R ? var_G : var_R, var_Min = var_B < m0 ? var_B : m0;
        const double x0 = var_R < var_G ? var_G : var_R, var_Max = x0 < var_B ? var_B : x0;
        const double del_Max = var_Max - var_Min;
        float h = 0.f, s, v = (float)var_Max;
        if (del_Max < 0.00001 && del_Max > -0.00001) s = 0.f;
        else {
            s = (float)(del_Max / (var_Max == 0.0 ? 1.0 : var_Max));
            if (var_R == var_Max) h = (float)((var_G - var_B) / del_Max);
            else if (var_G == var_Max) h = (float)(2.0 + (var_B - var_R) / del_Max);
            else if (var_B == var_Max) h = (float)(4.0 + (var_R - var_G) / del_Max);
            h /= 6.f;
            if (h < 0.f) h += 1.f;
            if (h > 1.f) h -= 1.f;
        }
        s *= 0.99f;
        const float h1 = h * 6.f;
        const int k = (int)h1;
        const float f = h1 - k;
        const float p = v * (1.f - s), q = v * (1.f - s * f), t = v * (1.f - s * (1.f - f));
        float r1, g1, b1;
        if (k == 1) { r1 = q; g1 = v; b1 = p; }
        else if (k == 2) { r1 = p; g1 = v; b1 = t; }
        else if (k == 3) { r1 = p; g1 = q; b1 = v; }
        else if (k == 4) { r1 = t; g1 = p; b1 = v; }
        else if (k == 5) { r1 = v; g1 = p; b1 = q; }
        else { r1 = v; g1 = t; b1 = p; }
        R[i] = r1 * 65535.0f; G[i] = g1 * 65535.0f; B[i] = b1 * 65535.0f;
    }
}

}  // namespace

int art_prophoto_blue_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip)
{
    art_prof_begin(ctx, "k_prophoto_blue");
    k_prophoto_blue<<<dim3((W + 255) / 256, std::min(H, 148 * 8)), 256, 0, ctx->stream>>>(r, g, b, ip, W, H);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_bw_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_bw_params* p)
{
    const bool gamma = p->gamma_r || p->gamma_g || p->gamma_b, cast = p->ulut || p->vlut;
    if (gamma && !(p->gamma_r && p->gamma_g && p->gamma_b)) return ctx->fail(ART_HP_ERR_INVALID, "black and white: the three gamma tables come together");
    if (cast && !(p->ulut && p->vlut && p->ws)) return ctx->fail(ART_HP_ERR_INVALID, "black and white: the colour cast needs ulut, vlut and ws");
    cudaStream_t st = ctx->stream;
    constexpr size_t N = 65536;
    void* blk = nullptr;
    int rc = art_pool_alloc(ctx, 5 * N * sizeof(float), &blk);
    if (rc) return rc;
    float* d = (float*)blk;
    const float* src[5] = {p->gamma_r, p->gamma_g, p->gamma_b, p->ulut, p->vlut};
    const float* dev[5] = {};
    for (int k = 0; k < 5; ++k) {
        if (!src[k]) continue;
        // pageable source: the call returns once the bytes are staged
        const cudaError_t e = cudaMemcpyAsync(d + k * N, src[k], N * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { art_pool_free(ctx, blk); return ctx->fail(ART_HP_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(e)); }
        dev[k] = d + k * N;
    }
    BwArgs a{};
    a.r = r; a.g = g; a.b = b; a.ip = ip; a.W = W; a.H = H;
    a.bwr = p->bwr; a.bwg = p->bwg; a.bwb = p->bwb; a.kcorec = p->kcorec;
    a.gr = dev[0]; a.gg = dev[1]; a.gb = dev[2]; a.ul = dev[3]; a.vl = dev[4];
    a.aligned = ip % 4 == 0 && ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    if (p->ws) { a.w0 = (float)p->ws[3]; a.w1 = (float)p->ws[4]; a.w2 = (float)p->ws[5]; }
    art_prof_begin(ctx, "k_bw");
    k_bw<<<dim3(((W + 3) / 4 + 127) / 128, std::min(H, 148 * 8)), 128, 0, st>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    art_pool_free(ctx, blk);
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
