// Per-pixel stages of the hot path (HBM-bound streaming kernels).
//
// scale_convert: RawImageSource::getImage's gain/clip step (reference rtengine/rawimagesource.cc
// L943-1025 at skip == 1: `rtot *= rm; if (doClip) rtot = CLIP(rtot)`) fused with the matrix branch of
// RawImageSource::colorSpaceConversion_ (L3184-3213: double coefficients times float samples, summed
// in double, rounded to float once).  One pass over the three planes: 12 B read + 12 B written per pixel
// where the reference makes two passes (48 B/px).
#include "ctx.h"

namespace {

struct ScArgs {
    float *r, *g, *b; size_t pitch;
    int W, H;
    float mul[3];
    int do_clip, do_mat;
    double mat[9];
};

__device__ __forceinline__ float clip65535(float a)
{   // rt_math.h L97-101: CLIP(a) = LIM(a, 0, MAXVALF) = max(0, min(a, 65535))
    const float m = a < 65535.f ? a : 65535.f;     // std::min(a, hi)
    return 0.f < m ? m : 0.f;                       // std::max(lo, m)
}

__device__ __forceinline__ void sc_pixel(const ScArgs& a, float& r, float& g, float& b)
{
    r *= a.mul[0]; g *= a.mul[1]; b *= a.mul[2];
    if (a.do_clip) { r = clip65535(r); g = clip65535(g); b = clip65535(b); }
    if (a.do_mat) {
        const double dr = r, dg = g, db = b;
        const float nr = (float)(a.mat[0] * dr + a.mat[1] * dg + a.mat[2] * db);
        const float ng = (float)(a.mat[3] * dr + a.mat[4] * dg + a.mat[5] * db);
        const float nb = (float)(a.mat[6] * dr + a.mat[7] * dg + a.mat[8] * db);
        r = nr; g = ng; b = nb;
    }
}

// one thread per 4 consecutive pixels of a row (float4), grid-stride over rows
__global__ void __launch_bounds__(256) k_scale_convert(ScArgs a)
{
    const int w4 = a.W >> 2;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t row = (size_t)y * a.pitch;
        if (x < w4) {
            float4* pr = reinterpret_cast<float4*>(a.r + row) + x;
            float4* pg = reinterpret_cast<float4*>(a.g + row) + x;
            float4* pb = reinterpret_cast<float4*>(a.b + row) + x;
            float4 r = *pr, g = *pg, b = *pb;
            sc_pixel(a, r.x, g.x, b.x); sc_pixel(a, r.y, g.y, b.y); sc_pixel(a, r.z, g.z, b.z); sc_pixel(a, r.w, g.w, b.w);
            *pr = r; *pg = g; *pb = b;
        } else if (x == w4) {
            for (int c = w4 * 4; c < a.W; ++c) {      // ragged tail of the row
                float r = a.r[row + c], g = a.g[row + c], b = a.b[row + c];
                sc_pixel(a, r, g, b);
                a.r[row + c] = r; a.g[row + c] = g; a.b[row + c] = b;
            }
        }
    }
}

// scalar fallback for planes whose base/pitch are not 16-byte aligned
__global__ void __launch_bounds__(256) k_scale_convert_scalar(ScArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t i = (size_t)y * a.pitch + x;
        float r = a.r[i], g = a.g[i], b = a.b[i];
        sc_pixel(a, r, g, b);
        a.r[i] = r; a.g[i] = g; a.b[i] = b;
    }
}

// RawImageSource::HLRecovery_blend for one pixel (rawimagesource.cc L3613-3747; hlRecovery calls it with maxval 65535): clipped pixels get their
// chroma (in an opponent space) scaled to the unclipped estimate's, faded in above half the lowest clip point.  The tail mixes float variables with
// double literals, so those expressions are evaluated in double and rounded once, as in the reference.
struct HlArgs { float clip[3], maxave, clippt, fixpt, maxval; };
__device__ __forceinline__ float hl_min(float a, float b) { return b < a ? b : a; }      // rtengine::min
__device__ __forceinline__ void hl_blend_pixel(const HlArgs& h, float& rin, float& gin, float& bin)
{
    const float trans[3][3] = {{1, 1, 1}, {1.7320508f, -1.7320508f, 0}, {-1, -1, 2}};
    const float itrans[3][3] = {{1, 0.8660254f, -0.5f}, {1, -0.8660254f, -0.5f}, {1, 0, 1}};
    float rgb[3] = {rin, gin, bin};
    if (!(rgb[0] > h.clippt) && !(rgb[1] > h.clippt) && !(rgb[2] > h.clippt)) return;
    float cam[2][3], lab[2][3], sum[2], lratio = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) { lratio += hl_min(rgb[c], h.clip[c]); cam[0][c] = rgb[c]; cam[1][c] = hl_min(cam[0][c], h.maxval); }
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            lab[i][c] = 0;
#pragma unroll
            for (int j = 0; j < 3; j++) lab[i][c] += trans[c][j] * cam[i][j];
        }
        sum[i] = 0;
#pragma unroll
        for (int c = 1; c < 3; c++) sum[i] += lab[i][c] * lab[i][c];
    }
    const float chratio = sqrtf(sum[1] / sum[0]);
    lab[0][1] *= chratio; lab[0][2] *= chratio;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cam[0][c] = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) cam[0][c] += itrans[c][j] * lab[0][j];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) rgb[c] = cam[0][c] / 3;
    if (rin > h.fixpt) { const float t = (hl_min(h.clip[0], rin) - h.fixpt) / (h.clip[0] - h.fixpt), f = t * t; rin = hl_min(h.maxave, f * rgb[0] + (1 - f) * rin); }
    if (gin > h.fixpt) { const float t = (hl_min(h.clip[1], gin) - h.fixpt) / (h.clip[1] - h.fixpt), f = t * t; gin = hl_min(h.maxave, f * rgb[1] + (1 - f) * gin); }
    if (bin > h.fixpt) { const float t = (hl_min(h.clip[2], bin) - h.fixpt) / (h.clip[2] - h.fixpt), f = t * t; bin = hl_min(h.maxave, f * rgb[2] + (1 - f) * bin); }
    const float tot = (rin + gin + bin);
    lratio /= tot;
    const float L = tot / 3 / lratio;
    const float C = lratio * 1.732050808 * (rin - gin);
    const float H = lratio * (2 * bin - rin - gin);
    rin = L - H / 6.0 + C / 3.464101615;
    gin = L - H / 6.0 - C / 3.464101615;
    bin = L + H / 3.0;
}

// out-of-place form for art_hp_develop: getImage reads the demosaiced planes from (border, border) and writes a
// (W - 2 border) x (H - 2 border) image (rawimagesource.cc L943-1025 with sx1 = sy1 = border, transformRect L664-700): gains / clip, the optional
// "Blend" highlight reconstruction, the coarse rotation of rotateLine (L57-87) and the mirrors (L1079-1086) as the store's address, then the matrix.
// W / H below are the SOURCE line geometry (imwidth / imheight); a quarter turn writes an H-wide, W-high image.
struct ScCropArgs { const float *sr, *sg, *sb; size_t sp; ScArgs d; int tran, hr; HlArgs hl; int skip, sx1, sy1, maxx, maxy; };      // skip > 1: the preview form, source planes un-offset
__global__ void __launch_bounds__(256) k_scale_convert_crop(ScCropArgs c)
{
    const ScArgs& a = c.d;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    const int rot = c.tran & 3, ow = (rot & 1) ? a.H : a.W, oh = (rot & 1) ? a.W : a.H;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        float r, g, b;
        if (c.skip <= 1) {
            const size_t i = (size_t)y * c.sp + x;
            r = c.sr[i]; g = c.sg[i]; b = c.sb[i];
        } else {
            // the preview form (L943-968): skip x skip box sum from (min(sy1 + skip y, maxy - skip), min(sx1 + skip x, maxx - skip)), rows outer, from 0
            const int i0 = min(c.sy1 + c.skip * y, c.maxy - c.skip), j0 = min(c.sx1 + c.skip * x, c.maxx - c.skip);
            r = g = b = 0.f;
            for (int m = 0; m < c.skip; ++m)
                for (int n = 0; n < c.skip; ++n) {
                    const size_t i = (size_t)(i0 + m) * c.sp + (j0 + n);
                    r += c.sr[i]; g += c.sg[i]; b += c.sb[i];
                }
        }
        r *= a.mul[0]; g *= a.mul[1]; b *= a.mul[2];
        if (a.do_clip) { r = clip65535(r); g = clip65535(g); b = clip65535(b); }
        if (c.hr) hl_blend_pixel(c.hl, r, g, b);
        if (a.do_mat) {
            const double dr = r, dg = g, db = b;
            const float nr = (float)(a.mat[0] * dr + a.mat[1] * dg + a.mat[2] * db);
            const float ng = (float)(a.mat[3] * dr + a.mat[4] * dg + a.mat[5] * db);
            const float nb = (float)(a.mat[6] * dr + a.mat[7] * dg + a.mat[8] * db);
            r = nr; g = ng; b = nb;
        }
        int row = y, col = x;
        if (rot == 2) { row = a.H - 1 - y; col = a.W - 1 - x; }
        else if (rot == 1) { row = x; col = a.H - 1 - y; }
        else if (rot == 3) { row = a.W - 1 - x; col = y; }
        if (c.tran & 8) col = ow - 1 - col;
        if (c.tran & 4) row = oh - 1 - row;
        const size_t o = (size_t)row * a.pitch + col;
        a.r[o] = r; a.g[o] = g; a.b[o] = b;
    }
}

// scaleColors, Bayer branch (rawimagesource.cc L2731-2772): black subtraction + per-CFA-channel scaling in
// place, and the per-colour maxima chmax[] (max is exact under any association, so a tree reduce is fine;
// values are >= 0, hence atomicMax on the int bit pattern orders them correctly).
struct ScaleColorsArgs {
    float* raw; size_t pitch; int W, H; unsigned filters;
    float black[4], mul[4];
    int* chmax_bits;          // 3 ints, zeroed by the caller
};

__device__ __forceinline__ unsigned fc_pw(unsigned filters, int row, int col)
{   // RawImage::FC, rtengine/rawimage.h L186-189
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}

__global__ void __launch_bounds__(256) k_scale_colors(ScaleColorsArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    float mx[3] = {0.f, 0.f, 0.f};
    if (x < a.W) {
        for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
            const int c = fc_pw(a.filters, y, x);
            const int c4 = (c == 1 && !(y & 1)) ? 3 : c;              // 0=R, 1=G1, 2=B, 3=G2
            float* p = a.raw + (size_t)y * a.pitch + x;
            const float d = *p - a.black[c4];
            const float val = (0.f < d ? d : 0.f) * a.mul[c4];       // max(0.f, raw - black) * mul
            *p = val;
            // a column holds at most two colours; keep the three running maxima branch-free
            mx[0] = (c == 0 && mx[0] < val) ? val : mx[0];
            mx[1] = (c == 1 && mx[1] < val) ? val : mx[1];
            mx[2] = (c == 2 && mx[2] < val) ? val : mx[2];
        }
    }
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = mx[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float w = __shfl_xor_sync(0xffffffffu, v, o); v = v < w ? w : v; }
        if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(a.chmax_bits + k, __float_as_int(v));
    }
}

// scaleColors, X-Trans branch (rawimagesource.cc L2795-2826): c = XTRANSFC(row, col)
struct ScaleColorsXtArgs { float* raw; size_t pitch; int W, H; int xt[36]; float black[3], mul[3]; int* chmax_bits; };
__global__ void __launch_bounds__(256) k_scale_colors_xtrans(const __grid_constant__ ScaleColorsXtArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    float mx[3] = {0.f, 0.f, 0.f};
    if (x < a.W) {
        for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
            const int c = a.xt[(y % 6) * 6 + (x % 6)];
            float* p = a.raw + (size_t)y * a.pitch + x;
            const float d = *p - (c == 0 ? a.black[0] : c == 1 ? a.black[1] : a.black[2]);
            const float val = (0.f < d ? d : 0.f) * (c == 0 ? a.mul[0] : c == 1 ? a.mul[1] : a.mul[2]);
            *p = val;
            mx[0] = (c == 0 && mx[0] < val) ? val : mx[0];
            mx[1] = (c == 1 && mx[1] < val) ? val : mx[1];
            mx[2] = (c == 2 && mx[2] < val) ? val : mx[2];
        }
    }
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = mx[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float w = __shfl_xor_sync(0xffffffffu, v, o); v = v < w ? w : v; }
        if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(a.chmax_bits + k, __float_as_int(v));
    }
}

}  // namespace

int art_scale_colors_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch,
                         const float black[4], const float mul[4], int* d_chmax_bits)
{
    ScaleColorsArgs a;
    a.raw = raw; a.pitch = pitch; a.W = W; a.H = H; a.filters = filters; a.chmax_bits = d_chmax_bits;
    for (int i = 0; i < 4; ++i) { a.black[i] = black[i]; a.mul[i] = mul[i]; }
    ART_CUDA(ctx, cudaMemsetAsync(d_chmax_bits, 0, 3 * sizeof(int), ctx->stream));
    dim3 grid((W + 255) / 256, std::min(H, 148 * 4));
    art_prof_begin(ctx, "k_scale_colors");
    k_scale_colors<<<grid, 256, 0, ctx->stream>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_scale_convert_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch,
                          const float mul[3], int doClip, const double* mat)
{
    ScArgs a;
    a.r = r; a.g = g; a.b = b; a.pitch = pitch; a.W = W; a.H = H;
    for (int i = 0; i < 3; ++i) a.mul[i] = mul[i];
    a.do_clip = doClip; a.do_mat = mat != nullptr;
    for (int i = 0; i < 9; ++i) a.mat[i] = mat ? mat[i] : 0.0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 && (pitch & 3) == 0;
    const int rows = std::min(H, 148 * 8);
    if (aligned) {
        dim3 grid(((W >> 2) + 1 + 255) / 256, rows);
        art_prof_begin(ctx, "k_scale_convert");
        k_scale_convert<<<grid, 256, 0, ctx->stream>>>(a);
    } else {
        dim3 grid((W + 255) / 256, rows);
        art_prof_begin(ctx, "k_scale_convert_scalar");
        k_scale_convert_scalar<<<grid, 256, 0, ctx->stream>>>(a);
    }
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_scale_convert_crop_dev(art_hp_ctx* ctx, int W, int H, const float* sr, const float* sg, const float* sb, size_t sp,
                               float* r, float* g, float* b, size_t pitch, const float mul[3], int doClip, const double* mat,
                               int tran, int hr_blend, const float* hlmax, int skip, int sx1, int sy1, int maxx, int maxy)
{
    ScCropArgs c;
    c.sr = sr; c.sg = sg; c.sb = sb; c.sp = sp;
    c.skip = skip; c.sx1 = sx1; c.sy1 = sy1; c.maxx = maxx; c.maxy = maxy;
    c.tran = tran; c.hr = hr_blend && hlmax;
    if (c.hr) {       // the line constants of HLRecovery_blend (L3623-3640), maxval = 65535
        const float minpt = std::min(std::min(hlmax[0], hlmax[1]), hlmax[2]);
        c.hl.maxave = (hlmax[0] + hlmax[1] + hlmax[2]) / 3;
        for (int k = 0; k < 3; ++k) c.hl.clip[k] = std::min(c.hl.maxave, hlmax[k]);
        const float clipthresh = 0.95, fixthresh = 0.5;
        c.hl.maxval = 65535.0f;
        c.hl.clippt = clipthresh * c.hl.maxval;
        c.hl.fixpt = fixthresh * minpt;
    } else c.hl = HlArgs{};
    ScArgs& a = c.d;
    a.r = r; a.g = g; a.b = b; a.pitch = pitch; a.W = W; a.H = H;
    for (int i = 0; i < 3; ++i) a.mul[i] = mul[i];
    a.do_clip = doClip; a.do_mat = mat != nullptr;
    for (int i = 0; i < 9; ++i) a.mat[i] = mat ? mat[i] : 0.0;
    dim3 grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_scale_convert_crop");
    k_scale_convert_crop<<<grid, 256, 0, ctx->stream>>>(c);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

// ------------------------------------------------------------------ ImProcFunctions::channelMixer (ipchmixer.cc L152-232), the per-pixel loop
namespace {
struct MixArgs { float* r; float* g; float* b; size_t pitch; int W, H; float m[9]; };
__global__ void __launch_bounds__(256) k_channel_mixer(MixArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    const bool vec = (x & ~3) + 4 <= a.W;        // the SSE2 loop `for (; x < W - 3; x += 4)` clamps with _mm_max_ps (NaN -> 0), the tail with rt_math.h max
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t o = (size_t)y * a.pitch + x;
        const float r = a.r[o], g = a.g[o], b = a.b[o];
        const float rmix = (r * a.m[0] + g * a.m[1] + b * a.m[2]);
        const float gmix = (r * a.m[3] + g * a.m[4] + b * a.m[5]);
        const float bmix = (r * a.m[6] + g * a.m[7] + b * a.m[8]);
        a.r[o] = vec ? (rmix > 0.f ? rmix : 0.f) : (rmix < 0.f ? 0.f : rmix);
        a.g[o] = vec ? (gmix > 0.f ? gmix : 0.f) : (gmix < 0.f ? 0.f : gmix);
        a.b[o] = vec ? (bmix > 0.f ? bmix : 0.f) : (bmix < 0.f ? 0.f : bmix);
    }
}
}  // namespace

int art_channel_mixer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const float m[9])
{
    MixArgs a;
    a.r = r; a.g = g; a.b = b; a.pitch = pitch; a.W = W; a.H = H;
    for (int i = 0; i < 9; ++i) a.m[i] = m[i];
    art_prof_begin(ctx, "k_channel_mixer");
    k_channel_mixer<<<dim3((W + 255) / 256, std::min(H, 148 * 8)), 256, 0, ctx->stream>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_scale_colors_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, float* raw, size_t pitch,
                                const float black[3], const float mul[3], int* d_chmax_bits)
{
    ScaleColorsXtArgs a;
    a.raw = raw; a.pitch = pitch; a.W = W; a.H = H; a.chmax_bits = d_chmax_bits;
    for (int i = 0; i < 36; ++i) a.xt[i] = xtrans36[i];
    for (int i = 0; i < 3; ++i) { a.black[i] = black[i]; a.mul[i] = mul[i]; }
    ART_CUDA(ctx, cudaMemsetAsync(d_chmax_bits, 0, 3 * sizeof(int), ctx->stream));
    art_prof_begin(ctx, "k_scale_colors_xtrans");
    k_scale_colors_xtrans<<<dim3((W + 255) / 256, std::min(H, 148 * 4)), 256, 0, ctx->stream>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
