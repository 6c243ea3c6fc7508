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

// out-of-place form for art_hp_develop: getImage reads the demosaiced planes from (border, border) and writes a
// (W - 2 border) x (H - 2 border) image (rawimagesource.cc L943-1025 with sx1 = sy1 = border, transformRect L664-700)
struct ScCropArgs { const float *sr, *sg, *sb; size_t sp; ScArgs d; };
__global__ void __launch_bounds__(256) k_scale_convert_crop(ScCropArgs c)
{
    const ScArgs& a = c.d;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t i = (size_t)y * c.sp + x, o = (size_t)y * a.pitch + x;
        float r = c.sr[i], g = c.sg[i], b = c.sb[i];
        sc_pixel(a, r, g, b);
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
                               float* r, float* g, float* b, size_t pitch, const float mul[3], int doClip, const double* mat)
{
    ScCropArgs c;
    c.sr = sr; c.sg = sg; c.sb = sb; c.sp = sp;
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
