// Lanczos resampler for sm_100a.
//
// Replaces ImProcFunctions::Lanczos (reference rtengine/ipresize.cc L38-207; the caller is ImProcFunctions::resize L365-394, the
// destination size comes from resizeScale L230-362: int(w * scale + 0.5)), without the Lab round trip around it (src->setMode(LAB) /
// dst->setMode(mode): per-pixel colour conversions, the colour chain's business).  a = 3 lobes, support = int(2 a / min(scale, 1)) + 1
// taps, weights a sin(pi x) sin(pi x / a) / (pi x)^2 through sleef's scalar xsinf, normalised per output row / column; an output row
// is interpolated vertically into a source-width line, then horizontally -- both sums in ascending tap order from 0, a product and a sum
// per tap (the reference's SSE2 4-column groups and its scalar tail accumulate in that same order).
//   k_lz_weights   one thread per output column (or row): first / last tap and the normalised weights (the sum over taps is serial, as
//                  in the reference)
//   k_lanczos      CTA = (one output row) x (a tile of output columns) x (plane): the vertical interpolation of exactly the source columns the
//                  tile's taps reach goes into shared memory (coalesced reads of `support` source rows, which neighbouring output rows
//                  share through L2), then one thread per output pixel sums its horizontal taps out of shared memory.  The source-width
//                  line of the reference never exists in HBM: 4 B read per source sample (x the vertical overlap L2 absorbs), 4 B written
//                  per output sample.
// Compiled with -fmad=false, IEEE division.
#include "ctx.h"

namespace {

__device__ __forceinline__ float xsinf_scalar(float d)
{   // sleef.h L993-1016; xrintf is cvtss2si (round to nearest even) on the reference's SSE2 build
    const int q = __float2int_rn(d * (float)0.31830988618379067154);
    d = q * (-0.78515625f * 4) + d;
    d = q * (-0.00024127960205078125f * 4) + d;
    d = q * (-6.3329935073852539062e-07f * 4) + d;
    d = q * (-4.9604681473525147339e-10f * 4) + d;
    const float s = d * d;
    if ((q & 1) != 0) d = -d;
    float u = 2.6083159809786593541503e-06f;
    u = u * s + -0.0001981069071916863322258f;
    u = u * s + 0.00833307858556509017944336f;
    u = u * s + -0.166666597127914428710938f;
    u = s * (u * d) + d;
    return u;
}

__device__ __forceinline__ float lanc(float x, float a)
{   // ipresize.cc L38-48
    if (x * x < 1e-6f) return 1.0f;
    if (x * x > a * a) return 0.0f;
    x = (float)3.14159265358979323846 * x;
    return a * xsinf_scalar(x) * xsinf_scalar(x / a) / (x * x);
}

// taps of every output sample along one axis (L82-109 for columns, L136-152 for rows)
__global__ void __launch_bounds__(128) k_lz_weights(int n_out, int n_src, float delta, float sc, float a, int support,
                                                    float* __restrict__ w_all, int* __restrict__ lo_all, int* __restrict__ hi_all)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const float x0 = ((float)j + 0.5f) * delta - 0.5f;
    float* w = w_all + (size_t)j * support;
    int lo = (int)floorf(x0 - a / sc) + 1; if (lo < 0) lo = 0;
    int hi = (int)floorf(x0 + a / sc) + 1; if (hi > n_src) hi = n_src;
    lo_all[j] = lo; hi_all[j] = hi;
    float ws = 0.0f;
    for (int k = 0; k < support; ++k) {
        float v = 0.f;
        if (lo + k < hi) {
            v = lanc(sc * (x0 - (float)(lo + k)), a);
            ws += v;
        }
        w[k] = v;
    }
    for (int k = 0; k < support; ++k) w[k] /= ws;
}

struct LzArgs {
    const float* src[3]; size_t sp;
    float* dst[3]; size_t dp;
    int sW, sH, dW, dH, support, tile;
    const float *wh, *wv;
    const int *jj0, *jj1, *ii0, *ii1;
};

__global__ void __launch_bounds__(256) k_lanczos(const __grid_constant__ LzArgs A)
{
    extern __shared__ float line[];
    const int i = blockIdx.y, c = blockIdx.z;
    const int j0 = blockIdx.x * A.tile, j1 = min(j0 + A.tile, A.dW);
    // source columns this tile's taps reach (jj0 / jj1 are non-decreasing in j)
    const int s0 = A.jj0[j0], s1 = A.jj1[j1 - 1];
    const int ii0 = A.ii0[i], ii1 = A.ii1[i];
    const float* __restrict__ wv = A.wv + (size_t)i * A.support;
    const float* __restrict__ src = A.src[c];
    for (int jj = s0 + threadIdx.x; jj < s1; jj += blockDim.x) {
        float v = 0.0f;
        for (int ii = ii0; ii < ii1; ++ii) v += wv[ii - ii0] * src[(size_t)ii * A.sp + jj];
        line[jj - s0] = v;
    }
    __syncthreads();
    for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        const float* __restrict__ wh = A.wh + (size_t)j * A.support;
        const int a0 = A.jj0[j], a1 = A.jj1[j];
        float v = 0.0f;
        for (int jj = a0; jj < a1; ++jj) v += wh[jj - a0] * line[jj - s0];
        A.dst[c][(size_t)i * A.dp + j] = v;
    }
}

}  // namespace

int art_lanczos_dev(art_hp_ctx* ctx, const float* s0, const float* s1, const float* s2, size_t sp, int sW, int sH,
                    float* d0, float* d1, float* d2, size_t dp, int dW, int dH, float scale)
{
    cudaStream_t st = ctx->stream;
    const float delta = 1.0f / scale;
    const float a = 3.0f;
    const float sc = std::min(scale, 1.0f);
    const int support = (int)(2.0f * a / sc) + 1;
    // shared-memory line of one tile: tile * delta source columns plus the taps' reach; 44 KB stays under the default dynamic limit
    const int max_line = 11264;
    int tile = 256;
    while (tile > 1 && (int)((double)tile * delta) + support + 4 > max_line) tile >>= 1;
    if ((int)((double)tile * delta) + support + 4 > max_line)
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "Lanczos scale %g: %d taps do not fit the shared-memory line", (double)scale, support);
    const size_t wbytes = round_up(((size_t)dW + dH) * support * sizeof(float), 256);
    const size_t ibytes = round_up((size_t)2 * (dW + dH) * sizeof(int), 256);
    int rc = art_reserve(ctx, ctx->d_small2, wbytes + ibytes);
    if (rc) return rc;
    float* wh = (float*)ctx->d_small2.p;
    float* wv = wh + (size_t)dW * support;
    int* jj0 = (int*)((char*)ctx->d_small2.p + wbytes);
    int *jj1 = jj0 + dW, *ii0 = jj1 + dW, *ii1 = ii0 + dH;
    art_prof_begin(ctx, "k_lz_weights");
    k_lz_weights<<<(dW + 127) / 128, 128, 0, st>>>(dW, sW, delta, sc, a, support, wh, jj0, jj1);
    k_lz_weights<<<(dH + 127) / 128, 128, 0, st>>>(dH, sH, delta, sc, a, support, wv, ii0, ii1);
    art_prof_end(ctx);
    LzArgs A{};
    A.src[0] = s0; A.src[1] = s1; A.src[2] = s2; A.sp = sp;
    A.dst[0] = d0; A.dst[1] = d1; A.dst[2] = d2; A.dp = dp;
    A.sW = sW; A.sH = sH; A.dW = dW; A.dH = dH; A.support = support; A.tile = tile;
    A.wh = wh; A.wv = wv; A.jj0 = jj0; A.jj1 = jj1; A.ii0 = ii0; A.ii1 = ii1;
    const size_t smem = ((size_t)((double)tile * delta) + support + 4) * sizeof(float);
    art_prof_begin(ctx, "k_lanczos");
    for (int y0 = 0; y0 < dH; y0 += 65535) {       // grid.y limit
        LzArgs B = A;
        const int ny = std::min(65535, dH - y0);
        for (int c = 0; c < 3; ++c) B.dst[c] = A.dst[c] + (size_t)y0 * dp;
        B.wv = A.wv + (size_t)y0 * support; B.ii0 = A.ii0 + y0; B.ii1 = A.ii1 + y0;
        k_lanczos<<<dim3((dW + tile - 1) / tile, ny, 3), 256, smem, st>>>(B);
        ctx->launches++;
    }
    art_prof_end(ctx);
    ctx->launches += 2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
