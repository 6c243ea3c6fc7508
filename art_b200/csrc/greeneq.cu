// Bayer green equilibration (preprocess) for sm_100a.
//
// Replaces (reference) rtengine/green_equil_RT.cc RawImageSource::green_equilibrate_global L37-89 and ::green_equilibrate L92-250, the SSE2 build.
//   * global: the two green phases are scaled to their common mean.  The sums are doubles added left to right along a row and then row after row
//     (the reference reduces the rows with OpenMP, whose order moves with the schedule; the one-thread order is the oracle's and this kernel's): one warp
//     per row stages coalesced chunks in shared memory and lane 0 adds them in order, one thread then adds the row sums in order.
//   * local: where the two green populations around a green site differ by more than the local texture explains, the site moves half way to a
//     gradient-weighted diagonal interpolation.  Reads go to a copy of the frame (the reference packs the green sites into a copy), writes to the
//     frame: every site is independent, one thread per site.  The 8-wide SSE2 groups and the scalar tail associate the two four-term sums and the
//     threshold product differently; both are kept, by column.
// Bit-identical to the oracle (tests/test_greeneq_gpu.py).  Compiled with -fmad=false.
#include "ctx.h"

namespace {

__device__ __forceinline__ unsigned fc_ge(unsigned filters, int row, int col) { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }

constexpr int GE_CHUNK = 256;
__global__ void __launch_bounds__(128) k_ge_rowsums(const float* __restrict__ raw, size_t pitch, int W, int H, unsigned filters, int border, double* __restrict__ rowsum)
{
    __shared__ float stage[4][GE_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = border + blockIdx.x * 4 + warp;
    if (i >= H - border) return;
    const int j0 = border + ((fc_ge(filters, i, border) & 1) ^ 1);
    const int n = j0 < W - border ? (W - border - j0 + 1) / 2 : 0;          // green sites j0, j0 + 2, ... < W - border
    const float* row = raw + (size_t)i * pitch;
    double acc = 0.;
    for (int k0 = 0; k0 < n; k0 += GE_CHUNK) {
        const int m = min(GE_CHUNK, n - k0);
        for (int k = lane; k < m; k += 32) stage[warp][k] = row[j0 + 2 * (k0 + k)];
        __syncwarp();
        if (lane == 0)
            for (int k = 0; k < m; ++k) acc += stage[warp][k];
        __syncwarp();
    }
    if (lane == 0) rowsum[i] = acc;
}

// corr[0] = corrg1 (even rows), corr[1] = corrg2 (odd rows), L63-75
__global__ void k_ge_corr(const double* __restrict__ rowsum, int W, int H, unsigned filters, int border, double* __restrict__ corr)
{
    if (threadIdx.x || blockIdx.x) return;
    int ng1 = 0, ng2 = 0;
    double avgg1 = 0., avgg2 = 0.;
    for (int i = border; i < H - border; i++) {
        const int ng = (W - 2 * border + (fc_ge(filters, i, border) & 1)) / 2;
        if (i & 1) { avgg2 += rowsum[i]; ng2 += ng; } else { avgg1 += rowsum[i]; ng1 += ng; }
    }
    if (ng1 == 0 || avgg1 == 0.0) { ng1 = 1; avgg1 = 1.0; }
    if (ng2 == 0 || avgg2 == 0.0) { ng2 = 1; avgg2 = 1.0; }
    corr[0] = (avgg1 / ng1 + avgg2 / ng2) / 2.0 / (avgg1 / ng1);
    corr[1] = (avgg1 / ng1 + avgg2 / ng2) / 2.0 / (avgg2 / ng2);
}

__global__ void __launch_bounds__(256) k_ge_scale(float* __restrict__ raw, size_t pitch, int W, int H, unsigned filters, int border, const double* __restrict__ corr)
{
    const int i = border + blockIdx.y;
    if (i >= H - border) return;
    const int j = border + ((fc_ge(filters, i, border) & 1) ^ 1) + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (j >= W - border) return;
    float* p = raw + (size_t)i * pitch + j;
    *p = (float)((double)*p * corr[i & 1]);          // rawData[i][j] *= corrg with a double corrg
}

struct GeArgs { const float* src; float* raw; size_t sp, pitch; int W, H; unsigned filters; float thresh; const float* tmap; size_t tp; };
__global__ void __launch_bounds__(256) k_ge_local(GeArgs a)
{
    const int rr = 4 + blockIdx.y;
    if (rr >= a.H - 4) return;
    const int width = a.W;
    const int c0 = 5 - (fc_ge(a.filters, rr, 2) & 1);
    const int cc = c0 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (cc >= width - 6) return;
    // columns taken by the 8-wide SSE2 loop: c0, c0 + 8, ... while cc < width - 12, each iteration covering cc, cc + 2, cc + 4, cc + 6
    const int vec_end = c0 < width - 12 ? c0 + (width - 12 - c0 + 7) / 8 * 8 : c0;
    const bool vec = cc < vec_end;
    // C(r, x): green site x of row r in the reference's packed copy = column 2 x + (first green column of the row)
    auto C = [&](int r, int x) { return a.src[(size_t)r * a.sp + 2 * x + ((fc_ge(a.filters, r, 0) & 1) ^ 1)]; };
    const float eps = 1.f;
    const float o1_1 = C(rr - 1, (cc - 1) >> 1), o1_2 = C(rr - 1, (cc + 1) >> 1), o1_3 = C(rr + 1, (cc - 1) >> 1), o1_4 = C(rr + 1, (cc + 1) >> 1);
    const float o2_1 = C(rr - 2, cc >> 1), o2_2 = C(rr + 2, cc >> 1), o2_3 = C(rr, (cc >> 1) - 1), o2_4 = C(rr, (cc >> 1) + 1);
    float d1, d2;
    if (vec) { d1 = ((o1_1 + o1_2) + o1_3) + o1_4; d2 = ((o2_1 + o2_2) + o2_3) + o2_4; }
    else { d1 = (o1_1 + o1_2) + (o1_3 + o1_4); d2 = (o2_1 + o2_2) + (o2_3 + o2_4); }
    const float c1 = (fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4) + fabsf(o1_2 - o1_4));
    const float c2 = (fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4) + fabsf(o2_2 - o2_4));
    const float tf = a.tmap ? a.tmap[(size_t)rr * a.tp + cc] : a.thresh;
    const bool hit = vec ? (c1 + c2) < (6.f * tf) * fabsf(d1 - d2) : (c1 + c2) < 6 * tf * fabsf(d1 - d2);
    if (!hit) return;
    const float gin = C(rr, cc >> 1);
    const float gmp2p2 = gin - C(rr + 2, (cc >> 1) + 1), gmm2m2 = gin - C(rr - 2, (cc >> 1) - 1);
    const float gmm2p2 = gin - C(rr - 2, (cc >> 1) + 1), gmp2m2 = gin - C(rr + 2, (cc >> 1) - 1);
    const float gse = o1_4 + 0.5f * gmp2p2, gnw = o1_1 + 0.5f * gmm2m2, gne = o1_2 + 0.5f * gmm2p2, gsw = o1_3 + 0.5f * gmp2m2;
    const float t1 = C(rr + 3, (cc + 3) >> 1) - o1_4, t2 = C(rr - 3, (cc - 3) >> 1) - o1_1, t3 = C(rr - 3, (cc + 3) >> 1) - o1_2, t4 = C(rr + 3, (cc - 3) >> 1) - o1_3;
    const float wtse = 1.f / (eps + gmp2p2 * gmp2p2 + t1 * t1);
    const float wtnw = 1.f / (eps + gmm2m2 * gmm2m2 + t2 * t2);
    const float wtne = 1.f / (eps + gmm2p2 * gmm2p2 + t3 * t3);
    const float wtsw = 1.f / (eps + gmp2m2 * gmp2m2 + t4 * t4);
    const float ginterp = (gse * wtse + gnw * wtnw + gne * wtne + gsw * wtsw) / (wtse + wtnw + wtne + wtsw);
    if (ginterp - gin < tf * (ginterp + gin)) a.raw[(size_t)rr * a.pitch + cc] = 0.5f * (ginterp + gin);
}

}  // namespace

int art_green_equilibrate_global_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch, int border)
{
    if (H - 2 * border <= 0 || W - 2 * border <= 0) return ART_HP_OK;
    int rc = art_reserve(ctx, ctx->d_small2, ((size_t)H + 8) * sizeof(double));
    if (rc) return rc;
    double* rowsum = (double*)ctx->d_small2.p;
    double* corr = rowsum + H;
    cudaStream_t st = ctx->stream;
    const int rows = H - 2 * border;
    art_prof_begin(ctx, "k_ge_rowsums");
    k_ge_rowsums<<<(rows + 3) / 4, 128, 0, st>>>(raw, pitch, W, H, filters, border, rowsum);
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_ge_corr");
    k_ge_corr<<<1, 32, 0, st>>>(rowsum, W, H, filters, border, corr);
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_ge_scale");
    k_ge_scale<<<dim3(((W - 2 * border + 1) / 2 + 255) / 256, rows), 256, 0, st>>>(raw, pitch, W, H, filters, border, corr);
    art_prof_end(ctx);
    ctx->launches += 3;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_green_equilibrate_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch, float thresh, const float* thresh_map, size_t map_pitch)
{
    if (H < 9 || W < 12) return ART_HP_OK;          // no site inside the 4-row / 6-column margins (the loops are empty)
    int rc = art_reserve(ctx, ctx->d_work, pitch * (size_t)H * sizeof(float));
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemcpyAsync(ctx->d_work.p, raw, pitch * (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    GeArgs a;
    a.src = (const float*)ctx->d_work.p; a.raw = raw; a.sp = pitch; a.pitch = pitch; a.W = W; a.H = H; a.filters = filters;
    a.thresh = thresh; a.tmap = thresh_map; a.tp = map_pitch;
    art_prof_begin(ctx, "k_ge_local");
    k_ge_local<<<dim3((W / 2 + 255) / 256, H - 8), 256, 0, st>>>(a);
    art_prof_end(ctx);
    ctx->launches += 1;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
