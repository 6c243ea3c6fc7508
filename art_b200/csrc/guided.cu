// Float box blur and the fast guided filter for sm_100a.
//
// art_boxblur_dev replaces rtengine::boxblur(float** src, float** dst, int radius, int W, int H, bool)
// (reference rtengine/boxblur.h L318-556): a running mean along rows (dividing by the window length) then
// along columns (multiplying by 1/len in the steady state).  The running mean is a serial float recurrence
// per line and is evaluated in the reference's order (bit-exact); the parallelism is across lines.  Both
// passes run out of place through a scratch plane, which removes the reference's ring buffers: the sample
// leaving the window is simply read from the untouched source plane.
// art_guided_dev replaces rtengine::guidedFilter (rtengine/guidedfilter.cc L80-241).
// Compiled with -fmad=false.
#include "ctx.h"

namespace {

// x2 / y2: a second plane of the same geometry filtered by the same launch (blockIdx.y = 1): the guided filter's box blurs come in independent pairs,
// and a launch has only one thread (or a quarter of a warp) per line
struct BoxArgs { const float* x; size_t xp; float* y; size_t yp; int W, H, radius; const float* x2; float* y2; };

// one line of n samples: x[k*xs] -> y[k*ys], x and y distinct planes
__device__ __forceinline__ void box_line(const float* x, size_t xs, float* y, size_t ys, int n, int radius, bool use_rlen)
{
    float len = radius + 1;
    float t = x[0];
    for (int j = 1; j <= radius; j++) t += x[j * xs];
    t /= len;
    y[0] = t;
    for (int c = 1; c <= radius; c++) {
        t = (t * len + x[(size_t)(c + radius) * xs]) / (len + 1);
        y[c * ys] = t;
        ++len;
    }
    const float rlen = 1.f / len;
    // steady state, loads software-pipelined 32 samples ahead: a launch has one thread per line (a few thousand threads), so the memory
    // latency of a step is only hidden by the loads the thread itself keeps in flight
    int c = radius + 1;
    const int cend = n - radius;
    constexpr int PF = 32;
    for (; c + PF - 1 < cend; c += PF) {
        float in[PF], old[PF];
        #pragma unroll
        for (int k = 0; k < PF; ++k) { in[k] = x[(size_t)(c + k + radius) * xs]; old[k] = x[(size_t)(c + k - radius - 1) * xs]; }
        #pragma unroll
        for (int k = 0; k < PF; ++k) {
            t = use_rlen ? t + (in[k] - old[k]) * rlen : t + (in[k] - old[k]) / len;
            y[(size_t)(c + k) * ys] = t;
        }
    }
    for (; c < cend; c++) {
        const float d = x[(size_t)(c + radius) * xs] - x[(size_t)(c - radius - 1) * xs];
        t = use_rlen ? t + d * rlen : t + d / len;
        y[(size_t)c * ys] = t;
    }
    for (c = n - radius; c < n; c++) {
        t = (t * len - x[(size_t)(c - radius - 1) * xs]) / (len - 1);
        y[(size_t)c * ys] = t;
        --len;
    }
}

__global__ void __launch_bounds__(64) k_box_h(BoxArgs a)      // thread per row (L350-381)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const float* X = blockIdx.y ? a.x2 : a.x;
    float* Y = blockIdx.y ? a.y2 : a.y;
    if (r < a.H) box_line(X + (size_t)r * a.xp, 1, Y + (size_t)r * a.yp, 1, a.W, a.radius, false);
}
// The horizontal pass for radius <= 111: a warp owns BH_ROWS = 2 rows and streams along them in 32-column tiles -- coalesced row segments into
// registers (the next tile's loads in flight while this one is processed), the samples parked in a shared-memory ring of `ring` columns
// (>= 2 radius + 33), the steady-state increments (x[c + r] - x[c - r - 1]) / len formed by all lanes, then lane k < 2 walks row k's chain (one
// dependent add per step) and the means leave through the same transposing tile.  Same operations in the same order as box_line; what changes is
// that every global access is a contiguous 128-byte row segment instead of 32 rows' worth of 4-byte samples (k_box_h: 0.46 TB/s at 45 MP).
constexpr int BH_ROWS = 2, BH_WARPS = 4, BH_SP = 36;
__global__ void __launch_bounds__(BH_WARPS * 32) k_box_h_tiles(BoxArgs a, int ring)
{
    extern __shared__ __align__(16) float bh_shm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RP = ring + 4, rmask = ring - 1;
    float* rg = bh_shm + (size_t)warp * BH_ROWS * (RP + BH_SP);
    float* dt = rg + BH_ROWS * RP;
    const int W = a.W, rad = a.radius;
    const int units = (a.H + BH_ROWS - 1) / BH_ROWS;
    const int ntiles = (W + 31) / 32, niter = (W + rad + 31) / 32;
    const float flen = (float)(2 * rad + 1);
    for (int u = blockIdx.x * BH_WARPS + warp; u < units; u += gridDim.x * BH_WARPS) {
        const int row0 = u * BH_ROWS, nrows = min(BH_ROWS, a.H - row0);
        const float* __restrict__ src = (blockIdx.y ? a.x2 : a.x) + (size_t)row0 * a.xp;
        float* __restrict__ dst = (blockIdx.y ? a.y2 : a.y) + (size_t)row0 * a.yp;
        float t = 0.f, len = (float)(rad + 1);
        float nxt[BH_ROWS];
#pragma unroll
        for (int r = 0; r < BH_ROWS; ++r) nxt[r] = (r < nrows && lane < W) ? src[(size_t)r * a.xp + lane] : 0.f;
        // iteration T consumes tile T (while there is one) and produces the outputs p = 32 T - rad + k, k = 0 .. 31
        for (int T = 0; T < niter; ++T) {
            const int col = 32 * T + lane;
            if (T < ntiles) {
                float cur[BH_ROWS];
#pragma unroll
                for (int r = 0; r < BH_ROWS; ++r) cur[r] = nxt[r];
                if (T + 1 < ntiles) {
                    const int ncol = col + 32;
#pragma unroll
                    for (int r = 0; r < BH_ROWS; ++r) nxt[r] = (r < nrows && ncol < W) ? src[(size_t)r * a.xp + ncol] : 0.f;
                }
#pragma unroll
                for (int r = 0; r < BH_ROWS; ++r) rg[r * RP + (col & rmask)] = cur[r];
                __syncwarp();
                const int tc = (col - 2 * rad - 1) & rmask;
#pragma unroll
                for (int r = 0; r < BH_ROWS; ++r) dt[r * BH_SP + lane] = (cur[r] - rg[r * RP + tc]) / flen;      // increment of position col - rad
            }
            __syncwarp();
            const int base = 32 * T - rad;
            if (lane < nrows) {
                const float* x = rg + lane * RP;
                float* o = dt + lane * BH_SP;
                const int k_main0 = max(0, rad + 1 - base), k_main1 = min(32, W - rad - base);      // steady state: rad < p < W - rad
                const int k_first = max(0, -base), k_last = min(32, W - base);
                for (int k = k_first; k < min(k_main0, k_last); ++k) {      // p <= rad: the first mean and the growing window
                    const int p = base + k;
                    if (p == 0) {
                        t = x[0];
                        for (int q = 1; q <= rad; q++) t += x[q & rmask];
                        t /= len;
                    } else {
                        t = (t * len + x[(p + rad) & rmask]) / (len + 1);
                        ++len;
                    }
                    o[k] = t;
                }
                {
                    int k = max(k_main0, k_first);
                    for (; k < k_main1 && (k & 3); ++k) { t = t + o[k]; o[k] = t; }
                    for (; k + 4 <= k_main1; k += 4) {
                        float4 v = *reinterpret_cast<float4*>(o + k);
                        t = t + v.x; v.x = t;
                        t = t + v.y; v.y = t;
                        t = t + v.z; v.z = t;
                        t = t + v.w; v.w = t;
                        *reinterpret_cast<float4*>(o + k) = v;
                    }
                    for (; k < k_main1; ++k) { t = t + o[k]; o[k] = t; }
                }
                for (int k = max(max(k_main1, k_main0), k_first); k < k_last; ++k) {        // p >= W - rad: the shrinking window
                    const int p = base + k;
                    t = (t * len - x[(p - rad - 1) & rmask]) / (len - 1);
                    --len;
                    o[k] = t;
                }
            }
            __syncwarp();
            const int p = base + lane;
            if (p >= 0 && p < W) {
#pragma unroll
                for (int r = 0; r < BH_ROWS; ++r) if (r < nrows) dst[(size_t)r * a.yp + p] = dt[r * BH_SP + lane];
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(64) k_box_v(BoxArgs a)     // thread per column (L383-553)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const float* X = blockIdx.y ? a.x2 : a.x;
    float* Y = blockIdx.y ? a.y2 : a.y;
    if (c < a.W) box_line(X + c, a.xp, Y + c, a.yp, a.H, a.radius, true);
}

struct RsArgs { const float* s; size_t sp; int Ws, Hs; float* d; size_t dp; int Wd, Hd; };

__device__ __forceinline__ float bilinear(const float* s, size_t sp, int W, int H, float x, float y)
{   // getBilinearValue, rescale.h L27-51
    const int xi = min((int)x, W - 1), yi = min((int)y, H - 1);
    const float xf = x - xi, yf = y - yi;
    const int xi1 = min(xi + 1, W - 1), yi1 = min(yi + 1, H - 1);
    const float bl = s[(size_t)yi * sp + xi], br = s[(size_t)yi * sp + xi1], tl = s[(size_t)yi1 * sp + xi], tr = s[(size_t)yi1 * sp + xi1];
    const float b = xf * br + (1.f - xf) * bl;
    const float t = xf * tr + (1.f - xf) * tl;
    return yf * t + (1.f - yf) * b;
}

__global__ void __launch_bounds__(256) k_resample(RsArgs a)   // f_subsample, guidedfilter.cc L144-159
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.Wd) return;
    const bool same = (a.Ws == a.Wd && a.Hs == a.Hd);
    const float col_scale = (float)a.Ws / (float)a.Wd, row_scale = (float)a.Hs / (float)a.Hd;
    for (int y = blockIdx.y; y < a.Hd; y += gridDim.y)
        a.d[(size_t)y * a.dp + x] = same ? a.s[(size_t)y * a.sp + x] : bilinear(a.s, a.sp, a.Ws, a.Hs, x * col_scale, y * row_scale);
}

struct EwArgs { float *a, *b; const float *c, *d; size_t n; float eps; int op; };

// op 0: a = a*b (corrIp = I1*p1 into b's plane: see call site); op 1: the four chained statistics
__global__ void __launch_bounds__(256) k_guided_ew(EwArgs e)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < e.n; k += (size_t)gridDim.x * blockDim.x) {
        if (e.op == 0) {
            const float I = e.a[k], p = e.b[k];
            e.b[k] = I * p;          // corrIp (guidedfilter.cc L189)
            e.a[k] = I * I;          // corrI  (L194)
        } else {
            // a-plane holds mean(corrI), b-plane mean(corrIp); c = meanI, d = meanp
            const float meanI = e.c[k], meanp = e.d[k];
            const float varI = e.a[k] - (meanI * meanI);          // L199 SUBMUL
            const float covIp = e.b[k] - (meanI * meanp);         // L203
            const float av = covIp / (varI + e.eps);              // L207 DIVEPSILON
            e.a[k] = av;
            e.b[k] = meanp - (av * meanI);                        // L211
        }
    }
}

struct UpArgs { const float *ma, *mb; size_t mp; int w, h; const float* I; size_t ip; float* q; size_t qp; int W, H; };

__global__ void __launch_bounds__(256) k_guided_up(UpArgs a)   // L222-238
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    const float col_scale = (float)a.w / (float)a.W, row_scale = (float)a.h / (float)a.H;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const float ymrs = y * row_scale;
        a.q[(size_t)y * a.qp + x] = bilinear(a.ma, a.mp, a.w, a.h, x * col_scale, ymrs) * a.I[(size_t)y * a.ip + x] + bilinear(a.mb, a.mp, a.w, a.h, x * col_scale, ymrs);
    }
}

// src2 / dst2 / tmp2: an optional second plane with the pitches of the first
int box_planes(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, float* tmp, size_t tp, int W, int H, int radius,
               const float* src2 = nullptr, float* dst2 = nullptr, float* tmp2 = nullptr)
{
    cudaStream_t st = ctx->stream;
    const unsigned np = src2 ? 2 : 1;
    if (radius == 0) {      // boxblur.h L322-335
        if (src != dst) ART_CUDA(ctx, cudaMemcpy2DAsync(dst, dp * sizeof(float), src, sp * sizeof(float), (size_t)W * sizeof(float), H, cudaMemcpyDeviceToDevice, st));
        if (src2 && src2 != dst2) ART_CUDA(ctx, cudaMemcpy2DAsync(dst2, dp * sizeof(float), src2, sp * sizeof(float), (size_t)W * sizeof(float), H, cudaMemcpyDeviceToDevice, st));
        return ART_HP_OK;
    }
    BoxArgs h{src, sp, tmp, tp, W, H, radius, src2, tmp2};
    int ring = 64;
    while (ring < 2 * radius + 33) ring *= 2;
    if (ring <= 256 && W > 2 * radius) {
        const int units = (H + BH_ROWS - 1) / BH_ROWS;
        const size_t smem = (size_t)BH_WARPS * BH_ROWS * (ring + 4 + BH_SP) * sizeof(float);       // <= 38 KB
        art_prof_begin(ctx, "k_box_h_tiles");
        k_box_h_tiles<<<dim3((units + BH_WARPS - 1) / BH_WARPS, np), BH_WARPS * 32, smem, st>>>(h, ring);
        art_prof_end(ctx);
    } else {
        art_prof_begin(ctx, "k_box_h");
        k_box_h<<<dim3((H + 63) / 64, np), 64, 0, st>>>(h);
        art_prof_end(ctx);
    }
    BoxArgs v{tmp, tp, dst, dp, W, H, radius, tmp2, dst2};
    art_prof_begin(ctx, "k_box_v");
    k_box_v<<<dim3((W + 63) / 64, np), 64, 0, st>>>(v);
    art_prof_end(ctx);
    ctx->launches += 2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

}  // namespace

int art_boxblur_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, int radius)
{
    const size_t tp = round_up((size_t)W, 32);
    int rc = art_reserve(ctx, ctx->d_scratch, tp * (size_t)H * sizeof(float));
    if (rc) return rc;
    return box_planes(ctx, src, sp, dst, dp, (float*)ctx->d_scratch.p, tp, W, H, radius);
}

int art_guided_subsampling(int w, int h, int r)
{   // calculate_subsampling, guidedfilter.cc L58-75
    if (r == 1) return 1;
    if (std::max(w, h) <= 600) return 1;
    for (int s = 5; s > 0; --s) if (r % s == 0) return s;
    return std::max(2, std::min(r / 2, 4));
}

int art_guided_dev(art_hp_ctx* ctx, const float* guide, size_t gp, const float* src, size_t sp, float* dst, size_t dp,
                   int W, int H, int r, float epsilon, int subsampling)
{
    cudaStream_t st = ctx->stream;
    if (subsampling <= 0) subsampling = art_guided_subsampling(W, H, r);
    const int w = W / subsampling, h = H / subsampling;
    const size_t p = round_up((size_t)w, 32), n = p * (size_t)h;
    int rc = art_reserve(ctx, ctx->d_scratch, 6 * n * sizeof(float));
    if (rc) return rc;
    float* I1 = (float*)ctx->d_scratch.p;
    float *p1 = I1 + n, *meanI = I1 + 2 * n, *meanp = I1 + 3 * n, *tmp = I1 + 4 * n, *tmp2 = I1 + 5 * n;
    const dim3 sgrid((w + 255) / 256, std::min(h, 148 * 4));
    art_prof_begin(ctx, "k_resample");
    k_resample<<<sgrid, 256, 0, st>>>(RsArgs{guide, gp, W, H, I1, p, w, h});
    k_resample<<<sgrid, 256, 0, st>>>(RsArgs{src, sp, W, H, p1, p, w, h});
    art_prof_end(ctx);
    ctx->launches += 2;
    int rad = (int)((float)r / subsampling);                          // f_mean's int rad <- float r1 (L161-169)
    rad = std::max(0, std::min(rad, (std::min(w, h) - 1) / 2 - 1));
    if ((rc = box_planes(ctx, I1, p, meanI, p, tmp, p, w, h, rad, p1, meanp, tmp2))) return rc;      // meanI, meanp: one launch pair
    const int ewgrid = 148 * 8;
    art_prof_begin(ctx, "k_guided_ew");
    k_guided_ew<<<ewgrid, 256, 0, st>>>(EwArgs{I1, p1, nullptr, nullptr, n, epsilon, 0});
    art_prof_end(ctx);
    ctx->launches++;
    if ((rc = box_planes(ctx, p1, p, p1, p, tmp, p, w, h, rad, I1, I1, tmp2))) return rc;      // mean(corrIp), mean(corrI)
    art_prof_begin(ctx, "k_guided_ew");
    k_guided_ew<<<ewgrid, 256, 0, st>>>(EwArgs{I1, p1, meanI, meanp, n, epsilon, 1});
    art_prof_end(ctx);
    ctx->launches++;
    if ((rc = box_planes(ctx, I1, p, I1, p, tmp, p, w, h, rad, p1, p1, tmp2))) return rc;      // meana, meanb
    art_prof_begin(ctx, "k_guided_up");
    k_guided_up<<<dim3((W + 255) / 256, std::min(H, 148 * 4)), 256, 0, st>>>(UpArgs{I1, p1, p, w, h, guide, gp, dst, dp, W, H});
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
