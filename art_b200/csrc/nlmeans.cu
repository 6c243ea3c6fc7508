// NL-means smoothing and the detail mask for sm_100a.
//
// Replaces (reference) rtengine/nlmeans.cc NLMeans L50-280, rtengine/FTblockDN.cc laplacian L1366-1403 and
// detail_mask L1408-1476, rtengine/rescale.h rescaleBilinear L27-77.  Bit-exact with the reference's SSE2 build:
//   * the fp32 integral image of every (tile, shift) is built with the reference's own recurrence
//     St[y][x] = (St[y][x-1] + St[y-1][x]) - (St[y-1][x-1] - score) -- no reassociation -- as a wavefront: one
//     thread per tile row, row y one column behind row y-1, the value handed down with a warp shuffle (shared
//     memory at warp seams), the whole 150x150 tile kept in shared memory for the patch-distance lookups;
//   * weights accumulate per pixel in the reference's shift order (each pixel is owned by one thread for the whole
//     tile, so the global-memory read-modify-write of SW/dst needs no atomics);
//   * pixels the reference handles in 4-wide SSE groups use the vector LUT interpolation (rtengine/LUT.h
//     L349-377), the trailing pixels of each tile row the scalar one (L437-459);
//   * the reference runs its tile loop with MXCSR.FTZ set (L158-159): results of arithmetic that are subnormal
//     become signed zeros while subnormal *inputs* are honoured.  fz() below does exactly that after every
//     arithmetic instruction of the tile loop (the kernels are compiled without -ftz, which would also flush inputs).
// Compiled with -fmad=false.
#include <cmath>
#include "ctx.h"
#include "sleef_dev.cuh"

namespace {

__device__ __forceinline__ float fz(float r)
{
    return fabsf(r) < 1.17549435e-38f ? __int_as_float(__float_as_int(r) & 0x80000000) : r;
}
__device__ __forceinline__ float lim(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }

// ------------------------------------------------------------------ rescaleBilinear (+ fused point ops of detail_mask)
// POST 0: plain; 1: xlin2log(v / p0, 50) (L1424-1428); 2: scurve(LIM01(v + p0)) (L1432-1447)
template <int POST>
__global__ void __launch_bounds__(256) k_rescale(const float* __restrict__ src, size_t sp, int Ws, int Hs,
                                                 float* __restrict__ dst, size_t dp, int Wd, int Hd, float p0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= Wd) return;
    const float col_scale = (float)Ws / (float)Wd;
    const float row_scale = (float)Hs / (float)Hd;
    const float fy = y * row_scale, fx = x * col_scale;
    const int xi = min((int)fx, Ws - 1), yi = min((int)fy, Hs - 1);
    const float xf = fx - xi, yf = fy - yi;
    const int xi1 = min(xi + 1, Ws - 1), yi1 = min(yi + 1, Hs - 1);
    const float bl = src[(size_t)yi * sp + xi], br = src[(size_t)yi * sp + xi1];
    const float tl = src[(size_t)yi1 * sp + xi], tr = src[(size_t)yi1 * sp + xi1];
    const float b = xf * br + (1.f - xf) * bl;
    const float t = xf * tr + (1.f - xf) * tl;
    float v = yf * t + (1.f - yf) * b;
    if (POST == 1) v = sleef::xlin2log_scalar(v / p0, 50.f);
    if (POST == 2) v = sleef::xlin2log_scalar(sleef::pow_F_scalar(lim(v + p0, 0.f, 1.f), 2.23f), 101.f);
    dst[(size_t)y * dp + x] = v;
}

__global__ void __launch_bounds__(256) k_laplacian(const float* __restrict__ src, float* __restrict__ dst, int W, int H,
                                                   float threshold, float ceiling, float factor)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float f = factor / ceiling;
    const int n = (y - 1 < 0) ? y + 1 : y - 1, s = (y + 1 >= H) ? y - 1 : y + 1;
    const int w = (x - 1 < 0) ? x + 1 : x - 1, e = (x + 1 >= W) ? x - 1 : x + 1;
#define G(yy, xx) fmaxf(src[(size_t)(yy) * W + (xx)], 0.f)
    const float v = -8.f * G(y, x) + G(n, x) + G(s, x) + G(y, w) + G(y, e) + G(n, w) + G(n, e) + G(s, w) + G(s, e);
#undef G
    dst[(size_t)y * W + x] = lim(fabsf(v) - threshold, 0.f, ceiling) * f;
}

__global__ void __launch_bounds__(256) k_fill(float* dst, size_t dp, int W, float v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < W) dst[(size_t)blockIdx.y * dp + x] = v;
}

// ------------------------------------------------------------------ NL-means set-up kernels
__global__ void __launch_bounds__(256) k_nlm_pad(const float* __restrict__ img, size_t ip, int W, int H,
                                                 float* __restrict__ src, int WW, int border, float factor)
{   // L102-109, including its `y >= H` (not H + border) clamp
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= WW) return;
    const int yy = y <= border ? 0 : y >= H ? H - 1 : y - border;
    const int xx = x <= border ? 0 : x >= W ? W - 1 : x - border;
    src[(size_t)y * WW + x] = img[(size_t)yy * ip + xx] / factor;
}
__global__ void __launch_bounds__(256) k_nlm_lut(float* lut, int n, float lutfactor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lut[i] = sleef::xexpf_scalar(-((float)i * lutfactor));
}
__global__ void __launch_bounds__(256) k_nlm_maskprep(float* mask, size_t n, float h2, float lutfactor)
{   // L133-137
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = (1.f / (mask[i] * h2)) / lutfactor;
}

// ------------------------------------------------------------------ the tile kernel
constexpr int TS = 150, NT = 512, LUTSZ = 8192;
struct NlmArgs {
    const float* src; int WW, HH;
    const float* mask;           // pitch W, already (1 / (mask h2)) / lutfactor
    float* dst; size_t dp;       // the image plane, zeroed: accumulates weight * src, then the estimate
    float* SW;                   // pitch W, zeroed
    const float* lut;
    int W, H, sr, pr, border, ntx;
    float factor;
    long long* dbg;              // optional: per-phase clock64 totals of CTA 0 (A, B, C), debugging aid
};

// x86 flush-to-zero arithmetic (results flushed, inputs honoured).  When both inputs are known not to be subnormal
// (they are results of earlier flushed operations) PTX's .ftz forms are exactly that; otherwise compute without .ftz
// and flush the result (fz).  add.ftz(x, -0) is fz(x) in one instruction.
__device__ __forceinline__ float addz(float x, float y) { float r; asm("add.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float subz(float x, float y) { float r; asm("sub.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float mulz(float x, float y) { float r; asm("mul.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float fz1(float x) { return addz(x, -0.f); }

__device__ __forceinline__ float lut_scalar1(const float* __restrict__ data, float index)
{   // LUT.h L437-459 (clip below and above)
    const int idx = (int)index;
    if (index < 0.f || !(index == index)) return __ldg(data);
    if (index > (float)(LUTSZ - 2)) return __ldg(data + LUTSZ - 1);
    const float diff = fz1(index - (float)idx);
    const float p1 = __ldg(data + idx);
    const float p2 = fz1(__ldg(data + idx + 1) - p1);
    return fz1(p1 + mulz(p2, diff));
}
__device__ __forceinline__ float lut_vector1(const float* __restrict__ data, float index)
{   // LUT.h L349-377
    const int idx = (int)fminf(fmaxf(index, 0.f), (float)(LUTSZ - 2));
    const float lower = __ldg(data + idx), upper = __ldg(data + idx + 1);
    const float diff = fz1(fminf(fmaxf(index, 0.f), (float)(LUTSZ - 1)) - (float)idx);
    return addz(fz1(diff * upper), fz1(subz(1.f, diff) * lower));
}

// One CTA per tile: 16 warps, two CTAs per SM.  Per shift: (A) scores into St, (B) the integral image, (C) weights.
// (B) is a wavefront over the first 5 warps: thread t owns tile row t and is at column s - t in step s, reading the
// row above from shared memory, with a 160-thread named barrier per step (tools/wavefront_probe.cu: 85 clk/step
// against 277 for a flag-synchronised warp pipeline and 61 for a single warp walking 32-row blocks).  With virtual
// zeros above and left of the tile and score(0,0) := 0 the reference's three recurrences (first row, first column,
// interior; L194-204) are the one interior formula, bit for bit.  The other 11 warps wait at the block barrier;
// (A) and (C) use all 16.
constexpr int NB_THREADS = 5 * 32;
__global__ void __launch_bounds__(NT, 2) k_nlm_tile(NlmArgs a)
{
    extern __shared__ float St[];        // [TS][TS]
    const int step = TS - 2 * a.border;
    const int tile_y = blockIdx.x / a.ntx, tile_x = blockIdx.x - tile_y * a.ntx;
    const int start_y = tile_y * step, end_y = min(start_y + TS, a.HH), TH = end_y - start_y;
    const int start_x = tile_x * step, end_x = min(start_x + TS, a.WW), TW = end_x - start_x;
    const int IH = TH - 2 * a.border, IW = TW - 2 * a.border;
    if (IH <= 0 || IW <= 0) return;      // the reference builds the integral images of such tiles and uses none of them
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pr = a.pr;
    constexpr int NW = NT / 32, CU = 5;  // CU * 32 >= TS: column chunks per row
    long long tA = 0, tB = 0, tC = 0, c0 = 0, c1 = 0;
    const bool timing = a.dbg && blockIdx.x == 0 && t == 0;

    for (int ty = -a.sr; ty <= a.sr; ++ty) {
        for (int tx = -a.sr; tx <= a.sr; ++tx) {
            // (A) scores (L186-189)
            if (timing) c0 = clock64();
            for (int yy = warp; yy < TH; yy += NW) {
                const float* r0 = a.src + (size_t)min(max(yy + start_y, 0), a.HH - 1) * a.WW;
                const float* r1 = a.src + (size_t)min(max(yy + ty + start_y, 0), a.HH - 1) * a.WW;
                float p[CU], q[CU];
#pragma unroll
                for (int k = 0; k < CU; ++k) {
                    const int xx = lane + 32 * k;
                    p[k] = __ldcg(r0 + min(max(xx + start_x, 0), a.WW - 1));
                    q[k] = __ldcg(r1 + min(max(xx + tx + start_x, 0), a.WW - 1));
                }
#pragma unroll
                for (int k = 0; k < CU; ++k) {
                    const int xx = lane + 32 * k;
                    const float d = fz1(p[k] - q[k]);
                    if (xx < TW) St[yy * TS + xx] = (yy | xx) ? mulz(d, d) : 0.f;
                }
            }
            __syncthreads();
            if (timing) { c1 = clock64(); tA += c1 - c0; c0 = c1; }
            // (B) integral image (L194-204)
            if (t < NB_THREADS) {
                const bool rowvalid = t < TH;
                float* myrow = St + min(t, TS - 1) * TS;
                const float* above = St + max(min(t, TS - 1) - 1, 0) * TS;
                float left = 0.f, up = 0.f, upl;
                float sc = myrow[0];
                const int nsteps = TW + TH - 1;
                for (int s = 0; s < nsteps; ++s) {
                    const int xx = s - t;
                    const bool act = rowvalid && (unsigned)xx < (unsigned)TW;
                    upl = up;
                    up = (act && t > 0) ? above[xx] : 0.f;
                    const float v = subz(addz(left, up), subz(upl, sc));
                    if (act) { myrow[xx] = v; left = v; }
                    if (rowvalid && (unsigned)(xx + 1) < (unsigned)TW) sc = myrow[xx + 1];
                    asm volatile("bar.sync 1, %0;" ::"n"(NB_THREADS) : "memory");
                }
            }
            __syncthreads();
            if (timing) { c1 = clock64(); tB += c1 - c0; c0 = c1; }
            // (C) weights (L207-247)
            for (int oy = warp; oy < IH; oy += NW) {
                const int y = start_y + oy, sty = a.border + oy;
                const float* srow = a.src + (size_t)(start_y + a.border + oy + ty) * a.WW + tx + start_x + a.border;
                const float* mrow = a.mask + (size_t)y * a.W + start_x;
                float* swrow = a.SW + (size_t)y * a.W + start_x;
                float* orow = a.dst + (size_t)y * a.dp + start_x;
                float m[CU], sw[CU], o[CU], sv[CU];
#pragma unroll
                for (int k = 0; k < CU; ++k) {
                    const int ox = min(lane + 32 * k, IW - 1);
                    m[k] = __ldcg(mrow + ox); sw[k] = __ldcg(swrow + ox); o[k] = __ldcg(orow + ox); sv[k] = __ldcg(srow + ox);
                }
                const float* sp = St + (sty + pr) * TS + a.border, * sm = St + (sty - pr) * TS + a.border;
#pragma unroll
                for (int k = 0; k < CU; ++k) {
                    const int ox = lane + 32 * k;
                    if (ox < IW) {
                        float dist2 = subz(subz(addz(sp[ox + pr], sm[ox - pr]), sp[ox - pr]), sm[ox + pr]);
                        dist2 = fmaxf(dist2, 0.f);
                        const float d = fz1(dist2 * m[k]);
                        const float weight = ((ox & ~3) < IW - 3) ? lut_vector1(a.lut, d) : lut_scalar1(a.lut, d);
                        __stcg(swrow + ox, fz1(sw[k] + weight));
                        __stcg(orow + ox, addz(o[k], fz1(weight * sv[k])));
                    }
                }
            }
            __syncthreads();
            if (timing) { c1 = clock64(); tC += c1 - c0; }
        }
    }
    if (timing) { a.dbg[0] = tA; a.dbg[1] = tB; a.dbg[2] = tC; }
    // final estimate (L252-273)
    for (int oy = warp; oy < IH; oy += NW) {
        const int y = start_y + oy;
        for (int ox = lane; ox < IW; ox += 32) {
            const int x = start_x + ox;
            float* o = a.dst + (size_t)y * a.dp + x;
            const float f = addz(1e-5f, __ldcg(a.SW + (size_t)y * a.W + x));
            *o = fz1(fz1(__ldcg(o) / f) * a.factor);
        }
    }
}

}  // namespace

int art_detail_mask_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* mask, size_t mp, int W, int H,
                        float scaling, float threshold, float ceiling, float factor, int blur_type, float blur,
                        float* scratch /* 2 * (W/4) * (H/4) floats */)
{
    cudaStream_t st = ctx->stream;
    const dim3 gfull((W + 255) / 256, H);
    if (W < 8 || H < 8) {
        k_fill<<<gfull, 256, 0, st>>>(mask, mp, W, 1.f);
        ctx->launches += 1;
        ART_CUDA(ctx, cudaGetLastError());
        return ART_HP_OK;
    }
    const int W4 = W / 4, H4 = H / 4;
    float* L2 = scratch;
    float* m2 = scratch + (size_t)W4 * H4;
    const dim3 gq((W4 + 255) / 256, H4);
    art_prof_begin(ctx, "k_detail_mask");
    k_rescale<1><<<gq, 256, 0, st>>>(src, sp, W, H, L2, W4, W4, H4, scaling);
    k_laplacian<<<gq, 256, 0, st>>>(L2, m2, W4, H4, threshold / scaling, ceiling / scaling, factor);
    k_rescale<2><<<gfull, 256, 0, st>>>(m2, W4, W4, H4, mask, mp, W, H, 1.f - factor);
    art_prof_end(ctx);
    ctx->launches += 3;
    ART_CUDA(ctx, cudaGetLastError());
    if (blur_type == 2) return art_gauss_dev(ctx, mask, mp, mask, mp, W, H, (double)blur);
    if (blur_type == 1 && (int)blur > 0)
        for (int i = 0; i < 3; ++i) { int rc = art_boxblur_dev(ctx, mask, mp, mask, mp, W, H, (int)blur); if (rc) return rc; }
    return ART_HP_OK;
}

int art_nlmeans_dev(art_hp_ctx* ctx, float* img, size_t ip, int W, int H, float normcoeff, int strength, int detail_thresh, float scale)
{
    if (!strength) return ART_HP_OK;
    if (!(scale > 0.f) || !(normcoeff > 0.f)) return ctx->fail(ART_HP_ERR_INVALID, "nlmeans: scale and normcoeff must be positive");
    const int search_radius = int(std::ceil(5.f / scale));
    const int patch_radius = int(std::ceil(2.f / scale));
    const int border = search_radius + patch_radius;
    if (2 * border >= TS) return ctx->fail(ART_HP_ERR_INVALID, "nlmeans: scale %g leaves no tile interior", (double)scale);
    const float ph = std::pow(float(strength) / 100.f, 0.9f) / 10.f / scale;
    const float h2 = ph * ph;
    const float amount = std::max(0.f, std::min(float(detail_thresh) / 100.f, 0.99f));
    const int WW = W + 2 * border, HH = H + 2 * border;
    const size_t n = (size_t)W * H, nq = (size_t)(W / 4) * (H / 4);
    const size_t floats = round_up(n, 64) * 2 + round_up((size_t)WW * HH, 64) + round_up(2 * nq, 64) + LUTSZ;
    int rc = art_reserve(ctx, ctx->d_work, floats * sizeof(float));
    if (rc) return rc;
    float* mask = (float*)ctx->d_work.p;
    float* SW = mask + round_up(n, 64);
    float* src = SW + round_up(n, 64);
    float* quarter = src + round_up((size_t)WW * HH, 64);
    float* lut = quarter + round_up(2 * nq, 64);
    cudaStream_t st = ctx->stream;

    if ((rc = art_detail_mask_dev(ctx, img, ip, mask, W, W, H, normcoeff, 1e-3f * normcoeff, normcoeff, amount, 2, 2.f / scale, quarter))) return rc;
    constexpr float lutfactor = 100.f / float(LUTSZ - 1);
    art_prof_begin(ctx, "k_nlm_setup");
    k_nlm_pad<<<dim3((WW + 255) / 256, HH), 256, 0, st>>>(img, ip, W, H, src, WW, border, normcoeff);
    k_nlm_lut<<<LUTSZ / 256, 256, 0, st>>>(lut, LUTSZ, lutfactor);
    k_nlm_maskprep<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mask, n, h2, lutfactor);
    ART_CUDA(ctx, cudaMemset2DAsync(img, ip * sizeof(float), 0, (size_t)W * sizeof(float), H, st));
    ART_CUDA(ctx, cudaMemsetAsync(SW, 0, n * sizeof(float), st));
    art_prof_end(ctx);

    const int stepsz = TS - 2 * border;
    NlmArgs a{};
    a.src = src; a.WW = WW; a.HH = HH; a.mask = mask; a.dst = img; a.dp = ip; a.SW = SW; a.lut = lut;
    a.W = W; a.H = H; a.sr = search_radius; a.pr = patch_radius; a.border = border; a.factor = normcoeff;
    a.ntx = int(std::ceil(float(WW) / stepsz));
    const int nty = int(std::ceil(float(HH) / stepsz));
    const size_t smem = (size_t)TS * TS * sizeof(float);
    if (!(ctx->attrs_set & art_hp_ctx::ATTR_NLM)) {
        ART_CUDA(ctx, cudaFuncSetAttribute(k_nlm_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attrs_set |= art_hp_ctx::ATTR_NLM;
    }
    if (getenv("ART_HP_NLM_PHASES")) {
        if (!ctx->d_nlm_dbg) ART_CUDA(ctx, cudaMalloc(&ctx->d_nlm_dbg, 3 * sizeof(long long)));
        a.dbg = (long long*)ctx->d_nlm_dbg;
    }
    art_prof_begin(ctx, "k_nlm_tile");
    k_nlm_tile<<<a.ntx * nty, NT, smem, st>>>(a);
    art_prof_end(ctx);
    ctx->launches += 4;
    ART_CUDA(ctx, cudaGetLastError());
    if (a.dbg) {
        long long h[3];
        ART_CUDA(ctx, cudaMemcpyAsync(h, a.dbg, sizeof h, cudaMemcpyDeviceToHost, st));
        ART_CUDA(ctx, cudaStreamSynchronize(st));
        fprintf(stderr, "[nlm phases, CTA 0, clocks] scores %lld  integral %lld  weights %lld\n", h[0], h[1], h[2]);
    }
    return ART_HP_OK;
}
