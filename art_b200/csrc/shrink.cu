// FTblockDN wavelet shrinkage for sm_100a.
//
// Replaces (reference rtengine/FTblockDN.cc) MadRgb L569-603, ShrinkAllL L638-726, ShrinkAllAB L729-839 and the
// drivers WaveletDenoiseAllL L1111-1167 (edge == 0) / WaveletDenoiseAllAB L1170-1221 on device-resident
// wavelet decompositions (wavelet.cu).  Everything stays in HBM: the MAD values live in a small device table
// that the shrink kernels read, so a whole denoise pass needs no host synchronisation.
//   * MAD: int32 histogram of |trunc(coeff)| (exact by construction: integer counts), hot low bins privatised in
//     shared memory, then a one-block prefix scan finds the median bin and interpolates like the reference; every
//     subband of a decomposition in one launch pair (grid.y = subband).
//   * shrink factor / apply: element-wise; coefficients the reference handles in its 4-wide SSE loops use the
//     vector xexpf (rtengine/sleefsseavx.h L1326-1345) and the vector expression association, the n % 4 tail
//     uses the scalar xexpf (rtengine/sleef.h L1247-1266) and the scalar association -- bit-exact.
//   * the local averaging is the flat boxblur (rtengine/boxblur.h L558-742) with its three column classes, fused with the
//     shrink factor (horizontal pass) and the apply step (vertical pass); all subbands of a channel run in one grid.
// Compiled with -fmad=false.
#include "ctx.h"
#include "sleef_dev.cuh"

struct WLevel { int w, h, w2, h2, skip, sub; float* band[4]; };
struct art_hp_wavelet {
    art_hp_ctx* ctx;
    int nlev, W, H, subsamp;
    WLevel lev[10];
    float* block;
    float* buf[2];
    float* coeff0;
    int consumed;
};

namespace {
using sleef::xexpf_scalar;
using sleef::xexpf_vector;

// ------------------------------------------------------------------ MAD (MadRgb, L569-603)
constexpr int NB = 65536, HOT = 4096;
// ------------------------------------------------------------------ shrink factors
struct ShArgs {
    float* c; const float* cL; const float* nv; float* sf; const float* sfd;
    const float* mad; int n; float lvlmul; float noisevar_ab; int useCCurve;
    const float* madab;
    int nv_uniform; float nv_value;      // the noise-variance map holds one value everywhere: it is passed instead of read
};

// WaveletDenoiseAll_BiShrinkAB's "simple" shrinkage of the levels below the coarsest one (L1046-1087): in place, no local averaging,
// the factor squared, mad_abr = (useNoiseCCurve ? noisevar_ab : SQR(noisevar_ab)) * madab
__global__ void __launch_bounds__(256) k_sf_AB_simple(ShArgs a)
{
    const float mad_L = a.mad[0];
    const float madab = a.madab[0];
    const float mad_abr = a.useCCurve ? a.noisevar_ab * madab : (a.noisevar_ab * a.noisevar_ab) * madab;
    const float rmadLm9 = 1.f / (mad_L * 9.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float xl = a.cL[i], xab = a.c[i];
        const float nvi = a.nv_uniform ? a.nv_value : a.nv[i];
        if ((i & ~3) < a.n - 3) {
            const float mad_ab = nvi * mad_abr;
            const float mag_ab = xab * xab;
            const float mag_L = (xl * xl) * rmadLm9;
            const float f = 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
            a.c[i] = xab * (f * f);
        } else {
            const float mag_L = xl * xl, mag_ab = xab * xab;
            const float f = 1.f - xexpf_scalar(-(mag_ab / (nvi * mad_abr)) - (mag_L / (9.f * mad_L)));
            a.c[i] = xab * (f * f);
        }
    }
}

// ------------------------------------------------------------------ batched shrinkage: every subband of a channel in one launch
// The local averaging is two running sums per subband (rows, then columns) whose fp32 association is part of the result, so a chain
// is serial -- but a channel has 15 subbands x (2732 rows | 4096 columns) of them at 45 MP, enough to fill the chip when they all run
// in ONE grid with one chain per thread:
//   k_shrink_sf the shrink factor sf (ShrinkAllL L669-684 / ShrinkAllAB L762-786), element-wise at full occupancy (two IEEE divisions
//               and one sleef exp per coefficient: the arithmetic-heavy part stays out of the latency-bound chain kernels)
//   k_shrink_h  the horizontal pass of the flat boxblur (boxblur.h L571-602), one chain per lane through shared-memory transposes
//   k_shrink_v  the vertical pass (boxblur.h L614-710, three column classes) fused with the apply step (L692-709 / L791-813): one
//               thread per column marching down the subband with eight rows of loads in flight, the blurred value never stored.
// Per coefficient: 4 (+4 L coefficient) B read, 8 B written; then 12 B read (+4 B trailing sample from L2), 4 B written.
// A job is one subband; the jobs of a batch may belong to different channels (a, b and L of a frame run in one grid: the chroma factors
// read the luminance coefficients in the first kernel, the luminance coefficients are only rewritten in the last one).
struct ShJob {
    float* c; const float* cL; const float* madL; const float* madab; float* sf; float* tmp; int W, H, rad; float lvlmul;
    const float* nv; float nv_value; float noisevar_ab; int nv_uniform, useCCurve, ab;      // the noise-variance map (or its one value), chroma parameters
};
constexpr int SH_MAXJOBS = 64;          // 2 x 3 x 8 chroma + 3 x 5 luminance subbands
struct ShBatch {
    ShJob job[SH_MAXJOBS]; int njobs;
    int unit0[SH_MAXJOBS + 1];            // first work unit of each job (prefix sums), per kernel
};
constexpr int SH_ROWS = 32, SH_RING = 64, SH_RP = SH_RING + 1, SH_SP = 36, SH_WARPS = 8;      // SH_SP: 16-byte aligned rows, conflict-free 128-bit reads by 8 lanes
constexpr size_t SH_SMEM = (size_t)SH_WARPS * SH_ROWS * (SH_RP + SH_SP) * sizeof(float);

struct SfJob { const float* c; const float* cL; const float* nv; float nv_value; bool nv_uniform, ab; };      // the fields the inner loop reads, in registers
__device__ __forceinline__ float sf_value(const SfJob& j, float levelFactor, float madab, float rmadLm9, float mad_L, size_t i, size_t n)
{
    const float nvi = j.nv_uniform ? j.nv_value : j.nv[i];
    const bool vec = (i & ~(size_t)3) + 3 < n;           // handled by a full 4-wide vector in the reference (for (i = 0; i < n - 3; i += 4))
    if (!j.ab) {
        const float eps = 0.01f;
        const float x = j.c[i];
        const float mag = x * x;
        if (vec) {
            const float mad = nvi * levelFactor;
            return mag / (mag + mad * xexpf_vector(-mag / (9.0f * mad)) + eps);
        }
        return mag / (mag + levelFactor * nvi * xexpf_scalar(-mag / (9 * levelFactor * nvi)) + eps);
    }
    const float xl = j.cL[i], xab = j.c[i];
    if (vec) {
        const float mad_ab = nvi * madab;
        const float mag_ab = xab * xab;
        const float mag_L = (xl * xl) * rmadLm9;
        return 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
    }
    const float mag_L = xl * xl, mag_ab = xab * xab;
    return (1.f - xexpf_scalar(-(mag_ab / (nvi * madab)) - (mag_L / (9.f * mad_L))));
}

// shrink factors of every subband of the batch: element-wise, grid.y = subband
__global__ void __launch_bounds__(256) k_shrink_sf(const __grid_constant__ ShBatch b)
{
    const ShJob& jb = b.job[blockIdx.y];
    const SfJob j{jb.c, jb.cL, jb.nv, jb.nv_value, jb.nv_uniform != 0, jb.ab != 0};
    float* __restrict__ sf = jb.sf;
    const size_t n = (size_t)jb.W * jb.H;
    const float mad_L = jb.madL[0];
    const float levelFactor = j.ab ? 0.f : mad_L * 5.f / jb.lvlmul;
    float madab = j.ab ? jb.madab[0] : 0.f;
    if (j.ab) madab = jb.useCCurve ? madab : madab * jb.noisevar_ab;
    const float rmadLm9 = 1.f / (mad_L * 9.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = sf_value(j, levelFactor, madab, rmadLm9, mad_L, i + k * stride, n);
#pragma unroll
        for (int k = 0; k < 4; ++k) sf[i + k * stride] = v[k];
    }
    for (; i < n; i += stride) sf[i] = sf_value(j, levelFactor, madab, rmadLm9, mad_L, i, n);
}

// horizontal pass of the flat boxblur over sf -> tmp (boxblur.h L571-602).  A warp owns SH_ROWS = 32 rows of one subband and streams along
// them in 32-column tiles: coalesced row segments into registers (the next tile's loads are in flight while this one is processed), the
// samples parked in a 64-column shared-memory ring, the per-step increments (x[n + rad] - x[n - rad - 1]) / len formed by all lanes, then
// lane r walks row r's chain -- one dependent add per step, every lane of the warp busy, four steps per 128-bit shared-memory access -- and
// the sums leave through the same transposing tile.  Units are dealt to CTAs round-robin so that a launch with fewer units than warp
// slots still spreads over every SM.
__global__ void __launch_bounds__(SH_WARPS * 32) k_shrink_h(const __grid_constant__ ShBatch b)
{
    extern __shared__ __align__(16) float shm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ring = shm + (size_t)warp * SH_ROWS * (SH_RP + SH_SP);
    float* dt = ring + SH_ROWS * SH_RP;
    const int total = b.unit0[b.njobs];
    for (int u = warp * gridDim.x + blockIdx.x; u < total; u += gridDim.x * SH_WARPS) {
        int ji = 0;
        while (u >= b.unit0[ji + 1]) ++ji;
        const ShJob& j = b.job[ji];
        const int W = j.W, H = j.H, rad = j.rad;
        const int row0 = (u - b.unit0[ji]) * SH_ROWS, nrows = min(SH_ROWS, H - row0);
        const float* __restrict__ src = j.sf + (size_t)row0 * W;
        float* __restrict__ dst = j.tmp + (size_t)row0 * W;
        const int ntiles = (W + 31) / 32;
        const bool chain = lane < nrows;
        float t = 0.f;
        int len = rad + 1;
        const float rlen = 1.f / (float)(2 * rad + 1);
        float nxt[SH_ROWS];
#pragma unroll
        for (int r = 0; r < SH_ROWS; ++r) nxt[r] = (r < nrows && lane < W) ? src[(size_t)r * W + lane] : 0.f;
        // iteration T consumes tile T and produces the output positions p = 32 T - rad + k, k = 0..31; one more iteration drains the ramp-down
        for (int T = 0; T <= ntiles; ++T) {
            const int col = 32 * T + lane;
            if (T < ntiles) {
                float cur[SH_ROWS];
#pragma unroll
                for (int r = 0; r < SH_ROWS; ++r) cur[r] = nxt[r];
                if (T + 1 < ntiles) {
                    const int ncol = col + 32;
#pragma unroll
                    for (int r = 0; r < SH_ROWS; ++r) nxt[r] = (r < nrows && ncol < W) ? src[(size_t)r * W + ncol] : 0.f;
                }
#pragma unroll
                for (int r = 0; r < SH_ROWS; ++r) ring[r * SH_RP + (col & (SH_RING - 1))] = cur[r];
                __syncwarp();
                const int tc = (col - 2 * rad - 1) & (SH_RING - 1);
#pragma unroll
                for (int r = 0; r < SH_ROWS; ++r) dt[r * SH_SP + lane] = (cur[r] - ring[r * SH_RP + tc]) * rlen;      // increment of position col - rad
            }
            __syncwarp();
            const int base = 32 * T - rad;
            if (chain) {
                const float* x = ring + lane * SH_RP;
                float* o = dt + lane * SH_SP;
                const int k_main0 = max(0, rad + 1 - base), k_main1 = min(32, W - rad - base);      // main region rad < p < W - rad
                const int k_first = max(0, -base), k_last = min(32, W - base);
                for (int k = k_first; k < min(k_main0, k_last); ++k) {      // p <= rad: the first sample and the ramp-up
                    const int p = base + k;
                    if (p == 0) {
                        t = x[0];
                        for (int q = 1; q <= rad; q++) t += x[q];
                        t = t / len;
                    } else {
                        t = (t * len + x[(p + rad) & (SH_RING - 1)]) / (len + 1);
                        len++;
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
                for (int k = max(max(k_main1, k_main0), k_first); k < k_last; ++k) {        // p >= W - rad: the ramp-down
                    const int p = base + k;
                    t = (t * len - x[(p - rad - 1) & (SH_RING - 1)]) / (len - 1);
                    len--;
                    o[k] = t;
                }
            }
            __syncwarp();
            const int p = base + lane;
            if (p >= 0 && p < W) {
#pragma unroll
                for (int r = 0; r < SH_ROWS; ++r) if (r < nrows) dst[(size_t)r * W + p] = dt[r * SH_SP + lane];
            }
            __syncwarp();
        }
    }
}

constexpr int SV_THREADS = 128, SV_UNROLL = 8;
__global__ void __launch_bounds__(SV_THREADS) k_shrink_v(const __grid_constant__ ShBatch b)
{
    const int total = b.unit0[b.njobs];        // units of 32 columns
    const int lane = threadIdx.x & 31;
    for (int u = blockIdx.x * (SV_THREADS / 32) + (threadIdx.x >> 5); u < total; u += gridDim.x * (SV_THREADS / 32)) {
        int ji = 0;
        while (u >= b.unit0[ji + 1]) ++ji;
        const ShJob& j = b.job[ji];
        const int W = j.W, H = j.H, rad = j.rad;
        const int col = (u - b.unit0[ji]) * 32 + lane;
        if (col >= W) continue;
        const size_t n = (size_t)W * H;
        const float* __restrict__ x = j.tmp + col;
        const float* __restrict__ sfp = j.sf + col;
        float* __restrict__ cp = j.c + col;
        const bool scalar_class = col >= W - (W % 4);       // boxblur.h L614-710: the W % 4 tail columns divide, the vector columns multiply by 1 / len
        const float eps = 0.01f;
        auto emit = [&](int row, float d, float s, float xc) {
            const size_t i = (size_t)row * W + col;
            cp[(size_t)row * W] = ((i & ~(size_t)3) + 3 < n) ? xc * (d * d + s * s) / (d + s + eps) : xc * ((d * d + s * s) / (d + s + eps));
        };
        float t;
        const float full = (float)(2 * rad + 1);
        // first row and the ramp-up
        if (!scalar_class) {
            float len = (float)(rad + 1);
            t = x[0];
            for (int i = 1; i <= rad; i++) t = t + x[(size_t)i * W];
            t = t / len;
            emit(0, t, sfp[0], cp[0]);
            for (int st = 1; st <= rad; st++) {
                const float lp1 = len + 1.f;
                t = (t * len + x[(size_t)(st + rad) * W]) / lp1;
                emit(st, t, sfp[(size_t)st * W], cp[(size_t)st * W]);
                len = lp1;
            }
        } else {
            int len = rad + 1;
            t = x[0] / len;
            for (int i = 1; i <= rad; i++) t += x[(size_t)i * W] / len;
            emit(0, t, sfp[0], cp[0]);
            for (int st = 1; st <= rad; st++) {
                t = (t * len + x[(size_t)(st + rad) * W]) / (len + 1);
                emit(st, t, sfp[(size_t)st * W], cp[(size_t)st * W]);
                len++;
            }
        }
        const float rlen = 1.f / full;
        const int first = rad + 1, last = H - rad;
        int st = first;
        for (; st + SV_UNROLL <= last; st += SV_UNROLL) {
            float lead[SV_UNROLL], trail[SV_UNROLL], sv[SV_UNROLL], cv[SV_UNROLL];
#pragma unroll
            for (int k = 0; k < SV_UNROLL; ++k) {
                lead[k] = x[(size_t)(st + k + rad) * W];
                trail[k] = x[(size_t)(st + k - rad - 1) * W];
                sv[k] = sfp[(size_t)(st + k) * W];
                cv[k] = cp[(size_t)(st + k) * W];
            }
#pragma unroll
            for (int k = 0; k < SV_UNROLL; ++k) {
                const float diff = lead[k] - trail[k];
                t = t + (scalar_class ? diff / full : diff * rlen);
                emit(st + k, t, sv[k], cv[k]);
            }
        }
        for (; st < last; ++st) {
            const float diff = x[(size_t)(st + rad) * W] - x[(size_t)(st - rad - 1) * W];
            t = t + (scalar_class ? diff / full : diff * rlen);
            emit(st, t, sfp[(size_t)st * W], cp[(size_t)st * W]);
        }
        if (!scalar_class) {
            float len = full;
            for (st = max(last, first); st < H; st++) {
                const float lm1 = len - 1.f;
                t = (t * len - x[(size_t)(st - rad - 1) * W]) / lm1;
                emit(st, t, sfp[(size_t)st * W], cp[(size_t)st * W]);
                len = lm1;
            }
        } else {
            int len = 2 * rad + 1;
            for (st = max(last, first); st < H; st++) {
                t = (t * len - x[(size_t)(st - rad - 1) * W]) / (len - 1);
                emit(st, t, sfp[(size_t)st * W], cp[(size_t)st * W]);
                len--;
            }
        }
    }
}

// The same pass with FOUR adjacent columns per thread (128-bit loads and stores, four independent chains per thread): for batches whose
// subbands all have W % 4 == 0 and 16-byte aligned planes -- then no column is in the reference's scalar tail class and every group of four
// coefficients is one of its full vectors, so the arithmetic below is the vector branch of k_shrink_v, component by component.  A warp streams
// 512 contiguous bytes per row and array instead of 128: fewer, fatter DRAM streams for the same bytes.
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_muls(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_divs(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ __forceinline__ float shrink_apply(float xc, float d, float s) { return xc * (d * d + s * s) / (d + s + 0.01f); }
__device__ __forceinline__ float4 f4_apply(float4 xc, float4 d, float4 s)
{
    return make_float4(shrink_apply(xc.x, d.x, s.x), shrink_apply(xc.y, d.y, s.y), shrink_apply(xc.z, d.z, s.z), shrink_apply(xc.w, d.w, s.w));
}
constexpr int SV4_UNROLL = 4;
__global__ void __launch_bounds__(SV_THREADS) k_shrink_v4(const __grid_constant__ ShBatch b)
{
    const int total = b.unit0[b.njobs];        // units of 128 columns
    const int lane = threadIdx.x & 31;
    for (int u = blockIdx.x * (SV_THREADS / 32) + (threadIdx.x >> 5); u < total; u += gridDim.x * (SV_THREADS / 32)) {
        int ji = 0;
        while (u >= b.unit0[ji + 1]) ++ji;
        const ShJob& j = b.job[ji];
        const int W = j.W, H = j.H, rad = j.rad;
        const int col = (u - b.unit0[ji]) * 128 + lane * 4;
        if (col >= W) continue;
        const size_t W4 = (size_t)(W >> 2);           // row stride in float4
        const float4* __restrict__ x = reinterpret_cast<const float4*>(j.tmp + col);
        const float4* __restrict__ sfp = reinterpret_cast<const float4*>(j.sf + col);
        float4* __restrict__ cp = reinterpret_cast<float4*>(j.c + col);
        const float full = (float)(2 * rad + 1);
        float4 t;
        {   // first row and the ramp-up (boxblur.h L614-640, vector columns)
            float len = (float)(rad + 1);
            t = x[0];
            for (int i = 1; i <= rad; i++) t = f4_add(t, x[(size_t)i * W4]);
            t = f4_divs(t, len);
            cp[0] = f4_apply(cp[0], t, sfp[0]);
            for (int st = 1; st <= rad; st++) {
                const float lp1 = len + 1.f;
                t = f4_divs(f4_add(f4_muls(t, len), x[(size_t)(st + rad) * W4]), lp1);
                cp[(size_t)st * W4] = f4_apply(cp[(size_t)st * W4], t, sfp[(size_t)st * W4]);
                len = lp1;
            }
        }
        const float rlen = 1.f / full;
        const int first = rad + 1, last = H - rad;
        int st = first;
        for (; st + SV4_UNROLL <= last; st += SV4_UNROLL) {
            float4 lead[SV4_UNROLL], trail[SV4_UNROLL], sv[SV4_UNROLL], cv[SV4_UNROLL];
#pragma unroll
            for (int k = 0; k < SV4_UNROLL; ++k) {
                lead[k] = x[(size_t)(st + k + rad) * W4];
                trail[k] = x[(size_t)(st + k - rad - 1) * W4];
                sv[k] = sfp[(size_t)(st + k) * W4];
                cv[k] = cp[(size_t)(st + k) * W4];
            }
#pragma unroll
            for (int k = 0; k < SV4_UNROLL; ++k) {
                t = f4_add(t, f4_muls(f4_sub(lead[k], trail[k]), rlen));
                cp[(size_t)(st + k) * W4] = f4_apply(cv[k], t, sv[k]);
            }
        }
        for (; st < last; ++st) {
            t = f4_add(t, f4_muls(f4_sub(x[(size_t)(st + rad) * W4], x[(size_t)(st - rad - 1) * W4]), rlen));
            cp[(size_t)st * W4] = f4_apply(cp[(size_t)st * W4], t, sfp[(size_t)st * W4]);
        }
        float len = full;
        for (st = max(last, first); st < H; st++) {      // the ramp-down
            const float lm1 = len - 1.f;
            t = f4_divs(f4_sub(f4_muls(t, len), x[(size_t)(st - rad - 1) * W4]), lm1);
            cp[(size_t)st * W4] = f4_apply(cp[(size_t)st * W4], t, sfp[(size_t)st * W4]);
            len = lm1;
        }
    }
}

// MadRgb of several subbands at once: grid.y = subband
struct MadBatch { const float* band[SH_MAXJOBS]; int n[SH_MAXJOBS]; int w[SH_MAXJOBS]; float* out[SH_MAXJOBS]; int njobs, square; };      // w = row length of the subband
__global__ void __launch_bounds__(512) k_mad_hist_all(const __grid_constant__ MadBatch mb, int* __restrict__ histo)
{
    __shared__ int hot[HOT];
    const int jb = blockIdx.y;
    const float* __restrict__ data = mb.band[jb];
    const size_t n = (size_t)mb.n[jb];
    int* __restrict__ h = histo + (size_t)jb * NB;
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) hot[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int v = abs(__float2int_rz(data[i]));      // abs(static_cast<int>(x))
        v = v < 65535 ? v : 65535;
        if (v < HOT) atomicAdd(&hot[v], 1); else atomicAdd(&h[v], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) if (hot[i]) atomicAdd(&h[i], hot[i]);
}
__global__ void __launch_bounds__(1024) k_mad_median_all(const __grid_constant__ MadBatch mb, const int* __restrict__ histo_all)
{
    __shared__ int part[1024];
    const int jb = blockIdx.x;
    const int* __restrict__ histo = histo_all + (size_t)jb * NB;
    const int n = mb.n[jb];
    const int t = threadIdx.x;
    int s = 0;
    for (int k = 0; k < NB / 1024; ++k) s += histo[t * (NB / 1024) + k];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        float r = 0.f;
        if (n > 1) {
            const int half = n / 2;
            int count = 0, seg = 0;
            while (seg < 1024 && count + part[seg] < half) { count += part[seg]; ++seg; }
            int median = seg * (NB / 1024);
            while (count < half) { count += histo[median]; ++median; }
            const int count_ = count - histo[median - 1];
            r = (float)((double)((median - 1) + (half - count_) / ((float)(count - count_))) / 0.6745);
        }
        mb.out[jb][0] = mb.square ? r * r : r;
    }
}

int blur_radius(int level, double scale) { const int r = (int)((level + 2) / scale); return r > 1 ? r : 1; }

// scratch of the batched path: [njobs histograms][njobs x (sf, tmp) planes]
struct BatchScratch { int* histo; float* planes; size_t np; };
int batch_scratch(art_hp_ctx* ctx, int njobs, size_t n, bool planes, BatchScratch& s)
{
    s.np = round_up(n, 64);
    const size_t hb = round_up((size_t)njobs * NB * sizeof(int), 256);
    int rc = art_reserve(ctx, ctx->d_scratch, hb + (planes ? 2 * (size_t)njobs * s.np * sizeof(float) : 0));
    if (rc) return rc;
    s.histo = (int*)ctx->d_scratch.p;
    s.planes = (float*)((char*)ctx->d_scratch.p + hb);
    return ART_HP_OK;
}

int mad_batch(art_hp_ctx* ctx, const MadBatch& mb, int* histo)
{
    if (mb.njobs == 0) return ART_HP_OK;
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemsetAsync(histo, 0, (size_t)mb.njobs * NB * sizeof(int), st));
    // One frame across GPUs: MadRgb is a statistic of the whole subband.  The int32 histograms are exact, so each rank counts the
    // coefficient rows it owns, the histograms are summed over the ranks and every rank takes the median of the frame's counts:
    // the same bits as the single-GPU frame wherever the coefficients are.  (Subbands are decimated once: subband row = image row / 2.)
    MadBatch hist = mb, med = mb;
    if (ctx->band.active) {
        for (int j = 0; j < mb.njobs; ++j) {
            const int r0 = ctx->band.own0 >> 1, r1 = (ctx->band.own1 + 1) >> 1;
            hist.band[j] = mb.band[j] + (size_t)r0 * mb.w[j];
            hist.n[j] = (r1 - r0) * mb.w[j];
            med.n[j] = ((ctx->band.H_full + 1) >> 1) * mb.w[j];
        }
    }
    art_prof_begin(ctx, "k_mad_hist_all");
    k_mad_hist_all<<<dim3(std::max(1, 148 * 4 / mb.njobs), mb.njobs), 512, 0, st>>>(hist, histo);
    art_prof_end(ctx);
    if (ctx->band.active) {
        art_prof_begin(ctx, "allreduce_mad_hist");
        const int arc = art_allreduce_i32(ctx, histo, (size_t)mb.njobs * NB);
        art_prof_end(ctx);
        if (arc) return arc;
    }
    art_prof_begin(ctx, "k_mad_median_all");
    k_mad_median_all<<<mb.njobs, 1024, 0, st>>>(med, histo);
    art_prof_end(ctx);
    ctx->launches += 2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int shrink_batch(art_hp_ctx* ctx, ShBatch& b)
{
    if (b.njobs == 0) return ART_HP_OK;
    cudaStream_t st = ctx->stream;
    if (!(ctx->attrs_set & art_hp_ctx::ATTR_SHRINK)) {
        ART_CUDA(ctx, cudaFuncSetAttribute(k_shrink_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH_SMEM));
        ctx->attrs_set |= art_hp_ctx::ATTR_SHRINK;
    }
    b.unit0[0] = 0;
    for (int i = 0; i < b.njobs; ++i) {
        const ShJob& j = b.job[i];
        if (j.rad > 15) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "blur radius %d above 15 (the horizontal pass keeps 2 rad + 1 <= 32 samples behind the tile)", j.rad);
        if (2 * j.rad + 1 > j.W || 2 * j.rad + 1 > j.H) return ctx->fail(ART_HP_ERR_INVALID, "blur radius %d does not fit a %dx%d subband", j.rad, j.W, j.H);
        b.unit0[i + 1] = b.unit0[i] + (j.H + SH_ROWS - 1) / SH_ROWS;
    }
    {
        size_t nmax = 0;
        for (int i = 0; i < b.njobs; ++i) nmax = std::max(nmax, (size_t)b.job[i].W * b.job[i].H);
        const int gx = (int)std::min<size_t>((nmax + 1023) / 1024, (size_t)std::max(1, ctx->sm_count * 8 / b.njobs));
        art_prof_begin(ctx, "k_shrink_sf");
        k_shrink_sf<<<dim3(gx, b.njobs), 256, 0, st>>>(b);
        art_prof_end(ctx);
    }
    art_prof_begin(ctx, "k_shrink_h");
    k_shrink_h<<<std::min(b.unit0[b.njobs], 2 * ctx->sm_count), SH_WARPS * 32, SH_SMEM, st>>>(b);     // 101 KB of shared memory: two CTAs per SM; units dealt round-robin over the CTAs
    art_prof_end(ctx);
    // four columns per thread when every subband allows it (W % 4 == 0, 16-byte aligned planes); ART_HP_SHRINK_V4 = 0 keeps the one-column kernel
    static const int use_v4 = [] { const char* e = getenv("ART_HP_SHRINK_V4"); return e ? atoi(e) : 1; }();
    bool v4 = use_v4 != 0;
    for (int i = 0; i < b.njobs && v4; ++i) {
        const ShJob& j = b.job[i];
        v4 = (j.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(j.c) | reinterpret_cast<uintptr_t>(j.sf) | reinterpret_cast<uintptr_t>(j.tmp)) & 15) == 0;
    }
    if (v4) {
        for (int i = 0; i < b.njobs; ++i) b.unit0[i + 1] = b.unit0[i] + (b.job[i].W + 127) / 128;
        art_prof_begin(ctx, "k_shrink_v");
        k_shrink_v4<<<std::min((b.unit0[b.njobs] + SV_THREADS / 32 - 1) / (SV_THREADS / 32), ctx->sm_count * 16), SV_THREADS, 0, st>>>(b);
        art_prof_end(ctx);
        ctx->launches += 3;
        ART_CUDA(ctx, cudaGetLastError());
        return ART_HP_OK;
    }
    for (int i = 0; i < b.njobs; ++i) b.unit0[i + 1] = b.unit0[i] + (b.job[i].W + 31) / 32;
    art_prof_begin(ctx, "k_shrink_v");
    k_shrink_v<<<std::min((b.unit0[b.njobs] + SV_THREADS / 32 - 1) / (SV_THREADS / 32), ctx->sm_count * 16), SV_THREADS, 0, st>>>(b);      // one unit per warp up to 9.5 k units: the scheduler hands the tail out CTA by CTA
    art_prof_end(ctx);
    ctx->launches += 3;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

}  // namespace

extern "C" {

// madL[lvl][dir-1] = SQR(MadRgb(level_coeffs(lvl)[dir])) for every level (FTblockDN.cc L2311-2320) into a device table
int art_hp_wavelet_mad_dev(art_hp_ctx* ctx, const art_hp_wavelet* w, float* d_madL)
{
    if (!ctx || !w || !d_madL) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    if (3 * w->nlev > SH_MAXJOBS) return ctx->fail(ART_HP_ERR_INVALID, "too many wavelet levels (%d)", w->nlev);
    BatchScratch bs;
    int rc = batch_scratch(ctx, 3 * w->nlev, 0, false, bs);
    if (rc) return rc;
    MadBatch mb{};
    mb.square = 1;
    for (int l = 0; l < w->nlev; ++l)
        for (int d = 1; d < 4; ++d) {
            mb.band[mb.njobs] = w->lev[l].band[d]; mb.n[mb.njobs] = w->lev[l].w2 * w->lev[l].h2; mb.w[mb.njobs] = w->lev[l].w2; mb.out[mb.njobs] = d_madL + 3 * l + (d - 1);
            mb.njobs++;
        }
    return mad_batch(ctx, mb, bs.histo);
}

}  // extern "C"

// internal forms: `uniform` != nullptr says the noise-variance map is that one value everywhere (the map pointer is then unused)
int art_wavelet_denoise_L(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* uniform, const float* d_madL, double scale);
int art_wavelet_denoise_AB(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom, const float* uniform,
                           const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale, int bishrink = 0);

extern "C" {

int art_hp_wavelet_denoise_L_dev(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* d_madL, double scale)
{
    if (!d_noisevarlum) return ART_HP_ERR_INVALID;
    return art_wavelet_denoise_L(ctx, wL, d_noisevarlum, nullptr, d_madL, scale);
}

}  // extern "C"

int art_wavelet_denoise_L(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* uniform, const float* d_madL, double scale)
{
    if (!ctx || !wL || (!d_noisevarlum && !uniform) || !d_madL || !(scale > 0)) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const int maxlvl = std::min(wL->nlev, 5);      // L1115
    BatchScratch bs;
    int rc = batch_scratch(ctx, 3 * maxlvl, (size_t)wL->lev[0].w2 * wL->lev[0].h2, true, bs);
    if (rc) return rc;
    ShBatch b{};
    for (int l = 0; l < maxlvl; ++l)
        for (int d = 1; d < 4; ++d) {
            const WLevel& L = wL->lev[l];
            ShJob& j = b.job[b.njobs];
            j.c = L.band[d]; j.cL = nullptr; j.madL = d_madL + 3 * l + (d - 1); j.madab = nullptr;
            j.sf = bs.planes + (size_t)(2 * b.njobs) * bs.np; j.tmp = j.sf + bs.np;
            j.W = L.w2; j.H = L.h2; j.rad = blur_radius(l, scale); j.lvlmul = (float)(l + 1);
            j.nv = d_noisevarlum; j.nv_uniform = uniform != nullptr; j.nv_value = uniform ? *uniform : 0.f; j.ab = 0;
            b.njobs++;
        }
    return shrink_batch(ctx, b);
}

extern "C" int art_hp_wavelet_denoise_AB_dev(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom,
                                             const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale)
{
    if (!d_noisevarchrom) return ART_HP_ERR_INVALID;
    return art_wavelet_denoise_AB(ctx, wL, wab, d_noisevarchrom, nullptr, d_madL, noisevar_ab, useNoiseCCurve, autoch, scale, 0);
}

// bishrink != 0: WaveletDenoiseAll_BiShrinkAB (L976-1108) instead of WaveletDenoiseAllAB: the coarsest level goes through ShrinkAllAB
// as usual (its MAD is the same whether computed up front or now), every finer level through the simple shrinkage
int art_wavelet_denoise_AB(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom, const float* uniform,
                           const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale, int bishrink)
{
    if (!ctx || !wL || !wab || (!d_noisevarchrom && !uniform) || !d_madL || !(scale > 0)) return ART_HP_ERR_INVALID;
    if (wL->nlev != wab->nlev || wL->W != wab->W || wL->H != wab->H) return ctx->fail(ART_HP_ERR_INVALID, "L and ab decompositions differ in shape");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    if (autoch && noisevar_ab <= 0.001f) noisevar_ab = 0.02f;     // L737-739
    if (!(noisevar_ab > 0.001f)) return ART_HP_OK;                // L761 (MadRgb is computed but unused)
    if (3 * wL->nlev > SH_MAXJOBS) return ctx->fail(ART_HP_ERR_INVALID, "too many wavelet levels (%d)", wL->nlev);
    const int nj = 3 * wL->nlev;
    BatchScratch bs;
    int rc = batch_scratch(ctx, nj, (size_t)wab->lev[0].w2 * wab->lev[0].h2, true, bs);
    if (rc) return rc;
    // MadRgb of every ab subband: njobs floats kept behind the L table's 24 entries? no -- in the context's small buffer
    if ((rc = art_reserve(ctx, ctx->d_small, 64 * sizeof(float)))) return rc;
    float* madab = (float*)ctx->d_small.p;
    MadBatch mb{};
    mb.square = 1;
    for (int l = 0; l < wL->nlev; ++l)
        for (int d = 1; d < 4; ++d) {
            const WLevel& L = wab->lev[l];
            mb.band[mb.njobs] = L.band[d]; mb.n[mb.njobs] = L.w2 * L.h2; mb.w[mb.njobs] = L.w2; mb.out[mb.njobs] = madab + mb.njobs;
            mb.njobs++;
        }
    if ((rc = mad_batch(ctx, mb, bs.histo))) return rc;
    ShBatch b{};
    for (int l = 0; l < wL->nlev; ++l)
        for (int d = 1; d < 4; ++d) {
            const WLevel& L = wab->lev[l];
            const int idx = 3 * l + (d - 1);
            if (bishrink && l != wL->nlev - 1) {
                // WaveletDenoiseAll_BiShrinkAB's "simple" shrinkage of the finer levels (L1046-1087): element-wise, in place
                ShArgs a{};
                a.c = L.band[d]; a.cL = wL->lev[l].band[d]; a.nv = d_noisevarchrom; a.mad = d_madL + idx; a.n = L.w2 * L.h2;
                a.noisevar_ab = noisevar_ab; a.useCCurve = useNoiseCCurve; a.madab = madab + idx;
                a.nv_uniform = uniform != nullptr; a.nv_value = uniform ? *uniform : 0.f;
                art_prof_begin(ctx, "k_sf_AB_simple");
                k_sf_AB_simple<<<std::min((a.n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(a);
                art_prof_end(ctx);
                ctx->launches++;
                continue;
            }
            ShJob& j = b.job[b.njobs];
            j.c = L.band[d]; j.cL = wL->lev[l].band[d]; j.madL = d_madL + idx; j.madab = madab + idx;
            j.sf = bs.planes + (size_t)(2 * b.njobs) * bs.np; j.tmp = j.sf + bs.np;
            j.W = L.w2; j.H = L.h2; j.rad = blur_radius(l, scale); j.lvlmul = 0.f;
            j.nv = d_noisevarchrom; j.nv_uniform = uniform != nullptr; j.nv_value = uniform ? *uniform : 0.f; j.ab = 1;
            j.noisevar_ab = noisevar_ab; j.useCCurve = useNoiseCCurve;
            b.njobs++;
        }
    ART_CUDA(ctx, cudaGetLastError());
    return shrink_batch(ctx, b);
}

// WaveletDenoiseAllAB of both chroma channels and (with_L) WaveletDenoiseAllL of the luminance in ONE batch: the MAD of the 2 x 15 chroma
// subbands in one launch pair, then sf / horizontal / vertical over up to 45 subbands.  Same arithmetic as the per-channel calls (the chroma
// factors read the luminance coefficients in k_shrink_sf, the luminance coefficients change in k_shrink_v only); what changes is the number
// of chains in flight: 61 k column chains per channel leave an SM with 13 warps, three channels fill it.
int art_wavelet_denoise_LAB(art_hp_ctx* ctx, art_hp_wavelet* wL, art_hp_wavelet* wa, art_hp_wavelet* wb, const float* d_noisevarchrom, const float* uniform_c,
                            float noisevar_a, float noisevar_b, int useNoiseCCurve, const float* d_noisevarlum, const float* uniform_l,
                            const float* d_madL, double scale, int with_L)
{
    if (!ctx || !wL || !wa || !wb || (!d_noisevarchrom && !uniform_c) || !d_madL || !(scale > 0)) return ART_HP_ERR_INVALID;
    if (with_L && !d_noisevarlum && !uniform_l) return ART_HP_ERR_INVALID;
    art_hp_wavelet* wab[2] = {wa, wb};
    const float nvab[2] = {noisevar_a, noisevar_b};
    for (int c = 0; c < 2; ++c)
        if (wL->nlev != wab[c]->nlev || wL->W != wab[c]->W || wL->H != wab[c]->H) return ctx->fail(ART_HP_ERR_INVALID, "L and ab decompositions differ in shape");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const int maxlvlL = std::min(wL->nlev, 5);
    if (2 * 3 * wL->nlev + 3 * maxlvlL > SH_MAXJOBS) return ctx->fail(ART_HP_ERR_INVALID, "too many wavelet levels (%d)", wL->nlev);
    BatchScratch bs;
    int rc = batch_scratch(ctx, 2 * 3 * wL->nlev + 3 * maxlvlL, (size_t)wL->lev[0].w2 * wL->lev[0].h2, true, bs);
    if (rc) return rc;
    if ((rc = art_reserve(ctx, ctx->d_small, 64 * sizeof(float)))) return rc;
    float* madab = (float*)ctx->d_small.p;
    MadBatch mb{};
    mb.square = 1;
    ShBatch b{};
    for (int c = 0; c < 2; ++c) {
        if (!(nvab[c] > 0.001f)) continue;                // L761
        for (int l = 0; l < wL->nlev; ++l)
            for (int d = 1; d < 4; ++d) {
                const WLevel& L = wab[c]->lev[l];
                const int idx = 3 * l + (d - 1);
                float* mslot = madab + mb.njobs;
                mb.band[mb.njobs] = L.band[d]; mb.n[mb.njobs] = L.w2 * L.h2; mb.w[mb.njobs] = L.w2; mb.out[mb.njobs] = mslot;
                mb.njobs++;
                ShJob& j = b.job[b.njobs];
                j.c = L.band[d]; j.cL = wL->lev[l].band[d]; j.madL = d_madL + idx; j.madab = mslot;
                j.W = L.w2; j.H = L.h2; j.rad = blur_radius(l, scale); j.lvlmul = 0.f;
                j.nv = d_noisevarchrom; j.nv_uniform = uniform_c != nullptr; j.nv_value = uniform_c ? *uniform_c : 0.f; j.ab = 1;
                j.noisevar_ab = nvab[c]; j.useCCurve = useNoiseCCurve;
                b.njobs++;
            }
    }
    if (with_L)
        for (int l = 0; l < maxlvlL; ++l)
            for (int d = 1; d < 4; ++d) {
                const WLevel& L = wL->lev[l];
                ShJob& j = b.job[b.njobs];
                j.c = L.band[d]; j.cL = nullptr; j.madL = d_madL + 3 * l + (d - 1); j.madab = nullptr;
                j.W = L.w2; j.H = L.h2; j.rad = blur_radius(l, scale); j.lvlmul = (float)(l + 1);
                j.nv = d_noisevarlum; j.nv_uniform = uniform_l != nullptr; j.nv_value = uniform_l ? *uniform_l : 0.f; j.ab = 0;
                b.njobs++;
            }
    for (int i = 0; i < b.njobs; ++i) { b.job[i].sf = bs.planes + (size_t)(2 * i) * bs.np; b.job[i].tmp = b.job[i].sf + bs.np; }
    if ((rc = mad_batch(ctx, mb, bs.histo))) return rc;
    return shrink_batch(ctx, b);
}
