// Fattal tone mapping for sm_100a: ImProcFunctions::dynamicRangeCompression -> ToneMapFattal02
// (reference rtengine/tmo_fattal02.cc L1053-1215) with tmo_fattal02 (L421-681), the Gaussian pyramid / gradient /
// attenuation helpers (L157-417), the DCT Poisson solver (L731-950) and denoise::Median_Denoise
// (rtengine/FTblockDN.cc L87-445).
//
// Everything is restated operation by operation (fp32 association, the double-promoted products at L593-594 and L902, the
// float rgbLuminance over the float TMatrix, the vector / scalar sleef lanes of the xlogf and xexpf row loops) and is bit-identical to the reference
// compiled in place, with two documented exceptions:
//   * the two 2-D REDFT00 transforms are FFTW calls in the reference (fftw3f is an external library).  Here they are an
//     fp64 mixed-radix FFT in shared memory (one CTA per row, in-place decimation in frequency, real-even unpacking),
//     rounded to float exactly where the reference's float plan stores its result;
//   * pow() of calculateFiMatrix (L397) is glibc powf in the reference; here it is the fp64 pow rounded to float.
// tests/test_fattal_gpu.py holds the result to 1e-4 relative against the oracle.
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>
#include <map>
#include <mutex>

namespace {

constexpr int NLEVELS = 7;            // tmo_fattal02.cc L543
constexpr int DIM_CAP = 1920;         // RT_dimension_cap, L147
constexpr int FFT_THREADS = 512;      // the largest CTA; the kernels stride by blockDim.x: the 8192-point axis runs 512 threads (one CTA per SM: 128 registers),
constexpr int FFT_THREADS_SMALL = 256;      // an axis with at most 384 radix-16 butterflies per fused pass 256 (two CTAs per SM overlap each other's barriers: the
                                            // solve along the 5632-point axis of the 45 MP frame went from 1.88 to 1.57 ms; the 8192-point passes lose 20 % at 256)
constexpr int MAX_STAGES = 16;

__device__ __forceinline__ float fmaxr(float a, float b) { return a < b ? b : a; }    // std::max(a, b)

// ------------------------------------------------------------------------------------------------------------------
// luminance, median, nearest + log
// ------------------------------------------------------------------------------------------------------------------
struct Ws3 { float y0, y1, y2; };

__device__ __forceinline__ float luminance(float r, float g, float b, const Ws3 ws)
{   // Color::rgbLuminance(r, g, b, TMatrix), color.h L203-207; TMatrix holds floats (iccstore.h L38): float arithmetic
    return r * ws.y0 + g * ws.y1 + b * ws.y2;
}

__global__ void k_fat_lum(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ B, size_t ip,
                          float* __restrict__ Y, int W, int H, Ws3 ws)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * ip + x;
    Y[(size_t)y * W + x] = fmaxr(luminance(R[i], G[i], B[i], ws), 1.f);               // L1086
}

__device__ __forceinline__ unsigned ord_bits(float f)
{   // order-preserving map float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_float(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// do_median_denoise<useUpperBound>, one iteration, src != dst (FTblockDN.cc L87-421): the median is a selection, so any
// exact selection gives the reference's value; here a bitwise bisection over the ordered bit patterns
__global__ void k_median(const float* __restrict__ src, size_t sp, float* __restrict__ dst, size_t dp, int W, int H,
                         int type, int border, int useUpper, float upper)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float v = src[(size_t)y * sp + x];
    float out = v;
    if (x >= border && x < W - border && y >= border && y < H - border && (!useUpper || v <= upper)) {
        unsigned a[81];
        int n = 0;
        for (int ii = -border; ii <= border; ++ii)
            for (int jj = -border; jj <= border; ++jj) {
                const int d = abs(ii) + abs(jj);
                const bool take = type == 0 ? d <= 1 : type == 2 ? d <= 2 : true;
                if (take) a[n++] = ord_bits(src[(size_t)(y + ii) * sp + x + jj]);
            }
        const int k = n / 2;
        unsigned prefix = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned cand = prefix | (1u << bit);
            int c = 0;
            for (int i = 0; i < n; ++i) c += a[i] < cand;
            if (c <= k) prefix = cand;
        }
        out = ord_float(prefix);
    }
    dst[(size_t)y * dp + x] = out;
}

// rescale_nearest(Yr, L) (L1125, rescale.h L80-106) fused with H = xlogf(L + eps) (L483-500)
__global__ void k_fat_nearest_log(const float* __restrict__ Yr, int W, int H, float* __restrict__ Hl, int w2, int h2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w2 || y >= h2) return;
    const int sy = y * H / h2, sx = x * W / w2;
    const float v = Yr[(size_t)sy * W + sx] + 1e-4f;
    Hl[(size_t)y * w2 + x] = ((x & ~3) < w2 - 3) ? sleef::xlogf_vector(v) : sleef::xlogf_scalar(v);
}

__device__ __forceinline__ float bilinear_at(const float* __restrict__ src, int Ws, int Hs, float col_scale, float row_scale, int x, int y)
{   // getBilinearValue(src, x * col_scale, y * row_scale), rescale.h L27-50
    const float fx = x * col_scale, fy = y * row_scale;
    const int xi = min((int)fx, Ws - 1), yi = min((int)fy, Hs - 1);
    const float xf = fx - xi, yf = fy - yi;
    const int xi1 = min(xi + 1, Ws - 1), yi1 = min(yi + 1, Hs - 1);
    const float bl = src[(size_t)yi * Ws + xi], br = src[(size_t)yi * Ws + xi1];
    const float tl = src[(size_t)yi1 * Ws + xi], tr = src[(size_t)yi1 * Ws + xi1];
    const float b = xf * br + (1.f - xf) * bl;
    const float t = xf * tr + (1.f - xf) * tl;
    return yf * t + (1.f - yf) * b;
}

__global__ void k_fat_bilinear(const float* __restrict__ src, int Ws, int Hs, float* __restrict__ dst, int Wd, int Hd)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= Wd || y >= Hd) return;
    dst[(size_t)y * Wd + x] = bilinear_at(src, Ws, Hs, (float)Ws / (float)Wd, (float)Hs / (float)Hd, x, y);
}

// ------------------------------------------------------------------------------------------------------------------
// pyramid helpers (images of at most 1920 x 1920)
// ------------------------------------------------------------------------------------------------------------------
// the [1 2 1]/4 separable blur of gaussianBlur (L179-247), both passes fused; Src(x, y) supplies the input sample
template <class Src>
__device__ __forceinline__ float blur3_at(const Src& I, int w, int h, int x, int y)
{
    auto T = [&](int yy) -> float {
        if (x == 0) return (3.f * I(0, yy) + I(1, yy)) * 0.25f;
        if (x == w - 1) return (3.f * I(w - 1, yy) + I(w - 2, yy)) * 0.25f;
        float t = 2.f * I(x, yy);
        t += I(x - 1, yy);
        t += I(x + 1, yy);
        return t * 0.25f;
    };
    if (y == 0) return (3.f * T(0) + T(1)) * 0.25f;
    if (y == h - 1) return (3.f * T(h - 1) + T(h - 2)) * 0.25f;
    float t = 2.f * T(y);
    t += T(y - 1);
    t += T(y + 1);
    return t * 0.25f;
}

__global__ void k_fat_blur(const float* __restrict__ in, float* __restrict__ out, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    auto I = [&](int xx, int yy) { return in[(size_t)yy * w + xx]; };
    out[(size_t)y * w + x] = (w < 3 || h < 3) ? I(x, y) : blur3_at(I, w, h, x, y);
}

__global__ void k_fat_down(const float* __restrict__ A, int aw, float* __restrict__ Bm, int w, int h)
{   // downSample, L157-177
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float p = A[(size_t)(2 * y) * aw + 2 * x];
    p += A[(size_t)(2 * y) * aw + 2 * x + 1];
    p += A[(size_t)(2 * y + 1) * aw + 2 * x];
    p += A[(size_t)(2 * y + 1) * aw + 2 * x + 1];
    Bm[(size_t)y * w + x] = p * 0.25f;
}

// calculateGradients, L285-320: G and per-block partial sums (double), reduced in a fixed order by k_fat_avg
__global__ void k_fat_grad(const float* __restrict__ Hm, float* __restrict__ G, int w, int h, float divider, double* __restrict__ partial)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    double g = 0.0;
    if (x < w && y < h) {
        const int n = (y == 0 ? 0 : y - 1), s = (y + 1 == h ? y : y + 1);
        const int wx = (x == 0 ? 0 : x - 1), e = (x + 1 == w ? x : x + 1);
        const float gx = Hm[(size_t)y * w + wx] - Hm[(size_t)y * w + e];
        const float gy = Hm[(size_t)s * w + x] - Hm[(size_t)n * w + x];
        const float v = sqrtf(gx * gx + gy * gy) / divider;
        G[(size_t)y * w + x] = v;
        g = (double)v;
    }
    __shared__ double red[256];
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    red[t] = g;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) red[t] += red[t + s];
        __syncthreads();
    }
    if (t == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}

__global__ void k_fat_avg(const double* __restrict__ partial, int n, int count, float* __restrict__ avg)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *avg = (float)(red[0] / count);
}

// one level of calculateFiMatrix (L359-417): fi[k] = blur(upSample(fi[k+1])) (or 1 at the coarsest level), then
// fi[k] *= pow((max(grad, 1e-4) + noise) / (alfa * avgGrad[k]), beta - 1) on the levels the attenuation applies to
__global__ void k_fat_fi(const float* __restrict__ coarse, int aw, int ah, const float* __restrict__ grad, float* __restrict__ out, int w, int h,
                         int top, int apply, float alfa, float beta, float noise, const float* __restrict__ avg)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float v;
    if (top) v = 1.0f;
    else {
        auto I = [&](int xx, int yy) {            // upSample, L324-341
            int ax = (int)(xx * 0.5f), ay = (int)(yy * 0.5f);
            ax = ax < aw ? ax : aw - 1;
            ay = ay < ah ? ay : ah - 1;
            return coarse[(size_t)ay * aw + ax];
        };
        v = (w < 3 || h < 3) ? I(x, y) : blur3_at(I, w, h, x, y);
    }
    if (apply) {
        const float a = alfa * *avg;
        const float gr = grad[(size_t)y * w + x];
        const float g = (gr < 1e-4f) ? (float)1e-4 : gr;
        const float value = (float)pow((double)((g + noise) / a), (double)(beta - 1.0f));
        v *= value;
    }
    out[(size_t)y * w + x] = v;
}

// attenuated gradient field and its divergence (L578-624) in one pass.  FI is either the full-size matrix or, when the
// image was capped at 1920 px, the small one sampled through rescaleBilinear on the fly (L556-566)
// The attenuation sample fi(x, y) (four reads of the small matrix and a bilinear blend when SCALED) and the log-luminance of a pixel are shared
// by the five gradient terms of its neighbours: a 32 x 8 tile evaluates each once for itself and its one-pixel rim into shared memory (1.33
// evaluations per pixel instead of 5); the gradient and divergence expressions are unchanged.
template <bool SCALED>
__global__ void __launch_bounds__(256) k_fat_div(const float* __restrict__ Hl, const float* __restrict__ FI, int fw, int fh, float* __restrict__ F, int w, int h)
{
    constexpr int TW = 32, TH = 8, SW_ = TW + 2, SH_ = TH + 2;
    __shared__ float sfi[SH_][SW_ + 1], shv[SH_][SW_ + 1];
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const float cs = (float)fw / (float)w, rs = (float)fh / (float)h;
    for (int i = threadIdx.y * TW + threadIdx.x; i < SW_ * SH_; i += TW * TH) {
        const int ly = i / SW_, lx = i - ly * SW_;
        const int gx_ = min(max(x0 - 1 + lx, 0), w - 1), gy_ = min(max(y0 - 1 + ly, 0), h - 1);      // rim cells outside the image are never read
        sfi[ly][lx] = SCALED ? bilinear_at(FI, fw, fh, cs, rs, gx_, gy_) : FI[(size_t)gy_ * w + gx_];
        shv[ly][lx] = Hl[(size_t)gy_ * w + gx_];
    }
    __syncthreads();
    if (x >= w || y >= h) return;
    auto fi = [&](int xx, int yy) -> float { return sfi[yy - y0 + 1][xx - x0 + 1]; };
    auto hv = [&](int xx, int yy) -> float { return shv[yy - y0 + 1][xx - x0 + 1]; };
    auto gx = [&](int xx, int yy) -> float {
        const int xp1 = (xx + 1 >= w ? w - 2 : xx + 1);
        return (float)((double)(hv(xp1, yy) - hv(xx, yy)) * 0.5 * (double)(fi(xp1, yy) + fi(xx, yy)));
    };
    auto gy = [&](int xx, int yy) -> float {
        const int yp1 = (yy + 1 >= h ? h - 2 : yy + 1);
        return (float)((double)(hv(xx, yp1) - hv(xx, yy)) * 0.5 * (double)(fi(xx, yp1) + fi(xx, yy)));
    };
    const float gxc = gx(x, y), gyc = gy(x, y);
    float v = gxc + gyc;
    if (x > 0) v -= gx(x - 1, y);
    if (y > 0) v -= gy(x, y - 1);
    if (x == 0) v += gxc;
    if (y == 0) v += gyc;
    F[(size_t)y * w + x] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// REDFT00 rows: fp64 FFT in shared memory
// ------------------------------------------------------------------------------------------------------------------
struct FftPlan {
    int N;                      // DCT-I of N + 1 samples = real-even DFT of length 2N = complex FFT of length N
    int nst;
    int radix[MAX_STAGES];
    int lgr[MAX_STAGES];        // log2(radix) for the power-of-two radices, -1 otherwise
    int lgM[MAX_STAGES];        // log2(M) of the stage (M = sub-transform length after the stage) when a power of two, else -1
    // execution list: one entry per trip through shared memory.  op = the stage's radix, or 16 for two consecutive radix-4 stages fused
    // (fft_stage44); oplg = log2(M) after the entry's last stage (or -1)
    int nops;
    int op[MAX_STAGES];
    int oplg[MAX_STAGES];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }

// shared-memory layout: XOR-fold swizzle of the 16-byte slot index inside aligned groups of 8, so that unit-stride,
// small-power-of-two-stride (late stages) and digit-reversed (unpacking) accesses all spread over the bank groups
__device__ __forceinline__ int swg(int i) { return ((i >> 3) ^ (i >> 6) ^ (i >> 9) ^ (i >> 12)) & 7; }      // GF(2)-linear in the bits of i
__device__ __forceinline__ int sw(int i) { return i ^ swg(i); }

template <int R, int NT>
__device__ __forceinline__ void fft_stage(double2* z, int N, int L, int lgM, const double2* __restrict__ tw)
{   // in-place decimation-in-frequency pass of radix R over sub-transforms of length L; tw[k] = exp(-i pi k / N)
    const int M = L / R;
    const int twstep = 2 * N / L;
    double2 wr[R];              // exp(-2 pi i q / R)
    if (R != 2 && R != 4) {
#pragma unroll
        for (int q = 0; q < R; ++q) wr[q] = tw[(2 * N / R) * q];
    }
    constexpr bool POW2 = (R == 2 || R == 4);
    int gq[R];                  // swizzle term of q * M: with power-of-two R and M the bit fields of blk, q and j are disjoint
    if (POW2 && lgM >= 0) {
#pragma unroll
        for (int q = 0; q < R; ++q) gq[q] = swg(q << lgM);
    }
    for (int t = threadIdx.x; t < N / R; t += NT) {
        int blk, j;
        if (lgM >= 0) { blk = t >> lgM; j = t & (M - 1); }
        else { blk = t / M; j = t - blk * M; }
        const int base = blk * L + j;
        int ad[R];
        if (POW2 && lgM >= 0) {
            const int gb = swg(base);
#pragma unroll
            for (int q = 0; q < R; ++q) ad[q] = (base + (q << lgM)) ^ (gb ^ gq[q]);
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) ad[q] = sw(base + q * M);
        }
        double2 x[R], yv[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = z[ad[q]];
        if (R == 2) {
            yv[0] = make_double2(x[0].x + x[1].x, x[0].y + x[1].y);
            yv[1] = make_double2(x[0].x - x[1].x, x[0].y - x[1].y);
        } else if (R == 4) {
            const double2 a = make_double2(x[0].x + x[2].x, x[0].y + x[2].y), b = make_double2(x[0].x - x[2].x, x[0].y - x[2].y);
            const double2 c = make_double2(x[1].x + x[3].x, x[1].y + x[3].y), d = make_double2(x[1].x - x[3].x, x[1].y - x[3].y);
            // w4 = -i: y1 = b - i d, y3 = b + i d
            yv[0] = make_double2(a.x + c.x, a.y + c.y);
            yv[2] = make_double2(a.x - c.x, a.y - c.y);
            yv[1] = make_double2(b.x + d.y, b.y - d.x);
            yv[3] = make_double2(b.x - d.y, b.y + d.x);
        } else {
#pragma unroll
            for (int pq = 0; pq < R; ++pq) {
                double2 s = x[0];
#pragma unroll
                for (int q = 1; q < R; ++q) {
                    const double2 m = cmul(x[q], wr[(pq * q) % R]);
                    s.x += m.x; s.y += m.y;
                }
                yv[pq] = s;
            }
        }
        z[ad[0]] = yv[0];
        if (M > 1) {
            const double2 w1 = tw[twstep * j];
            double2 w = w1;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                z[ad[q]] = cmul(yv[q], w);
                if (q + 1 < R) w = cmul(w, w1);
            }
        } else {
#pragma unroll
            for (int q = 1; q < R; ++q) z[ad[q]] = yv[q];
        }
    }
}

// Two consecutive radix-4 passes (sub-transform lengths L and L / 4, both powers of two) in one trip through shared memory: the
// thread owns the 16 points {base + (4 a + b) L / 16} and keeps them in registers between the passes.  Operation for
// operation what fft_stage<4> computes twice (same butterflies, same twiddle products in the same order), so the
// transform's bits do not change; the number of shared-memory passes and CTA barriers of the power-of-two part halves.
__device__ __forceinline__ void radix4(const double2 (&x)[4], double2 (&y)[4])
{
    const double2 a = make_double2(x[0].x + x[2].x, x[0].y + x[2].y), b = make_double2(x[0].x - x[2].x, x[0].y - x[2].y);
    const double2 c = make_double2(x[1].x + x[3].x, x[1].y + x[3].y), d = make_double2(x[1].x - x[3].x, x[1].y - x[3].y);
    y[0] = make_double2(a.x + c.x, a.y + c.y);
    y[2] = make_double2(a.x - c.x, a.y - c.y);
    y[1] = make_double2(b.x + d.y, b.y - d.x);
    y[3] = make_double2(b.x - d.y, b.y + d.x);
}
template <int NT>
__device__ __forceinline__ void fft_stage44(double2* z, int N, int L, int lgM16, const double2* __restrict__ tw)
{
    const int M16 = 1 << lgM16;
    const int tws1 = 2 * N / L, tws2 = 4 * tws1;
    int gq[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) gq[q] = swg(q << lgM16);
    for (int t = threadIdx.x; t < N / 16; t += NT) {
        const int j = t & (M16 - 1);
        const int base = ((t >> lgM16) << (lgM16 + 4)) + j;
        const int gb = swg(base);
        double2 x[4][4];            // [a][b]
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q >> 2][q & 3] = z[(base + (q << lgM16)) ^ (gb ^ gq[q])];
        // pass 1: radix 4 over a for each b, twiddles exp(-2 pi i d (b M16 + j) / L)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double2 in[4] = {x[0][b], x[1][b], x[2][b], x[3][b]};
            double2 y[4];
            radix4(in, y);
            const double2 w1 = tw[tws1 * ((b << lgM16) + j)];
            double2 w = w1;
            x[0][b] = y[0];
#pragma unroll
            for (int d = 1; d < 4; ++d) {
                x[d][b] = cmul(y[d], w);
                if (d + 1 < 4) w = cmul(w, w1);
            }
        }
        // pass 2: radix 4 over b for each d, twiddles exp(-2 pi i f j / (L / 4))
        const double2 w1 = tw[tws2 * j];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            double2 y[4];
            radix4(x[d], y);
            z[(base + ((4 * d) << lgM16)) ^ (gb ^ gq[4 * d])] = y[0];
            if (M16 > 1) {
                double2 w = w1;
#pragma unroll
                for (int f = 1; f < 4; ++f) {
                    z[(base + ((4 * d + f) << lgM16)) ^ (gb ^ gq[4 * d + f])] = cmul(y[f], w);
                    if (f + 1 < 4) w = cmul(w, w1);
                }
            } else {
#pragma unroll
                for (int f = 1; f < 4; ++f) z[(base + ((4 * d + f) << lgM16)) ^ (gb ^ gq[4 * d + f])] = y[f];
            }
        }
    }
}

template <int NT>
__device__ __forceinline__ void fft_inplace(double2* z, const FftPlan& P, const double2* __restrict__ tw)
{
    int L = P.N;
    for (int s = 0; s < P.nops; ++s) {
        const int r = P.op[s];
        switch (r) {
            case 2: fft_stage<2, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 3: fft_stage<3, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 4: fft_stage<4, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 5: fft_stage<5, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 7: fft_stage<7, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 11: fft_stage<11, NT>(z, P.N, L, P.oplg[s], tw); break;
            case 16: fft_stage44<NT>(z, P.N, L, P.oplg[s], tw); break;
            default: fft_stage<13, NT>(z, P.N, L, P.oplg[s], tw); break;
        }
        L /= r;
        __syncthreads();
    }
}

// where output bin k of the in-place transform sits (digit reversal over the stage radices); tabulated once per plan
__device__ __forceinline__ int fft_pos(const FftPlan& P, int k)
{
    int pos = 0, L = P.N;
    for (int s = 0; s < P.nst; ++s) {
        const int r = P.radix[s];
        int q, d, M;
        if (P.lgr[s] >= 0) { q = k >> P.lgr[s]; d = k & (r - 1); } else { q = k / r; d = k - q * r; }
        if (P.lgM[s] >= 0) { M = 1 << P.lgM[s]; pos += d << P.lgM[s]; } else { M = L / r; pos += d * M; }
        k = q;
        L = M;
    }
    return pos;
}
__global__ void k_fat_postab(int* tab, FftPlan P)
{   // tab[k] = swizzled slot of bin k for k in [0, N], bin N aliasing bin 0
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= P.N) tab[k] = sw(fft_pos(P, k == P.N ? 0 : k));
}

// pack the even extension of x[0..N] (stride 1) into z: z[m] = y[2m] + i y[2m+1], y[j] = x[j <= N ? j : 2N - j]
template <int NT, class T>
__device__ __forceinline__ void dct_pack(double2* z, const T* x, int N)
{
    for (int m = threadIdx.x; m < N; m += NT) {
        const int a = 2 * m, b = 2 * m + 1;
        z[sw(m)] = make_double2((double)x[a <= N ? a : 2 * N - a], (double)x[b <= N ? b : 2 * N - b]);
    }
    __syncthreads();
}

// Y_k of the real-even sequence from the packed transform: E_k + exp(-i pi k / N) O_k, real part
__device__ __forceinline__ double dct_unpack(const double2* z, const FftPlan& P, const double2* __restrict__ tw, const int* __restrict__ postab, int k)
{
    const double2 zk = z[postab[k]];
    const double2 zc = z[postab[P.N - k]];
    const double2 w = tw[k];
    return fma(w.y, 0.5 * (zk.x - zc.x), fma(w.x, 0.5 * (zk.y + zc.y), 0.5 * (zk.x + zc.x)));
}

// MODE 0: float rows -> double rows.            (first half of transform_normal2ev, L781-788)
// MODE 1: double rows -> the rest of solve_pde_fft along this axis: finish the forward transform, round to float as the
//         reference's float plan does, scale (L790-809), divide by the eigenvalue sums (L899-906), pre-scale for the
//         inverse transform (L742-756) and run the inverse transform along the same axis -> double rows.
// MODE 2: double rows -> float rows through xexpf with the row loop's vector / scalar lanes (L647-664).
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fat_dct(const void* in, size_t ipitch, void* out, size_t opitch, int nrows, FftPlan P,
          const double2* __restrict__ tw, const int* __restrict__ postab, const double* __restrict__ lam_k, const double* __restrict__ lam_row, float factor)
{
    extern __shared__ double2 z[];
    const int N = P.N;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        if (MODE == 0) dct_pack<NT>(z, (const float*)in + (size_t)row * ipitch, N);
        else dct_pack<NT>(z, (const double*)in + (size_t)row * ipitch, N);
        fft_inplace<NT>(z, P, tw);
        if (MODE == 0) {
            double* o = (double*)out + (size_t)row * opitch;
            for (int k = threadIdx.x; k <= N; k += NT) o[k] = dct_unpack(z, P, tw, postab, k);
        } else if (MODE == 2) {
            float* o = (float*)out + (size_t)row * opitch;
            const int width = N + 1;
            for (int k = threadIdx.x; k <= N; k += NT) {
                const float v = (float)dct_unpack(z, P, tw, postab, k);
                o[k] = ((k & ~3) < width - 3) ? sleef::xexpf_vector(v) : sleef::xexpf_scalar(v);
            }
        } else {
            // this CTA's row is image column x = row; k runs over image rows y
            double* o = (double*)out + (size_t)row * opitch;
            const bool xedge = row == 0 || row == nrows - 1;
            const double lx = lam_row[row];
            for (int k = threadIdx.x; k <= N; k += NT) {
                float t = (float)dct_unpack(z, P, tw, postab, k);
                const bool yedge = k == 0 || k == N;
                t *= factor;
                if (yedge) t *= 0.5f;
                if (xedge) t *= 0.5f;
                t = (float)((double)t / (lam_k[k] + lx));
                if (row == 0 && k == 0) t = 0.f;
                if (!yedge && !xedge) t *= 0.25f;
                else if (!(yedge && xedge)) t *= 0.5f;
                o[k] = (double)t;
            }
            __syncthreads();                 // all of z consumed, all of the row written
            dct_pack<NT>(z, (const double*)o, N);
            fft_inplace<NT>(z, P, tw);
            for (int k = threadIdx.x; k <= N; k += NT) o[k] = dct_unpack(z, P, tw, postab, k);      // z holds the whole transform: the row can be overwritten as it is unpacked
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FFT_THREADS)
k_dct_dd(const double* in, size_t ipitch, double* out, size_t opitch, int nrows, FftPlan P, const double2* __restrict__ tw, const int* __restrict__ postab)
{
    extern __shared__ double2 z[];
    const int N = P.N;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        dct_pack<FFT_THREADS>(z, in + (size_t)row * ipitch, N);
        fft_inplace<FFT_THREADS>(z, P, tw);
        double vals[(14336 + FFT_THREADS) / FFT_THREADS];
        int c = 0;
        for (int k = threadIdx.x; k <= N; k += FFT_THREADS) vals[c++] = dct_unpack(z, P, tw, postab, k);
        c = 0;
        double* o = out + (size_t)row * opitch;
        for (int k = threadIdx.x; k <= N; k += FFT_THREADS) o[k] = vals[c++];
        __syncthreads();
    }
}
__global__ void k_round(const double* __restrict__ in, size_t ip, float* __restrict__ out, int n0, int n1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < n1 && y < n0) out[(size_t)y * n1 + x] = (float)in[(size_t)y * ip + x];
}

__global__ void k_fat_transpose(const double* __restrict__ in, size_t ipitch, double* __restrict__ out, size_t opitch, int rows, int cols)
{
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * ipitch + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * opitch + r] = tile[threadIdx.x][j];
    }
}

__global__ void k_fat_twiddles(double2* tw, int N)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * N) return;
    double s, c;
    sincospi((double)k / (double)N, &s, &c);
    tw[k] = make_double2(c, -s);
}

// ------------------------------------------------------------------------------------------------------------------
// thumbnails, median / shadow statistics, final application
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_fat_thumbs(const float* __restrict__ Yr, int W, int H, const float* __restrict__ L, int w2, int h2,
                             float* __restrict__ ta, float* __restrict__ tb, int ww, int hh)
{   // rescale_nearest(Yr, tmp), rescale_nearest(L, tmp): L1150, L1158
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ww || y >= hh) return;
    ta[y * ww + x] = Yr[(size_t)(y * H / hh) * W + x * W / ww];
    tb[y * ww + x] = L[(size_t)(y * h2 / hh) * w2 + x * w2 / ww];
}

// k-th smallest (0-based) of n ordered keys: four 8-bit histogram passes, whole CTA
__device__ unsigned block_select(const float* __restrict__ v, int n, int k, unsigned* hist, unsigned* sh)
{
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned key = ord_bits(v[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned acc = 0, b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[b] > (unsigned)k) break;
                acc += hist[b];
            }
            sh[0] = b; sh[1] = acc;
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        mask |= 255u << shift;
        k -= (int)sh[1];
        __syncthreads();
    }
    return prefix;
}

// sum, in ascending order, of the oidx + 1 smallest samples (the loops at L1153-1156 and L1162-1165 over the sorted array)
__device__ float block_low_sum(const float* __restrict__ v, int n, int oidx, unsigned* hist, unsigned* sh, float* list, int cap)
{
    const unsigned tkey = block_select(v, n, oidx, hist, sh);
    if (threadIdx.x == 0) sh[2] = 0;
    for (int i = threadIdx.x; i < cap; i += blockDim.x) list[i] = __int_as_float(0x7f800000);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (ord_bits(v[i]) < tkey) list[atomicAdd(&sh[2], 1u)] = v[i];
    __syncthreads();
    const int cnt = (int)sh[2];
    for (int k = 2; k <= cap; k <<= 1)                 // bitonic sort of the samples below the threshold
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = list[i], b = list[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { list[i] = b; list[ixj] = a; }
                }
            }
            __syncthreads();
        }
    float s = 0.f;
    if (threadIdx.x == 0) {
        const float t = ord_float(tkey);
        for (int i = 0; i < cnt; ++i) s += list[i];
        for (int i = cnt; i <= oidx; ++i) s += t;
        s /= oidx;
    }
    return s;       // valid on thread 0
}

__global__ void __launch_bounds__(1024) k_fat_stats(const float* __restrict__ ta, const float* __restrict__ tb, int sz, float* __restrict__ params)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned sh[4];
    __shared__ float list[2048];
    const int idx = sz / 2;
    const int oidx = max(1, min((int)(sz * 0.05f + 0.5f), sz - 1));        // L1148
    const float oldMedian = ord_float(block_select(ta, sz, idx, hist, sh));
    const float old_min = block_low_sum(ta, sz, oidx, hist, sh, list, 2048);
    __syncthreads();
    const float newMedian = ord_float(block_select(tb, sz, idx, hist, sh));
    const float new_min = block_low_sum(tb, sz, oidx, hist, sh, list, 2048);
    if (threadIdx.x == 0) {
        params[0] = (oldMedian == 0.f || newMedian == 0.f) ? 65535.f : (oldMedian / newMedian);    // L1160
        params[1] = old_min - new_min;                                                               // L1167
    }
}

__global__ void k_fat_apply(float* __restrict__ R, float* __restrict__ G, float* __restrict__ B, size_t ip, int w, int h,
                            const float* __restrict__ Yr, const float* __restrict__ L, int w2, int h2, const float* __restrict__ params,
                            int satcontrol, Ws3 ws)
{   // L1172-1213
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float hr = (float)h2 / (float)h, wr = (float)w2 / (float)w;
    const float scale = params[0], offset = params[1];
    const int yy = min((int)(y * hr + 1), h2 - 1), xx = min((int)(x * wr + 1), w2 - 1);
    const float Y = fmaxr(Yr[(size_t)y * w + x], 1e-4f);
    const float l = fmaxr(L[(size_t)yy * w2 + xx], 1e-4f) * (scale / Y);
    const size_t i = (size_t)y * ip + x;
    float r = R[i], g = G[i], b = B[i], s = 1.f;
    if (l > 1.f) {
        r = fmaxr(r * l - offset, r);
        g = fmaxr(g * l - offset, g);
        b = fmaxr(b * l - offset, b);
        if (satcontrol) s = sleef::xexpf_scalar(0.3f * sleef::xlogf_scalar(1.f / l));      // pow_F
    } else {
        r *= l; g *= l; b *= l;
        if (satcontrol) s = sleef::xexpf_scalar(0.3f * sleef::xlogf_scalar(l));
    }
    if (satcontrol && s != 1.f) {
        const float ll = luminance(r, g, b, ws);
        const float rl = r - ll, gl = g - ll, bl = b - ll;
        r = ll + s * rl; g = ll + s * gl; b = ll + s * bl;
    }
    R[i] = r; G[i] = g; B[i] = b;
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
int find_fast_dim(int dim)
{   // L998-1050
    unsigned v = (unsigned)dim;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    const int d1 = (int)v;
    const int d[12] = {d1 / 128 * 65, d1 / 64 * 33, d1 / 512 * 273, d1 / 16 * 9, d1 / 8 * 5, d1 / 16 * 11,
                       d1 / 128 * 91, d1 / 4 * 3, d1 / 64 * 49, d1 / 16 * 13, d1 / 8 * 7, d1};
    for (int i = 0; i < 12; ++i) if (d[i] >= dim) return d[i];
    return dim;
}

bool make_plan(int N, FftPlan* P)
{
    P->N = N; P->nst = 0;
    int n = N;
    auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    const int odd[5] = {13, 11, 7, 5, 3};
    for (int f : odd)
        while (n % f == 0) { if (P->nst >= MAX_STAGES) return false; P->radix[P->nst++] = f; n /= f; }
    while (n % 4 == 0) { if (P->nst >= MAX_STAGES) return false; P->radix[P->nst++] = 4; n /= 4; }
    if (n % 2 == 0) { if (P->nst >= MAX_STAGES) return false; P->radix[P->nst++] = 2; n /= 2; }
    int L = N;
    for (int s = 0; s < P->nst; ++s) { L /= P->radix[s]; P->lgr[s] = lg(P->radix[s]); P->lgM[s] = lg(L); }
    P->nops = 0;
    for (int s = 0; s < P->nst; ++s) {
        if (P->radix[s] == 4 && s + 1 < P->nst && P->radix[s + 1] == 4 && P->lgM[s + 1] >= 0) { P->op[P->nops] = 16; P->oplg[P->nops++] = P->lgM[s + 1]; ++s; }
        else { P->op[P->nops] = P->radix[s]; P->oplg[P->nops++] = P->lgM[s]; }
    }
    return n == 1;
}

struct AxisTables { double2* tw = nullptr; double* lam = nullptr; int* postab = nullptr; };
std::mutex g_tab_mu;
std::map<std::pair<int, int>, AxisTables> g_tabs;       // (device, n) -> tables; a handful of sizes per process

int axis_tables(art_hp_ctx* ctx, int n, AxisTables* out)
{
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_pair(ctx->device, n);
    auto it = g_tabs.find(key);
    if (it != g_tabs.end()) { *out = it->second; return ART_HP_OK; }
    const int N = n - 1;
    AxisTables t;
    ART_CUDA(ctx, cudaMalloc(&t.tw, sizeof(double2) * 2 * (size_t)N));
    ART_CUDA(ctx, cudaMalloc(&t.lam, sizeof(double) * (size_t)n));
    ART_CUDA(ctx, cudaMalloc(&t.postab, sizeof(int) * (size_t)n));
    FftPlan P;
    if (!make_plan(N, &P)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "no FFT plan for length %d", N);
    k_fat_postab<<<(n + 255) / 256, 256, 0, ctx->stream>>>(t.postab, P);
    k_fat_twiddles<<<(2 * N + 255) / 256, 256, 0, ctx->stream>>>(t.tw, N);
    ctx->launches++;
    std::vector<double> lam(n);
    for (int i = 0; i < n; ++i) {            // get_lambda, L813-823
        const double s = std::sin((double)i / (2 * (n - 1)) * 3.14159265358979323846);
        lam[i] = -4.0 * (s * s);
    }
    ART_CUDA(ctx, cudaMemcpyAsync(t.lam, lam.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // lam is a local
    g_tabs[key] = t;
    *out = t;
    return ART_HP_OK;
}

inline dim3 grid2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

int art_fattal_fast_dim(int dim) { return find_fast_dim(dim); }

int art_median_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, int type, int useUpper, float upper)
{
    static const int border_of[6] = {1, 1, 2, 2, 3, 4};
    const dim3 b(32, 8);
    art_prof_begin(ctx, "k_median");
    k_median<<<grid2(W, H, b), b, 0, ctx->stream>>>(src, sp, dst, dp, W, H, type, border_of[type], useUpper, upper);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_fattal_dev(art_hp_ctx* ctx, float* R, float* G, float* B, size_t ip, int W, int H, int threshold, int amount, int satcontrol,
                   const double* ws9)
{
    cudaStream_t st = ctx->stream;
    const int detail_level = 3;                                            // L1056
    float alpha = 1.f;
    if (threshold < 0) alpha += (threshold * 0.9f) / 100.f;
    else if (threshold > 0) alpha += threshold / 100.f;
    const float beta = 1.f - (amount * 0.3f) / 100.f;
    if (alpha <= 0 || beta <= 0) return ART_HP_OK;                         // L1068-1070
    const float noise = alpha * 0.01f;
    const int w2 = find_fast_dim(W) + 1, h2 = find_fast_dim(H) + 1;
    const int N1 = w2 - 1, N0 = h2 - 1;
    FftPlan P1, P0;
    if (!make_plan(N1, &P1) || !make_plan(N0, &P0)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "no FFT plan for %d x %d", w2, h2);
    const size_t smem1 = sizeof(double2) * round_up((size_t)N1, 8), smem0 = sizeof(double2) * round_up((size_t)N0, 8);
    if (std::max(smem1, smem0) > 227u * 1024u || std::max(N1, N0) > 14336)
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "padded side %d exceeds the shared-memory transform (max 14336)", std::max(N1, N0));
    int ww, hh;
    if (W >= H) { const float ratio = 200.f / W; ww = 200; hh = (int)(ratio * H); }
    else { const float ratio = 200.f / H; hh = 200; ww = (int)(ratio * W); }
    const int sz = ww * hh;
    if (sz < 2) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "aspect ratio of %dx%d leaves an empty 200-px thumbnail", W, H);

    // geometry of the capped pyramid
    int pw[NLEVELS], ph[NLEVELS];
    int sw = w2, sh = h2;
    const bool scaled = std::max(w2, h2) > DIM_CAP;
    if (scaled) {
        const float s = (float)DIM_CAP / (float)std::max(w2, h2);
        sw = (int)((float)(size_t)w2 * s);
        sh = (int)((float)(size_t)h2 * s);
    }
    pw[0] = sw; ph[0] = sh;
    for (int k = 1; k < NLEVELS; ++k) {
        if (pw[k - 1] > 2 && ph[k - 1] > 2) { pw[k] = pw[k - 1] / 2; ph[k] = ph[k - 1] / 2; }
        else { pw[k] = pw[k - 1]; ph[k] = ph[k - 1]; }
    }
    // workspace
    const size_t n = (size_t)W * H, n2 = (size_t)w2 * h2, ns = (size_t)sw * sh;
    const size_t pa = round_up((size_t)w2, 16), pb = round_up((size_t)h2, 16);
    size_t pyr_total = 0;
    for (int k = 0; k < NLEVELS; ++k) pyr_total += round_up((size_t)pw[k] * ph[k], 64);
    const size_t nblocks0 = (size_t)((sw + 31) / 32) * ((sh + 7) / 8);
    size_t bytes = 0;
    auto add = [&bytes](size_t b) { const size_t o = bytes; bytes += round_up(b, 256); return o; };
    const size_t oY0 = add(n * 4), oYr = add(n * 4), oH = add(n2 * 4), oF = add(n2 * 4);
    const size_t oA = add(pa * h2 * 8), oB = add(pb * w2 * 8);
    const size_t oPyr = add(pyr_total * 4), oGrad = add(pyr_total * 4), oFi = add(pyr_total * 4), oTmp = add(round_up(ns, 64) * 4);
    const size_t oPart = add(nblocks0 * 8), oAvg = add(64 * 4), oTa = add((size_t)sz * 4), oTb = add((size_t)sz * 4), oPar = add(64);
    int rc = art_reserve(ctx, ctx->d_fattal, bytes);
    if (rc) return rc;
    char* base = (char*)ctx->d_fattal.p;
    float *Y0 = (float*)(base + oY0), *Yr = (float*)(base + oYr), *Hl = (float*)(base + oH), *F = (float*)(base + oF);
    double *A = (double*)(base + oA), *Bt = (double*)(base + oB);
    float *pyr0 = (float*)(base + oPyr), *grad0 = (float*)(base + oGrad), *fi0 = (float*)(base + oFi), *tmp = (float*)(base + oTmp);
    double* partial = (double*)(base + oPart);
    float *avg = (float*)(base + oAvg), *ta = (float*)(base + oTa), *tb = (float*)(base + oTb), *params = (float*)(base + oPar);
    float *pyr[NLEVELS], *grad[NLEVELS], *fi[NLEVELS];
    {
        size_t o = 0;
        for (int k = 0; k < NLEVELS; ++k) { pyr[k] = pyr0 + o; grad[k] = grad0 + o; fi[k] = fi0 + o; o += round_up((size_t)pw[k] * ph[k], 64); }
    }
    AxisTables t1, t0;
    if ((rc = axis_tables(ctx, w2, &t1))) return rc;
    if ((rc = axis_tables(ctx, h2, &t0))) return rc;

    const Ws3 ws = {(float)ws9[3], (float)ws9[4], (float)ws9[5]};
    const dim3 b(32, 8);
#define FAT_LAUNCH(name, kern, grid, block, smem, ...)              \
    do {                                                            \
        art_prof_begin(ctx, name);                                  \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);               \
        art_prof_end(ctx);                                          \
        ctx->launches++;                                            \
    } while (0)

    // luminance, shadow median (L1086-1117)
    FAT_LAUNCH("k_fat_lum", k_fat_lum, grid2(W, H, b), b, 0, R, G, B, ip, Y0, W, H, ws);
    {
        const float r = (float)std::max(W, H) / (float)DIM_CAP;
        const int med = r >= 3 ? 4 : r >= 2 ? 3 : r >= 1 ? 2 : 1;
        if ((rc = art_median_dev(ctx, Y0, W, Yr, W, W, H, med, 1, 65.535f))) return rc;
    }
    // H = log(nearest(Yr) + eps); capped copy; pyramid; gradients (L483-553)
    FAT_LAUNCH("k_fat_nearest_log", k_fat_nearest_log, grid2(w2, h2, b), b, 0, Yr, W, H, Hl, w2, h2);
    const float* Hs = Hl;
    if (scaled) {
        FAT_LAUNCH("k_fat_bilinear", k_fat_bilinear, grid2(sw, sh, b), b, 0, Hl, w2, h2, pyr[0], sw, sh);
        Hs = pyr[0];
    }
    {
        const float* level = Hs;
        for (int k = 0; k < NLEVELS; ++k) {
            if (k > 0) {
                FAT_LAUNCH("k_fat_blur", k_fat_blur, grid2(pw[k - 1], ph[k - 1], b), b, 0, level, tmp, pw[k - 1], ph[k - 1]);
                if (pw[k - 1] > 2 && ph[k - 1] > 2)
                    FAT_LAUNCH("k_fat_down", k_fat_down, grid2(pw[k], ph[k], b), b, 0, tmp, pw[k - 1], pyr[k], pw[k], ph[k]);
                else
                    ART_CUDA(ctx, cudaMemcpyAsync(pyr[k], tmp, sizeof(float) * (size_t)pw[k] * ph[k], cudaMemcpyDeviceToDevice, st));
                level = pyr[k];
            }
            const dim3 g = grid2(pw[k], ph[k], b);
            FAT_LAUNCH("k_fat_grad", k_fat_grad, g, b, 0, level, grad[k], pw[k], ph[k], (float)std::pow(2.0f, k + 1), partial);
            FAT_LAUNCH("k_fat_avg", k_fat_avg, 1, 256, 0, partial, (int)(g.x * g.y), pw[k] * ph[k], avg + k);
        }
    }
    // attenuation matrix, coarse to fine (L359-417)
    for (int k = NLEVELS - 1; k >= 0; --k) {
        const int apply = ((k >= detail_level || k == NLEVELS - 1) && beta != 1.f) ? 1 : 0;
        const int top = k == NLEVELS - 1;
        FAT_LAUNCH("k_fat_fi", k_fat_fi, grid2(pw[k], ph[k], b), b, 0, top ? nullptr : fi[k + 1], top ? 0 : pw[k + 1], top ? 0 : ph[k + 1],
                   grad[k], fi[k], pw[k], ph[k], top, apply, alpha, beta, noise, avg + k);
    }
    // divergence of the attenuated gradient field (L556-624)
    if (scaled) FAT_LAUNCH("k_fat_div", k_fat_div<true>, grid2(w2, h2, b), b, 0, Hl, fi[0], sw, sh, F, w2, h2);
    else FAT_LAUNCH("k_fat_div", k_fat_div<false>, grid2(w2, h2, b), b, 0, Hl, fi[0], sw, sh, F, w2, h2);

    // Poisson solve (L869-950) and exponentiation (L647-664)
    if (!(ctx->attrs_set & art_hp_ctx::ATTR_FATTAL)) {
        cudaFuncSetAttribute(k_fat_dct<0, FFT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_fat_dct<1, FFT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_fat_dct<2, FFT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_fat_dct<0, FFT_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        cudaFuncSetAttribute(k_fat_dct<1, FFT_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        cudaFuncSetAttribute(k_fat_dct<2, FFT_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
        ctx->attrs_set |= art_hp_ctx::ATTR_FATTAL;
    }
    auto dct_grid = [&](size_t smem, int rows) {
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200u * 1024u) / std::max<size_t>(smem, 1)));
        return std::min(rows, ctx->sm_count * per_sm);
    };
    const float factor = 1.0f / ((h2 - 1) * (w2 - 1));
    const dim3 tb32(32, 8);
    auto fft_threads = [](const FftPlan& P) { return P.N / 16 <= 384 ? FFT_THREADS_SMALL : FFT_THREADS; };
    const bool small1 = fft_threads(P1) == FFT_THREADS_SMALL && smem1 <= 113u * 1024u, small0 = fft_threads(P0) == FFT_THREADS_SMALL && smem0 <= 113u * 1024u;
    if (small1)
        FAT_LAUNCH("k_fat_dct_rows", (k_fat_dct<0, FFT_THREADS_SMALL>), dct_grid(smem1, h2), FFT_THREADS_SMALL, smem1, (const void*)F, (size_t)w2, (void*)A, pa, h2, P1, t1.tw, t1.postab,
                   (const double*)nullptr, (const double*)nullptr, 0.f);
    else
        FAT_LAUNCH("k_fat_dct_rows", (k_fat_dct<0, FFT_THREADS>), dct_grid(smem1, h2), FFT_THREADS, smem1, (const void*)F, (size_t)w2, (void*)A, pa, h2, P1, t1.tw, t1.postab,
                   (const double*)nullptr, (const double*)nullptr, 0.f);
    FAT_LAUNCH("k_fat_transpose", k_fat_transpose, dim3((w2 + 31) / 32, (h2 + 31) / 32), tb32, 0, A, pa, Bt, pb, h2, w2);
    if (small0)
        FAT_LAUNCH("k_fat_dct_solve", (k_fat_dct<1, FFT_THREADS_SMALL>), dct_grid(smem0, w2), FFT_THREADS_SMALL, smem0, (const void*)Bt, pb, (void*)Bt, pb, w2, P0, t0.tw, t0.postab,
                   (const double*)t0.lam, (const double*)t1.lam, factor);
    else
        FAT_LAUNCH("k_fat_dct_solve", (k_fat_dct<1, FFT_THREADS>), dct_grid(smem0, w2), FFT_THREADS, smem0, (const void*)Bt, pb, (void*)Bt, pb, w2, P0, t0.tw, t0.postab,
                   (const double*)t0.lam, (const double*)t1.lam, factor);
    FAT_LAUNCH("k_fat_transpose", k_fat_transpose, dim3((h2 + 31) / 32, (w2 + 31) / 32), tb32, 0, Bt, pb, A, pa, w2, h2);
    float* L = F;      // F is dead once the first row pass has read it
    if (small1)
        FAT_LAUNCH("k_fat_dct_exp", (k_fat_dct<2, FFT_THREADS_SMALL>), dct_grid(smem1, h2), FFT_THREADS_SMALL, smem1, (const void*)A, pa, (void*)L, (size_t)w2, h2, P1, t1.tw, t1.postab,
                   (const double*)nullptr, (const double*)nullptr, 0.f);
    else
        FAT_LAUNCH("k_fat_dct_exp", (k_fat_dct<2, FFT_THREADS>), dct_grid(smem1, h2), FFT_THREADS, smem1, (const void*)A, pa, (void*)L, (size_t)w2, h2, P1, t1.tw, t1.postab,
                   (const double*)nullptr, (const double*)nullptr, 0.f);

    // median / shadow statistics on 200-px thumbnails, final application (L1129-1213)
    FAT_LAUNCH("k_fat_thumbs", k_fat_thumbs, grid2(ww, hh, b), b, 0, Yr, W, H, L, w2, h2, ta, tb, ww, hh);
    FAT_LAUNCH("k_fat_stats", k_fat_stats, 1, 1024, 0, ta, tb, sz, params);
    FAT_LAUNCH("k_fat_apply", k_fat_apply, grid2(W, H, b), b, 0, R, G, B, ip, W, H, Yr, L, w2, h2, params, satcontrol, ws);
#undef FAT_LAUNCH
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

// 2-D REDFT00 alone (n0 rows x n1 cols, contiguous float in / out): the transform of tmo_fattal02.cc L768-772 by itself,
// for the parity test against the definition
int art_redft00_2d_dev(art_hp_ctx* ctx, const float* in, float* out, int n0, int n1)
{
    cudaStream_t st = ctx->stream;
    FftPlan P1, P0;
    if (n0 < 3 || n1 < 3 || !make_plan(n1 - 1, &P1) || !make_plan(n0 - 1, &P0) || std::max(n0, n1) - 1 > 14336)
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "no shared-memory FFT plan for %d x %d", n0, n1);
    const size_t pa = round_up((size_t)n1, 16), pb = round_up((size_t)n0, 16);
    int rc = art_reserve(ctx, ctx->d_fattal, (round_up(pa * n0, 32) + pb * n1) * 8 + 512);
    if (rc) return rc;
    double* A = (double*)ctx->d_fattal.p;
    double* Bt = A + round_up(pa * n0, 32);
    AxisTables t1, t0;
    if ((rc = axis_tables(ctx, n1, &t1))) return rc;
    if ((rc = axis_tables(ctx, n0, &t0))) return rc;
    cudaFuncSetAttribute(k_fat_dct<0, FFT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_dct_dd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem1 = sizeof(double2) * round_up((size_t)(n1 - 1), 8), smem0 = sizeof(double2) * round_up((size_t)(n0 - 1), 8);
    k_fat_dct<0, FFT_THREADS><<<std::min(n0, ctx->sm_count), FFT_THREADS, smem1, st>>>(in, (size_t)n1, A, pa, n0, P1, t1.tw, t1.postab, nullptr, nullptr, 0.f);
    k_fat_transpose<<<dim3((n1 + 31) / 32, (n0 + 31) / 32), dim3(32, 8), 0, st>>>(A, pa, Bt, pb, n0, n1);
    k_dct_dd<<<std::min(n1, ctx->sm_count), FFT_THREADS, smem0, st>>>(Bt, pb, Bt, pb, n1, P0, t0.tw, t0.postab);
    k_fat_transpose<<<dim3((n0 + 31) / 32, (n1 + 31) / 32), dim3(32, 8), 0, st>>>(Bt, pb, A, pa, n1, n0);
    k_round<<<dim3((n1 + 31) / 32, (n0 + 7) / 8), dim3(32, 8), 0, st>>>(A, pa, out, n0, n1);
    ctx->launches += 5;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
