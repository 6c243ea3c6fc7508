// gaussianBlur (GAUSS_STANDARD, no box buffer) for sm_100a.
//
// Replaces gaussianBlur / gaussianBlurImpl (reference rtengine/gauss.cc L1387-1574) for the branches an
// x86-64 build takes with gausstype == GAUSS_STANDARD: copy (sigma < 0.25), gauss3x3 (L129-174) or the
// separable 3-tap pair (L446-526) for sigma < 0.6, the Young-van Vliet recursive filter with Triggs-Sdika
// boundaries in float (gaussHorizontalSse L554-665, gaussVerticalSse L716-856; the H%4 leftover rows and
// W%8 leftover columns go through the scalar loops that compute in double) for sigma < 25 and in double
// (gaussHorizontal L669-713, gaussVertical L1148-1225) above.
//
// The recursive filter is a serial recurrence along each line and must be evaluated in the reference's
// order to be bit-exact; the parallelism is across lines.  Vertical pass: one thread per column (coalesced
// by construction), software-prefetched loads.  Horizontal pass: one warp per 32 rows walking the row in
// 32-column tiles that are transposed through shared memory, so global traffic stays coalesced while each
// lane runs its own row's recurrence.  The causal pass stores its output in the destination plane (float)
// or in a double scratch plane (sigma >= 25), the anticausal pass sweeps back over it.
// GAUSS_MULT / GAUSS_DIV in the recursive branch (L1490-1511: gaussVerticalSsemult L860-999, gaussVerticalSsediv L1002-1144, the forms
// deconvsharpening reaches above sigma 1.15) are the same recurrence with another store: dst *= v, or dst = div / (v > 0 ? v : 1)
// clamped at 0 (not on the three boundary rows of the 8-column groups, L1078-1091) -- an epilogue of the vertical kernel.
// Compiled with -fmad=false: bit-identical to the reference build.
#include "ctx.h"

#include <cmath>

namespace {

struct Coef {
    float Bf, b1f, b2f, b3f, Mf[9];      // float-lane form (coefficients rounded to float, L569-573)
    double B, b1, b2, b3, M[9];          // scalar-remainder / double form
};

void yvv_factors(double sigma, double& b1, double& b2, double& b3, double& B, double M[9])
{   // calculateYvVFactors<double>, gauss.cc L94-126
    double q;
    if (sigma < 2.5) q = 3.97156 - 4.14554 * std::sqrt(1.0 - 0.26891 * sigma);
    else q = 0.98711 * sigma - 0.96330;
    double b0 = 1.57825 + 2.44413 * q + 1.4281 * q * q + 0.422205 * q * q * q;
    b1 = 2.44413 * q + 2.85619 * q * q + 1.26661 * q * q * q;
    b2 = -1.4281 * q * q - 1.26661 * q * q * q;
    b3 = 0.422205 * q * q * q;
    B = 1.0 - (b1 + b2 + b3) / b0;
    b1 /= b0; b2 /= b0; b3 /= b0;
    M[0] = -b3 * b1 + 1.0 - b3 * b3 - b2;
    M[1] = (b3 + b1) * (b2 + b3 * b1);
    M[2] = b3 * (b1 + b3 * b2);
    M[3] = b1 + b3 * b2;
    M[4] = -(b2 - 1.0) * (b2 + b3 * b1);
    M[5] = -(b3 * b1 + b3 * b3 + b2 - 1.0) * b3;
    M[6] = b3 * b1 + b2 + b1 * b1 - b2 * b2;
    M[7] = b1 * b2 + b3 * b2 * b2 - b1 * b3 * b3 - b3 * b3 * b3 - b3 * b2 + b3;
    M[8] = b3 * (b1 + b3 * b2);
}

// ---- line state machines.  MODE 0: float lanes; 1: scalar remainder (double math, float storage); 2: double
template <int MODE> struct Acc;
template <> struct Acc<0> { using T = float; };
template <> struct Acc<1> { using T = float; };    // state variables are the float scratch values
template <> struct Acc<2> { using T = double; };

template <int MODE>
struct Line {
    using T = typename Acc<MODE>::T;
    T r, m2, m3;        // tmp[j-1], tmp[j-2], tmp[j-3] going forward; tmp[j+1], [j+2], [j+3] going back
    float x0;           // first sample
    __device__ __forceinline__ T first(const Coef& c, float x)
    {
        x0 = x;
        if (MODE == 0) m3 = x * (c.Bf + c.b1f + c.b2f + c.b3f);
        else if (MODE == 1) m3 = (float)(x * (c.B + c.b1 + c.b2 + c.b3));
        else m3 = c.B * x + c.b1 * x + c.b2 * x + c.b3 * x;
        return m3;
    }
    __device__ __forceinline__ T second(const Coef& c, float x)
    {
        if (MODE == 0) m2 = x * c.Bf + m3 * c.b1f + x0 * (c.b2f + c.b3f);
        else if (MODE == 1) m2 = (float)(c.B * x + c.b1 * m3 + x0 * (c.b2 + c.b3));
        else m2 = c.B * x + c.b1 * m3 + c.b2 * x0 + c.b3 * x0;
        return m2;
    }
    __device__ __forceinline__ T third(const Coef& c, float x)
    {
        if (MODE == 0) r = x * c.Bf + m2 * c.b1f + m3 * c.b2f + x0 * c.b3f;
        else if (MODE == 1) r = (float)(c.B * x + c.b1 * m2 + c.b2 * m3 + c.b3 * x0);
        else r = c.B * x + c.b1 * m2 + c.b2 * m3 + c.b3 * x0;
        return r;
    }
    // j >= 3: new = B x + b1 tmp[j-1] + b2 tmp[j-2] + b3 tmp[j-3]
    __device__ __forceinline__ T fwd(const Coef& c, float x)
    {
        T n;
        if (MODE == 0) n = x * c.Bf + r * c.b1f + m2 * c.b2f + m3 * c.b3f;
        else if (MODE == 1) n = (float)(c.B * x + c.b1 * r + c.b2 * m2 + c.b3 * m3);
        else n = c.B * x + c.b1 * r + c.b2 * m2 + c.b3 * m3;
        m3 = m2; m2 = r; r = n;
        return n;
    }
    // Triggs-Sdika boundary at the line end; xl = last input sample.  On entry r, m2, m3 = tmp[n-1], tmp[n-2], tmp[n-3];
    // returns the three final outputs o1 = out[n-1], o2 = out[n-2], o3 = out[n-3] and primes the backward state.
    __device__ __forceinline__ void boundary(const Coef& c, float xl, T& o1, T& o2, T& o3)
    {
        if (MODE == 0) {
            const float p1 = xl + c.Mf[6] * (r - xl) + c.Mf[7] * (m2 - xl) + c.Mf[8] * (m3 - xl);
            const float p0 = xl + c.Mf[3] * (r - xl) + c.Mf[4] * (m2 - xl) + c.Mf[5] * (m3 - xl);
            o1 = xl + c.Mf[0] * (r - xl) + c.Mf[1] * (m2 - xl) + c.Mf[2] * (m3 - xl);
            o2 = c.Bf * m2 + c.b1f * o1 + c.b2f * p0 + c.b3f * p1;
            o3 = c.Bf * m3 + c.b1f * o2 + c.b2f * o1 + c.b3f * p0;
        } else if (MODE == 1) {
            const float m1 = (float)(xl + c.M[0] * (r - xl) + c.M[1] * (m2 - xl) + c.M[2] * (m3 - xl));
            const float p0 = (float)(xl + c.M[3] * (r - xl) + c.M[4] * (m2 - xl) + c.M[5] * (m3 - xl));
            const float p1 = (float)(xl + c.M[6] * (r - xl) + c.M[7] * (m2 - xl) + c.M[8] * (m3 - xl));
            o1 = m1;
            o2 = (float)(c.B * m2 + c.b1 * o1 + c.b2 * p0 + c.b3 * p1);
            o3 = (float)(c.B * m3 + c.b1 * o2 + c.b2 * o1 + c.b3 * p0);
        } else {
            const double m1 = xl + c.M[0] * (r - xl) + c.M[1] * (m2 - xl) + c.M[2] * (m3 - xl);
            const double p0 = xl + c.M[3] * (r - xl) + c.M[4] * (m2 - xl) + c.M[5] * (m3 - xl);
            const double p1 = xl + c.M[6] * (r - xl) + c.M[7] * (m2 - xl) + c.M[8] * (m3 - xl);
            o1 = m1;
            o2 = c.B * m2 + c.b1 * o1 + c.b2 * p0 + c.b3 * p1;
            o3 = c.B * m3 + c.b1 * o2 + c.b2 * o1 + c.b3 * p0;
        }
        r = o3; m2 = o2; m3 = o1;      // out[j+1], out[j+2], out[j+3] for j = n-4
    }
    // j <= n-4: out[j] = B tmp[j] + b1 out[j+1] + b2 out[j+2] + b3 out[j+3]
    __device__ __forceinline__ T bwd(const Coef& c, T t)
    {
        T n;
        if (MODE == 0) n = t * c.Bf + r * c.b1f + m2 * c.b2f + m3 * c.b3f;
        else if (MODE == 1) n = (float)(c.B * t + c.b1 * r + c.b2 * m2 + c.b3 * m3);
        else n = c.B * t + c.b1 * r + c.b2 * m2 + c.b3 * m3;
        m3 = m2; m2 = r; r = n;
        return n;
    }
};

struct GArgs {
    const float* src; size_t sp;
    float* dst; size_t dp;
    double* dscr; size_t dsp;       // double scratch plane (sigma >= 25), pitch in doubles
    int W, H;
    int big;                        // 1: all-double form
    float* cs; size_t csp;          // vertical pass: plane that takes the causal output (dst; the source plane for GAUSS_MULT)
    const float* div; size_t vp;    // GAUSS_DIV numerator
    Coef c;
};

// ------------------------------------------------------------------ vertical pass: thread per column
// EPI 0: dst = v; 1 (GAUSS_MULT): dst *= v; 2 (GAUSS_DIV): dst = div / (v > 0 ? v : 1), clamped at 0
constexpr int VR = 96, VT = 64;       // rows of a column in flight (cp.async ring depth), threads per CTA; 2 rings x 96 rows x 64 threads x 4 B = 48 KB of dynamic shared memory
template <int MODE, int EPI>
__device__ __forceinline__ void vline(const GArgs& a, int col, float* vring)
{
    using T = typename Acc<MODE>::T;
    Line<MODE> L;
    const int H = a.H;
    const float* x = a.src + col;
    float* y = a.dst + col;
    float* cs = a.cs + col;
    const float* e = EPI == 1 ? y : (EPI == 2 ? a.div + col : nullptr);      // second operand of the epilogue
    const size_t ep = EPI == 1 ? a.dp : a.vp;
    double* d = MODE == 2 ? a.dscr + col : nullptr;
    auto emit = [&](int j, float v, float ev, bool boundary_row) {
        float* o = y + (size_t)j * a.dp;
        if (EPI == 0) *o = v;
        else if (EPI == 1) *o = ev * v;
        else {
            float q = ev / (v > 0.f ? v : 1.f);
            if (MODE == 1) q = q < 0.f ? 0.f : q;                  // rtengine::max(q, 0.f), every row (L1139-1141)
            else if (!boundary_row) q = q > 0.f ? q : 0.f;         // vmaxf(q, ZEROV), rows < H-3 only (L1103-1104)
            *o = q;
        }
    };
    if (MODE != 2) {
        // A frame has only W chains (one or two warps per SM), so the memory latency has to be hidden by depth: every thread keeps
        // VR - 1 rows of its column in flight through cp.async into a private shared-memory ring (ncu, register prefetch of 24 rows:
        // 6 % occupancy, 90 % of the issue slots idle waiting for the loads).  The input of row j + VR - 1 is read before the output
        // of row j is stored: in-place safe, as before.
        float* ring = vring + threadIdx.x;                 // slot s of this thread: ring[s * blockDim.x]
        const unsigned rbase = (unsigned)__cvta_generic_to_shared(ring);
        const unsigned rstep = blockDim.x * sizeof(float);
        auto fetch = [&](const float* g, int slot, bool ok) {
            if (ok) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(rbase + slot * rstep), "l"(g) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int k = 0; k < VR - 1; ++k) fetch(x + (size_t)k * a.sp, k, k < H);
        float xl = 0.f;
        for (int j = 0; j < H; ++j) {
            const int jn = j + VR - 1;
            fetch(x + (size_t)jn * a.sp, jn % VR, jn < H);
            asm volatile("cp.async.wait_group %0;" :: "n"(VR - 1) : "memory");
            const float xv = ring[(j % VR) * blockDim.x];
            T t;
            if (j == 0) t = L.first(a.c, xv);
            else if (j == 1) t = L.second(a.c, xv);
            else if (j == 2) t = L.third(a.c, xv);
            else t = L.fwd(a.c, xv);
            cs[(size_t)j * a.csp] = (float)t;
            xl = xv;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        T o1, o2, o3;
        L.boundary(a.c, xl, o1, o2, o3);
        emit(H - 1, (float)o1, EPI ? e[(size_t)(H - 1) * ep] : 0.f, true);
        emit(H - 2, (float)o2, EPI ? e[(size_t)(H - 2) * ep] : 0.f, true);
        emit(H - 3, (float)o3, EPI ? e[(size_t)(H - 3) * ep] : 0.f, true);
        // anticausal sweep over the stored causal output (rows H-4 .. 0); the epilogue operand rides in a second ring
        float* ring2 = ring + VR * blockDim.x;
        const unsigned r2base = rbase + VR * rstep;
        auto fetch2 = [&](int j, int slot) {
            if (j >= 0) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(rbase + slot * rstep), "l"(cs + (size_t)j * a.csp) : "memory");
                if (EPI) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(r2base + slot * rstep), "l"(e + (size_t)j * ep) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int k = 0; k < VR - 1; ++k) fetch2(H - 4 - k, k);
        int it = 0;
        for (int j = H - 4; j >= 0; --j, ++it) {
            fetch2(j - (VR - 1), (it + VR - 1) % VR);
            asm volatile("cp.async.wait_group %0;" :: "n"(VR - 1) : "memory");
            const T t = (T)ring[(it % VR) * blockDim.x];
            const float ev = EPI ? ring2[(it % VR) * blockDim.x] : 0.f;
            emit(j, (float)L.bwd(a.c, t), ev, false);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    // all-double form (sigma >= 25): register prefetch; the input of row j+PF is loaded before the output of row j is stored (in-place safe)
    constexpr int PF = 8;
    float q[PF];
    #pragma unroll
    for (int k = 0; k < PF; ++k) q[k] = (k < H) ? x[(size_t)k * a.sp] : 0.f;
    float xl = 0.f;
    for (int j0 = 0; j0 < H; j0 += PF) {
        #pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int j = j0 + k;
            if (j < H) {
                const float xv = q[k];
                const int jn = j + PF;
                q[k] = (jn < H) ? x[(size_t)jn * a.sp] : 0.f;
                T t;
                if (j == 0) t = L.first(a.c, xv);
                else if (j == 1) t = L.second(a.c, xv);
                else if (j == 2) t = L.third(a.c, xv);
                else t = L.fwd(a.c, xv);
                if (MODE == 2) d[(size_t)j * a.dsp] = t; else cs[(size_t)j * a.csp] = (float)t;
                xl = xv;
            }
        }
    }
    T o1, o2, o3;
    L.boundary(a.c, xl, o1, o2, o3);
    emit(H - 1, (float)o1, EPI ? e[(size_t)(H - 1) * ep] : 0.f, true);
    emit(H - 2, (float)o2, EPI ? e[(size_t)(H - 2) * ep] : 0.f, true);
    emit(H - 3, (float)o3, EPI ? e[(size_t)(H - 3) * ep] : 0.f, true);
    // anticausal sweep over the stored causal output
    T p[PF];
    float pe[PF];
    #pragma unroll
    for (int k = 0; k < PF; ++k) {
        const int j = H - 4 - k;
        p[k] = (j >= 0) ? (MODE == 2 ? (T)d[(size_t)j * a.dsp] : (T)cs[(size_t)j * a.csp]) : (T)0;
        pe[k] = (EPI && j >= 0) ? e[(size_t)j * ep] : 0.f;
    }
    for (int j0 = H - 4; j0 >= 0; j0 -= PF) {
        #pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int j = j0 - k;
            if (j >= 0) {
                const T t = p[k];
                const float ev = pe[k];
                const int jn = j - PF;
                p[k] = (jn >= 0) ? (MODE == 2 ? (T)d[(size_t)jn * a.dsp] : (T)cs[(size_t)jn * a.csp]) : (T)0;
                pe[k] = (EPI && jn >= 0) ? e[(size_t)jn * ep] : 0.f;
                emit(j, (float)L.bwd(a.c, t), ev, false);
            }
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(VT) k_gauss_v(GArgs a)
{
    extern __shared__ float vring[];
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.W) return;
    if (a.big) vline<2, EPI>(a, col, vring);
    else if (col < a.W - (a.W % 8)) vline<0, EPI>(a, col, vring);  // 8-column vector groups (L750)
    else vline<1, EPI>(a, col, vring);                             // scalar remainder (L843)
}

// ------------------------------------------------------------------ horizontal pass: warp per 32 rows
template <int MODE>
__device__ __forceinline__ void hline(const GArgs& a, int row0, int lane, float (*tile)[33], double (*dtile)[33])
{
    using T = typename Acc<MODE>::T;
    Line<MODE> L;
    const int W = a.W;
    const int nrows = min(32, a.H - row0);
    const bool mine = lane < nrows;
    float xl = 0.f;
    // causal sweep, tile by tile: the next tile's 32 row segments are loaded into registers before this tile's serial chain runs
    float nxt[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) nxt[r] = (r < nrows && lane < W) ? a.src[(size_t)(row0 + r) * a.sp + lane] : 0.f;
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int nc = min(32, W - c0);
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[r][lane] = nxt[r];
        if (c0 + 32 < W) {
            const int cn = c0 + 32 + lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) nxt[r] = (r < nrows && cn < W) ? a.src[(size_t)(row0 + r) * a.sp + cn] : 0.f;
        }
        __syncwarp();
        if (mine) {
            if (MODE != 2 && nc == 32) {
                // the lane's 32 samples into registers first: the recurrence then runs on registers, its shared-memory loads off the dependency chain
                float xr[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) xr[k] = tile[lane][k];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    T t;
                    if (k < 3 && c0 == 0) t = k == 0 ? L.first(a.c, xr[k]) : k == 1 ? L.second(a.c, xr[k]) : L.third(a.c, xr[k]);
                    else t = L.fwd(a.c, xr[k]);
                    tile[lane][k] = (float)t;
                }
                xl = xr[31];
            } else {
                for (int k = 0; k < nc; ++k) {
                    const int j = c0 + k;
                    const float xv = tile[lane][k];
                    T t;
                    if (j == 0) t = L.first(a.c, xv);
                    else if (j == 1) t = L.second(a.c, xv);
                    else if (j == 2) t = L.third(a.c, xv);
                    else t = L.fwd(a.c, xv);
                    if (MODE == 2) dtile[lane][k] = t; else tile[lane][k] = (float)t;
                    xl = xv;
                }
            }
        }
        __syncwarp();
        for (int r = 0; r < nrows; ++r)
            if (lane < nc) {
                if (MODE == 2) a.dscr[(size_t)(row0 + r) * a.dsp + c0 + lane] = dtile[r][lane];
                else a.dst[(size_t)(row0 + r) * a.dp + c0 + lane] = tile[r][lane];
            }
        __syncwarp();
    }
    T o1 = 0, o2 = 0, o3 = 0;
    if (mine) L.boundary(a.c, xl, o1, o2, o3);
    // anticausal sweep: tiles from the right; the last three columns come from the boundary step
    const int ntile = (W + 31) / 32;
    if (MODE != 2) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int cn = (ntile - 1) * 32 + lane;
            nxt[r] = (r < nrows && cn < W) ? a.dst[(size_t)(row0 + r) * a.dp + cn] : 0.f;
        }
    }
    for (int tix = ntile - 1; tix >= 0; --tix) {
        const int c0 = tix * 32, nc = min(32, W - c0);
        if (MODE == 2) {
            for (int r = 0; r < nrows; ++r)
                if (lane < nc) dtile[r][lane] = a.dscr[(size_t)(row0 + r) * a.dsp + c0 + lane];
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) tile[r][lane] = nxt[r];
            if (tix > 0) {
                const int cn = c0 - 32 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r) nxt[r] = (r < nrows) ? a.dst[(size_t)(row0 + r) * a.dp + cn] : 0.f;
            }
        }
        __syncwarp();
        if (mine) {
            if (MODE != 2 && nc == 32 && c0 + 32 <= W - 3) {
                float xr[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) xr[k] = tile[lane][k];
#pragma unroll
                for (int k = 31; k >= 0; --k) tile[lane][k] = (float)L.bwd(a.c, (T)xr[k]);
            } else {
                for (int k = nc - 1; k >= 0; --k) {
                    const int j = c0 + k;
                    T v;
                    if (j == W - 1) v = o1;
                    else if (j == W - 2) v = o2;
                    else if (j == W - 3) v = o3;
                    else v = L.bwd(a.c, MODE == 2 ? (T)dtile[lane][k] : (T)tile[lane][k]);
                    tile[lane][k] = (float)v;
                }
            }
        }
        __syncwarp();
        for (int r = 0; r < nrows; ++r) if (lane < nc) a.dst[(size_t)(row0 + r) * a.dp + c0 + lane] = tile[r][lane];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(64) k_gauss_h(GArgs a)
{
    __shared__ float tile[2][32][33];
    __shared__ double dtile[2][32][33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * 2 + w) * 32;
    if (row0 >= a.H) return;
    // rows of a warp can mix forms (the last H%4 rows use the scalar remainder): run each form over the
    // whole warp with the rows of the other form masked out
    if (a.big) { hline<2>(a, row0, lane, tile[w], dtile[w]); return; }
    const int nfl = a.H - (a.H % 4);                          // rows < nfl are in 4-row vector groups (L586)
    if (row0 + 32 <= nfl) { hline<0>(a, row0, lane, tile[w], dtile[w]); return; }
    // mixed warp: split by masking -- the float rows first, then the remainder rows
    GArgs b = a;
    if (row0 < nfl) { b.H = nfl; hline<0>(b, row0, lane, tile[w], dtile[w]); }
    b = a;
    const int r1 = max(row0, nfl);
    hline<1>(b, r1, lane, tile[w], dtile[w]);
}

// ------------------------------------------------------------------ sigma < 0.6
struct G3Args {
    const float* src; size_t sp; float* dst; size_t dp; int W, H;
    float c0, c1, c2, b0, b1;
};

__global__ void __launch_bounds__(256) k_gauss3x3(G3Args a)
{   // gauss3x3, L129-174
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.W) return;
    for (int i = blockIdx.y; i < a.H; i += gridDim.y) {
        const float* s = a.src + (size_t)i * a.sp;
        float v;
        const bool er = (i == 0 || i == a.H - 1), ec = (j == 0 || j == a.W - 1);
        if (er && ec) v = s[j];
        else if (er) v = a.b1 * (s[j - 1] + s[j + 1]) + a.b0 * s[j];
        else if (ec) { const float* u = s - a.sp; const float* d = s + a.sp; v = a.b1 * (u[j] + d[j]) + a.b0 * s[j]; }
        else {
            const float* u = s - a.sp;
            const float* d = s + a.sp;
            v = a.c2 * (u[j - 1] + u[j + 1] + d[j - 1] + d[j + 1]) + a.c1 * (u[j] + s[j - 1] + s[j + 1] + d[j]) + a.c0 * s[j];
        }
        a.dst[(size_t)i * a.dp + j] = v;
    }
}

// separable 3-tap pair (gaussHorizontal3 L446-464, gaussVertical3 L467-526): pass 0 src -> dst (rows), pass 1 src -> dst (columns)
__global__ void __launch_bounds__(256) k_gauss3sep(G3Args a, int vertical)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.W) return;
    for (int i = blockIdx.y; i < a.H; i += gridDim.y) {
        const float* s = a.src + (size_t)i * a.sp;
        float v;
        if (!vertical) v = (j == 0 || j == a.W - 1) ? s[j] : (a.c1 * (s[j - 1] + s[j + 1]) + a.c0 * s[j]);
        else if (i == 0 || i == a.H - 1) v = s[j];
        else { const float* u = s - a.sp; const float* d = s + a.sp; v = a.c1 * (d[j] + u[j]) + s[j] * a.c0; }
        a.dst[(size_t)i * a.dp + j] = v;
    }
}

}  // namespace

int art_gauss_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, double sigma)
{
    cudaStream_t st = ctx->stream;
    const bool inplace = (src == dst);
    if (sigma < 0.25) {       // L1438-1444
        if (!inplace) ART_CUDA(ctx, cudaMemcpy2DAsync(dst, dp * sizeof(float), src, sp * sizeof(float), (size_t)W * sizeof(float), H, cudaMemcpyDeviceToDevice, st));
        return ART_HP_OK;
    }
    const dim3 pgrid((W + 255) / 256, std::min(H, 148 * 4));
    if (sigma < 0.6) {
        G3Args g{src, sp, dst, dp, W, H, 0, 0, 0, 0, 0};
        if (!inplace) {       // L1446-1478
            double c0 = 1.0, c1 = std::exp(-0.5 * ((1.0 / sigma) * (1.0 / sigma))), c2 = std::exp(-((1.0 / sigma) * (1.0 / sigma)));
            const double sum = c0 + 4.0 * (c1 + c2);
            c0 /= sum; c1 /= sum; c2 /= sum;
            double b1 = std::exp(-1.0 / (2.0 * sigma * sigma));
            const double bsum = 2.0 * b1 + 1.0;
            b1 /= bsum;
            g.c0 = (float)c0; g.c1 = (float)c1; g.c2 = (float)c2; g.b0 = (float)(1.0 / bsum); g.b1 = (float)b1;
            art_prof_begin(ctx, "k_gauss3x3");
            k_gauss3x3<<<pgrid, 256, 0, st>>>(g);
            art_prof_end(ctx);
            ctx->launches++;
        } else {              // L1479-1487: horizontal into a scratch plane, vertical back
            double c1d = std::exp(-1.0 / (2.0 * sigma * sigma));
            const double csum = 2.0 * c1d + 1.0;
            c1d /= csum;
            g.c1 = (float)c1d; g.c0 = (float)(1.0 / csum);
            const size_t tp = round_up((size_t)W, 32);
            int rc = art_reserve(ctx, ctx->d_scratch, tp * (size_t)H * sizeof(float));
            if (rc) return rc;
            float* tmp = (float*)ctx->d_scratch.p;
            G3Args h = g; h.dst = tmp; h.dp = tp;
            art_prof_begin(ctx, "k_gauss3sep");
            k_gauss3sep<<<pgrid, 256, 0, st>>>(h, 0);
            G3Args v = g; v.src = tmp; v.sp = tp;
            k_gauss3sep<<<pgrid, 256, 0, st>>>(v, 1);
            art_prof_end(ctx);
            ctx->launches += 2;
        }
        ART_CUDA(ctx, cudaGetLastError());
        return ART_HP_OK;
    }
    GArgs a;
    a.src = src; a.sp = sp; a.dst = dst; a.dp = dp; a.W = W; a.H = H; a.dscr = nullptr; a.dsp = 0;
    a.cs = dst; a.csp = dp; a.div = nullptr; a.vp = 0;
    a.big = sigma >= 25.0;
    double b1, b2, b3, B, M[9];
    if (!a.big) {
        const float sigf = (float)sigma;         // gaussHorizontalSse(..., const float sigma)
        yvv_factors(sigf, b1, b2, b3, B, M);
        for (int i = 0; i < 9; ++i) {            // L559-563
            M[i] *= (1.0 + b2 + (b1 - b3) * b3);
            M[i] /= (1.0 + b1 - b2 + b3) * (1.0 - b1 - b2 - b3);
        }
    } else {
        yvv_factors(sigma, b1, b2, b3, B, M);
        for (int i = 0; i < 9; ++i) M[i] /= (1.0 + b1 - b2 + b3) * (1.0 + b2 + (b1 - b3) * b3);     // L674-677
        a.dsp = round_up((size_t)W, 32);
        int rc = art_reserve(ctx, ctx->d_scratch, a.dsp * (size_t)H * sizeof(double));
        if (rc) return rc;
        a.dscr = (double*)ctx->d_scratch.p;
    }
    a.c.B = B; a.c.b1 = b1; a.c.b2 = b2; a.c.b3 = b3;
    a.c.Bf = (float)B; a.c.b1f = (float)b1; a.c.b2f = (float)b2; a.c.b3f = (float)b3;
    for (int i = 0; i < 9; ++i) { a.c.M[i] = M[i]; a.c.Mf[i] = (float)M[i]; }
    art_prof_begin(ctx, "k_gauss_h");
    k_gauss_h<<<(H + 63) / 64, 64, 0, st>>>(a);
    art_prof_end(ctx);
    GArgs v = a;
    v.src = dst; v.sp = dp;                      // vertical runs in place on the horizontal result (L1529-1530)
    art_prof_begin(ctx, "k_gauss_v");
    k_gauss_v<0><<<(W + VT - 1) / VT, VT, 2 * VR * VT * sizeof(float), st>>>(v);
    art_prof_end(ctx);
    ctx->launches += 2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

// gaussianBlur(src, dst, W, H, sigma, nullptr, GAUSS_MULT (type 1) / GAUSS_DIV (type 2), div) in the recursive branch (L1490-1511),
// 0.6 <= sigma < 25, src != dst.  GAUSS_MULT filters src horizontally IN PLACE like the reference (L1496) and then keeps the
// vertical pass's causal output there; GAUSS_DIV leaves src alone.
int art_gauss_divmult_dev(art_hp_ctx* ctx, float* src, size_t sp, float* dst, size_t dp, const float* div, size_t vp, int W, int H, double sigma, int type)
{
    if (!(sigma >= 0.6) || sigma >= 25.0 || W < 4 || H < 4 || src == dst || (type != 1 && type != 2) || (type == 2 && !div))
        return ctx->fail(ART_HP_ERR_INVALID, "recursive GAUSS_MULT / GAUSS_DIV: needs 0.6 <= sigma < 25, W, H >= 4, src != dst (got sigma %.3f, %dx%d, type %d)", sigma, W, H, type);
    cudaStream_t st = ctx->stream;
    GArgs a;
    a.W = W; a.H = H; a.dscr = nullptr; a.dsp = 0; a.big = 0;
    double b1, b2, b3, B, M[9];
    const float sigf = (float)sigma;
    yvv_factors(sigf, b1, b2, b3, B, M);
    for (int i = 0; i < 9; ++i) {
        M[i] *= (1.0 + b2 + (b1 - b3) * b3);
        M[i] /= (1.0 + b1 - b2 + b3) * (1.0 - b1 - b2 - b3);
    }
    a.c.B = B; a.c.b1 = b1; a.c.b2 = b2; a.c.b3 = b3;
    a.c.Bf = (float)B; a.c.b1f = (float)b1; a.c.b2f = (float)b2; a.c.b3f = (float)b3;
    for (int i = 0; i < 9; ++i) { a.c.M[i] = M[i]; a.c.Mf[i] = (float)M[i]; }
    float* hp = type == 1 ? src : dst;                // plane holding the horizontal result
    const size_t hpp = type == 1 ? sp : dp;
    a.src = src; a.sp = sp; a.dst = hp; a.dp = hpp; a.cs = hp; a.csp = hpp; a.div = nullptr; a.vp = 0;
    art_prof_begin(ctx, "k_gauss_h");
    k_gauss_h<<<(H + 63) / 64, 64, 0, st>>>(a);
    art_prof_end(ctx);
    GArgs v = a;
    v.src = hp; v.sp = hpp; v.cs = hp; v.csp = hpp; v.dst = dst; v.dp = dp; v.div = div; v.vp = vp;
    art_prof_begin(ctx, "k_gauss_v");
    if (type == 1) k_gauss_v<1><<<(W + VT - 1) / VT, VT, 2 * VR * VT * sizeof(float), st>>>(v);
    else k_gauss_v<2><<<(W + VT - 1) / VT, VT, 2 * VR * VT * sizeof(float), st>>>(v);
    art_prof_end(ctx);
    ctx->launches += 2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
