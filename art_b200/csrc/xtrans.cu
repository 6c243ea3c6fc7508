// X-Trans demosaic (Markesteijn, 1 or 3 passes): RawImageSource::xtrans_interpolate, cielab and xtransborder_interpolate
// (reference rtengine/xtrans_demosaic.cc L42-116, L122-173, L181-969).
//
// Design.  The reference grid is kept (tiles of 114 from (3, 3), stride 98).  One persistent 1024-thread CTA per SM walks the
// tiles of the frame (static striding) and owns (a) a tile slab in HBM / L2 with exactly the reference's buffer layout and (b)
// 208 KB of shared memory holding ONE direction plane (114 x 114 x 3 floats) plus the green min/max table.  The direction planes
// of a pass never read each other, so a plane is loaded once, all passes and all their steps run on it in shared memory, it is
// stored as direction p (pass 0) and p + 4 (passes 1-2), and its CIELab / YPbPr conversion (in place) and derivative are taken
// before the next plane comes in.  Every step is a data-parallel loop over the tile followed by a CTA barrier; which sites a step
// visits, with which hexagon / colour / direction plane, is computed per pixel instead of with the reference's running column
// toggles (the test-suite's CPU restatement is organised the same way and is pinned bit-exact to the reference).  The small
// per-sensor tables are indexed differently in every lane and therefore live in shared memory, not in the constant bank.
//
// The layout has to be the reference's because its sub-buffers alias (homo and the green min/max table over lab, homosum
// over drv) and the 5x5 homogeneity sums next to the image border read homo bytes that the homogeneity step never wrote,
// i.e. bytes of Lab / YPbPr floats and of min/max floats; those sums may pass 255, where the reference's 16-wide SSE2 path
// saturates and its scalar tail wraps.  The slab is cleared at the start of every tile: the canonical ("det") reference.
//
// Algorithmic bytes: 4 B read + 12 B written per pixel = 16 B/px.  The kernel is latency / L2 bound, not HBM
// bound.  Phase times at 6240 x 4160, 3-pass CIELab (tools/xtrans_phase_probe.py): mosaic + first greens 1.1 ms, passes 4.2 ms,
// Lab + derivatives 4.9 ms (three cube-root LUT gathers per pixel and direction into a 327 KB table that no longer fits beside
// 208 KB of shared memory in L1), homogeneity maps / 5x5 sums / selection 2.9 ms.
#include "ctx.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int TS = 114, TSH = TS / 2;
constexpr int XT_THREADS = 1024;
constexpr size_t XT_SMEM = ((size_t)TS * TS * 3 + (size_t)TS * TSH * 2) * sizeof(float);     // one direction plane + the green min/max table
constexpr int CBRT_N = 0x14000;

struct XtArgs {
    const float* raw; size_t rp;
    float *R, *G, *B; size_t op;
    int W, H, passes, ndir, useCieLab;
    int ntx, nty;
    float* slabs; size_t slab_floats;
    const float* cbrt;
    float xyz_cam[9];
    signed char hexv[3][3][8], hexh[3][3][8];      // allhex as (v, h) pairs: image offset h + v * rp, tile offset h + v * TS
    unsigned char xt[6][6];
    unsigned char rshift[3];
    int sgrow, sgcol;
    int stop;       // tools only (ART_XT_STOP): leave every tile after phase `stop` (0 = run everything)
};

// The small per-sensor tables are indexed by (row % 3, col % 3) / (row % 6, col % 6), i.e. with a different index in every lane:
// from kernel parameters (constant bank) such loads serialise per distinct address, so the kernel keeps a copy in shared memory.
struct XtTab {
    unsigned char xt[6][6];
    unsigned char rshift[3];
    signed char hexv[3][3][8], hexh[3][3][8];
};
__device__ __forceinline__ int fcol(const XtTab& a, int row, int col) { return a.xt[row % 6][col % 6]; }
__device__ __forceinline__ bool isgreen(const XtTab& a, int row, int col) { return a.xt[row % 3][col % 3] & 1; }
__device__ __forceinline__ float limf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float sqrf(float v) { return v * v; }
__device__ __forceinline__ float lut_i(const float* __restrict__ t, int idx) { return t[idx < 0 ? 0 : (idx > CBRT_N - 1 ? CBRT_N - 1 : idx)]; }

// for (ri, ci) over an NR x NC rectangle, thread-strided in row-major order, without a division per iteration
#define XT_FOR2(ri, ci, NR, NC)                                                                                                   \
    for (int nr_ = (NR), nc_ = (NC), ok_ = (nr_ > 0 && nc_ > 0), ri = ok_ ? tid / nc_ : nr_, ci = ok_ ? tid % nc_ : 0,             \
             dr_ = ok_ ? XT_THREADS / nc_ : 0, dc_ = ok_ ? XT_THREADS % nc_ : 0;                                                  \
         ri < nr_; ci += dc_, ri += dr_ + (ci >= nc_), ci -= (ci >= nc_) ? nc_ : 0)

// rgb[d][r][c][ch] of the slab
#define RGB(d, r, c, ch) rgb[(((size_t)(d) * TS + (r)) * TS + (c)) * 3 + (ch)]

__global__ void __launch_bounds__(XT_THREADS) k_xtrans(const XtArgs a)
{
    extern __shared__ __align__(16) float xsm[];
    __shared__ XtTab tb;
    const int tid = threadIdx.x;
    if (tid < 36) tb.xt[tid / 6][tid % 6] = a.xt[tid / 6][tid % 6];
    if (tid < 3) tb.rshift[tid] = a.rshift[tid];
    if (tid < 72) { (&tb.hexv[0][0][0])[tid] = (&a.hexv[0][0][0])[tid]; (&tb.hexh[0][0][0])[tid] = (&a.hexh[0][0][0])[tid]; }
    __syncthreads();
    const int ndir = a.ndir, W = a.W, H = a.H;
    float* const buffer = a.slabs + (size_t)blockIdx.x * a.slab_floats;
    float* const lab = buffer + TS * TS * (ndir * 3);                 // [3][TS-8][TS-8]
    float* const drv = buffer + TS * TS * (ndir * 3 + 3);             // [ndir][TS-10][TS-10]
    unsigned char* const homo = reinterpret_cast<unsigned char*>(lab);       // [ndir][TS][TS]
    float2* const gmm = reinterpret_cast<float2*>(lab);                      // [TS][TSH] {min, max}
    constexpr int LW = TS - 8, DW = TS - 10;
    constexpr int LAB_VEC_COLS = (LW - 3 + 3) / 4 * 4;               // columns covered by `for (j = 0; j < labWidth - 3; j += 4)`

    for (int t = blockIdx.x; t < a.ntx * a.nty; t += gridDim.x) {
        const int top = 3 + (t / a.ntx) * (TS - 16), left = 3 + (t % a.ntx) * (TS - 16);
        int mrow = min(top + TS, H - 3), mcol = min(left + TS, W - 3);
        const int rows = mrow - top, cols = mcol - left;
        float* rgb = buffer;

        {   // the canonical reference clears its tile buffer; what is read before it is written are the four mosaic planes
            // (channels the mosaic does not fill, rows / columns the passes do not reach) and the lab area (as homo bytes)
            float4* b4 = reinterpret_cast<float4*>(buffer);
            for (int i = tid; i < 4 * TS * TS * 3 / 4; i += XT_THREADS) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* l4 = reinterpret_cast<float4*>(lab);
            for (int i = tid; i < 3 * LW * LW / 4; i += XT_THREADS) l4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();

        // mosaic into rgb[0..3]; green min / max (L319-408) and green along the four directions (L421-475) at non-green sites
        XT_FOR2(r, c, rows, cols) {
            const int row = top + r, col = left + c;
            const float* pix = a.raw + (size_t)row * a.rp + col;
            const int f = fcol(tb, row, col);
            const float v = pix[0];
#pragma unroll
            for (int d = 0; d < 4; ++d) RGB(d, r, c, f) = v;
            if (f == 1) continue;
            int src = col;
            const bool rs = tb.rshift[row % 3];
            if (!rs && !isgreen(tb, row, col + 5) && col - 1 >= left) src = col - 1;      // second pixel of a horizontal pair
            float mn = FLT_MAX, mx = 0.f;
            {
                const signed char* hv = tb.hexv[row % 3][src % 3];
                const signed char* hh = tb.hexh[row % 3][src % 3];
                const float* sp = a.raw + (size_t)row * a.rp + src;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const float val = sp[hh[k] + (ptrdiff_t)hv[k] * (ptrdiff_t)a.rp];
                    mn = mn < val ? mn : val;
                    mx = mx > val ? mx : val;
                }
            }
            gmm[r * TSH + (c >> 1)] = make_float2(mn, mx);
            const signed char* hv = tb.hexv[row % 3][col % 3];
            const signed char* hh = tb.hexh[row % 3][col % 3];
            ptrdiff_t hex[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) hex[k] = hh[k] + (ptrdiff_t)hv[k] * (ptrdiff_t)a.rp;
            float color[4];
            color[0] = 0.6796875f * (pix[hex[1]] + pix[hex[0]]) - 0.1796875f * (pix[2 * hex[1]] + pix[2 * hex[0]]);
            color[1] = 0.87109375f * pix[hex[3]] + pix[hex[2]] * 0.12890625f + 0.359375f * (pix[0] - pix[-hex[2]]);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                color[2 + k] = 0.640625f * pix[hex[4 + k]] + 0.359375f * pix[-2 * hex[4 + k]] +
                               0.12890625f * (2.f * pix[0] - pix[3 * hex[4 + k]] - pix[-3 * hex[4 + k]]);
            const int flip = rs ? 0 : 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) RGB(k ^ flip, r, c, 1) = limf(color[k], mn, mx);
        }
        __syncthreads();

        if (a.stop == 1) { __syncthreads(); continue; }
        // ---- the passes, ONE DIRECTION PLANE AT A TIME IN SHARED MEMORY.  The four planes a pass works on never read each other
        // (the solitary-green step carries values from direction d-1 to d only within a plane), and passes 1 and 2 start from
        // pass 0's planes (`memcpy(rgb += 4, buffer, ...)`, L479-481): plane p is loaded once, pass 0 runs and is stored as
        // direction p, passes 1-2 run on the same shared-memory copy and are stored as direction p + 4.
        {
            float* const plane = xsm;                                              // [TS][TS][3]
            float2* const sgmm = reinterpret_cast<float2*>(xsm + TS * TS * 3);     // [TS][TSH]
            if (a.passes > 1) {
                const float4* s4 = reinterpret_cast<const float4*>(gmm);
                float4* d4 = reinterpret_cast<float4*>(sgmm);
                for (int i = tid; i < TS * TSH * 2 / 4; i += XT_THREADS) d4[i] = s4[i];
            }
            const int sol_row0 = (top - a.sgrow + 4) / 3 * 3 + a.sgrow, sol_col0 = (left - a.sgcol + 4) / 3 * 3 + a.sgcol;
            const int sol_nr = sol_row0 < mrow - 2 ? (mrow - 2 - sol_row0 + 2) / 3 : 0, sol_nc = sol_col0 < mcol - 2 ? (mcol - 2 - sol_col0 + 2) / 3 : 0;
            // Derivatives of the direction plane held in shared memory, in CIELab (L657-683; cielab L65-116) or YPbPr (L684-741).
            // The conversion is done in place over the plane's own pixels; the lab buffer in the slab is written for the LAST
            // direction only: nothing but its final content is ever read (as homo bytes next to the image border).
            auto labdrv = [&](int dd) {
                if (a.useCieLab) {
                    const int n = (rows - 8) * LW;
                    for (int i = tid; i < n; i += XT_THREADS) {
                        const int r = i / LW, c = i - r * LW;
                        float* px = plane + ((4 + r) * TS + 4 + c) * 3;
                        const float p0 = px[0], p1 = px[1], p2 = px[2];
                        float fx, fy, fz, L;
                        if (c < LAB_VEC_COLS) {     // 4-wide SSE2 groups (j < labWidth - 3): index rounded to nearest even by _mm_cvtps_epi32
                            const float X = p0 * a.xyz_cam[0] + p1 * a.xyz_cam[1] + p2 * a.xyz_cam[2];
                            const float Y = p0 * a.xyz_cam[3] + p1 * a.xyz_cam[4] + p2 * a.xyz_cam[5];
                            const float Z = p0 * a.xyz_cam[6] + p1 * a.xyz_cam[7] + p2 * a.xyz_cam[8];
                            fx = lut_i(a.cbrt, __float2int_rn(X)); fy = lut_i(a.cbrt, __float2int_rn(Y)); fz = lut_i(a.cbrt, __float2int_rn(Z));
                            L = 116.f * fy - 16.f;
                        } else {                    // scalar tail: 0.5 added first, index truncated
                            float x0 = 0.5f, x1 = 0.5f, x2 = 0.5f;
                            x0 += a.xyz_cam[0] * p0; x1 += a.xyz_cam[3] * p0; x2 += a.xyz_cam[6] * p0;
                            x0 += a.xyz_cam[1] * p1; x1 += a.xyz_cam[4] * p1; x2 += a.xyz_cam[7] * p1;
                            x0 += a.xyz_cam[2] * p2; x1 += a.xyz_cam[5] * p2; x2 += a.xyz_cam[8] * p2;
                            fx = lut_i(a.cbrt, (int)x0); fy = lut_i(a.cbrt, (int)x1); fz = lut_i(a.cbrt, (int)x2);
                            L = 116 * fy - 16;
                        }
                        px[0] = L; px[1] = 500.f * (fx - fy); px[2] = 200.f * (fy - fz);
                    }
                } else {
                    const int nr = rows - 8, nc = cols - 8;
                    XT_FOR2(r, c, nr, nc) {
                        float* px = plane + ((4 + r) * TS + 4 + c) * 3;
                        const float p0 = px[0], p1 = px[1], p2 = px[2];
                        const float y = 0.2627f * p0 + 0.6780f * p1 + 0.0593f * p2;
                        px[0] = y; px[1] = (p2 - y) * 0.56433f; px[2] = (p0 - y) * 0.67815f;
                    }
                }
                __syncthreads();
                {
                    const int q = dd & 3;           // dir[d & 3] = {1, ts, ts + 1, ts - 1} as a (row, column) step
                    const int fo = (q == 0 ? 1 : (q == 1 ? TS : (q == 2 ? TS + 1 : TS - 1))) * 3;
                    const int nr = rows - 10, nc = cols - 10;
                    XT_FOR2(ri, ci, nr, nc) {
                        const int r = 5 + ri, c = 5 + ci;
                        const float* l = plane + (r * TS + c) * 3;
                        float v;
                        if (a.useCieLab) {
                            const float g = 2 * l[0] - l[fo] - l[-fo];
                            v = sqrf(g) + sqrf((2 * l[1] - l[fo + 1] - l[-fo + 1] + g * 2.1551724f)) + sqrf((2 * l[2] - l[fo + 2] - l[-fo + 2] - g * 0.86206896f));
                        } else {
                            v = sqrf(2 * l[0] - l[fo] - l[-fo]) + sqrf(2 * l[1] - l[fo + 1] - l[-fo + 1]) + sqrf(2 * l[2] - l[fo + 2] - l[-fo + 2]);
                        }
                        drv[(size_t)dd * DW * DW + (r - 5) * DW + (c - 5)] = v;
                    }
                }
                if (dd == ndir - 1) {               // the lab buffer's final content
                    const int nr = rows - 8, nc = a.useCieLab ? LW : cols - 8;
                    XT_FOR2(r, c, nr, nc) {
                        const float* px = plane + ((4 + r) * TS + 4 + c) * 3;
                        lab[r * LW + c] = px[0]; lab[LW * LW + r * LW + c] = px[1]; lab[2 * LW * LW + r * LW + c] = px[2];
                    }
                }
                __syncthreads();
            };
            for (int p = 0; p < 4; ++p) {
                {
                    const float4* s4 = reinterpret_cast<const float4*>(buffer + (size_t)p * TS * TS * 3);
                    float4* d4 = reinterpret_cast<float4*>(plane);
                    for (int i = tid; i < TS * TS * 3 / 4; i += XT_THREADS) d4[i] = s4[i];
                }
                __syncthreads();
                for (int pass = 0; pass < a.passes; ++pass) {
                    if (pass) {             // recalculate green from interpolated values of closer pixels, L483-522
                        const int nr = rows - 4, nc = cols - 4;
                        XT_FOR2(ri, ci, nr, nc) {
                            const int r = 2 + ri, c = 2 + ci, row = top + r, col = left + c;
                            const int f = fcol(tb, row, col);
                            if (f == 1) continue;
                            const int q = p ^ (tb.rshift[row % 3] ? 0 : 1);          // plane (d - 2) ^ flip == p  <=>  d = q + 2, d in 3 .. 5
                            if (q == 0) continue;
                            const int d = q + 2;
                            const int hx = (tb.hexh[row % 3][col % 3][d] + tb.hexv[row % 3][col % 3][d] * TS) * 3;
                            const float2 mm = sgmm[r * TSH + (c >> 1)];
                            float* rix = plane + (r * TS + c) * 3;
                            const float val = 0.33333333f * (rix[-2 * hx + 1] + 2 * (rix[hx + 1] - rix[hx + f]) - rix[-2 * hx + f]) + rix[f];
                            rix[1] = limf(val, mm.x, mm.y);
                        }
                        __syncthreads();
                    }
                    // red and blue for solitary green pixels, L524-561: plane 0 <- d = 0, plane 1 <- d = 1, plane 2 <- d = 2, 3, plane 3 <- d = 4, 5
                    XT_FOR2(ri, ci, sol_nr, sol_nc) {
                        const int row = sol_row0 + 3 * ri, col = sol_col0 + 3 * ci;
                        const int h0 = fcol(tb, row, col + 1);
                        float* rix = plane + ((row - top) * TS + (col - left)) * 3;
                        const int dfirst = p < 2 ? p : 2 * p - 2, dlast = p < 2 ? p : 2 * p - 1;
                        float color[3][2], diff[2] = {0.f, 0.f};
                        for (int d = dfirst; d <= dlast; ++d) {
                            const int k = d - dfirst;
                            int h = h0 ^ ((d & 1) << 1);
                            const int o1 = (d & 1) ? TS : 1;
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int o = (o1 << c) * 3;
                                const float g = rix[1] + rix[1] - rix[o + 1] - rix[-o + 1];
                                color[h][k] = g + rix[o + h] + rix[-o + h];
                                if (d > 1) diff[k] += sqrf(rix[o + 1] - rix[-o + 1] - rix[o + h] + rix[-o + h]) + sqrf(g);
                                h ^= 2;
                            }
                            if (d > 2 && (d & 1))
                                if (diff[0] < diff[1]) { color[0][1] = color[0][0]; color[2][1] = color[2][0]; }
                            if ((d & 1) || d < 2) {
                                rix[0] = 0.5f * color[0][k];
                                rix[2] = 0.5f * color[2][k];
                            }
                        }
                    }
                    __syncthreads();
                    {   // red for blue pixels and vice versa, L563-603
                        const int nr = rows - 6, nc = cols - 6;
                        XT_FOR2(ri, ci, nr, nc) {
                            const int r = 3 + ri, c = 3 + ci, row = top + r, col = left + c;
                            const int fc = fcol(tb, row, col);
                            if (fc == 1) continue;
                            const int f = 2 - fc;
                            const int cs = ((row - a.sgrow) % 3) ? TS : 1;
                            const int hs = 3 * (cs ^ TS ^ 1);
                            const int co = cs * 3, ho = hs * 3;
                            float* rix = plane + (r * TS + c) * 3;
                            const bool usec = p > 1 || ((p ^ cs) & 1) ||
                                              ((fabsf(rix[1] - rix[co + 1]) + fabsf(rix[1] - rix[-co + 1])) < 2.f * (fabsf(rix[1] - rix[ho + 1]) + fabsf(rix[1] - rix[-ho + 1])));
                            const int io = usec ? co : ho;
                            rix[f] = rix[1] + 0.5f * (rix[io + f] + rix[-io + f] - rix[io + 1] - rix[-io + 1]);
                        }
                    }
                    __syncthreads();
                    if (2 * p < ndir) {     // red and blue for the 2x2 blocks of green, L605-650: `for (d = 0; d < ndir; d += 2)` reaches planes d / 2
                        const int nr = rows - 4, nc = cols - 4;
                        XT_FOR2(ri, ci, nr, nc) {
                            const int r = 2 + ri, c = 2 + ci, row = top + r, col = left + c;
                            if (!((row - a.sgrow) % 3) || !((col - a.sgcol) % 3)) continue;
                            const signed char* hv = tb.hexv[row % 3][col % 3];
                            const signed char* hh = tb.hexh[row % 3][col % 3];
                            float* rix = plane + (r * TS + c) * 3;
                            const int d = 2 * p;
                            const int h0 = hh[d] + hv[d] * TS, h1 = hh[d + 1] + hv[d + 1] * TS;
                            const float* p0 = rix + h0 * 3;
                            const float* p1 = rix + h1 * 3;
                            if (h0 + h1) {
                                const float g = 3 * rix[1] - 2 * p0[1] - p1[1];
                                rix[0] = (g + 2 * p0[0] + p1[0]) * 0.33333333f;
                                rix[2] = (g + 2 * p0[2] + p1[2]) * 0.33333333f;
                            } else {
                                const float g = 2 * rix[1] - p0[1] - p1[1];
                                rix[0] = (g + p0[0] + p1[0]) * 0.5f;
                                rix[2] = (g + p0[2] + p1[2]) * 0.5f;
                            }
                        }
                        __syncthreads();
                    }
                    if (pass == 0 || pass == a.passes - 1) {
                        float4* d4 = reinterpret_cast<float4*>(buffer + (size_t)(pass == 0 ? p : p + 4) * TS * TS * 3);
                        const float4* s4 = reinterpret_cast<const float4*>(plane);
                        for (int i = tid; i < TS * TS * 3 / 4; i += XT_THREADS) d4[i] = s4[i];
                        __syncthreads();
                    }
                }
                if (a.stop == 2) continue;
                if (a.passes > 1) {         // shared memory holds direction p + 4; direction p comes back from the slab afterwards
                    labdrv(p + 4);
                    const float4* s4 = reinterpret_cast<const float4*>(buffer + (size_t)p * TS * TS * 3);
                    float4* d4 = reinterpret_cast<float4*>(plane);
                    for (int i = tid; i < TS * TS * 3 / 4; i += XT_THREADS) d4[i] = s4[i];
                    __syncthreads();
                }
                labdrv(p);
            }
        }

        rgb = buffer;
        mrow = rows;
        mcol = cols;
        if (a.stop == 2) { __syncthreads(); continue; }

        if (a.stop == 3) { __syncthreads(); continue; }
        // The homogeneity maps and their 5x5 sums live in shared memory (2 x ndir x 114 x 114 bytes: exactly the space the direction
        // plane and the min/max table used).  The maps are seeded with the bytes of the slab's lab area, which is what the
        // reference's `homo` alias holds wherever the homogeneity step does not write.
        unsigned char* const shomo = reinterpret_cast<unsigned char*>(xsm);
        unsigned char* const shsum = shomo + (size_t)ndir * TS * TS;
        {
            const uint4* s4 = reinterpret_cast<const uint4*>(homo);
            uint4* d4 = reinterpret_cast<uint4*>(shomo);
            for (int i = tid; i < ndir * TS * TS / 16; i += XT_THREADS) d4[i] = s4[i];
        }
        __syncthreads();
        {   // homogeneity maps, L743-811
            const int nr = mrow - 12, nc = mcol - 12;
            XT_FOR2(ri, ci, nr, nc) {
                const int r = 6 + ri, c = 6 + ci;
                const float* dp = drv + (r - 5) * DW + (c - 5);
                float tr = dp[0] < dp[DW * DW] ? dp[0] : dp[DW * DW];
                for (int d = 2; d < ndir; ++d) tr = dp[(size_t)d * DW * DW] < tr ? dp[(size_t)d * DW * DW] : tr;
                tr *= 8;
                for (int d = 0; d < ndir; ++d) {
                    const float* q = dp + (size_t)d * DW * DW;
                    int cnt = 0;
#pragma unroll
                    for (int v = -1; v <= 1; ++v)
#pragma unroll
                        for (int h = -1; h <= 1; ++h) cnt += q[v * DW + h] <= tr ? 1 : 0;
                    shomo[(size_t)d * TS * TS + r * TS + c] = (unsigned char)cnt;
                }
            }
        }
        __syncthreads();

        if (a.stop == 4) { __syncthreads(); continue; }
        if (H - top < TS + 4) mrow = H - top + 2;
        if (W - left < TS + 4) mcol = W - left + 2;
        const int startrow = min(top, 8), startcol = min(left, 8);

        {   // 5x5 sums of the homogeneity maps, L822-868: saturating 16-wide groups, wrapping scalar tail on the last row
            const int nr = mrow - 8 - startrow, nc = mcol - 8 - startcol;
            const int n = nr > 0 && nc > 0 ? nr * nc : 0;
            for (int i = tid; i < n * ndir; i += XT_THREADS) {
                const int d = i / n, j = i - d * n;
                const int r = startrow + j / nc, c = startcol + j % nc;
                const int endcol = r < mrow - 9 ? mcol - 8 : mcol - 23;
                const int vec_end = endcol > startcol ? startcol + (endcol - startcol + 15) / 16 * 16 : startcol;
                const unsigned char* base = shomo + (size_t)d * TS * TS + r * TS + c;
                int sum = 0;
#pragma unroll
                for (int v = -2; v <= 2; ++v)
#pragma unroll
                    for (int h = -2; h <= 2; ++h) sum += base[v * TS + h];
                shsum[(size_t)d * TS * TS + r * TS + c] = (unsigned char)(c < vec_end ? min(sum, 255) : sum);
            }
        }
        __syncthreads();

        if (a.stop == 5) { __syncthreads(); continue; }
        {   // maximum minus an eighth (L870-911) and the average of the most homogeneous directions (L914-949)
            const int nr = mrow - 8 - startrow, nc = mcol - 8 - startcol;
            XT_FOR2(ri, ci, nr, nc) {
                const int r = startrow + ri, c = startcol + ci;
                unsigned char hm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const unsigned char* hs = shsum + r * TS + c;
                unsigned char maxval = hs[0];
                for (int d = 1; d < ndir; ++d) { const unsigned char v = hs[(size_t)d * TS * TS]; maxval = maxval < v ? v : maxval; }
                maxval -= maxval >> 3;
#pragma unroll
                for (int d = 0; d < 4; ++d) hm[d] = hs[(size_t)d * TS * TS];
                for (int d = 4; d < ndir; ++d) {
                    hm[d] = hs[(size_t)d * TS * TS];
                    if (hm[d - 4] < hm[d]) hm[d - 4] = 0;
                    else if (hm[d - 4] > hm[d]) hm[d] = 0;
                }
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int d = 0; d < ndir; ++d)
                    if (hm[d] >= maxval) {
                        const float* p = &RGB(d, r, c, 0);
                        a0 += p[0]; a1 += p[1]; a2 += p[2];
                        a3 += 1.f;
                    }
                const size_t o = (size_t)(r + top) * a.op + c + left;
                const float vr = a0 / a3, vg = a1 / a3, vb = a2 / a3;
                a.R[o] = 0.f < vr ? vr : 0.f;
                a.G[o] = 0.f < vg ? vg : 0.f;
                a.B[o] = 0.f < vb ? vb : 0.f;
            }
        }
        __syncthreads();
    }
}

// xtransborder_interpolate, L122-173: one thread per pixel of the outer ring
__global__ void k_xtrans_border(const XtArgs a, int border)
{
    const int W = a.W, H = a.H;
    const long ring_top = (long)border * W, side = (long)(H - 2 * border) * 2 * border;
    const long n = 2 * ring_top + side;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int row, col;
    if (i < ring_top) { row = (int)(i / W); col = (int)(i % W); }
    else if (i < ring_top + side) {
        const long j = i - ring_top;
        row = border + (int)(j / (2 * border));
        const int k = (int)(j % (2 * border));
        col = k < border ? k : W - 2 * border + k;
    } else { const long j = i - ring_top - side; row = H - border + (int)(j / W); col = (int)(j % W); }
    float sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int y = max(0, row - 1), v = row == 0 ? 0 : -1; y <= min(row + 1, H - 1); ++y, ++v)
        for (int x = max(0, col - 1), h = col == 0 ? 0 : -1; x <= min(col + 1, W - 1); ++x, ++h) {
            const float w = (v == 0 && h == 0) ? 0.f : ((v == 0 || h == 0) ? 0.5f : 0.25f);
            const int f = a.xt[y % 6][x % 6];
            sum[f] += a.raw[(size_t)y * a.rp + x] * w;
            sum[f + 3] += w;
        }
    const size_t o = (size_t)row * a.op + col;
    const float p = a.raw[(size_t)row * a.rp + col];
    switch (a.xt[row % 6][col % 6]) {
    case 0: a.R[o] = p; a.G[o] = sum[1] / sum[4]; a.B[o] = sum[2] / sum[5]; break;
    case 1:
        if (sum[3] == 0.f) { a.R[o] = p; a.G[o] = p; a.B[o] = p; }
        else { a.R[o] = sum[0] / sum[3]; a.G[o] = p; a.B[o] = sum[2] / sum[5]; }
        break;
    default: a.R[o] = sum[0] / sum[3]; a.G[o] = sum[1] / sum[4]; a.B[o] = p;
    }
}

}  // namespace

int art_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int* xtrans36, const float* rgb_cam12,
                   const float* raw, size_t rp, float* R, float* G, float* B, size_t op)
{
    static const short orth[12] = {1, 0, 0, 1, -1, 0, 0, -1, 1, 0, 0, 1};
    static const short patt[2][16] = {{0, 1, 0, -1, 2, 0, -1, 0, 1, 1, 1, -1, 0, 0, 0, 0}, {0, 1, 0, -2, 1, 0, -2, 0, 1, 1, -2, -2, 1, -1, -1, 1}};
    static const float xyz_rgb[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
    static const float d65_white[3] = {0.950456, 1, 1.088754};
    cudaStream_t st = ctx->stream;
    XtArgs a{};
    a.raw = raw; a.rp = rp; a.R = R; a.G = G; a.B = B; a.op = op; a.W = W; a.H = H;
    a.passes = passes; a.ndir = 4 << (passes > 1); a.useCieLab = useCieLab;
    { const char* e = getenv("ART_XT_STOP"); a.stop = e ? atoi(e) : 0; }
    for (int i = 0; i < 36; ++i) a.xt[i / 6][i % 6] = (unsigned char)xtrans36[i];
    for (int i = 0; i < 3; i++)          // L224-232, float arithmetic
        for (int j = 0; j < 3; j++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += xyz_rgb[i][k] * rgb_cam12[k * 4 + j] / d65_white[i];
            a.xyz_cam[i * 3 + j] = s;
        }
    auto green = [&](int r, int c) { return a.xt[r % 3][c % 3] & 1; };
    for (int row = 0; row < 3; row++)    // L235-266
        for (int col = 0; col < 3; col++) {
            const int gint = green(row, col);
            for (int ng = 0, d = 0; d < 10; d += 2) {
                if (green(row + orth[d] + 6, col + orth[d + 2] + 6)) ng = 0; else ng++;
                if (ng == 4) { a.sgrow = row; a.sgcol = col; }
                if (ng == gint + 1)
                    for (int c = 0; c < 8; c++) {
                        a.hexv[row][col][c ^ (gint * 2 & d)] = (signed char)(orth[d] * patt[gint][c * 2] + orth[d + 1] * patt[gint][c * 2 + 1]);
                        a.hexh[row][col][c ^ (gint * 2 & d)] = (signed char)(orth[d + 2] * patt[gint][c * 2] + orth[d + 3] * patt[gint][c * 2 + 1]);
                    }
            }
        }
    for (int row = 0; row < 3; row++) {  // L280-291
        int greencount = 0;
        for (int col = 0; col < 3; col++) greencount += green(row, col);
        a.rshift[row] = (greencount == 2);
    }
    int rc;
    if (!ctx->xt_cbrt_ready) {           // cielab's cbrt LUT, L44-63 (host libm cbrt, as in the reference)
        if ((rc = art_reserve(ctx, ctx->d_xt_cbrt, CBRT_N * sizeof(float)))) return rc;
        std::vector<float> t(CBRT_N);
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0;
        for (int i = 0; i < CBRT_N; i++) {
            const double r = i / 65535.0;
            t[i] = (float)(r > eps ? std::cbrt(r) : (kappa * r + 16.0) / 116.0);
        }
        ART_CUDA(ctx, cudaMemcpyAsync(ctx->d_xt_cbrt.p, t.data(), CBRT_N * sizeof(float), cudaMemcpyHostToDevice, st));
        ART_CUDA(ctx, cudaStreamSynchronize(st));      // t goes out of scope
        ctx->xt_cbrt_ready = true;
    }
    a.cbrt = (const float*)ctx->d_xt_cbrt.p;
    a.nty = H - 19 > 3 ? (H - 19 - 3 + (TS - 16) - 1) / (TS - 16) : 0;
    a.ntx = W - 19 > 3 ? (W - 19 - 3 + (TS - 16) - 1) / (TS - 16) : 0;
    const int ntiles = a.ntx * a.nty;
    if (ntiles > 0) {
        a.slab_floats = round_up((size_t)TS * TS * (a.ndir * 4 + 3) + 128, 32);
        const int grid = std::min(ntiles, ctx->sm_count);       // one 1024-thread CTA (208 KB of shared memory) per SM
        if (!(ctx->attrs_set & art_hp_ctx::ATTR_XTRANS)) {
            ART_CUDA(ctx, cudaFuncSetAttribute(k_xtrans, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XT_SMEM));
            ctx->attrs_set |= art_hp_ctx::ATTR_XTRANS;
        }
        if ((rc = art_reserve(ctx, ctx->d_scratch, (size_t)grid * a.slab_floats * sizeof(float)))) return rc;
        a.slabs = (float*)ctx->d_scratch.p;
        art_prof_begin(ctx, "k_xtrans");
        k_xtrans<<<grid, XT_THREADS, XT_SMEM, st>>>(a);
        art_prof_end(ctx);
        ctx->launches++;
    }
    const int border = passes > 1 ? 8 : 11;
    {
        const long n = 2L * border * W + (long)std::max(0, H - 2 * border) * 2 * border;
        art_prof_begin(ctx, "k_xtrans_border");
        k_xtrans_border<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, border);
        art_prof_end(ctx);
        ctx->launches++;
    }
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
