// AMaZE demosaic for sm_100a.
//
// Replaces RawImageSource::amaze_demosaic_RT (reference rtengine/amaze_demosaic_RT.cc L41-1595), the
// SSE2 code path (what x86-64 builds execute).  See DESIGN.md "AMaZE" for the full rationale.
//
// The reference result depends on its 160x160 tile grid (stride 128, origin -16): three passes update a
// buffer in place while sweeping down the tile, the vector loops touch columns outside their nominal
// range, and the per-tile scratch sub-buffers alias each other so that a few passes read bytes last
// written under another name.  Parity therefore needs the same grid, the same lane groups and the same
// scratch layout.  What is re-designed is the execution: instead of one thread walking one tile through
// ~20 passes, every pass is a kernel over ALL tiles of a band ("tile space": one 1.45 MB slab per tile
// in HBM/L2, identical layout to the reference's per-thread block), one thread per lane.  The three
// in-place passes are row recurrences that are parallel across columns and tiles:
//   * hcd (L536-566): within a row, lanes 2,3 of a vector see old neighbours, lanes 0,1 see the
//     previous vector's new lanes 2,3 -> one thread per (lanes 2,3 | next lanes 0,1) unit, rows parallel;
//   * vcd (L568-577): recurrence down rows of equal parity, columns independent -> one thread per
//     (column, parity) marching down the tile with the previous result in a register;
//   * hvwt (L958-964) and pmwt (L1213-1221): row r needs row r-1's new values at the diagonal
//     neighbours -> one warp per tile marching down, rows exchanged through shared memory.
// Compiled with -fmad=false, IEEE division: bit-identical to the reference build.
#include "ctx.h"

#include <climits>

namespace {

constexpr int TS = 160, TSH = 80;
constexpr int v1 = TS, v2 = 2 * TS, v3 = 3 * TS, p1 = -TS + 1, p2 = -2 * TS + 2, p3 = -3 * TS + 3,
              m1 = TS + 1, m2 = 2 * TS + 2, m3 = 3 * TS + 3;

// scratch slab layout, bytes (amaze_demosaic_RT.cc L124-174; cldf = 2 -> 128-byte gaps)
constexpr size_t FULL = (size_t)TS * TS * 4, HALF = (size_t)TS * TSH * 4, GAP = 128;
constexpr size_t OFF_RGBGREEN = 0;
constexpr size_t OFF_DELHVSQSUM = OFF_RGBGREEN + FULL + GAP;
constexpr size_t OFF_DIRWTS0 = OFF_DELHVSQSUM + FULL + GAP;
constexpr size_t OFF_DIRWTS1 = OFF_DIRWTS0 + FULL + GAP;
constexpr size_t OFF_VCD = OFF_DIRWTS1 + FULL + GAP;
constexpr size_t OFF_HCD = OFF_VCD + FULL + GAP;
constexpr size_t OFF_VCDALT = OFF_HCD + FULL + GAP;
constexpr size_t OFF_HCDALT = OFF_VCDALT + FULL + GAP;
constexpr size_t OFF_CDDIFFSQ = OFF_HCDALT + FULL + GAP;
constexpr size_t OFF_HVWT = OFF_CDDIFFSQ + FULL + 2 * GAP;
constexpr size_t OFF_DGINTV = OFF_HVWT + HALF + GAP;
constexpr size_t OFF_DGINTH = OFF_DGINTV + FULL + GAP;
constexpr size_t OFF_DGRBSQ1M = OFF_DGINTH + FULL + GAP;
constexpr size_t OFF_DGRBSQ1P = OFF_DGRBSQ1M + HALF + GAP;
constexpr size_t OFF_CFA = OFF_DGRBSQ1P + HALF + GAP;
constexpr size_t OFF_NYQUIST = OFF_CFA + FULL + GAP;
constexpr size_t OFF_NYQUTEST = OFF_NYQUIST + (size_t)TS * TSH + GAP;
constexpr size_t SLAB_BYTES = 14 * (size_t)4 * TS * TS + (size_t)TS * TSH + 18 * GAP;   // 1,448,704 (multiple of 256)
static_assert(SLAB_BYTES % 256 == 0, "slab stride must keep 128-byte alignment");

struct Slab {
    char* base;
    __device__ __forceinline__ float* f(size_t off) const { return reinterpret_cast<float*>(base + off); }
    __device__ __forceinline__ float* rgbgreen() const { return f(OFF_RGBGREEN); }
    __device__ __forceinline__ float* delhvsqsum() const { return f(OFF_DELHVSQSUM); }
    __device__ __forceinline__ float* dirwts0() const { return f(OFF_DIRWTS0); }
    __device__ __forceinline__ float* dirwts1() const { return f(OFF_DIRWTS1); }
    __device__ __forceinline__ float* vcd() const { return f(OFF_VCD); }
    __device__ __forceinline__ float* hcd() const { return f(OFF_HCD); }
    __device__ __forceinline__ float* vcdalt() const { return f(OFF_VCDALT); }
    __device__ __forceinline__ float* hcdalt() const { return f(OFF_HCDALT); }
    __device__ __forceinline__ float* cddiffsq() const { return f(OFF_CDDIFFSQ); }
    __device__ __forceinline__ float* hvwt() const { return f(OFF_HVWT); }
    __device__ __forceinline__ float* Dgrb0() const { return f(OFF_VCDALT); }                    // L148
    __device__ __forceinline__ float* Dgrb1() const { return f(OFF_VCDALT) + TS * TSH; }
    __device__ __forceinline__ float* delp() const { return f(OFF_CDDIFFSQ); }                   // L150
    __device__ __forceinline__ float* delm() const { return f(OFF_CDDIFFSQ + HALF + GAP); }      // L152 (= rbint, L154)
    __device__ __forceinline__ float* rbint() const { return delm(); }
    __device__ __forceinline__ float* dgintv() const { return f(OFF_DGINTV); }
    __device__ __forceinline__ float* Dgrb2() const { return f(OFF_DGINTV); }                    // L156: {h,v} pairs
    __device__ __forceinline__ float* dginth() const { return f(OFF_DGINTH); }
    __device__ __forceinline__ float* Dgrbsq1m() const { return f(OFF_DGRBSQ1M); }
    __device__ __forceinline__ float* Dgrbsq1p() const { return f(OFF_DGRBSQ1P); }
    __device__ __forceinline__ float* cfa() const { return f(OFF_CFA); }
    __device__ __forceinline__ float* pmwt() const { return f(OFF_DELHVSQSUM); }                 // L167
    __device__ __forceinline__ float* rbm() const { return f(OFF_VCD); }                         // L169
    __device__ __forceinline__ float* rbp() const { return f(OFF_VCD + HALF + GAP); }            // L170
    __device__ __forceinline__ unsigned char* nyquist() const { return reinterpret_cast<unsigned char*>(base + OFF_NYQUIST); }
    __device__ __forceinline__ unsigned char* nyquist2() const { return reinterpret_cast<unsigned char*>(base + OFF_CDDIFFSQ); }  // L173
};

struct AmzArgs {
    const float* raw; size_t rp;
    float *R, *G, *B; size_t op;
    int W, H; unsigned filters;
    int ntx;          // tiles per tile-row
    int ty0;          // first tile-row of this band
    int ntiles;       // tiles in this band
    char* slabs;      // ntiles slabs
    int* bbox;        // ntiles x 4: min row, max row, min col, max col of nyquist flags
    float clip_pt, clip_pt8;
    int ex, ey;
};

struct Geo { int top, left, rr1, cc1, rrmin, ccmin, rrmax, ccmax; };

__device__ __forceinline__ Geo geo(const AmzArgs& a, int t)
{
    Geo g;
    const int ty = a.ty0 + t / a.ntx, tx = t % a.ntx;
    g.top = -16 + ty * (TS - 32);
    g.left = -16 + tx * (TS - 32);
    const int bottom = min(g.top + TS, a.H + 16), right = min(g.left + TS, a.W + 16);   // L186-192
    g.rr1 = bottom - g.top;
    g.cc1 = right - g.left;
    g.rrmin = g.top < 0 ? 16 : 0;                                                      // L195-198
    g.ccmin = g.left < 0 ? 16 : 0;
    g.rrmax = bottom > a.H ? a.H - g.top : g.rr1;
    g.ccmax = right > a.W ? a.W - g.left : g.cc1;
    return g;
}
__device__ __forceinline__ Slab slab(const AmzArgs& a, int t) { return Slab{a.slabs + (size_t)t * SLAB_BYTES}; }

__device__ __forceinline__ unsigned fc(unsigned filters, int row, int col)
{   // RawImage::FC, rtengine/rawimage.h L186-189
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}
__device__ __forceinline__ float sq(float x) { return x * x; }
__device__ __forceinline__ float vminf(float a, float b) { return a < b ? a : b; }     // _mm_min_ps
__device__ __forceinline__ float vmaxf(float a, float b) { return a > b ? a : b; }     // _mm_max_ps
__device__ __forceinline__ float vintpf(float a, float b, float c) { return a * b + (1.f - a) * c; }   // sleefsseavx.h L1435
__device__ __forceinline__ float vmedian(float a, float b, float c) { return vmaxf(vminf(a, b), vminf(c, vmaxf(a, b))); }
__device__ __forceinline__ float stdmax(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ int sat8(int x) { return x > 127 ? 127 : (x < -128 ? -128 : x); }

constexpr float eps = 1e-5f, epssq = 1e-10f, arthresh = 0.75f;

// ------------------------------------------------------------------ pass 0: tile fill (L206-334)
// The nine loops of the reference in their order, a __syncthreads between loops that may overwrite
// each other.  Flat indices are used on purpose: border rows >= 160 run into the bytes after
// cfa / rgbgreen and border columns >= 160 wrap into the next row exactly as in the reference.
__global__ void __launch_bounds__(256) k_fill(AmzArgs a)
{
    const int t = blockIdx.x;
    const Geo g = geo(a, t);
    const Slab s = slab(a, t);
    float* cfa = s.cfa();
    float* rgbgreen = s.rgbgreen();
    const int W = a.W, H = a.H;
    const int tid = threadIdx.x, nt = blockDim.x;
#define PUT(idx, r, c) do { const float v_ = a.raw[(size_t)(r) * a.rp + (c)] / 65535.f; cfa[idx] = v_; rgbgreen[idx] = v_; } while (0)
    const int wc = g.ccmax - g.ccmin;
    if (g.rrmin > 0)
        for (int i = tid; i < 16 * wc; i += nt) { const int rr = i / wc, cc = g.ccmin + i % wc; PUT(rr * TS + cc, 32 - rr + g.top, cc + g.left); }
    for (int i = tid; i < (g.rrmax - g.rrmin) * wc; i += nt) { const int rr = g.rrmin + i / wc, cc = g.ccmin + i % wc; PUT(rr * TS + cc, rr + g.top, cc + g.left); }
    if (g.rrmax < g.rr1)
        for (int i = tid; i < 16 * wc; i += nt) { const int rr = i / wc, cc = g.ccmin + i % wc; PUT((g.rrmax + rr) * TS + cc, H - rr - 2, g.left + cc); }
    __syncthreads();
    const int hr = g.rrmax - g.rrmin;
    if (g.ccmin > 0)
        for (int i = tid; i < hr * 16; i += nt) { const int rr = g.rrmin + i / 16, cc = i % 16; PUT(rr * TS + cc, rr + g.top, 32 - cc + g.left); }
    __syncthreads();
    if (g.ccmax < g.cc1)
        for (int i = tid; i < hr * 16; i += nt) { const int rr = g.rrmin + i / 16, cc = i % 16; PUT(rr * TS + g.ccmax + cc, g.top + rr, W - cc - 2); }
    __syncthreads();
    if (g.rrmin > 0 && g.ccmin > 0)
        for (int i = tid; i < 256; i += nt) { const int rr = i / 16, cc = i % 16; PUT(rr * TS + cc, 32 - rr, 32 - cc); }
    __syncthreads();
    if (g.rrmax < g.rr1 && g.ccmax < g.cc1)
        for (int i = tid; i < 256; i += nt) { const int rr = i / 16, cc = i % 16; PUT((g.rrmax + rr) * TS + g.ccmax + cc, H - rr - 2, W - cc - 2); }
    __syncthreads();
    if (g.rrmin > 0 && g.ccmax < g.cc1)
        for (int i = tid; i < 256; i += nt) { const int rr = i / 16, cc = i % 16; PUT(rr * TS + g.ccmax + cc, 32 - rr, W - cc - 2); }
    __syncthreads();
    if (g.rrmax < g.rr1 && g.ccmin > 0)
        for (int i = tid; i < 256; i += nt) { const int rr = i / 16, cc = i % 16; PUT((g.rrmax + rr) * TS + cc, H - rr - 2, 32 - cc); }
#undef PUT
    if (tid < 4) a.bbox[4 * t + tid] = (tid == 0 || tid == 2) ? INT_MAX : 0;
}

// Full-resolution passes: block (160 columns, 4 rows); grid.x = row chunks, grid.y = tile.
constexpr int FR_ROWS = 4;
// R/B-site passes: block (80 half-columns, 4 rows).
constexpr int HR_ROWS = 4;

// ------------------------------------------------------------------ pass 1: gradients (L342-350)
__global__ void __launch_bounds__(TS * FR_ROWS) k_grad(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 2 + blockIdx.x * FR_ROWS + threadIdx.y, cc = threadIdx.x;
    if (rr >= g.rr1 - 2 || (cc & ~3) >= g.cc1) return;          // vector groups start at 0,4,.. < cc1
    const Slab s = slab(a, t);
    const float* cfa = s.cfa();
    const int i = rr * TS + cc;
    const float c0 = cfa[i];
    const float delh = fabsf(cfa[i + 1] - cfa[i - 1]);
    const float delv = fabsf(cfa[i + v1] - cfa[i - v1]);
    s.dirwts1()[i] = eps + fabsf(cfa[i + 2] - c0) + fabsf(c0 - cfa[i - 2]) + delh;
    s.dirwts0()[i] = eps + fabsf(cfa[i + v2] - c0) + fabsf(c0 - cfa[i - v2]) + delv;
    s.delhvsqsum()[i] = sq(delh) + sq(delv);
}

// ------------------------------------------------------------------ pass 2: directional G estimates (L380-431)
__global__ void __launch_bounds__(TS * FR_ROWS) k_dirinterp(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 4 + blockIdx.x * FR_ROWS + threadIdx.y, cc = threadIdx.x;
    if (rr >= g.rr1 - 4 || cc < 4 || (cc & ~3) >= g.cc1 - 7) return;   // groups start at 4,8,.. < cc1-7
    const Slab s = slab(a, t);
    const float* cfa = s.cfa();
    const float* dirwts0 = s.dirwts0();
    const float* dirwts1 = s.dirwts1();
    const int i = rr * TS + cc;
    const float sgn = (fc(a.filters, rr, cc) & 1) ? -1.f : 1.f;        // sgnv: +1 at R/B sites, -1 at G sites
    const float cfav = cfa[i];
    const float cru = cfa[i - v1] * (dirwts0[i - v2] + dirwts0[i]) / (dirwts0[i - v2] * (eps + cfav) + dirwts0[i] * (eps + cfa[i - v2]));
    const float crd = cfa[i + v1] * (dirwts0[i + v2] + dirwts0[i]) / (dirwts0[i + v2] * (eps + cfav) + dirwts0[i] * (eps + cfa[i + v2]));
    const float crl = cfa[i - 1] * (dirwts1[i - 2] + dirwts1[i]) / (dirwts1[i - 2] * (eps + cfav) + dirwts1[i] * (eps + cfa[i - 2]));
    const float crr = cfa[i + 1] * (dirwts1[i + 2] + dirwts1[i]) / (dirwts1[i + 2] * (eps + cfav) + dirwts1[i] * (eps + cfa[i + 2]));
    const float guha = cfa[i - v1] + 0.5f * (cfav - cfa[i - v2]);
    const float gdha = cfa[i + v1] + 0.5f * (cfav - cfa[i + v2]);
    const float glha = cfa[i - 1] + 0.5f * (cfav - cfa[i - 2]);
    const float grha = cfa[i + 1] + 0.5f * (cfav - cfa[i + 2]);
    float guar = fabsf(1.f - cru) < arthresh ? cfav * cru : guha;
    float gdar = fabsf(1.f - crd) < arthresh ? cfav * crd : gdha;
    float glar = fabsf(1.f - crl) < arthresh ? cfav * crl : glha;
    float grar = fabsf(1.f - crr) < arthresh ? cfav * crr : grha;
    const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
    const float vwt = dirwts0[i - v1] / (dirwts0[i + v1] + dirwts0[i - v1]);
    const float Ginthha = vintpf(hwt, grha, glha);
    const float Gintvha = vintpf(vwt, gdha, guha);
    const float hcdaltv = sgn * (Ginthha - cfav);
    const float vcdaltv = sgn * (Gintvha - cfav);
    s.hcdalt()[i] = hcdaltv;
    s.vcdalt()[i] = vcdaltv;
    const bool clip = (cfav > a.clip_pt8) || (Gintvha > a.clip_pt8) || (Ginthha > a.clip_pt8);
    if (clip) { guar = guha; gdar = gdha; glar = glha; grar = grha; }
    s.vcd()[i] = clip ? vcdaltv : sgn * (vintpf(vwt, gdar, guar) - cfav);
    s.hcd()[i] = clip ? hcdaltv : sgn * (vintpf(hwt, grar, glar) - cfav);
    s.dgintv()[i] = vminf(sq(guha - gdha), sq(guar - gdar));
    s.dginth()[i] = vminf(sq(glha - grha), sq(glar - grar));
}

// one lane of the variance select + saturation bound (L542-566 for hcd, L546-577 for vcd):
// x = this lane's value, xm/xp its -2/+2 neighbours along the direction, same for the alt plane,
// c = cfa, cm/cp = cfa neighbours at -1/+1 along the direction.
__device__ __forceinline__ float bound_lane(float x, float xm, float xp, float alt, float altm, float altp,
                                            float c, float cm, float cp, float sgn, float clip_pt)
{
    const float nsgn = -sgn, sgn3 = sgn + sgn + sgn;
    const float var = sq(xm - x) + sq(xm - xp) + sq(x - xp);
    const float altvar = sq(altm - alt) + sq(altm - altp) + sq(alt - altp);
    x = altvar < var ? alt : x;
    const float Gint = sgn * x + c;
    const float temp2 = sgn3 * x;
    const float wt = 1.f + temp2 / (eps + Gint + c);
    const bool mask = nsgn * x > 0.f;
    const float old = x;
    const float temp = nsgn * (c - vmedian(Gint, cm, cp));
    x = (temp2 < -(c + Gint)) ? temp : vintpf(wt, x, temp);
    x = mask ? x : old;
    x = (Gint > clip_pt) ? temp : x;
    return x;
}

// ------------------------------------------------------------------ pass 3a: hcd in place (L536-566)
// thread = unit u of a row: columns a..a+3 with a = 2+4u: (lanes 2,3 of vector u-1 | lanes 0,1 of vector u)
__global__ void __launch_bounds__(40 * 4) k_hcd(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 4 + blockIdx.x * 4 + threadIdx.y, u = threadIdx.x;
    const int G = g.cc1 > 8 ? (g.cc1 - 8 + 3) / 4 : 0;          // vectors start at 4,8,.. < cc1-4
    const bool row_ok = rr < g.rr1 - 4;
    const bool has_prev = row_ok && u >= 1 && u - 1 < G;          // lanes 2,3 of vector u-1 (columns a, a+1)
    const bool has_cur = row_ok && u < G;                         // lanes 0,1 of vector u (columns a+2, a+3)
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
    const Slab s = slab(a, t);
    float* hcd = s.hcd();
    const int i = rr * TS + 2 + 4 * u;                            // column a
    if (has_prev || has_cur) {
        const float* hcdalt = s.hcdalt();
        const float* cfa = s.cfa();
        float h[8], al[8], c[6];
        #pragma unroll
        for (int k = 0; k < 8; ++k) { h[k] = hcd[i - 2 + k]; al[k] = hcdalt[i - 2 + k]; }   // columns a-2 .. a+5
        #pragma unroll
        for (int k = 0; k < 6; ++k) c[k] = cfa[i - 1 + k];                                 // columns a-1 .. a+4
        const int col = 2 + 4 * u;
        const float s0 = (fc(a.filters, rr, col) & 1) ? -1.f : 1.f, s1 = -s0;
        float l0 = h[2], l1 = h[3];                               // what lanes 0,1 of vector u see at columns a, a+1
        if (has_prev) {
            n0 = bound_lane(h[2], h[0], h[4], al[2], al[0], al[4], c[1], c[0], c[2], s0, a.clip_pt);
            n1 = bound_lane(h[3], h[1], h[5], al[3], al[1], al[5], c[2], c[1], c[3], s1, a.clip_pt);
            l0 = n0; l1 = n1;
        }
        if (has_cur) {
            n2 = bound_lane(h[4], l0, h[6], al[4], al[2], al[6], c[3], c[2], c[4], s0, a.clip_pt);
            n3 = bound_lane(h[5], l1, h[7], al[5], al[3], al[7], c[4], c[3], c[5], s1, a.clip_pt);
        }
    }
    __syncthreads();     // every old value of these rows has been read
    if (has_prev) { hcd[i] = n0; hcd[i + 1] = n1; }
    if (has_cur) { hcd[i + 2] = n2; hcd[i + 3] = n3; }
}

// ------------------------------------------------------------------ pass 3b: vcd recurrence + cddiffsq (L546-578)
// thread = (column, row parity) marching down the tile
__global__ void __launch_bounds__(TS * 2) k_vcd(AmzArgs a)
{
    const int t = blockIdx.x;
    const Geo g = geo(a, t);
    const int cc = threadIdx.x, par = threadIdx.y;
    const int G = g.cc1 > 8 ? (g.cc1 - 8 + 3) / 4 : 0;
    if (cc < 4 || cc >= 4 + 4 * G) return;
    const Slab s = slab(a, t);
    float* vcd = s.vcd();
    const float* vcdalt = s.vcdalt();
    const float* hcd = s.hcd();
    const float* cfa = s.cfa();
    float* cddiffsq = s.cddiffsq();
    int rr = 4 + par;
    if (rr >= g.rr1 - 4) return;
    int i = rr * TS + cc;
    float xm = vcd[i - v2];               // rows 2,3 are never updated: "new" == old there
    float x = vcd[i];
    float altm = vcdalt[i - v2], alt = vcdalt[i];
    // Only `xm` is carried from row to row: the operands of VCD_U rows are fetched together (the loads of a row never
    // alias a store of this thread's earlier rows: stores go to vcd[i] / cddiffsq[i], loads to vcd[i + v2] and read-only planes).
    constexpr int VCD_U = 8;
    for (; rr + 2 * (VCD_U - 1) < g.rr1 - 4; rr += 2 * VCD_U, i += v2 * VCD_U) {
        float xp[VCD_U], ap[VCD_U], c0[VCD_U], cu[VCD_U], cd[VCD_U], hh[VCD_U];
        #pragma unroll
        for (int k = 0; k < VCD_U; ++k) {
            const int ik = i + k * v2;
            xp[k] = vcd[ik + v2]; ap[k] = vcdalt[ik + v2];
            c0[k] = cfa[ik]; cu[k] = cfa[ik - v1]; cd[k] = cfa[ik + v1];
            hh[k] = hcd[ik];
        }
        #pragma unroll
        for (int k = 0; k < VCD_U; ++k) {
            const int ik = i + k * v2;
            const float sgn = (fc(a.filters, rr + 2 * k, cc) & 1) ? -1.f : 1.f;
            const float nv = bound_lane(x, xm, xp[k], alt, altm, ap[k], c0[k], cu[k], cd[k], sgn, a.clip_pt);
            vcd[ik] = nv;
            cddiffsq[ik] = sq(nv - hh[k]);
            xm = nv; x = xp[k]; altm = alt; alt = ap[k];
        }
    }
    for (; rr < g.rr1 - 4; rr += 2, i += v2) {
        const float xp = vcd[i + v2], altp = vcdalt[i + v2];
        const float sgn = (fc(a.filters, rr, cc) & 1) ? -1.f : 1.f;
        const float nv = bound_lane(x, xm, xp, alt, altm, altp, cfa[i], cfa[i - v1], cfa[i + v1], sgn, a.clip_pt);
        vcd[i] = nv;
        cddiffsq[i] = sq(nv - hcd[i]);
        xm = nv; x = xp; altm = alt; alt = altp;
    }
}

// R/B-site lane decoding shared by the half-resolution passes: lane j of a row whose vector loop is
// `for (indx = rr*ts + c0 + p; indx < rr*ts + cc1 - cend; indx += 8)`, 4 sites (stride 2) per vector.
__device__ __forceinline__ bool rb_lane(int j, int c0p, int bound, int* cc)
{
    const int start = c0p + 8 * (j >> 2);
    *cc = c0p + 2 * j;
    return start < bound;
}

// ------------------------------------------------------------------ pass 4: hvwt (L681-723)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_hvwt(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 6 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 6) return;
    int cc;
    if (!rb_lane(threadIdx.x, 6 + (fc(a.filters, rr, 2) & 1), g.cc1 - 6, &cc) || cc >= TS) return;
    const Slab s = slab(a, t);
    const float *vcd = s.vcd(), *hcd = s.hcd(), *dirwts0 = s.dirwts0(), *dirwts1 = s.dirwts1(), *dgintv = s.dgintv(), *dginth = s.dginth();
    const int i = rr * TS + cc;
    float temp = vcd[i];
    const float uave = temp + vcd[i - v1] + vcd[i - v2] + vcd[i - v3];
    const float dave = temp + vcd[i + v1] + vcd[i + v2] + vcd[i + v3];
    float Dgrbvvaru = sq(temp - uave) + sq(vcd[i - v1] - uave) + sq(vcd[i - v2] - uave) + sq(vcd[i - v3] - uave);
    float Dgrbvvard = sq(temp - dave) + sq(vcd[i + v1] - dave) + sq(vcd[i + v2] - dave) + sq(vcd[i + v3] - dave);
    const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
    const float vwt = dirwts0[i - v1] / (dirwts0[i - v1] + dirwts0[i + v1]);
    temp = hcd[i];
    const float lave = temp + (hcd[i - 3] + hcd[i - 2]) + hcd[i - 1];
    const float rave = temp + (hcd[i + 1] + hcd[i + 2]) + hcd[i + 3];
    float Dgrbhvarl = sq(temp - lave) + sq(hcd[i - 1] - lave) + sq(hcd[i - 2] - lave) + sq(hcd[i - 3] - lave);
    float Dgrbhvarr = sq(temp - rave) + sq(hcd[i + 1] - rave) + sq(hcd[i + 2] - rave) + sq(hcd[i + 3] - rave);
    const float vcdvar = epssq + vintpf(vwt, Dgrbvvard, Dgrbvvaru);
    const float hcdvar = epssq + vintpf(hwt, Dgrbhvarr, Dgrbhvarl);
    Dgrbvvaru = dgintv[i - v1] + dgintv[i - v2];
    Dgrbvvard = dgintv[i + v1] + dgintv[i + v2];
    Dgrbhvarl = dginth[i - 2] + dginth[i - 1];
    Dgrbhvarr = dginth[i + 1] + dginth[i + 2];
    const float vcdvar1 = epssq + dgintv[i] + vintpf(vwt, Dgrbvvard, Dgrbvvaru);
    const float hcdvar1 = epssq + dginth[i] + vintpf(hwt, Dgrbhvarr, Dgrbhvarl);
    const float varwt = hcdvar / (vcdvar + hcdvar);
    const float diffwt = hcdvar1 / (vcdvar1 + hcdvar1);
    const bool dec = ((0.5f - varwt) * (0.5f - diffwt) > 0.f) && (fabsf(0.5f - diffwt) < fabsf(0.5f - varwt));
    s.hvwt()[i >> 1] = dec ? varwt : diffwt;
}

// ------------------------------------------------------------------ pass 5/6: nyquist test + flags + bbox (L789-864)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_nyqtest(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 6 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 6) return;
    const int c0 = 6 + (fc(a.filters, rr, 2) & 1);
    const int cc = c0 + 2 * threadIdx.x;
    if (cc >= g.cc1 - 6) return;                 // only sites the flag loop (L853) visits matter
    // vector loop covers starts c0 + 8k < cc1 - 7; the rest of the row goes through the scalar form
    const int nvec = g.cc1 - 7 > c0 ? (g.cc1 - 7 - c0 + 7) / 8 : 0;
    const bool vec = (int)(threadIdx.x >> 2) < nvec;
    const Slab s = slab(a, t);
    const float* cddiffsq = s.cddiffsq();
    const float* d = s.delhvsqsum();
    const int i = rr * TS + cc;
    const float gaussodd0 = 0.14659727707323927f, gaussodd1 = 0.103592713382435f, gaussodd2 = 0.0732036125103057f, gaussodd3 = 0.0365543548389495f;
    const float nyqthresh = 0.5f;
    const float gg0 = nyqthresh * 0.07384411893421103f, gg1 = nyqthresh * 0.06207511968171489f, gg2 = nyqthresh * 0.0521818194747806f,
                gg3 = nyqthresh * 0.03687419286733595f, gg4 = nyqthresh * 0.03099732204057846f, gg5 = nyqthresh * 0.018413194161458882f;
    const float g1sum = vec ? (d[i - v1] + d[i - 1] + d[i + 1] + d[i + v1]) : (d[i - v1] + d[i + 1] + d[i - 1] + d[i + v1]);
    const float val =
        (gaussodd0 * cddiffsq[i] +
         gaussodd1 * (cddiffsq[i - m1] + cddiffsq[i + p1] + cddiffsq[i - p1] + cddiffsq[i + m1]) +
         gaussodd2 * (cddiffsq[i - v2] + cddiffsq[i - 2] + cddiffsq[i + 2] + cddiffsq[i + v2]) +
         gaussodd3 * (cddiffsq[i - m2] + cddiffsq[i + p2] + cddiffsq[i - p2] + cddiffsq[i + m2])) -
        (gg0 * d[i] +
         gg1 * g1sum +
         gg2 * (d[i - m1] + d[i + p1] + d[i - p1] + d[i + m1]) +
         gg3 * (d[i - v2] + d[i - 2] + d[i + 2] + d[i + v2]) +
         gg4 * (d[i - v2 - 1] + d[i - v2 + 1] + d[i - TS - 2] + d[i - TS + 2] + d[i + TS - 2] + d[i + TS + 2] + d[i + v2 - 1] + d[i + v2 + 1]) +
         gg5 * (d[i - m2] + d[i + p2] + d[i - p2] + d[i + m2]));
    if (val > 0.f) {
        s.nyquist()[i >> 1] = 1;
        int* bb = a.bbox + 4 * t;
        atomicMin(bb + 0, rr); atomicMax(bb + 1, rr); atomicMin(bb + 2, cc); atomicMax(bb + 3, cc);
    }
}

struct NyBox { bool on; int r0, r1, c0, c1; };
__device__ __forceinline__ NyBox nybox(const AmzArgs& a, int t, const Geo& g)
{   // L847-876
    const int* bb = a.bbox + 4 * t;
    NyBox b;
    int nystartrow = bb[0], nyendrow = bb[1], nystartcol = bb[2], nyendcol = bb[3];
    if (nyendrow == 0) { nystartrow = 0; nystartcol = TS + 1; }      // nothing flagged
    b.on = nystartrow != nyendrow && nystartcol != nyendcol;
    nyendrow++; nyendcol++;
    nystartcol -= (nystartcol & 1);
    b.r0 = max(8, nystartrow); b.r1 = min(g.rr1 - 8, nyendrow);
    b.c0 = max(8, nystartcol); b.c1 = min(g.cc1 - 8, nyendcol);
    return b;
}

// ------------------------------------------------------------------ pass 7/8: nyquist2 majority + area interpolation (L877-953)
__global__ void __launch_bounds__(256) k_nyquist2(AmzArgs a)
{
    const int t = blockIdx.x;
    const Geo g = geo(a, t);
    const NyBox b = nybox(a, t, g);
    if (!b.on) return;
    const Slab s = slab(a, t);
    const unsigned char* nyquist = s.nyquist();
    unsigned char* nyquist2 = s.nyquist2();
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = tid; k < (TS - 8) * TSH / 4; k += nt) reinterpret_cast<unsigned*>(nyquist2 + 4 * TSH)[k] = 0u;   // L877
    __syncthreads();
    // majority over the 8 quincunx neighbours; whole rows, 16 half-sites per vector (L885-900)
    const int nb = ((g.cc1 + 31) / 32) * 16;       // bytes per row the vector loop covers
    for (int k = tid; k < (b.r1 - b.r0) * nb; k += nt) {
        const int rr = b.r0 + k / nb, h = k % nb;
        const int indx = rr * TS;
#define NQ(o) ((int)(signed char)nyquist[((indx + (o)) >> 1) + h])
        int t1 = sat8(NQ(-v2) + NQ(-m1));
        int t2 = sat8(NQ(p1) + NQ(-2));
        const int t3 = sat8(NQ(2) + NQ(-p1));
        const int t4 = sat8(NQ(m1) + NQ(v2));
#undef NQ
        t1 = sat8(t1 + t3);
        t2 = sat8(t2 + t4);
        t1 = sat8(t1 + t2);
        unsigned char val = nyquist[(indx >> 1) + h];
        val = t1 > 4 ? 1 : val;
        val = t1 < 4 ? 0 : val;
        nyquist2[(indx >> 1) + h] = val;
    }
    __syncthreads();
    // area interpolation at flagged sites of the box (L918-953)
    const float* cfa = s.cfa();
    float* hvwt = s.hvwt();
    const int wbox = (b.c1 - b.c0 + 1) / 2 + 1;
    for (int k = tid; k < (b.r1 - b.r0) * wbox; k += nt) {
        const int rr = b.r0 + k / wbox;
        const int cc = b.c0 + (fc(a.filters, rr, 2) & 1) + 2 * (k % wbox);
        if (cc >= b.c1) continue;
        const int indx = rr * TS + cc;
        if (!nyquist2[indx >> 1]) continue;
        float sumcfa = 0.f, sumh = 0.f, sumv = 0.f, sumsqh = 0.f, sumsqv = 0.f, areawt = 0.f;
        for (int ii = -6; ii < 7; ii += 2) {
            int indx1 = indx + (ii * TS) - 6;
            for (int jj = -6; jj < 7; jj += 2, indx1 += 2)
                if (nyquist2[indx1 >> 1]) {
                    const float cfatemp = cfa[indx1];
                    sumcfa += cfatemp;
                    sumh += (cfa[indx1 - 1] + cfa[indx1 + 1]);
                    sumv += (cfa[indx1 - v1] + cfa[indx1 + v1]);
                    sumsqh += sq(cfatemp - cfa[indx1 - 1]) + sq(cfatemp - cfa[indx1 + 1]);
                    sumsqv += sq(cfatemp - cfa[indx1 - v1]) + sq(cfatemp - cfa[indx1 + v1]);
                    areawt += 1.f;
                }
        }
        sumh = sumcfa - 0.5f * sumh;
        sumv = sumcfa - 0.5f * sumv;
        areawt = 0.5f * areawt;
        const float hcdvar = epssq + fabsf(areawt * sumsqh - sumh * sumh);
        const float vcdvar = epssq + fabsf(areawt * sumsqv - sumv * sumv);
        hvwt[indx >> 1] = hcdvar / (vcdvar + hcdvar);
    }
}

// ------------------------------------------------------------------ row recurrences (L962-964, L1218-1221)
// One warp per tile marches down rows [r_first, rr1 - r_first).  PMWT=false: hvwt (scalar loop, columns
// 8+p .. < cc1-8); PMWT=true: pmwt (vector loop, starts 10+p+8k < cc1-10, 4 sites per vector).
// Rows are exchanged through three shared-memory row buffers (previous-new, current, next-old).
template <bool PMWT>
__global__ void __launch_bounds__(128) k_rowrec(AmzArgs a)
{
    __shared__ float rows[4][3][TSH];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + w;
    if (t >= a.ntiles) return;
    const Geo g = geo(a, t);
    const Slab s = slab(a, t);
    float* buf = PMWT ? s.pmwt() : s.hvwt();
    const int rf = PMWT ? 10 : 8;
    if (g.rr1 - rf <= rf) return;
    float* prev = rows[w][0];
    float* cur = rows[w][1];
    float* next = rows[w][2];
    for (int h = lane; h < TSH; h += 32) { prev[h] = buf[(rf - 1) * TSH + h]; cur[h] = buf[rf * TSH + h]; }
    // the old values of the rows below are fetched RR_PF rows ahead of the row being decided (a row is only ever written by its own
    // step, so a value fetched early is the value the step would have read); only the decided row travels through shared memory
    constexpr int RR_PF = 6;
    const int rend = g.rr1 - rf;
    float q[RR_PF][3];
    auto fetch = [&](int row, float (&dst)[3]) {
        #pragma unroll
        for (int m = 0; m < 3; ++m) { const int h = lane + 32 * m; dst[m] = (row < TS && h < TSH) ? buf[row * TSH + h] : 0.f; }
    };
    #pragma unroll
    for (int k = 0; k < RR_PF; ++k) fetch(rf + 1 + k, q[k]);
    __syncwarp();
    for (int rr0 = rf; rr0 < rend; rr0 += RR_PF) {
        #pragma unroll
        for (int k = 0; k < RR_PF; ++k) {
            const int rr = rr0 + k;
            if (rr >= rend) break;
            #pragma unroll
            for (int m = 0; m < 3; ++m) { const int h = lane + 32 * m; if (h < TSH) next[h] = q[k][m]; }
            fetch(rr + 1 + RR_PF, q[k]);
            __syncwarp();
            const int p = fc(a.filters, rr, 2) & 1;
            float nv[3];
            bool act[3];
            #pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int j = lane + 32 * m;                 // site number in the row
                const int cc = rf + p + 2 * j;
                act[m] = PMWT ? ((rf + p + 8 * (j >> 2)) < g.cc1 - rf && cc < TS) : (cc < g.cc1 - rf);
                nv[m] = 0.f;
                if (act[m]) {
                    // diagonal neighbours: row above (new) and row below (old), columns cc-1 and cc+1
                    const int hl = (cc - 1) >> 1, hr = (cc + 1) >> 1;
                    const float alt = 0.25f * (prev[hl] + prev[hr] + next[hl] + next[hr]);
                    const float x = cur[cc >> 1];
                    nv[m] = fabsf(0.5f - x) < fabsf(0.5f - alt) ? alt : x;
                }
            }
            __syncwarp();
            #pragma unroll
            for (int m = 0; m < 3; ++m)
                if (act[m]) {
                    const int cc = rf + p + 2 * (lane + 32 * m);
                    cur[cc >> 1] = nv[m];
                    buf[(rr * TS + cc) >> 1] = nv[m];
                }
            __syncwarp();
            float* tmp = prev; prev = cur; cur = next; next = tmp;
        }
    }
}

// ------------------------------------------------------------------ pass 9b: G at R/B sites (L967-973)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_green(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 8 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 8) return;
    const int cc = 8 + (fc(a.filters, rr, 2) & 1) + 2 * threadIdx.x;
    if (cc >= g.cc1 - 8) return;
    const Slab s = slab(a, t);
    const int i = rr * TS + cc, k = i >> 1;
    float* rgbgreen = s.rgbgreen();
    const float hv = s.hvwt()[k];
    const float dg = hv * s.vcd()[i] + (1.f - hv) * s.hcd()[i];        // intp, rt_math.h L110
    s.Dgrb0()[k] = dg;
    const float gr = s.cfa()[i] + dg;
    rgbgreen[i] = gr;
    const unsigned char nq = s.nyquist2()[k];
    float* D2 = s.Dgrb2();
    // neighbours are G sites: rgbgreen there still holds cfa (never written by this pass)
    D2[2 * k] = nq ? sq(gr - 0.5f * (rgbgreen[i - 1] + rgbgreen[i + 1])) : 0.f;
    D2[2 * k + 1] = nq ? sq(gr - 0.5f * (rgbgreen[i - v1] + rgbgreen[i + v1])) : 0.f;
}

// ------------------------------------------------------------------ pass 10: Nyquist refinement (L980-999)
__global__ void __launch_bounds__(256) k_nyqrefine(AmzArgs a)
{
    const int t = blockIdx.x;
    const Geo g = geo(a, t);
    const NyBox b = nybox(a, t, g);
    if (!b.on) return;
    const Slab s = slab(a, t);
    const unsigned char* nyquist2 = s.nyquist2();
    const float* D2 = s.Dgrb2();
    const float gq0 = 0.169917f, gq1 = 0.108947f, gq2 = 0.069855f, gq3 = 0.0287182f;
    const int wbox = (b.c1 - b.c0 + 1) / 2 + 1;
    for (int k = threadIdx.x; k < (b.r1 - b.r0) * wbox; k += blockDim.x) {
        const int rr = b.r0 + k / wbox;
        const int cc = b.c0 + (fc(a.filters, rr, 2) & 1) + 2 * (k % wbox);
        if (cc >= b.c1) continue;
        const int indx = rr * TS + cc;
        if (!nyquist2[indx >> 1]) continue;
#define D2H(o) D2[2 * ((indx + (o)) >> 1)]
#define D2V(o) D2[2 * ((indx + (o)) >> 1) + 1]
        const float gvarh = epssq + (gq0 * D2H(0) +
                                     gq1 * (D2H(-m1) + D2H(p1) + D2H(-p1) + D2H(m1)) +
                                     gq2 * (D2H(-v2) + D2H(-2) + D2H(2) + D2H(v2)) +
                                     gq3 * (D2H(-m2) + D2H(p2) + D2H(-p2) + D2H(m2)));
        const float gvarv = epssq + (gq0 * D2V(0) +
                                     gq1 * (D2V(-m1) + D2V(p1) + D2V(-p1) + D2V(m1)) +
                                     gq2 * (D2V(-v2) + D2V(-2) + D2V(2) + D2V(v2)) +
                                     gq3 * (D2V(-m2) + D2V(p2) + D2V(-p2) + D2V(m2)));
#undef D2H
#undef D2V
        const float dg = (s.hcd()[indx] * gvarv + s.vcd()[indx] * gvarh) / (gvarv + gvarh);
        s.Dgrb0()[indx >> 1] = dg;
        s.rgbgreen()[indx] = s.cfa()[indx] + dg;
    }
}

// ------------------------------------------------------------------ pass 11: diagonal gradients (L1004-1026)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_diag(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 6 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 6) return;
    int cc;
    if (!rb_lane(threadIdx.x, 6, g.cc1 - 6, &cc) || cc >= TS) return;   // pairs start at even column 6
    const Slab s = slab(a, t);
    const float* cfa = s.cfa();
    const int i = rr * TS + cc;
    const bool odd = (fc(a.filters, rr, 2) & 1);        // false: R/B at even columns
    const int gi = odd ? i : i + 1;                     // G member of the column pair
    const int xi = odd ? i + 1 : i;                     // R/B member
    const float tv = cfa[gi];
    const float sp = sq(tv - cfa[gi - p1]) + sq(tv - cfa[gi + p1]);
    const float dp = fabsf(cfa[xi + p1] - cfa[xi - p1]);
    const float dm = fabsf(cfa[xi + m1] - cfa[xi - m1]);
    const float sm = sq(tv - cfa[gi - m1]) + sq(tv - cfa[gi + m1]);
    s.delp()[i >> 1] = dp;
    s.delm()[i >> 1] = dm;
    s.Dgrbsq1m()[i >> 1] = sm;
    s.Dgrbsq1p()[i >> 1] = sp;
}

// ------------------------------------------------------------------ pass 12: rbm, rbp, pmwt (L1057-1121)
__device__ __forceinline__ float diag_est(float cfav, float t1, float t2)
{
    const float r = (t1 + t1) / (eps + cfav + t2);
    return fabsf(1.f - r) < arthresh ? cfav * r : t1 + 0.5f * (cfav - t2);
}
__device__ __forceinline__ float diag_bound(float rb, float cfav, float ca, float cb, float clip_pt)
{
    const float t1 = vmedian(rb, ca, cb);
    const float wt = ((cfav - rb) + (cfav - rb)) / (eps + rb + cfav);
    float t2 = vintpf(wt, rb, t1);
    t2 = (rb + rb < cfav) ? t1 : t2;
    t2 = (rb < cfav) ? t2 : rb;
    return (t2 > clip_pt) ? vmedian(t2, ca, cb) : t2;
}
__global__ void __launch_bounds__(TSH * HR_ROWS) k_rbdiag(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 8 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 8) return;
    int cc;
    if (!rb_lane(threadIdx.x, 8 + (fc(a.filters, rr, 2) & 1), g.cc1 - 8, &cc) || cc >= TS) return;
    const Slab s = slab(a, t);
    const float *cfa = s.cfa(), *delm = s.delm(), *delp = s.delp(), *Dm = s.Dgrbsq1m(), *Dp = s.Dgrbsq1p();
    const int i = rr * TS + cc, k = i >> 1;
    const float cfav = cfa[i];
    const float gausseven0 = 0.13719494435797422f, gausseven1 = 0.05640252782101291f;
    {
        const float rbse = diag_est(cfav, cfa[i + m1], cfa[i + m2]);
        const float rbnw = diag_est(cfav, cfa[i - m1], cfa[i - m2]);
        const float t1 = eps + delm[k];
        const float wtse = t1 + delm[(i + m1) >> 1] + delm[(i + m2) >> 1];
        const float wtnw = t1 + delm[(i - m1) >> 1] + delm[(i - m2) >> 1];
        const float rbmv = (wtse * rbnw + wtnw * rbse) / (wtse + wtnw);
        const float rbpre = diag_bound(rbmv, cfav, cfa[i - m1], cfa[i + m1], a.clip_pt);
        const float rbne = diag_est(cfav, cfa[i + p1], cfa[i + p2]);
        const float rbsw = diag_est(cfav, cfa[i - p1], cfa[i - p2]);
        const float u1 = eps + delp[k];
        const float wtne = u1 + delp[(i + p1) >> 1] + delp[(i + p2) >> 1];
        const float wtsw = u1 + delp[(i - p1) >> 1] + delp[(i - p2) >> 1];
        const float rbpv = (wtne * rbsw + wtsw * rbne) / (wtne + wtsw);
        const float rbppost = diag_bound(rbpv, cfav, cfa[i - p1], cfa[i + p1], a.clip_pt);
        const float rbvarm = epssq + (gausseven0 * (Dm[(i - v1) >> 1] + Dm[(i - 1) >> 1] + Dm[(i + 1) >> 1] + Dm[(i + v1) >> 1]) +
                                      gausseven1 * (Dm[(i - v2 - 1) >> 1] + Dm[(i - v2 + 1) >> 1] + Dm[(i - 2 - v1) >> 1] + Dm[(i + 2 - v1) >> 1] +
                                                    Dm[(i - 2 + v1) >> 1] + Dm[(i + 2 + v1) >> 1] + Dm[(i + v2 - 1) >> 1] + Dm[(i + v2 + 1) >> 1]));
        const float pm = rbvarm / ((epssq + (gausseven0 * (Dp[(i - v1) >> 1] + Dp[(i - 1) >> 1] + Dp[(i + 1) >> 1] + Dp[(i + v1) >> 1]) +
                                             gausseven1 * (Dp[(i - v2 - 1) >> 1] + Dp[(i - v2 + 1) >> 1] + Dp[(i - 2 - v1) >> 1] + Dp[(i + 2 - v1) >> 1] +
                                                           Dp[(i - 2 + v1) >> 1] + Dp[(i + 2 + v1) >> 1] + Dp[(i + v2 - 1) >> 1] + Dp[(i + v2 + 1) >> 1]))) + rbvarm);
        // rbm/rbp live in the memory of vcd, pmwt in the memory of delhvsqsum: both are dead by now
        s.rbm()[k] = rbpre;
        s.rbp()[k] = rbppost;
        s.pmwt()[k] = pm;
    }
}

// ------------------------------------------------------------------ pass 13b: rbint (L1222)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_rbint(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 10 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 10) return;
    int cc;
    if (!rb_lane(threadIdx.x, 10 + (fc(a.filters, rr, 2) & 1), g.cc1 - 10, &cc) || cc >= TS) return;
    const Slab s = slab(a, t);
    const int i = rr * TS + cc, k = i >> 1;
    s.rbint()[k] = 0.5f * (s.cfa()[i] + vintpf(s.pmwt()[k], s.rbp()[k], s.rbm()[k]));
}

// ------------------------------------------------------------------ pass 14: G via R+B (L1241-1294)
__device__ __forceinline__ float card_est(float rbintv, float cn, float rbn)
{
    const float cr = (cn + cn) / (eps + rbintv + rbn);
    const float g1 = rbintv * cr;
    const float g2 = cn + 0.5f * (rbintv - rbn);
    return fabsf(1.f - cr) < arthresh ? g1 : g2;
}
__device__ __forceinline__ float card_bound(float Gint, float rbintv, float ca, float cb, float clip_pt)
{
    float Gint1 = vmedian(Gint, ca, cb);
    const float wt = ((rbintv - Gint) + (rbintv - Gint)) / (eps + Gint + rbintv);
    const float Gint2 = vintpf(wt, Gint, Gint1);
    Gint1 = ((Gint + Gint) < rbintv) ? Gint1 : Gint2;
    Gint = (Gint < rbintv) ? Gint1 : Gint;
    return (Gint > clip_pt) ? vmedian(Gint, ca, cb) : Gint;
}
// One thread per half-column k of a row: first the G-via-R+B lane that lives there (if any), then the
// split of G-B from G-R for B rows (L1382-1386) -- the two touch only their own site, so running the
// split right behind the lane is the same as running the two passes back to back.
__global__ void __launch_bounds__(TSH * HR_ROWS) k_greenrb(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 12 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 12) return;
    const Slab s = slab(a, t);
    const int k = rr * TSH + threadIdx.x;                       // half index of this thread's column pair
    const int p = fc(a.filters, rr, 2) & 1;
    // ---- pass 14 lane: site cc = 12 + p + 2 j  <=>  half column 6 + j
    const int j = (int)threadIdx.x - 6;
    int cc;
    if (j >= 0 && rb_lane(j, 12 + p, g.cc1 - 12, &cc) && cc < TS) {
        const int i = rr * TS + cc;
        const float hv = s.hvwt()[k];
        if (fabsf(0.5f - s.pmwt()[k]) >= fabsf(0.5f - hv)) {      // copymask lane
            const float *cfa = s.cfa(), *rbint = s.rbint(), *dirwts0 = s.dirwts0(), *dirwts1 = s.dirwts1();
            const float rbintv = rbint[k];
            const float gu = card_est(rbintv, cfa[i - v1], rbint[k - v1]);
            const float gd = card_est(rbintv, cfa[i + v1], rbint[k + v1]);
            float Gintv = (dirwts0[i - v1] * gd + dirwts0[i + v1] * gu) / (dirwts0[i + v1] + dirwts0[i - v1]);
            Gintv = card_bound(Gintv, rbintv, cfa[i - v1], cfa[i + v1], a.clip_pt);
            const float gl = card_est(rbintv, cfa[i - 1], rbint[k - 1]);
            const float gr = card_est(rbintv, cfa[i + 1], rbint[k + 1]);
            float Ginth = (dirwts1[i - 1] * gr + dirwts1[i + 1] * gl) / (dirwts1[i - 1] + dirwts1[i + 1]);
            Ginth = card_bound(Ginth, rbintv, cfa[i - 1], cfa[i + 1], a.clip_pt);
            const float greenv = vintpf(hv, Gintv, Ginth);
            s.rgbgreen()[i] = greenv;
            s.Dgrb0()[k] = greenv - cfa[i];
        }
    }
    // ---- pass 15: B rows are rr = 13 - ey, 15 - ey, ...; half columns [(13-ex)>>1, (cc1-12)>>1)
    if (rr >= 13 - a.ey && ((rr - (13 - a.ey)) & 1) == 0 &&
        k >= ((rr * TS + 13 - a.ex) >> 1) && k < ((rr * TS + g.cc1 - 12) >> 1)) {
        // The reference also stores Dgrb[0] = 0 here (L1385).  That zero is dead: pass 16 overwrites it at every
        // site pass 17 reads, and nothing else reads Dgrb[0] at B sites.  Measured on B200 the extra store made
        // this kernel 3x slower (1.38 ms vs 0.47 ms per 45 MP frame), so it is dropped; the scratch-block
        // comparison in tests/test_amaze_gpu.py masks exactly those cells.
        s.Dgrb1()[k] = s.Dgrb0()[k];
    }
}

// ------------------------------------------------------------------ pass 16: chrominance interpolation (L1394-1408)
__global__ void __launch_bounds__(TSH * HR_ROWS) k_chroma(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 14 + blockIdx.x * HR_ROWS + threadIdx.y;
    if (rr >= g.rr1 - 14) return;
    const int c0 = 14 + (fc(a.filters, rr, 2) & 1);
    int cc;
    if (!rb_lane(threadIdx.x, c0, g.cc1 - 14, &cc) || cc >= TS) return;
    const Slab s = slab(a, t);
    const int c = 1 - (int)fc(a.filters, rr, c0) / 2;
    float* D = c ? s.Dgrb1() : s.Dgrb0();
    const int i = rr * TS + cc;
#define DG(o) D[(i + (o)) >> 1]
    const float tempv = eps + fabsf(DG(-m1) - DG(m1));
    const float temp2v = eps + fabsf(DG(p1) - DG(-p1));
    const float wtnw = 1.f / (tempv + fabsf(DG(-m1) - DG(-m3)) + fabsf(DG(m1) - DG(-m3)));
    const float wtne = 1.f / (temp2v + fabsf(DG(p1) - DG(p3)) + fabsf(DG(-p1) - DG(p3)));
    const float wtsw = 1.f / (temp2v + fabsf(DG(-p1) - DG(m3)) + fabsf(DG(p1) - DG(-p3)));
    const float wtse = 1.f / (tempv + fabsf(DG(m1) - DG(-p3)) + fabsf(DG(-m1) - DG(m3)));
    const float res = (wtnw * (1.325f * DG(-m1) - 0.175f * DG(-m3) - 0.075f * (DG(-m1 - 2) + DG(-m1 - v2))) +
                       wtne * (1.325f * DG(p1) - 0.175f * DG(p3) - 0.075f * (DG(p1 + 2) + DG(p1 + v2))) +
                       wtsw * (1.325f * DG(-p1) - 0.175f * DG(-p3) - 0.075f * (DG(-p1 - 2) + DG(-p1 - v2))) +
                       wtse * (1.325f * DG(m1) - 0.175f * DG(m3) - 0.075f * (DG(m1 + 2) + DG(m1 + v2)))) / (wtnw + wtne + wtsw + wtse);
#undef DG
    // reads touch only sites of the opposite R/B type (all offsets are odd-row/odd-col), writes only own type
    D[i >> 1] = res;
}

// ------------------------------------------------------------------ pass 17/18: write R, G, B (L1441-1565)
__global__ void __launch_bounds__(TS * FR_ROWS) k_write(AmzArgs a)
{
    const int t = blockIdx.y;
    const Geo g = geo(a, t);
    const int rr = 16 + blockIdx.x * FR_ROWS + threadIdx.y, cc = threadIdx.x;
    if (rr >= g.rr1 - 16 || cc < 16 || cc >= g.cc1 - 16) return;
    const Slab s = slab(a, t);
    const float *hvwt = s.hvwt(), *D0 = s.Dgrb0(), *D1 = s.Dgrb1();
    const int indx = rr * TS + cc;
    const float gr = s.rgbgreen()[indx];
    float r, b;
    if (fc(a.filters, rr, cc) & 1) {
        const float hu = hvwt[(indx - v1) >> 1], hr = hvwt[(indx + 1) >> 1], hl = hvwt[(indx - 1) >> 1], hd = hvwt[(indx + v1) >> 1];
        const float temp = 1.f / (hu + 2.f - hr - hl + hd);
        r = gr - (hu * D0[(indx - v1) >> 1] + (1.f - hr) * D0[(indx + 1) >> 1] + (1.f - hl) * D0[(indx - 1) >> 1] + hd * D0[(indx + v1) >> 1]) * temp;
        b = gr - (hu * D1[(indx - v1) >> 1] + (1.f - hr) * D1[(indx + 1) >> 1] + (1.f - hl) * D1[(indx - 1) >> 1] + hd * D1[(indx + v1) >> 1]) * temp;
    } else {
        r = gr - D0[indx >> 1];
        b = gr - D1[indx >> 1];
    }
    const size_t o = (size_t)(rr + g.top) * a.op + (g.left + cc);
    a.R[o] = stdmax(0.f, 65535.f * r);
    a.B[o] = stdmax(0.f, 65535.f * b);
    a.G[o] = stdmax(0.f, 65535.f * gr);
}

}  // namespace

// number of passes the debug hook can stop after (tests/test_amaze_gpu.py compares slabs pass by pass)
extern "C" int art_hpdbg_amaze_num_passes(void) { return 18; }

static int amaze_band(art_hp_ctx* ctx, AmzArgs a, int stop_after)
{
    cudaStream_t st = ctx->stream;
    const int nt = a.ntiles;
    int pass = 0;
#define LAUNCH(kern, grid, block)                                      \
    do {                                                               \
        art_prof_begin(ctx, #kern);                                    \
        kern<<<grid, block, 0, st>>>(a);                               \
        art_prof_end(ctx);                                             \
        ctx->launches++;                                               \
        if (++pass == stop_after) { ART_CUDA(ctx, cudaGetLastError()); return ART_HP_OK; } \
    } while (0)
    // The oracle's deterministic variant zeroes a tile's scratch before the tile starts.  Clearing less is not safe: which sub-buffers
    // are read before the tile has written them (lanes overrunning a vector loop, cells read under an aliased name) depends on the
    // tile geometry AND the data (a poisoning sweep of the oracle found delhvsqsum, vcd .. cddiffsq, nyquist always, dginth and cfa
    // for some frames), so the whole slab is cleared.
    art_prof_begin(ctx, "memset_slabs");
    ART_CUDA(ctx, cudaMemsetAsync(a.slabs, 0, (size_t)nt * SLAB_BYTES, st));
    art_prof_end(ctx);
    const dim3 fr_block(TS, FR_ROWS), hr_block(TSH, HR_ROWS);
    const dim3 fr_grid((TS + FR_ROWS - 1) / FR_ROWS, nt), hr_grid((TS + HR_ROWS - 1) / HR_ROWS, nt);
    LAUNCH(k_fill, nt, 256);                                           // 1
    LAUNCH(k_grad, fr_grid, fr_block);                                 // 2
    LAUNCH(k_dirinterp, fr_grid, fr_block);                            // 3
    LAUNCH(k_hcd, dim3(TS / 4, nt), dim3(40, 4));                      // 4
    LAUNCH(k_vcd, nt, dim3(TS, 2));                                    // 5
    LAUNCH(k_hvwt, hr_grid, hr_block);                                 // 6
    LAUNCH(k_nyqtest, hr_grid, hr_block);                              // 7
    LAUNCH(k_nyquist2, nt, 256);                                       // 8
    LAUNCH(k_rowrec<false>, (nt + 3) / 4, 128);                        // 9
    LAUNCH(k_green, hr_grid, hr_block);                                // 10
    LAUNCH(k_nyqrefine, nt, 256);                                      // 11
    LAUNCH(k_diag, hr_grid, hr_block);                                 // 12
    LAUNCH(k_rbdiag, hr_grid, hr_block);                               // 13
    LAUNCH(k_rowrec<true>, (nt + 3) / 4, 128);                         // 14
    LAUNCH(k_rbint, hr_grid, hr_block);                                // 15
    LAUNCH(k_greenrb, hr_grid, hr_block);                              // 16 (+ the split, reference pass 15)
    if (++pass == stop_after) { ART_CUDA(ctx, cudaGetLastError()); return ART_HP_OK; }   // 17: kept for the debug numbering
    LAUNCH(k_chroma, hr_grid, hr_block);                               // 18
    if (stop_after == 0) {
        art_prof_begin(ctx, "k_write");
        k_write<<<fr_grid, fr_block, 0, st>>>(a);
        art_prof_end(ctx);
        ctx->launches++;
    }
#undef LAUNCH
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

static int amaze_setup(art_hp_ctx* ctx, AmzArgs& a, int W, int H, unsigned filters, const float* raw, size_t rp,
                       float* R, float* G, float* B, size_t op, double initialGain, int* nty_out)
{
    a.raw = raw; a.rp = rp; a.R = R; a.G = G; a.B = B; a.op = op; a.W = W; a.H = H; a.filters = filters;
    a.clip_pt = (float)(1.0 / initialGain);      // L53-54
    a.clip_pt8 = (float)(0.8 / initialGain);
    auto FC = [filters](int r, int c) { return (filters >> ((((r) << 1 & 14) + ((c) & 1)) << 1)) & 3u; };
    if (FC(0, 0) == 1) { if (FC(0, 1) == 0) { a.ey = 0; a.ex = 1; } else { a.ey = 1; a.ex = 0; } }      // L70-86
    else { if (FC(0, 0) == 0) { a.ey = 0; a.ex = 0; } else { a.ey = 1; a.ex = 1; } }
    a.ntx = (W + 16 + (TS - 32) - 1) / (TS - 32);       // lefts -16 + 128 k < W  (L183)
    *nty_out = (H + 16 + (TS - 32) - 1) / (TS - 32);
    return ART_HP_OK;
}

// tile rows per band: bounded by the scratch budget (default 6 GiB; ART_HP_AMAZE_SCRATCH_MB overrides)
static int band_rows(int ntx, int nty)
{
    size_t budget = (size_t)6144 << 20;
    if (const char* e = getenv("ART_HP_AMAZE_SCRATCH_MB")) { const long v = atol(e); if (v > 0) budget = (size_t)v << 20; }
    long rows = (long)(budget / (SLAB_BYTES * (size_t)ntx));
    if (rows < 1) rows = 1;
    if (rows > nty) rows = nty;
    return (int)rows;
}

int art_amaze_dev_banded(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                         float* R, float* G, float* B, size_t op, double initialGain, int border,
                         int band_tile_rows, art_band_cb cb, void* user, int row_begin, int row_end)
{
    AmzArgs a;
    int nty = 0;
    amaze_setup(ctx, a, W, H, filters, raw, rp, R, G, B, op, initialGain, &nty);
    int rows = band_rows(a.ntx, nty);
    if (band_tile_rows > 0 && band_tile_rows < rows) rows = band_tile_rows;
    int rc = art_reserve(ctx, ctx->d_scratch, (size_t)rows * a.ntx * SLAB_BYTES + (size_t)rows * a.ntx * 4 * sizeof(int) + 256);
    if (rc) return rc;
    a.slabs = (char*)ctx->d_scratch.p;
    a.bbox = (int*)(a.slabs + (size_t)rows * a.ntx * SLAB_BYTES);
    // tile rows covering the requested output rows (tile row ty writes rows [128 ty, 128 ty + 128))
    const int ty_begin = row_begin / (TS - 32), ty_end = std::min(nty, (row_end + (TS - 32) - 1) / (TS - 32));
    for (int ty0 = ty_begin; ty0 < ty_end; ty0 += rows) {
        const int nr = std::min(rows, ty_end - ty0);
        a.ty0 = ty0;
        a.ntiles = nr * a.ntx;
        // tile row ty reads raw rows [128*ty - 16, 128*ty + 144) plus, at the frame edges, mirror rows that
        // lie inside [0, 33) (top) and [H-17, H) (bottom); it writes image rows [128*ty, 128*ty + 128)
        // (L1441: rr in [16, rr1-16), row = rr + top)
        if (cb && (rc = cb(user, 0, 0, std::min(H, std::max(33, (TS - 32) * (ty0 + nr) + 16))))) return rc;
        if ((rc = amaze_band(ctx, a, 0))) return rc;
        if (cb && (rc = cb(user, 1, std::min(H, (TS - 32) * ty0), std::min(H, (TS - 32) * (ty0 + nr))))) return rc;
    }
    if (border < 4) return art_border_dev(ctx, W, H, filters, 3, raw, rp, R, G, B, op, row_begin, row_end);      // L1587-1589
    return ART_HP_OK;
}

int art_amaze_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                  float* R, float* G, float* B, size_t op, double initialGain, int border, int row_begin, int row_end)
{
    return art_amaze_dev_banded(ctx, W, H, filters, raw, rp, R, G, B, op, initialGain, border, 0, nullptr, nullptr, row_begin, row_end);
}

// Debug hook (not part of the ABI header; used by tests to localise a divergence): run passes
// 1..stop_after on the whole frame as ONE band and copy tile `tile`'s slab to the host.
extern "C" int art_hpdbg_amaze_slab(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* d_raw, size_t rp,
                                    double initialGain, int stop_after, int tile, void* host_slab, size_t host_bytes)
{
    if (!ctx || !d_raw || !host_slab || host_bytes < SLAB_BYTES) return ART_HP_ERR_INVALID;
    AmzArgs a;
    int nty = 0;
    amaze_setup(ctx, a, W, H, filters, d_raw, rp, nullptr, nullptr, nullptr, 0, initialGain, &nty);
    const int nt = a.ntx * nty;
    if (tile < 0 || tile >= nt || stop_after < 1 || stop_after > 18) return ART_HP_ERR_INVALID;
    int rc = art_reserve(ctx, ctx->d_scratch, (size_t)nt * SLAB_BYTES + (size_t)nt * 4 * sizeof(int) + 256);
    if (rc) return rc;
    a.slabs = (char*)ctx->d_scratch.p;
    a.bbox = (int*)(a.slabs + (size_t)nt * SLAB_BYTES);
    a.ty0 = 0;
    a.ntiles = nt;
    if ((rc = amaze_band(ctx, a, stop_after))) return rc;
    ART_CUDA(ctx, cudaMemcpyAsync(host_slab, a.slabs + (size_t)tile * SLAB_BYTES, SLAB_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}
