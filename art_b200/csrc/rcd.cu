// RCD demosaic + border interpolation for sm_100a.
//
// Replaces RawImageSource::rcd_demosaic (reference rtengine/rcd_demosaic.cc L51-347) and
// RawImageSource::border_interpolate2 (rtengine/demosaic_algos.cc L200-353).
//
// Design (DESIGN.md "RCD"): the reference walks 194x194 tiles at stride 176 and its result
// depends on that grid (VH_Dir is only defined on tile-local [4,T-4)x[4,C-4) and reads as 0 one
// step outside it, L149-166/L201).  We keep the grid: each CTA owns an 88x88 quarter of one
// reference tile's 176x176 output, stages the 108x108 CFA window it depends on (dependency reach is
// exactly 10 pixels) in shared memory, runs the seven RCD stages through four shared-memory planes
// with one __syncthreads between dependent stages, and stores R,G,B straight to HBM.  No
// intermediate ever touches global memory: HBM traffic is the 4 B/px read (+ halo) and 12 B/px write.
//
// Arithmetic: fp32, the reference's expression association verbatim; this file is compiled with
// -fmad=false (no FMA contraction) and IEEE division so results are bit-identical to the x86-64
// reference build.
#include "ctx.h"

#include <algorithm>

namespace {

constexpr int TS = 194;   // reference tileSize        (rcd_demosaic.cc L84)
constexpr int TB = 9;     // tileBorder == rcdBorder   (L82-83)
constexpr int TN = 176;   // tileSizeN                 (L85)
constexpr int SUB = 88;   // output rows/cols per CTA (TN / 2)
constexpr int HALO = 10;  // dependency reach of one output pixel
constexpr int RW = SUB + 2 * HALO;   // 108: region rows/cols (max)
constexpr int PS = 112;              // shared-memory row stride (floats)
constexpr size_t SMEM_BYTES = 4ull * RW * PS * sizeof(float);

__device__ __forceinline__ unsigned fc(unsigned filters, int row, int col)
{   // RawImage::FC, rtengine/rawimage.h L186-189
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float stdmax(float a, float b) { return a < b ? b : a; }   // std::max
__device__ __forceinline__ float lim01(float a) { const float m = a < 1.f ? a : 1.f; return 0.f < m ? m : 0.f; }
__device__ __forceinline__ float intp(float a, float b, float c) { return a * b + (1.f - a) * c; }   // rt_math.h L110
// squared colour-difference high-pass along step s (rcd_demosaic.cc L140, L151, L215-216)
__device__ __forceinline__ float hpf(const float* p, int s)
{
    return sqr((p[-3 * s] - p[-s] - p[s] + p[3 * s]) - 3.f * (p[-2 * s] + p[2 * s]) + 6.f * p[0]);
}

struct RcdArgs {
    const float* raw; size_t rp;
    float *R, *G, *B; size_t op;
    int W, H; unsigned filters;
    int ntw;
    int tr0;          // first reference tile row of this launch
};

template <int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, 1) rcd_kernel(RcdArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float* const A = smem;                 // cfa
    float* const Bv = A + RW * PS;         // VH_Dir
    float* const X = Bv + RW * PS;         // R/B slot: lpf then PQ_Dir; partner slot (col^1): interpolated G; first: vertical HPF
    float* const E = X + RW * PS;          // horizontal HPF, then P (odd col) / Q (odd col - 1) HPF, then R/B slot: the opposite colour

    const float eps = 1e-5f, epssq = 1e-10f, scale = 65536.f;   // L90-92
    const int tid = threadIdx.x;
    // which reference tile and which quarter of it
    const int tc = blockIdx.x >> 1, sc = blockIdx.x & 1;
    const int tr = a.tr0 + (blockIdx.y >> 1), sr = blockIdx.y & 1;
    const int R0 = tr * TN, C0 = tc * TN;
    const int T = min(TS, a.H - R0), C = min(TS, a.W - C0);        // tile rows / cols (L113-124)
    if (T <= 2 * TB || C <= 2 * TB) return;                         // nothing to write (L114-121 + empty write loop)
    // output range of this CTA, tile-local
    const int or0 = TB + sr * SUB, or1 = min(or0 + SUB, T - TB);
    const int oc0 = TB + sc * SUB, oc1 = min(oc0 + SUB, C - TB);
    if (or0 >= or1 || oc0 >= oc1) return;
    // region staged in shared memory, tile-local [rr0, rr0+NR) x [cc0, cc0+NC); cc0 even so that
    // column parity (hence CFA colour and the slot pairing) is the same in every coordinate system
    const int rr0 = max(0, or0 - HALO), cc0 = max(0, oc0 - HALO) & ~1;
    const int NR = min(T, or1 + HALO) - rr0, NC = min(C, oc1 + HALO) - cc0;
    const int gr0 = R0 + rr0, gc0 = C0 + cc0;                        // image coordinates of region origin

    // ---- fill (L126-132): cfa = LIM01(raw / 65536)
    for (int i = tid; i < NR * RW; i += NTHREADS) {
        const int r = i / RW, c = i - r * RW;
        if (c < NC) A[r * PS + c] = lim01(a.raw[(size_t)(gr0 + r) * a.rp + gc0 + c] / scale);
    }
    __syncthreads();

    // ---- step 1.1: squared vertical / horizontal HPF (L138-155) into X / E
    for (int i = tid; i < NR * RW; i += NTHREADS) {
        const int r = i / RW, c = i - r * RW;
        if (c >= NC) continue;
        const int tr_ = rr0 + r, tc_ = cc0 + c;
        const float* p = A + r * PS + c;
        if (r >= 3 && r < NR - 3 && tr_ >= 3 && tr_ < T - 3 && tc_ >= 4 && tc_ < C - 4) X[r * PS + c] = hpf(p, PS);
        if (c >= 3 && c < NC - 3 && tr_ >= 4 && tr_ < T - 4 && tc_ >= 3 && tc_ < C - 3) E[r * PS + c] = hpf(p, 1);
    }
    __syncthreads();
    // ---- step 1.2: VH_Dir (L156-162); 0 wherever the reference tile does not define it
    for (int i = tid; i < NR * RW; i += NTHREADS) {
        const int r = i / RW, c = i - r * RW;
        if (c >= NC) continue;
        const int tr_ = rr0 + r, tc_ = cc0 + c;
        float v = 0.f;
        if (r >= 4 && r < NR - 4 && c >= 4 && c < NC - 4 && tr_ >= 4 && tr_ < T - 4 && tc_ >= 4 && tc_ < C - 4) {
            const int k = r * PS + c;
            const float vs = stdmax(epssq, X[k - PS] + X[k] + X[k + PS]);
            const float hs = stdmax(epssq, E[k - 1] + E[k] + E[k + 1]);
            v = vs / (vs + hs);
        }
        Bv[r * PS + c] = v;
    }
    __syncthreads();

    // ---- step 2: low-pass at R/B sites (L169-175) -> X[R/B slot]; step 4.0: P/Q HPF (L213-218) -> E
    for (int i = tid; i < NR * (RW / 2); i += NTHREADS) {
        const int r = i / (RW / 2), j = i - r * (RW / 2);
        const int tr_ = rr0 + r;
        {   // lpf
            const int c = 2 * j + (fc(a.filters, gr0 + r, 0) & 1);
            const int tc_ = cc0 + c;
            if (c >= 1 && c < NC - 1 && r >= 1 && r < NR - 1 && tr_ >= 2 && tr_ < T - 2 && tc_ >= 2 && tc_ < C - 2) {
                const int k = r * PS + c;
                X[k] = A[k] + 0.5f * (A[k - PS] + A[k + PS] + A[k - 1] + A[k + 1])
                       + 0.25f * (A[k - PS - 1] + A[k - PS + 1] + A[k + PS - 1] + A[k + PS + 1]);
            }
        }
        {   // P/Q on odd columns of every row; 0 where the reference does not compute them
            const int c = 2 * j + 1;
            const int tc_ = cc0 + c;
            if (c < NC) {
                const int k = r * PS + c;
                float pv = 0.f, qv = 0.f;
                if (r >= 3 && r < NR - 3 && c >= 3 && c < NC - 3 && tr_ >= 3 && tr_ < T - 3 && tc_ >= 3 && tc_ < C - 3) {
                    pv = hpf(A + k, PS + 1);
                    qv = hpf(A + k, PS - 1);
                }
                E[k] = pv;
                E[k - 1] = qv;
            }
        }
    }
    __syncthreads();

    // ---- step 3: G at R/B sites (L178-206) -> X[partner slot]
    for (int i = tid; i < NR * (RW / 2); i += NTHREADS) {
        const int r = i / (RW / 2), j = i - r * (RW / 2);
        const int c = 2 * j + (fc(a.filters, gr0 + r, 0) & 1);
        const int tr_ = rr0 + r, tc_ = cc0 + c;
        if (!(c >= 4 && c < NC - 4 && r >= 4 && r < NR - 4 && tr_ >= 4 && tr_ < T - 4 && tc_ >= 4 && tc_ < C - 4)) continue;
        const int k = r * PS + c;
        const float x = A[k];
        const float ng = eps + (fabsf(A[k - PS] - A[k + PS]) + fabsf(x - A[k - 2 * PS])) + (fabsf(A[k - PS] - A[k - 3 * PS]) + fabsf(A[k - 2 * PS] - A[k - 4 * PS]));
        const float sg = eps + (fabsf(A[k - PS] - A[k + PS]) + fabsf(x - A[k + 2 * PS])) + (fabsf(A[k + PS] - A[k + 3 * PS]) + fabsf(A[k + 2 * PS] - A[k + 4 * PS]));
        const float wg = eps + (fabsf(A[k - 1] - A[k + 1]) + fabsf(x - A[k - 2])) + (fabsf(A[k - 1] - A[k - 3]) + fabsf(A[k - 2] - A[k - 4]));
        const float eg = eps + (fabsf(A[k - 1] - A[k + 1]) + fabsf(x - A[k + 2])) + (fabsf(A[k + 1] - A[k + 3]) + fabsf(A[k + 2] - A[k + 4]));
        const float l = X[k];
        const float ne = A[k - PS] * (l + l) / (eps + l + X[k - 2 * PS]);
        const float se = A[k + PS] * (l + l) / (eps + l + X[k + 2 * PS]);
        const float we = A[k - 1] * (l + l) / (eps + l + X[k - 2]);
        const float ee = A[k + 1] * (l + l) / (eps + l + X[k + 2]);
        const float ve = (sg * ne + ng * se) / (ng + sg);
        const float he = (wg * ee + eg * we) / (eg + wg);
        const float cv = Bv[k];
        const float nv = 0.25f * ((Bv[k - PS - 1] + Bv[k - PS + 1]) + (Bv[k + PS - 1] + Bv[k + PS + 1]));
        const float d = fabsf(0.5f - cv) < fabsf(0.5f - nv) ? nv : cv;
        X[r * PS + (c ^ 1)] = intp(d, he, ve);
    }
    __syncthreads();

    // ---- step 4.1: PQ_Dir at R/B sites (L221-227) -> X[R/B slot] (lpf is dead).
    // The reference keeps P/Q_CDiff_Hpf at half resolution (index indx/2, written from odd columns of
    // every row); a reader at (r,c) therefore sees the writers (r-1,a), (r,b), (r+1,a+2) for P and
    // (r-1,a+2), (r,b), (r+1,a) for Q with a = 2*((c-1)/2)+1, b = 2*(c/2)+1.
    for (int i = tid; i < NR * (RW / 2); i += NTHREADS) {
        const int r = i / (RW / 2), j = i - r * (RW / 2);
        const int c = 2 * j + (fc(a.filters, gr0 + r, 0) & 1);
        const int tr_ = rr0 + r, tc_ = cc0 + c;
        if (!(c >= 4 && c < NC - 4 && r >= 4 && r < NR - 4 && tr_ >= 4 && tr_ < T - 4 && tc_ >= 4 && tc_ < C - 4)) continue;
        const int ca = 2 * ((c - 1) >> 1) + 1, cb = 2 * (c >> 1) + 1;
        const float ps = stdmax(epssq, E[(r - 1) * PS + ca] + E[r * PS + cb] + E[(r + 1) * PS + ca + 2]);
        const float qs = stdmax(epssq, E[(r - 1) * PS + ca + 1] + E[r * PS + cb - 1] + E[(r + 1) * PS + ca - 1]);
        X[r * PS + c] = ps / (ps + qs);
    }
    __syncthreads();

    // ---- step 4.2: the opposite colour at R/B sites (L230-258) -> E[R/B slot]
    for (int i = tid; i < NR * (RW / 2); i += NTHREADS) {
        const int r = i / (RW / 2), j = i - r * (RW / 2);
        const int c = 2 * j + (fc(a.filters, gr0 + r, 0) & 1);
        const int tr_ = rr0 + r, tc_ = cc0 + c;
        if (c >= NC) continue;
        float ck = 0.f;
        if (c >= 4 && c < NC - 4 && r >= 4 && r < NR - 4 && tr_ >= 4 && tr_ < T - 4 && tc_ >= 4 && tc_ < C - 4) {
            const int k = r * PS + c;
            // diagonal neighbours of an R/B site carry colour (2 - fc) natively: rgb[c] there is cfa
            const float cv = X[k];
            const float nv = 0.25f * (X[k - PS - 1] + X[k - PS + 1] + X[k + PS - 1] + X[k + PS + 1]);
            const float d = (fabsf(0.5f - cv) < fabsf(0.5f - nv)) ? nv : cv;
            const float g0 = X[r * PS + (c ^ 1)];
            // rgb[1] at (r+-2, c+-2): R/B sites like the centre -> interpolated G in the partner slot
            const float gnw = X[(r - 2) * PS + ((c - 2) ^ 1)], gne = X[(r - 2) * PS + ((c + 2) ^ 1)];
            const float gsw = X[(r + 2) * PS + ((c - 2) ^ 1)], gse = X[(r + 2) * PS + ((c + 2) ^ 1)];
            const float nwg = eps + fabsf(A[k - PS - 1] - A[k + PS + 1]) + fabsf(A[k - PS - 1] - A[k - 3 * PS - 3]) + fabsf(g0 - gnw);
            const float neg = eps + fabsf(A[k - PS + 1] - A[k + PS - 1]) + fabsf(A[k - PS + 1] - A[k - 3 * PS + 3]) + fabsf(g0 - gne);
            const float swg = eps + fabsf(A[k - PS + 1] - A[k + PS - 1]) + fabsf(A[k + PS - 1] - A[k + 3 * PS - 3]) + fabsf(g0 - gsw);
            const float seg = eps + fabsf(A[k - PS - 1] - A[k + PS + 1]) + fabsf(A[k + PS + 1] - A[k + 3 * PS + 3]) + fabsf(g0 - gse);
            // rgb[1] at the diagonal neighbours (R/B sites): partner slots
            const float nwe = A[k - PS - 1] - X[(r - 1) * PS + ((c - 1) ^ 1)];
            const float nee = A[k - PS + 1] - X[(r - 1) * PS + ((c + 1) ^ 1)];
            const float swe = A[k + PS - 1] - X[(r + 1) * PS + ((c - 1) ^ 1)];
            const float see = A[k + PS + 1] - X[(r + 1) * PS + ((c + 1) ^ 1)];
            const float pe = (nwg * see + seg * nwe) / (nwg + seg);
            const float qe = (neg * swe + swg * nee) / (neg + swg);
            ck = g0 + intp(d, qe, pe);
        }
        E[r * PS + c] = ck;   // P/Q (read in step 4.1, before the barrier above) are dead
    }
    __syncthreads();

    // ---- step 4.3 (R,B at G sites, L261-302) fused with the write-out (L305-316)
    const int ow = oc1 - oc0, oh = or1 - or0;
    for (int i = tid; i < oh * SUB; i += NTHREADS) {
        const int orow = i / SUB, ocol = i - orow * SUB;
        if (ocol >= ow) continue;
        const int r = or0 + orow - rr0, c = oc0 + ocol - cc0;          // region-local
        const int grow = R0 + or0 + orow, gcol = C0 + oc0 + ocol;      // image
        const int k = r * PS + c;
        const unsigned col = fc(a.filters, grow, gcol);
        float red, green, blue;
        if (col == 1) {
            const float cv = Bv[k];
            const float nv = 0.25f * ((Bv[k - PS - 1] + Bv[k - PS + 1]) + (Bv[k + PS - 1] + Bv[k + PS + 1]));
            const float d = (fabsf(0.5f - cv) < fabsf(0.5f - nv)) ? nv : cv;
            const float g = A[k];
            const float n1 = eps + fabsf(g - A[k - 2 * PS]);
            const float s1 = eps + fabsf(g - A[k + 2 * PS]);
            const float w1 = eps + fabsf(g - A[k - 2]);
            const float e1 = eps + fabsf(g - A[k + 2]);
            const float gn = X[(r - 1) * PS + (c ^ 1)], gs = X[(r + 1) * PS + (c ^ 1)];
            const float gw = X[r * PS + ((c - 1) ^ 1)], ge = X[r * PS + ((c + 1) ^ 1)];
            const unsigned hcol = fc(a.filters, grow, gcol + 1);       // colour of the horizontal neighbours
            float outc[3];
            #pragma unroll
            for (int kk = 0; kk <= 2; kk += 2) {
                // plane kk at vertical neighbours (colour 2-hcol natively) and horizontal neighbours (hcol natively)
                const float* pv = (kk == (int)hcol) ? E : A;    // vertical sites: native iff kk != hcol
                const float* ph = (kk == (int)hcol) ? A : E;
                const float vn1 = pv[k - PS], vs1 = pv[k + PS], vn3 = pv[k - 3 * PS], vs3 = pv[k + 3 * PS];
                const float hw1 = ph[k - 1], he1 = ph[k + 1], hw3 = ph[k - 3], he3 = ph[k + 3];
                const float sn = fabsf(vn1 - vs1);
                const float ew = fabsf(hw1 - he1);
                const float ngr = n1 + sn + fabsf(vn1 - vn3);
                const float sgr = s1 + sn + fabsf(vs1 - vs3);
                const float wgr = w1 + ew + fabsf(hw1 - hw3);
                const float egr = e1 + ew + fabsf(he1 - he3);
                const float ne = vn1 - gn, se = vs1 - gs, we = hw1 - gw, ee = he1 - ge;
                const float ve = (ngr * se + sgr * ne) / (ngr + sgr);
                const float he = (egr * we + wgr * ee) / (egr + wgr);
                outc[kk] = g + intp(d, he, ve);
            }
            red = outc[0]; green = g; blue = outc[2];
        } else {
            const float nat = A[k], oth = E[k];
            green = X[r * PS + (c ^ 1)];
            red = (col == 0) ? nat : oth;
            blue = (col == 0) ? oth : nat;
        }
        const size_t o = (size_t)grow * a.op + gcol;
        a.R[o] = stdmax(0.f, red * scale);
        a.G[o] = stdmax(0.f, green * scale);
        a.B[o] = stdmax(0.f, blue * scale);
    }
}

// border_interpolate2 (demosaic_algos.cc L200-353): one thread per ring pixel.
__global__ void border_kernel(const float* __restrict__ raw, size_t rp, float* __restrict__ R, float* __restrict__ G,
                              float* __restrict__ B, size_t op, int W, int H, unsigned filters, int bord, int row_begin, int row_end)
{
    // ring pixels enumerated as: full rows [0,bord) and [H-bord,H) (W each), then for the middle rows
    // the 2*bord edge columns
    const long long nfull = 2ll * bord * W;
    const long long nside = (long long)(H - 2 * bord) * 2 * bord;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nfull + nside) return;
    int i, j;
    if (t < nfull) {
        const int rr = (int)(t / W);
        j = (int)(t - (long long)rr * W);
        i = rr < bord ? rr : H - 2 * bord + rr;
    } else {
        const long long u = t - nfull;
        const int rr = (int)(u / (2 * bord));
        const int cc = (int)(u - (long long)rr * 2 * bord);
        i = bord + rr;
        j = cc < bord ? cc : W - 2 * bord + cc;
    }
    if (i < row_begin || i >= row_end) return;
    float sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i1 = i - 1; i1 < i + 2; ++i1)
        for (int j1 = j - 1; j1 < j + 2; ++j1)
            if (i1 > -1 && i1 < H && j1 > -1 && j1 < W) {
                const unsigned k = fc(filters, i1, j1);
                const float v = raw[(size_t)i1 * rp + j1];
                // sum[k] += v; sum[k+3]++  -- predicated to keep the array in registers
                sum[0] += (k == 0) ? v : 0.f; sum[3] += (k == 0) ? 1.f : 0.f;
                sum[1] += (k == 1) ? v : 0.f; sum[4] += (k == 1) ? 1.f : 0.f;
                sum[2] += (k == 2) ? v : 0.f; sum[5] += (k == 2) ? 1.f : 0.f;
            }
    const unsigned k = fc(filters, i, j);
    const float x = raw[(size_t)i * rp + j];
    const size_t o = (size_t)i * op + j;
    if (k == 1) { R[o] = sum[0] / sum[3]; G[o] = x; B[o] = sum[2] / sum[5]; }
    else {
        G[o] = sum[1] / sum[4];
        if (k == 0) { R[o] = x; B[o] = sum[2] / sum[5]; }
        else { R[o] = sum[0] / sum[3]; B[o] = x; }
    }
}

}  // namespace

int art_border_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, int bord, const float* raw, size_t rp,
                   float* R, float* G, float* B, size_t op, int row_begin, int row_end)
{
    const long long n = 2ll * bord * W + (long long)(H - 2 * bord) * 2 * bord;
    const int bs = 256;
    art_prof_begin(ctx, "border_kernel");
    border_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, ctx->stream>>>(raw, rp, R, G, B, op, W, H, filters, bord, row_begin, row_end);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_rcd_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                float* R, float* G, float* B, size_t op, int row_begin, int row_end)
{
    if (!(ctx->attrs_set & art_hp_ctx::ATTR_RCD)) {
        ART_CUDA(ctx, cudaFuncSetAttribute(rcd_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        ART_CUDA(ctx, cudaFuncSetAttribute(rcd_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        ctx->attrs_set |= art_hp_ctx::ATTR_RCD;
    }
    const char* e = getenv("ART_HP_RCD_THREADS");
    const int nthreads = (e && atoi(e) == 512) ? 512 : 1024;
    const int nth = H / TN + ((H % TN) ? 1 : 0), ntw = W / TN + ((W % TN) ? 1 : 0);   // L86-87
    // reference tile row tr writes image rows [176 tr + 9, 176 tr + 185) (L305-316); bands are cut at 176 k + 9
    const int tr_begin = row_begin <= TB ? 0 : (row_begin - TB) / TN;
    const int tr_end = row_end >= H ? nth : std::min(nth, (row_end - TB) / TN);
    if (tr_end <= tr_begin) return art_border_dev(ctx, W, H, filters, TB, raw, rp, R, G, B, op, row_begin, row_end);
    RcdArgs a{raw, rp, R, G, B, op, W, H, filters, ntw, tr_begin};
    dim3 grid(2 * ntw, 2 * (tr_end - tr_begin));
    art_prof_begin(ctx, "rcd_kernel");
    if (nthreads == 512) rcd_kernel<512><<<grid, 512, SMEM_BYTES, ctx->stream>>>(a);
    else rcd_kernel<1024><<<grid, 1024, SMEM_BYTES, ctx->stream>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return art_border_dev(ctx, W, H, filters, TB, raw, rp, R, G, B, op, row_begin, row_end);   // L342
}
