// Dual demosaic and its flat-region demosaicers for sm_100a.
//
// Replaces (reference rtengine/)
//   RawImageSource::vng4_demosaic + vng4interpolate_row_redblue   vng4_demosaic_RT.cc L32-397
//   RawImageSource::dual_demosaic_RT after the first demosaicer   dual_demosaic_RT.cc L39-152
//   Color::RGB2L                                                  color.cc L1343-1380
//   buildBlendMask incl. its automatic contrast threshold         rt_algo.cc L65-170, L317-494
//   bayer_bilinear_demosaic(blend, ...)                           bayer_bilinear_demosaic.cc L33-75
//   fast_xtrans_interpolate_blend                                 xtrans_demosaic.cc L1033-1092
//
// VNG4.  The reference fills a 4-channel image bilinearly, then walks it with a per-phase gradient program.  Here:
//   k_vng4_fill    one thread per pixel: the three missing channels from the raw 3x3 neighbourhood (the reference's in-place fill only
//                  ever reads native samples, so it is a pure function of the raw plane), one 16-byte store per pixel
//   k_vng4_green   32x8 pixel tiles + 2-pixel halo of the 4-channel image in shared memory; the 16 phase programs (at most 64 terms of
//                  |a - b| << w into one or two of eight gradients) built on the host from the CFA, 8 KB in shared memory; gradient sums
//                  in term order, threshold min + max / 2, green from the directions under it
//   k_vng4_rb      red / blue from colour differences against that green
//   border_interpolate2(3) from rcd.cu.
// Dual demosaic.  L of the first demosaicer's frame (SSE2 group-of-four LUT semantics via a warp vote of the four lanes of a group),
// contrast -> blend factor (vector / scalar sleef exp by column class), Gaussian blur (gauss.cu), then one blend kernel per second demosaicer.
// The automatic threshold is a search for the flattest tile: tile averages / variances keep the reference's four SSE lane sums (one thread
// per lane, the lane sum serial as in the reference), a (value, index) lexicographic minimum reproduces "first strictly smaller wins", and
// the whole decision chain stays on the device (a state word per pass; kernels of a pass that is not taken return at once), so the entry
// never synchronises.
// Compiled with -fmad=false, IEEE division / sqrt.
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>

namespace {

__device__ __host__ __forceinline__ unsigned fc4(unsigned f, int row, int col) { return (f >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
__device__ __forceinline__ float max0(float v) { return 0.f < v ? v : 0.f; }
__device__ __forceinline__ float intp(float a, float b, float c) { return a * b + (1.f - a) * c; }

// ------------------------------------------------------------------ VNG4
struct VTerm { signed char dy1, dx1, dy2, dx2; unsigned char color, wshift, gmask, pad; };
struct VHood { signed char dy, dx; unsigned char has_g, pad; };
struct VProg { VTerm term[16][64]; VHood hood[16][8]; int nterm[16]; };

const signed char VNG_TERMS[64 * 6] = {
    -2, -2, +0, -1, 0, 0x01, -2, -2, +0, +0, 1, 0x01, -2, -1, -1, +0, 0, 0x01,
    -2, -1, +0, -1, 0, 0x02, -2, -1, +0, +0, 0, 0x03, -2, -1, +0, +1, 1, 0x01,
    -2, +0, +0, -1, 0, 0x06, -2, +0, +0, +0, 1, 0x02, -2, +0, +0, +1, 0, 0x03,
    -2, +1, -1, +0, 0, 0x04, -2, +1, +0, -1, 1, 0x04, -2, +1, +0, +0, 0, 0x06,
    -2, +1, +0, +1, 0, 0x02, -2, +2, +0, +0, 1, 0x04, -2, +2, +0, +1, 0, 0x04,
    -1, -2, -1, +0, 0, (signed char)0x80, -1, -2, +0, -1, 0, 0x01, -1, -2, +1, -1, 0, 0x01,
    -1, -2, +1, +0, 1, 0x01, -1, -1, -1, +1, 0, (signed char)0x88, -1, -1, +1, -2, 0, 0x40,
    -1, -1, +1, -1, 0, 0x22, -1, -1, +1, +0, 0, 0x33, -1, -1, +1, +1, 1, 0x11,
    -1, +0, -1, +2, 0, 0x08, -1, +0, +0, -1, 0, 0x44, -1, +0, +0, +1, 0, 0x11,
    -1, +0, +1, -2, 1, 0x40, -1, +0, +1, -1, 0, 0x66, -1, +0, +1, +0, 1, 0x22,
    -1, +0, +1, +1, 0, 0x33, -1, +0, +1, +2, 1, 0x10, -1, +1, +1, -1, 1, 0x44,
    -1, +1, +1, +0, 0, 0x66, -1, +1, +1, +1, 0, 0x22, -1, +1, +1, +2, 0, 0x10,
    -1, +2, +0, +1, 0, 0x04, -1, +2, +1, +0, 1, 0x04, -1, +2, +1, +1, 0, 0x04,
    +0, -2, +0, +0, 1, (signed char)0x80, +0, -1, +0, +1, 1, (signed char)0x88, +0, -1, +1, -2, 0, 0x40,
    +0, -1, +1, +0, 0, 0x11, +0, -1, +2, -2, 0, 0x40, +0, -1, +2, -1, 0, 0x20,
    +0, -1, +2, +0, 0, 0x30, +0, -1, +2, +1, 1, 0x10, +0, +0, +0, +2, 1, 0x08,
    +0, +0, +2, -2, 1, 0x40, +0, +0, +2, -1, 0, 0x60, +0, +0, +2, +0, 1, 0x20,
    +0, +0, +2, +1, 0, 0x30, +0, +0, +2, +2, 1, 0x10, +0, +1, +1, +0, 0, 0x44,
    +0, +1, +1, +2, 0, 0x10, +0, +1, +2, -1, 1, 0x40, +0, +1, +2, +0, 0, 0x60,
    +0, +1, +2, +1, 0, 0x20, +0, +1, +2, +2, 0, 0x10, +1, -2, +1, +0, 0, (signed char)0x80,
    +1, -1, +1, +1, 0, (signed char)0x88, +1, +0, +1, +2, 0, 0x08, +1, +0, +2, -1, 0, 0x40,
    +1, +0, +2, +1, 0, 0x10
};
const signed char VNG_CHOOD[16] = {-1, -1, -1, 0, -1, +1, 0, +1, +1, +1, +1, 0, +1, -1, 0, -1};

// the gradient program of each of the 8 x 2 phases (vng4_demosaic_RT.cc L224-281)
void vng4_program(unsigned prefilters, VProg& P)
{
    memset(&P, 0, sizeof P);
    for (int row = 0; row < 8; row++)
        for (int col = 0; col < 2; col++) {
            const int ph = row * 2 + col;
            const signed char* cp = VNG_TERMS;
            int n = 0;
            for (int t = 0; t < 64; t++) {
                const int y1 = *cp++, x1 = *cp++, y2 = *cp++, x2 = *cp++, weight = *cp++, grads = (unsigned char)*cp++;
                const unsigned color = fc4(prefilters, row + y1, col + x1);
                if (fc4(prefilters, row + y2, col + x2) != color) continue;
                const int diag = (fc4(prefilters, row, col + 1) == color && fc4(prefilters, row + 1, col) == color) ? 2 : 1;
                if (std::abs(y1 - y2) == diag && std::abs(x1 - x2) == diag) continue;
                VTerm& T = P.term[ph][n++];
                T.dy1 = (signed char)y1; T.dx1 = (signed char)x1; T.dy2 = (signed char)y2; T.dx2 = (signed char)x2;
                T.color = (unsigned char)color; T.wshift = (unsigned char)weight; T.gmask = (unsigned char)grads;
            }
            P.nterm[ph] = n;
            cp = VNG_CHOOD;
            const unsigned color = fc4(prefilters, row, col);
            for (int g = 0; g < 8; g++) {
                const int y = *cp++, x = *cp++;
                P.hood[ph][g].dy = (signed char)y; P.hood[ph][g].dx = (signed char)x;
                P.hood[ph][g].has_g = (fc4(prefilters, row + y, col + x) != color && fc4(prefilters, row + y * 2, col + x * 2) == color) ? 1 : 0;
            }
        }
}

__global__ void __launch_bounds__(256) k_vng4_fill(const float* __restrict__ raw, size_t rp, float4* __restrict__ image, int W, int H, unsigned prefilters)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= W) return;
    const unsigned own = fc4(prefilters, row, col);
    const float x0 = raw[(size_t)row * rp + col];
    float pix[4] = {0.f, 0.f, 0.f, 0.f};
    if (row >= 1 && row < H - 1 && col >= 1 && col < W - 1) {
        float sum[4] = {0.f, 0.f, 0.f, 0.f}, wsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int y = -1; y <= 1; y++)
#pragma unroll
            for (int x = -1; x <= 1; x++) {
                const int shift = (y == 0) + (x == 0);
                if (shift == 2) continue;
                const unsigned color = fc4(prefilters, row + y, col + x);
                const float v = raw[(size_t)(row + y) * rp + (col + x)] * (float)(1 << shift), w = (float)(1 << shift);
#pragma unroll
                for (unsigned c = 0; c < 4; ++c) { sum[c] = color == c ? sum[c] + v : sum[c]; wsum[c] = color == c ? wsum[c] + w : wsum[c]; }
            }
#pragma unroll
        for (unsigned c = 0; c < 4; ++c) pix[c] = sum[c] * (1.f / wsum[c]);
    }
#pragma unroll
    for (unsigned c = 0; c < 4; ++c) pix[c] = own == c ? x0 : pix[c];
    image[(size_t)row * W + col] = make_float4(pix[0], pix[1], pix[2], pix[3]);
}

constexpr int VT_W = 32, VT_H = 8, VT_SW = VT_W + 4, VT_SH = VT_H + 4;
__global__ void __launch_bounds__(VT_W * VT_H) k_vng4_green(const float4* __restrict__ image, const VProg* __restrict__ prog, float* __restrict__ green, size_t gp,
                                                            int W, int H, unsigned prefilters)
{
    __shared__ float tile[VT_SH][VT_SW][4];
    __shared__ VTerm terms[16][64];
    __shared__ VHood hood[16][8];
    __shared__ int nterm[16];
    const int tid = threadIdx.y * VT_W + threadIdx.x;
    {
        const uint2* src = reinterpret_cast<const uint2*>(&prog->term[0][0]);
        uint2* dst = reinterpret_cast<uint2*>(&terms[0][0]);
        for (int i = tid; i < 16 * 64; i += VT_W * VT_H) dst[i] = src[i];
        const unsigned* hs = reinterpret_cast<const unsigned*>(&prog->hood[0][0]);
        unsigned* hd = reinterpret_cast<unsigned*>(&hood[0][0]);
        for (int i = tid; i < 16 * 8; i += VT_W * VT_H) hd[i] = hs[i];
        if (tid < 16) nterm[tid] = prog->nterm[tid];
    }
    const int col0 = blockIdx.x * VT_W - 2, row0 = blockIdx.y * VT_H - 2;
    for (int i = tid; i < VT_SH * VT_SW; i += VT_W * VT_H) {
        const int ly = i / VT_SW, lx = i % VT_SW;
        const int r = row0 + ly, c = col0 + lx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0 && r < H && c >= 0 && c < W) v = image[(size_t)r * W + c];
        tile[ly][lx][0] = v.x; tile[ly][lx][1] = v.y; tile[ly][lx][2] = v.z; tile[ly][lx][3] = v.w;
    }
    __syncthreads();
    const int col = col0 + 2 + threadIdx.x, row = row0 + 2 + threadIdx.y;
    if (row < 2 || row >= H - 2 || col < 2 || col >= W - 2) return;
    const int ly = threadIdx.y + 2, lx = threadIdx.x + 2;
    int color = (int)fc4(prefilters, row, col);
    const int ph = (row & 7) * 2 + (col & 1);
    float gval[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int nt = nterm[ph];
    for (int t = 0; t < nt; ++t) {
        const VTerm T = terms[ph][t];
        const float diff = fabsf(tile[ly + T.dy1][lx + T.dx1][T.color] - tile[ly + T.dy2][lx + T.dx2][T.color]) * (float)(1 << T.wshift);
#pragma unroll
        for (int g = 0; g < 8; ++g) gval[g] = (T.gmask >> g & 1) ? gval[g] + diff : gval[g];
    }
    float mn = gval[0], mx = gval[0];
#pragma unroll
    for (int g = 1; g < 8; g++) { if (gval[g] < mn) mn = gval[g]; if (mx < gval[g]) mx = gval[g]; }
    const float thold = mn + mx * 0.5f;
    float sum0 = 0.f, sum1 = 0.f;
    const float greenval = tile[ly][lx][color];
    int num = 0;
    if (color & 1) {
        color ^= 2;
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (gval[g] <= thold) {
                const VHood h = hood[ph][g];
                if (h.has_g) sum0 += greenval + tile[ly + 2 * h.dy][lx + 2 * h.dx][color ^ 2];
                sum1 += tile[ly + h.dy][lx + h.dx][color];
                num++;
            }
        sum0 *= 0.5f;
    } else {
#pragma unroll
        for (int g = 0; g < 8; g++)
            if (gval[g] <= thold) {
                const VHood h = hood[ph][g];
                if (h.has_g) sum0 += greenval + tile[ly + 2 * h.dy][lx + 2 * h.dx][color];
                sum1 += tile[ly + h.dy][lx + h.dx][1] + tile[ly + h.dy][lx + h.dx][3];
                num++;
            }
    }
    green[(size_t)row * gp + col] = max0(greenval + (sum1 - sum0) / (float)(2 * num));
}

// vng4interpolate_row_redblue (L32-55) for rows / columns 3 .. n-4; `filters` has the two greens collapsed
__global__ void __launch_bounds__(256) k_vng4_rb(const float* __restrict__ raw, size_t rp, float* __restrict__ red, const float* __restrict__ green,
                                                 float* __restrict__ blue, size_t op, int W, int H, unsigned filters)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y + 3;
    if (j < 3 || j >= W - 3 || i >= H - 3) return;
    float *ar = red + (size_t)i * op, *ab = blue + (size_t)i * op;
    const float *pg = green + (size_t)(i - 1) * op, *cg = green + (size_t)i * op, *ng = green + (size_t)(i + 1) * op;
    if (fc4(filters, i, 0) == 2 || fc4(filters, i, 1) == 2) { float* t = ar; ar = ab; ab = t; }
#define RAW(r, c) raw[(size_t)(r) * rp + (c)]
    if (fc4(filters, i, j) != 1) {
        ar[j] = RAW(i, j);
        float rb = (RAW(i - 1, j - 1) - pg[j - 1] + RAW(i + 1, j - 1) - ng[j - 1]);
        rb += (RAW(i - 1, j + 1) - pg[j + 1] + RAW(i + 1, j + 1) - ng[j + 1]);
        ab[j] = max0(cg[j] + rb * 0.25f);
    } else {
        ar[j] = max0(cg[j] + (RAW(i, j - 1) - cg[j - 1] + RAW(i, j + 1) - cg[j + 1]) / 2);
        ab[j] = max0(cg[j] + (RAW(i - 1, j) - pg[j] + RAW(i + 1, j) - ng[j]) / 2);
    }
#undef RAW
}

// ------------------------------------------------------------------ Color::RGB2L
__device__ __forceinline__ float vmaxf_(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float vminf_(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }
__device__ __forceinline__ float lut_v(const float* __restrict__ data, int size, float index)
{   // LUT.h L349-377
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}
__device__ __forceinline__ float xyz2lab_y(const float* __restrict__ cachefy, float f)
{   // Color::computeXYZ2LabY, color.cc L1262-1274; LUTf::operator[](float) with LUT_CLIP_BELOW, LUT.h L437-459
    const double kappa = 24389.0 / 27.0;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * (kappa * (double)f / 65535.f));
    if (f > 65535.f) return 327.68f * (116.f * sleef::xcbrtf_scalar(f / 65535.f) - 16.f);
    int idx = (int)f;
    if (f > 65534.f) idx = 65534;
    const float diff = f - (float)idx;
    const float p1 = cachefy[idx];
    const float p2 = cachefy[idx + 1] - p1;
    return p1 + p2 * diff;
}

__global__ void __launch_bounds__(256) k_rgb2l(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ B, size_t ip,
                                               float* __restrict__ L, size_t lp, int W, int H, float w0, float w1, float w2, const float* __restrict__ cachefy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        float yv = 0.f;
        if (x < W) { const size_t i = (size_t)y * ip + x; yv = w0 * R[i] + w1 * G[i] + w2 * B[i]; }
        // the SSE2 loop `for (i = 0; i < W - 3; i += 4)`: a group takes the LUT route unless one of its four Y leaves [0, 65535]
        int slow = (x < W && (yv > 65535.f || yv < 0.f)) ? 1 : 0;
        slow |= __shfl_xor_sync(0xffffffffu, slow, 1);
        slow |= __shfl_xor_sync(0xffffffffu, slow, 2);
        if (x >= W) continue;
        const bool vec = (x & ~3) < W - 3;
        L[(size_t)y * lp + x] = (vec && !slow) ? lut_v(cachefy, 65536, yv) : xyz2lab_y(cachefy, yv);
    }
}

// ------------------------------------------------------------------ buildBlendMask
// device state of the automatic threshold
struct AcState { float thr; int done; float minvar; int minI, minJ; int y0, x0, ny, nx; };

__global__ void __launch_bounds__(256) k_dual_contrast(const float* __restrict__ L, size_t lp, float* __restrict__ blend, size_t bp, int W, int H,
                                                       float thr_value, const AcState* __restrict__ st, float amount, float scale)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float thr = st ? st->thr : thr_value;
    const int i = min(max(x, 2), W - 3);                        // left / right border columns copy column 2 / W-3
    const bool vec = 2 + ((i - 2) & ~3) < W - 5;                // the 4-wide loop `for (i = 2; i < W - 5; i += 4)` covers this column
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        if (thr == 0.f) { blend[(size_t)y * bp + x] = amount; continue; }
        const int j = min(max(y, 2), H - 3);                    // upper / lower border rows copy row 2 / H-3
        const float* row = L + (size_t)j * lp;
        const float a = row[i + 1] - row[i - 1], b = row[i + lp] - row[i - (ptrdiff_t)lp];
        const float c = row[i + 2] - row[i - 2], d = row[i + 2 * lp] - row[i - 2 * (ptrdiff_t)lp];
        const float contrast = sqrtf(a * a + b * b + c * c + d * d) * scale;
        const float arg = 16.f - 16.f * contrast / thr;
        const float e = vec ? sleef::xexpf_vector(arg) : sleef::xexpf_scalar(arg);
        blend[(size_t)y * bp + x] = amount * (1.f / (1.f + e));
    }
}

// blend factor a consumer uses: the blurred mask, or `amount` (1) everywhere when the threshold came out as zero (rt_algo.cc L417-422: no blur then)
__device__ __forceinline__ float blend_at(const float* __restrict__ blurred, const float* __restrict__ unblurred, const AcState* __restrict__ st, size_t o)
{
    return (st && st->thr == 0.f) ? unblurred[o] : blurred[o];
}

// ---- automatic contrast threshold (rt_algo.cc L65-170, L317-414)
// one thread per (tile, SSE lane): lane k sums columns tileX + 4 g + k, rows outer, groups inner -- the order of the reference's vector accumulator;
// tile sizes are 80 and 40, so the scalar tail of the reference loops is empty (asserted on the host)
__global__ void __launch_bounds__(128) k_ac_scores(const float* __restrict__ lum, size_t lp, int ts, int skip, int ntw, int nth, int pass, int local,
                                                   float minLum, float maxLum, float* __restrict__ scores, const AcState* __restrict__ st)
{
    if (st->done) return;
    int oy = 0, ox = 0;
    if (local) { ntw = st->nx; nth = st->ny; oy = st->y0; ox = st->x0; }
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long tile = gid >> 2;
    const int k = (int)(gid & 3);
    const bool live = tile < (long)ntw * nth;
    float score = INFINITY;
    const int ti = live ? (int)(tile / ntw) : 0, tj = live ? (int)(tile % ntw) : 0;
    const int tileY = oy + ti * skip, tileX = ox + tj * skip;
    const float* base = lum + (size_t)tileY * lp + tileX + k;
    float v = 0.f;
    if (live)
        for (int y = 0; y < ts; ++y) {
            const float* r = base + (size_t)y * lp;
            for (int x = 0; x < ts; x += 4) v += r[x];
        }
    // vhadd: (v0 + v2) + (v1 + v3); the four lanes of a tile are four consecutive lanes of the warp
    float a = v + __shfl_xor_sync(0xffffffffu, v, 2);          // lanes 0,2 hold v0 + v2; lanes 1,3 hold v1 + v3 (commutative, same bits)
    float s = a + __shfl_xor_sync(0xffffffffu, a, 1);          // (v0 + v2) + (v1 + v3) on even lanes; (v1 + v3) + (v0 + v2) on odd ones: same bits
    const float avg = (0.f + s) / (float)(ts * ts);
    float w = 0.f;
    if (live)
        for (int y = 0; y < ts; ++y) {
            const float* r = base + (size_t)y * lp;
            for (int x = 0; x < ts; x += 4) { const float t = r[x] - avg; w += t * t; }
        }
    a = w + __shfl_xor_sync(0xffffffffu, w, 2);
    s = a + __shfl_xor_sync(0xffffffffu, a, 1);
    const float var = (0.f + s) / ((float)(ts * ts) * avg);
    if (live && k == 0) {
        if (avg < minLum || avg > maxLum) score = INFINITY;
        else score = var < 0.5f ? INFINITY : var;
        scores[tile] = score;
    }
}

// first strictly smaller value in raster order = lexicographic minimum of (value, index); NaN never wins (v < minvar is false)
__global__ void __launch_bounds__(1024) k_ac_argmin(const float* __restrict__ scores, int ntw, int nth, int local, AcState* __restrict__ st)
{
    if (st->done) return;
    __shared__ float sv[1024];
    __shared__ long si[1024];
    if (local) { ntw = st->nx; nth = st->ny; }
    const long n = (long)ntw * nth;
    float bv = INFINITY;
    long bi = 0x7fffffffffffffffL;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = scores[i];
        if (v < bv) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const float ov = sv[threadIdx.x + s];
            const long oi = si[threadIdx.x + s];
            if (ov < sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool found = sv[0] < INFINITY;          // nothing below +inf: minvar stays +inf, minI = minJ = 0
        st->minvar = sv[0];
        st->minI = found ? (int)(si[0] / ntw) : 0;
        st->minJ = found ? (int)(si[0] % ntw) : 0;
    }
}

// after pass 1's grid search: the pixel-by-pixel search window around the best 40 x 40 tile (L375-380)
__global__ void k_ac_window(AcState* st, int W, int H, int ts, int skip)
{
    if (st->done) return;
    const int minY = skip * st->minI, minX = skip * st->minJ;
    const int y0 = max(minY - skip, 0), x0 = max(minX - skip, 0);
    const int y1 = min(minY + skip, H - ts), x1 = min(minX + skip, W - ts);
    st->y0 = y0; st->x0 = x0; st->ny = y1 - y0 + 1; st->nx = x1 - x0 + 1;
}

// calcContrastThreshold (L112-170) on the chosen tile; mode 0: after pass 0 (taken when minvar <= 1), mode 1: after pass 1's local search
__global__ void __launch_bounds__(512) k_ac_threshold(const float* __restrict__ lum, size_t lp, int ts, int skip, int mode, float factor, AcState* __restrict__ st)
{
    extern __shared__ float bl[];            // (ts - 4)^2 contrasts, then 100 x 4 lane sums
    if (st->done) return;
    int tileY, tileX;
    if (mode == 0) {
        if (!(st->minvar <= 1.f)) return;
        tileY = skip * st->minI; tileX = skip * st->minJ;
    } else {
        if (!(st->minvar <= 8.f)) {
            __syncthreads();
            if (threadIdx.x == 0) { st->thr = 0.f; st->done = 1; }
            return;
        }
        tileY = st->y0 + st->minI; tileX = st->x0 + st->minJ;
    }
    const int n = ts - 4;
    const float scale = 0.0625f / 327.68f * factor;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
        const int jj = i / n, ii = i % n;
        const float* p = lum + (size_t)(tileY + 2 + jj) * lp + (tileX + 2 + ii);
        const float a = p[1] - p[-1], b = p[lp] - p[-(ptrdiff_t)lp], c = p[2] - p[-2], d = p[2 * lp] - p[-2 * (ptrdiff_t)lp];
        bl[i] = sqrtf(a * a + b * b + c * c + d * d) * scale;
    }
    float* lanes = bl + n * n;               // [100][4]
    __syncthreads();
    const int nvec = (ts - 7 + 3) / 4;       // groups of `for (i = 0; i < ts - 7; i += 4)`; 4 nvec == n for ts = 80, 40
    for (int u = threadIdx.x; u < 99 * 4; u += blockDim.x) {
        const int c = 1 + u / 4, k = u & 3;
        const float thr = c / 100.f;
        float v = 0.f;
        for (int j = 0; j < n; ++j)
            for (int g = 0; g < nvec; ++g) v += 1.f / (1.f + sleef::xexpf_vector(16.f - 16.f * bl[j * n + 4 * g + k] / thr));
        lanes[c * 4 + k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float limit = (float)(n * n) / 100.f;
        int c;
        for (c = 1; c < 100; ++c) {
            const float* v = lanes + c * 4;
            const float sum = 0.f + ((v[0] + v[2]) + (v[1] + v[3]));
            if (sum <= limit) break;
        }
        st->thr = c / 100.f;
        st->done = 1;
    }
}

__global__ void k_ac_init(AcState* st, float thr) { st->thr = thr; st->done = 0; st->minvar = INFINITY; st->minI = st->minJ = 0; st->y0 = st->x0 = 0; st->ny = st->nx = 0; }

// ------------------------------------------------------------------ the second demosaicers, mixed in by the blend factor
__global__ void __launch_bounds__(256) k_bilinear_blend(const float* __restrict__ raw, size_t rp, const float* __restrict__ blend, const float* __restrict__ flat, size_t bp,
                                                        const AcState* __restrict__ st, float* __restrict__ red, float* __restrict__ green, float* __restrict__ blue,
                                                        size_t op, int W, int H, unsigned filters)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y + 1;
    if (i >= H - 1 || x >= W) return;
    const int j0 = 2 - (int)(fc4(filters, i, 1) & 1);
    if (x < j0) return;
    const bool first = ((x - j0) & 1) == 0;
    const int j = first ? x : x - 1;         // the pair's first column
    if (j >= W - 2) return;
    float *n1 = red, *n2 = blue;
    if (fc4(filters, i, 0) == 2 || fc4(filters, i, 1) == 2) { float* t = n1; n1 = n2; n2 = t; }
#define RAW(r, c) raw[(size_t)(r) * rp + (c)]
    const size_t o = (size_t)i * op + x;
    const float b = blend_at(blend, flat, st, (size_t)i * bp + x);
    if (first) {
        green[o] = intp(b, green[o], RAW(i, j));
        n1[o] = intp(b, n1[o], (RAW(i, j - 1) + RAW(i, j + 1)) * 0.5f);
        n2[o] = intp(b, n2[o], (RAW(i - 1, j) + RAW(i + 1, j)) * 0.5f);
    } else {
        green[o] = intp(b, green[o], ((RAW(i - 1, j + 1) + RAW(i, j)) + (RAW(i, j + 2) + RAW(i + 1, j + 1))) * 0.25f);
        n1[o] = intp(b, n1[o], RAW(i, j + 1));
        n2[o] = intp(b, n2[o], ((RAW(i - 1, j) + RAW(i - 1, j + 2)) + (RAW(i + 1, j) + RAW(i + 1, j + 2))) * 0.25f);
    }
#undef RAW
}

__global__ void __launch_bounds__(256) k_mix3(const float* __restrict__ blend, const float* __restrict__ flat, size_t bp, const AcState* __restrict__ st,
                                              float* __restrict__ red, float* __restrict__ green, float* __restrict__ blue, size_t op,
                                              const float* __restrict__ tr, const float* __restrict__ tg, const float* __restrict__ tb, size_t tp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * op + x, t = (size_t)y * tp + x;
        const float b = blend_at(blend, flat, st, (size_t)y * bp + x);
        red[o] = intp(b, red[o], tr[t]);
        green[o] = intp(b, green[o], tg[t]);
        blue[o] = intp(b, blue[o], tb[t]);
    }
}

struct XtCfa { int m[36]; };
__global__ void __launch_bounds__(256) k_xtrans_fast_blend(const float* __restrict__ raw, size_t rp, const float* __restrict__ blend, const float* __restrict__ flat, size_t bp,
                                                           const AcState* __restrict__ st, float* __restrict__ red, float* __restrict__ green, float* __restrict__ blue,
                                                           size_t op, int W, int H, const __grid_constant__ XtCfa cfa)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y + 8;
    if (row >= H - 8 || col < 8 || col >= W - 8) return;
#define FCOL(r, c) cfa.m[((r) % 6) * 6 + ((c) % 6)]
    float sum[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int v = -1; v <= 1; v++)
#pragma unroll
        for (int h = -1; h <= 1; h++) {
            const float wgt = (v == 0 && h == 0) ? 0.f : ((v == 0 || h == 0) ? 0.5f : 0.25f);
            const int f = FCOL(row + v, col + h);
            const float t = raw[(size_t)(row + v) * rp + (col + h)] * wgt;
#pragma unroll
            for (int c = 0; c < 3; ++c) sum[c] = f == c ? sum[c] + t : sum[c];
        }
    const size_t o = (size_t)row * op + col;
    const float bl = blend_at(blend, flat, st, (size_t)row * bp + col), x = raw[(size_t)row * rp + col];
    switch (FCOL(row, col)) {
    case 0:
        red[o] = intp(bl, red[o], x); green[o] = intp(bl, green[o], sum[1] * 0.5f); blue[o] = intp(bl, blue[o], sum[2]);
        break;
    case 1:
        green[o] = intp(bl, green[o], x);
        if (FCOL(row, col - 1) == FCOL(row, col + 1)) { red[o] = intp(bl, red[o], sum[0]); blue[o] = intp(bl, blue[o], sum[2]); }
        else { red[o] = intp(bl, red[o], sum[0] * 1.3333333f); blue[o] = intp(bl, blue[o], sum[2] * 1.3333333f); }
        break;
    case 2:
        red[o] = intp(bl, red[o], sum[0]); green[o] = intp(bl, green[o], sum[1] * 0.5f); blue[o] = intp(bl, blue[o], x);
        break;
    }
#undef FCOL
}

struct DualTabs { float cachefy[65537]; VProg prog; unsigned prog_filters; };

int dual_tables(art_hp_ctx* ctx, const float** cachefy)
{
    int rc = art_reserve(ctx, ctx->d_dual_tabs, round_up(sizeof(float) * 65537, 256) + round_up(sizeof(VProg), 256) + round_up(sizeof(AcState), 256));      // the three slots below
    if (rc) return rc;
    float* d = (float*)ctx->d_dual_tabs.p;
    if (!ctx->dual_tabs_ready) {       // Color::cachefy, color.cc L205-233 (host libm cbrt, as in the reference)
        std::vector<float> cfy(65537);
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
        const int epsmaxint = (int)(MAXVALF * eps);
        int i = 0;
        for (; i <= epsmaxint; i++) cfy[i] = (float)(327.68 * (kappa * i / MAXVALF));
        for (; i < 65536; i++) cfy[i] = (float)(327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0));
        cfy[65536] = cfy[65535];
        ART_CUDA(ctx, cudaMemcpyAsync(d, cfy.data(), sizeof(float) * 65537, cudaMemcpyHostToDevice, ctx->stream));
        ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));       // the vector goes out of scope
        ctx->dual_tabs_ready = true;
    }
    *cachefy = d;
    return ART_HP_OK;
}
VProg* dual_prog_slot(art_hp_ctx* ctx) { return (VProg*)((char*)ctx->d_dual_tabs.p + round_up(sizeof(float) * 65537, 256)); }
AcState* dual_state_slot(art_hp_ctx* ctx) { return (AcState*)((char*)dual_prog_slot(ctx) + round_up(sizeof(VProg), 256)); }

}  // namespace

// device address of the threshold the last art_dual_blend_dev on this context used (null before the first call)
const float* art_dual_threshold_slot(art_hp_ctx* ctx) { return ctx->d_dual_tabs.p ? &dual_state_slot(ctx)->thr : nullptr; }

// RawImageSource::vng4_demosaic: raw -> red / green / blue (all written, border 3 by border_interpolate2)
int art_vng4_dev(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* raw, size_t rp, float* R, float* G, float* B, size_t op)
{
    cudaStream_t st = ctx->stream;
    const unsigned filters = prefilters & ~((prefilters & 0x55555555u) << 1);      // the two greens collapsed (vng4_demosaic_RT.cc L86)
    const float* cachefy = nullptr;
    int rc = dual_tables(ctx, &cachefy);
    if (rc) return rc;
    VProg* d_prog = dual_prog_slot(ctx);
    if (!ctx->dual_prog_ready || ctx->dual_prog_filters != prefilters) {
        static thread_local VProg P;
        vng4_program(prefilters, P);
        ART_CUDA(ctx, cudaMemcpyAsync(d_prog, &P, sizeof P, cudaMemcpyHostToDevice, st));
        ART_CUDA(ctx, cudaStreamSynchronize(st));        // P is rebuilt by the next call
        ctx->dual_prog_ready = true;
        ctx->dual_prog_filters = prefilters;
    }
    void* blk = nullptr;
    if ((rc = art_pool_alloc(ctx, (size_t)W * H * sizeof(float4), &blk))) return rc;
    float4* image = (float4*)blk;
    art_prof_begin(ctx, "k_vng4_fill");
    k_vng4_fill<<<dim3((W + 255) / 256, H), 256, 0, st>>>(raw, rp, image, W, H, prefilters);
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_vng4_green");
    k_vng4_green<<<dim3((W + VT_W - 1) / VT_W, (H + VT_H - 1) / VT_H), dim3(VT_W, VT_H), 0, st>>>(image, d_prog, G, op, W, H, prefilters);
    art_prof_end(ctx);
    if (H > 6) {
        art_prof_begin(ctx, "k_vng4_rb");
        k_vng4_rb<<<dim3((W + 255) / 256, H - 6), 256, 0, st>>>(raw, rp, R, G, B, op, W, H, filters);
        art_prof_end(ctx);
    }
    ctx->launches += 3;
    art_pool_free(ctx, blk);
    ART_CUDA(ctx, cudaGetLastError());
    return art_border_dev(ctx, W, H, filters, 3, raw, rp, R, G, B, op, 0, H);
}

// dual_demosaic_RT after the first demosaicer (R / G / B hold its frame).  second: 0 bilinear (Bayer), 1 VNG4 (Bayer; cfa = prefilters),
// 2 fast X-Trans (xtrans36).  contrast in percent; auto_contrast != 0: buildBlendMask's own threshold.  d_threshold_out (optional, device):
// receives the threshold used (contrast / 100).
int art_dual_blend_dev(art_hp_ctx* ctx, int second, int W, int H, unsigned cfa, const int* xtrans36, const float* raw, size_t rp,
                       float* R, float* G, float* B, size_t op, double contrast, int auto_contrast, float* d_threshold_out)
{
    cudaStream_t st = ctx->stream;
    const float* cachefy = nullptr;
    int rc = dual_tables(ctx, &cachefy);
    if (rc) return rc;
    AcState* state = dual_state_slot(ctx);
    const size_t lp = round_up((size_t)W, 32), pl = lp * H;
    const int nplanes = second == 1 ? 6 : 3;
    void* blk = nullptr;
    if ((rc = art_pool_alloc(ctx, pl * nplanes * sizeof(float), &blk))) return rc;
    float *L = (float*)blk, *flat = L + pl, *blend = flat + pl;
    const dim3 blk256(256), grid((W + 255) / 256, std::min(H, 1024));
    // L95-106
    art_prof_begin(ctx, "k_rgb2l");
    k_rgb2l<<<grid, blk256, 0, st>>>(R, G, B, op, L, lp, W, H, (float)0.212671, (float)0.715160, (float)0.072169, cachefy);     // dual_demosaic_RT.cc L95-99: double literals into a float matrix
    art_prof_end(ctx);
    ctx->launches++;
    const float contrastf = (float)(contrast / 100.0);
    k_ac_init<<<1, 1, 0, st>>>(state, contrastf);
    ctx->launches++;
    if (auto_contrast) {
        const float minLum = 2000.f, maxLum = 20000.f;       // luminance_factor 1
        size_t nscores = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const int ts = 80 / (pass + 1), skip = pass == 0 ? ts : ts / 4;
            const long n = (long)std::max(0, W / skip - 3 * pass) * std::max(0, H / skip - 3 * pass);
            nscores = std::max(nscores, (size_t)std::max(n, (long)(2 * skip + 1) * (2 * skip + 1)));
        }
        if ((rc = art_reserve(ctx, ctx->d_small2, nscores * sizeof(float) + 256))) { art_pool_free(ctx, blk); return rc; }
        float* scores = (float*)ctx->d_small2.p;
        art_prof_begin(ctx, "auto_contrast");
        for (int pass = 0; pass < 2; ++pass) {
            const int ts = 80 / (pass + 1), skip = pass == 0 ? ts : ts / 4;
            const int ntw = W / skip - 3 * pass, nth = H / skip - 3 * pass;
            const long nt = (long)std::max(0, ntw) * std::max(0, nth);
            if (nt > 0) {
                k_ac_scores<<<(unsigned)((nt * 4 + 127) / 128), 128, 0, st>>>(L, lp, ts, skip, ntw, nth, pass, 0, minLum, maxLum, scores, state);
                ctx->launches++;
            }
            k_ac_argmin<<<1, 1024, 0, st>>>(scores, std::max(0, ntw), std::max(0, nth), 0, state);
            ctx->launches++;
            const size_t smem = ((size_t)(ts - 4) * (ts - 4) + 400) * sizeof(float);
            if (pass == 0) {
                k_ac_threshold<<<1, 512, smem, st>>>(L, lp, ts, skip, 0, 1.f, state);
                ctx->launches++;
            } else {
                k_ac_window<<<1, 1, 0, st>>>(state, W, H, ts, skip);
                const long nl = (long)(2 * skip + 1) * (2 * skip + 1);
                k_ac_scores<<<(unsigned)((nl * 4 + 127) / 128), 128, 0, st>>>(L, lp, ts, 1, 0, 0, pass, 1, minLum, maxLum, scores, state);
                k_ac_argmin<<<1, 1024, 0, st>>>(scores, 0, 0, 1, state);
                k_ac_threshold<<<1, 512, smem, st>>>(L, lp, ts, skip, 1, 1.f, state);
                ctx->launches += 4;
            }
        }
        art_prof_end(ctx);
    }
    art_prof_begin(ctx, "k_dual_contrast");
    k_dual_contrast<<<grid, blk256, 0, st>>>(L, lp, flat, lp, W, H, contrastf, state, 1.f, 0.0625f / 327.68f * 1.f);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    // gaussianBlur(blend, blend, W, H, 2.0); when the threshold is zero the consumers read the unblurred plane (= amount everywhere)
    if ((rc = art_gauss_dev(ctx, flat, lp, blend, lp, W, H, 2.0))) { art_pool_free(ctx, blk); return rc; }
    if (d_threshold_out) ART_CUDA(ctx, cudaMemcpyAsync(d_threshold_out, &state->thr, sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (second == 0) {
        if (H > 2) {
            art_prof_begin(ctx, "k_bilinear_blend");
            k_bilinear_blend<<<dim3((W + 255) / 256, H - 2), 256, 0, st>>>(raw, rp, blend, flat, lp, state, R, G, B, op, W, H, cfa);
            art_prof_end(ctx);
            ctx->launches++;
        }
    } else if (second == 1) {
        float *tr = blend + pl, *tg = tr + pl, *tb = tg + pl;
        if ((rc = art_vng4_dev(ctx, W, H, cfa, raw, rp, tr, tg, tb, lp))) { art_pool_free(ctx, blk); return rc; }
        art_prof_begin(ctx, "k_mix3");
        k_mix3<<<grid, blk256, 0, st>>>(blend, flat, lp, state, R, G, B, op, tr, tg, tb, lp, W, H);
        art_prof_end(ctx);
        ctx->launches++;
    } else {
        XtCfa c;
        for (int i = 0; i < 36; ++i) c.m[i] = xtrans36[i];
        if (H > 16) {
            art_prof_begin(ctx, "k_xtrans_fast_blend");
            k_xtrans_fast_blend<<<dim3((W + 255) / 256, H - 16), 256, 0, st>>>(raw, rp, blend, flat, lp, state, R, G, B, op, W, H, c);
            art_prof_end(ctx);
            ctx->launches++;
        }
    }
    art_pool_free(ctx, blk);
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
