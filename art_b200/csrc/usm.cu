// Unsharp-mask sharpening: the "usm" route of ImProcFunctions::doSharpening (reference rtengine/ipsharpen.cc L711-790).
//
//   k_usm_lum      get_luminance (rt_algo.cc L942-956) fused with apply_gamma<false>(Y, 1, 3) (ipsharpen.cc L46-78) of the
//                  copy unsharp_mask works on: reads R, G, B once, writes Y and gamma(Y)
//   k_usm_contrast buildBlendMask's contrast sigmoid (rt_algo.cc L416-476): 4-pixel SSE2 groups from column 2 use the vector
//                  xexpf, the row tails the scalar one; the replicated two-pixel border is the clamped interior sample
//   art_gauss_dev  gaussianBlur(blend, blend, sigma = 2 / sqrt(scale)) in place and gaussianBlur(Y', b2, radius / scale)
//   k_usm_apply    unsharp_mask's threshold loop (L270-283, Threshold<int>::multiply in double, procparams.h L476-502),
//                  apply_gamma<true>, and multiply(rgb, YY, Y) (rt_algo.cc L958-975): reads Y', b2, blend, Y, R, G, B, writes R, G, B
//
// Algorithmic bytes per pixel: 12 + 8 (lum) + 4 + 4 (contrast) + 28 + 12 (apply) + the two Gaussians (8 each at the 3x3 / one
// IIR sweep pair) ~ 84-110 B/px (~110).  All kernels are streaming: HBM-bound.
// Bit-identical to the reference's SSE2 build (no FMA contraction, IEEE division and sqrt).
#include "ctx.h"

#include <vector>
#include "sleef_dev.cuh"

namespace {

__device__ __forceinline__ float maxr(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float minr(float a, float b) { return b < a ? b : a; }

__device__ __forceinline__ float lut_clip(const float* __restrict__ data, float index)
{   // LUT.h L437-459, 65536 entries, LUT_CLIP_BELOW | LUT_CLIP_ABOVE
    if (index < 0.f || !(index == index)) return data[0];
    if (index > 65534.f) return data[65535];
    const int idx = (int)index;
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}

__device__ __forceinline__ float gamma_apply(const float* __restrict__ glut, float l, float gamma)
{   // ipsharpen.cc L66-74, pivot 1
    if (l >= 0.f && l < 65536.f) return lut_clip(glut, l);
    l = sleef::pow_F_scalar(maxr(l / 65535.f, 1e-18f), gamma) * 1.f;
    return l * 65535.f;
}

// glut of apply_gamma<false> (slot 0, gamma 1/3) and apply_gamma<true> (slot 1, gamma 3), L56-62
__global__ void k_usm_tables(float* __restrict__ t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536) return;
    const float d = 65535.f * 1.f;
    const float gf = 1.f / 3.f, gr = 3.f;
    float a = 0.f, b = 0.f;
    if (i) {
        a = sleef::pow_F_scalar((float)i / d, gf) * 1.f; a *= 65535.f;
        b = sleef::pow_F_scalar((float)i / d, gr) * 1.f; b *= 65535.f;
    }
    t[i] = a;
    t[65536 + i] = b;
}

__global__ void __launch_bounds__(256) k_usm_lum(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ B, size_t ip,
                                                 float* __restrict__ Y, float* __restrict__ YY, size_t yp, int W, int H,
                                                 float w0, float w1, float w2, const float* __restrict__ glut)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x, o = (size_t)y * yp + x;
        const float l = R[i] * w0 + G[i] * w1 + B[i] * w2;       // Color::rgbLuminance over the float TMatrix
        Y[o] = l;
        YY[o] = gamma_apply(glut, l, 1.f / 3.f);
    }
}

__global__ void __launch_bounds__(256) k_usm_contrast(const float* __restrict__ L, size_t lp, float* __restrict__ blend, size_t bp, int W, int H,
                                                      float cpar, float s_scale, float amount, float scale)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float thr = sleef::pow_F_scalar(cpar, 1.2f) * s_scale;   // doSharpening L727, same sleef steps as the reference's pow_F
    const int i = min(max(x, 2), W - 3);                        // left / right border columns copy column 2 / W-3
    const bool vec = 2 + ((i - 2) & ~3) < W - 5;                // the 4-wide loop `for (i = 2; i < W - 5; i += 4)` covers this column
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
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

__global__ void k_usm_fill(float* __restrict__ p, size_t pp, int W, int H, float v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) p[(size_t)y * pp + x] = v;
}

struct Thr { double bl, tl, br, tr; };

__device__ __forceinline__ float threshold_multiply(const Thr t, float x, float y_max)
{   // Threshold<int>(bl, tl, br, tr, false)::multiply<float, float, float>, procparams.h L476-502
    const double val = x;
    if (val == t.br && t.br == t.tr) return y_max;
    if (val >= t.br) return 0.f;
    if (val > t.tr) return (float)((double)y_max * (1.0 - (val - t.tr) / (t.br - t.tr)));
    if (val >= t.tl) return y_max;
    if (val > t.bl) return (float)((double)y_max * (val - t.bl) / (t.tl - t.bl));
    return 0.f;
}

__global__ void __launch_bounds__(256) k_usm_apply(float* __restrict__ R, float* __restrict__ G, float* __restrict__ B, size_t ip,
                                                   const float* __restrict__ Y, const float* __restrict__ YY, const float* base, const float* __restrict__ b2,
                                                   const float* __restrict__ blend, size_t yp, int W, int H, int amount, Thr thr,
                                                   const float* __restrict__ glut_rev)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x, o = (size_t)y * yp + x;
        const float yy = YY[o], bl = blend[o], den = Y[o];
        const float diff = base[o] - b2[o];                      // base: YY itself, or its bilateral-filtered copy (edgesonly, L263-265)
        const float delta = threshold_multiply(thr, minr(fabsf(diff), 2000.f), amount * diff * 0.01f);
        float v = bl * (yy + delta) + (1.f - bl) * yy;           // intp(blend, Y + delta, Y)
        v = gamma_apply(glut_rev, v, 3.f);
        if (den > 0.f) {
            const float f = v / den;
            R[i] *= f; G[i] *= f; B[i] *= f;
        }
    }
}

// unsharp_mask with halo control: sharpenHaloCtrl (ipsharpen.cc L80-141) over a copy of the gamma-encoded luminance, then the same
// tail as k_usm_apply.  The reference slides a two-entry max / min queue along each row (both entries start at 0); here every
// pixel recomputes the three 3x3-window statistics of columns j-2, j-1, j from the 5x5 neighbourhood it holds in registers.
__global__ void __launch_bounds__(256) k_usm_apply_halo(float* __restrict__ R, float* __restrict__ G, float* __restrict__ B, size_t ip,
                                                        const float* __restrict__ Y, const float* __restrict__ YY, const float* base, const float* __restrict__ b2,
                                                        const float* __restrict__ blend, size_t yp, int W, int H, int amount, int halo_amount, Thr thr,
                                                        const float* __restrict__ glut_rev)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float scl = (100.f - halo_amount) * 0.01f;
    const float sharpFac = amount * 0.01f;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x, o = (size_t)y * yp + x;
        const float labL = YY[o], den = Y[o];
        float v = labL;
        if (y >= 2 && y < H - 2 && x >= 2 && x < W - 2) {
            // nL[y-2 .. y+2][x-2 .. x+2]; columns left of 0 are never used (their windows belong to j < 2, which count as 0)
            float n[5][5];
#pragma unroll
            for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                for (int dx = 0; dx < 5; ++dx) n[dy][dx] = base[o + (ptrdiff_t)(dy - 2) * (ptrdiff_t)yp + (dx - 2)];
            float mx[3], mn[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {           // window statistics of column j = x - 2 + k (columns k .. k + 2 of n)
                float np[3];
#pragma unroll
                for (int q = 0; q < 3; ++q)         // rows q .. q + 2 of n: np1 (rows y-2..y), np2, np3
                    np[q] = 2.f * (n[q][k] + n[q][k + 1] + n[q][k + 2] + n[q + 1][k] + n[q + 1][k + 1] + n[q + 1][k + 2] + n[q + 2][k] + n[q + 2][k + 1] + n[q + 2][k + 2]) / 27.f +
                            n[q + 1][k + 1] / 3.f;
                mx[k] = maxr(maxr(np[0], np[1]), np[2]);
                mn[k] = minr(minr(np[0], np[1]), np[2]);
            }
            const float max1 = x - 2 >= 2 ? mx[0] : 0.f, max2 = x - 1 >= 2 ? mx[1] : 0.f;
            const float min1 = x - 2 >= 2 ? mn[0] : 0.f, min2 = x - 1 >= 2 ? mn[1] : 0.f;
            float max_ = maxr(maxr(max1, max2), mx[2]), min_ = minr(minr(min1, min2), mn[2]);
            if (max_ < labL) max_ = labL;
            if (min_ > labL) min_ = labL;
            const float diff = n[2][2] - b2[o];
            const float delta = threshold_multiply(thr, minr(fabsf(diff), 2000.f), sharpFac * diff);
            float newL = labL + delta;
            if (newL > max_) newL = max_ + (newL - max_) * scl;
            else if (newL < min_) newL = min_ - (min_ - newL) * scl;
            const float bl = blend[o];
            v = bl * newL + (1.f - bl) * labL;
        }
        v = gamma_apply(glut_rev, v, 3.f);
        if (den > 0.f) {
            const float f = v / den;
            R[i] *= f; G[i] *= f; B[i] *= f;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// edgesonly: bilateral<float, float> (bilateral2.h L38-547).  One of 21 fixed integer kernels (3x3 .. 11x11, chosen by sigma in 0.1
// steps) weighted by a range LUT ec[d + 65536] = exp(-d^2 / (2 sens^2)) * scale that is read with LUTf's interpolating float index;
// numerator and denominator are summed in float, row-major over src[i - a][j - b], a, b = -h .. h; the h-pixel border is copied.
// ---------------------------------------------------------------------------------------------------------------------------
struct BlK { int half; float q[36]; };       // quarter kernel (h + 1) x (h + 1), row-major, corner first: the BL_OPERn arguments

__device__ __forceinline__ float bl_lut(const float* __restrict__ ec, float index)
{   // LUT.h L437-459, 0x20000 entries, LUT_CLIP_BELOW | LUT_CLIP_ABOVE
    if (index < 0.f || !(index == index)) return ec[0];
    if (index > 131070.f) return ec[131071];
    const int idx = (int)index;
    const float diff = index - (float)idx;
    const float p1 = ec[idx];
    const float p2 = ec[idx + 1] - p1;
    return p1 + p2 * diff;
}

__global__ void __launch_bounds__(256) k_usm_bilateral(BlK k, const float* __restrict__ src, float* __restrict__ dst, size_t yp, int W, int H,
                                                       const float* __restrict__ ec)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int h = k.half, hw = h + 1;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        if (y < h || x < h || y >= H - h || x >= W - h) { dst[o] = src[o]; continue; }        // BL_END, L52-57
        const float c = src[o];
        float v = 0.f, den = 0.f;
        bool first = true;
        for (int a = -h; a <= h; ++a) {
            const float* row = src + o - (ptrdiff_t)a * (ptrdiff_t)yp;
            const float* qr = k.q + (h - abs(a)) * hw;
            for (int b = -h; b <= h; ++b) {
                const float coef = qr[h - abs(b)];
                const float s = row[-b];
                const float e = bl_lut(ec, s - c + 65536.0f);
                const float tv = coef * (s * e), td = coef * e;
                if (first) { v = tv; den = td; first = false; }
                else { v = v + tv; den = den + td; }
            }
        }
        dst[o] = v / den;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// "rld": RL-deconvolution sharpening, doSharpening L747-771 without the corner boost: markImpulse (rt_algo.cc L497-591),
// deconvsharpening (ipsharpen.cc L144-230) over gaussianBlur's GAUSS_DIV / GAUSS_MULT forms for sigma <= 1.15 (gauss.cc L177-443:
// 3x3 with its border rules, 5x5, 7x7 with the `* c21` slip of L302 / L402 kept), multiply.  20 iterations x 2 stencil passes.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rld_copy_lum(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ B, size_t ip,
                                                      float* __restrict__ Y, size_t yp, int W, int H, float w0, float w1, float w2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x;
        Y[(size_t)y * yp + x] = R[i] * w0 + G[i] * w1 + B[i] * w2;
    }
}

__global__ void __launch_bounds__(256) k_rld_impulse(const float* __restrict__ src, const float* __restrict__ lpf, size_t yp, unsigned char* __restrict__ imp,
                                                     int W, int H, float impthrDiv24)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const float hpfabs = fabsf(src[(size_t)y * yp + x] - lpf[(size_t)y * yp + x]);
        float hfnbrave = 0.f;
        for (int i1 = max(0, y - 2); i1 <= min(y + 2, H - 1); i1++)
            for (int j1 = max(0, x - 2); j1 <= min(x + 2, W - 1); j1++)
                hfnbrave += fabsf(src[(size_t)i1 * yp + j1] - lpf[(size_t)i1 * yp + j1]);
        imp[(size_t)y * yp + x] = hpfabs > ((hfnbrave - hpfabs) * impthrDiv24);
    }
}

struct RldK { int size; float c[8]; float c3[5]; };      // 5x5: c21 c20 c11 c10 c00; 7x7: c31 c30 c22 c21 c20 c11 c10 c00; 3x3: c0 c1 c2 b0 b1

__device__ __forceinline__ float rld_conv(const RldK& k, const float* __restrict__ s, ptrdiff_t p, int W, int H, int x, int y, bool& inner)
{
#define S(dy, dx) s[(dy) * p + (dx)]
    if (k.size == 0) { inner = false; return 0.f; }          // recursive forms: the blur ran in art_gauss_divmult_dev
    if (k.size == 3) {
        inner = true;
        const bool top = (y == 0 || y == H - 1), side = (x == 0 || x == W - 1);
        if (top && side) return S(0, 0);
        if (top) return k.c3[4] * (S(0, -1) + S(0, 1)) + k.c3[3] * S(0, 0);
        if (side) return k.c3[4] * (S(-1, 0) + S(1, 0)) + k.c3[3] * S(0, 0);
        return k.c3[2] * (S(-1, -1) + S(-1, 1) + S(1, -1) + S(1, 1)) + k.c3[1] * (S(-1, 0) + S(0, -1) + S(0, 1) + S(1, 0)) + k.c3[0] * S(0, 0);
    }
    if (k.size == 5) {
        inner = y >= 2 && y < H - 2 && x >= 2 && x < W - 2;
        if (!inner) return 0.f;
        return k.c[0] * (S(-2, -1) + S(-2, 1) + S(-1, -2) + S(-1, 2) + S(1, -2) + S(1, 2) + S(2, -1) + S(2, 1)) +
               k.c[1] * (S(-2, 0) + S(0, -2) + S(0, 2) + S(2, 0)) +
               k.c[2] * (S(-1, -1) + S(-1, 1) + S(1, -1) + S(1, 1)) +
               k.c[3] * (S(-1, 0) + S(0, -1) + S(0, 1) + S(1, 0)) +
               k.c[4] * S(0, 0);
    }
    inner = y >= 3 && y < H - 3 && x >= 3 && x < W - 3;
    if (!inner) return 0.f;
    const float c31 = k.c[0], c30 = k.c[1], c22 = k.c[2], c21 = k.c[3], c20 = k.c[4], c11 = k.c[5], c10 = k.c[6], c00 = k.c[7];
    return c31 * (S(-3, -1) + S(-3, 1) + S(-1, -3) + S(-1, 3) + S(1, -3) + S(1, 3) + S(3, -1) + S(3, 1)) +
           c30 * (S(-3, 0) + S(0, -3) + S(0, 3) + S(3, 0)) +
           c22 * (S(-2, -2) + S(-2, 2) + S(2, -2) + S(2, 2)) +
           c21 * (S(-2, -1) + S(-2, 1) * c21 + S(-1, -2) + S(-1, 2) + S(1, -2) + S(1, 2) + S(2, -1) + S(2, 1)) +
           c20 * (S(-2, 0) + S(0, -2) + S(0, 2) + S(2, 0)) +
           c11 * (S(-1, -1) + S(-1, 1) + S(1, -1) + S(1, 1)) +
           c10 * (S(-1, 0) + S(0, -1) + S(0, 1) + S(1, 0)) +
           c00 * S(0, 0);
#undef S
}

__global__ void __launch_bounds__(256) k_rld_init(const float* __restrict__ Y, float* __restrict__ lum, float* __restrict__ tmpI, float* __restrict__ out,
                                                  size_t yp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        const float l = Y[o] + 1000.f;
        lum[o] = l;
        tmpI[o] = maxr(l, 0.f);
        out[o] = __int_as_float(0x7fc00000);
    }
}

// gaussianBlur(tmpI, tmp, sigma, nullptr, GAUSS_DIV, luminance)
__global__ void __launch_bounds__(256) k_rld_div(RldK k, const float* __restrict__ tmpI, const float* __restrict__ lum, float* __restrict__ tmp, size_t yp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        bool inner;
        const float t = rld_conv(k, tmpI + o, (ptrdiff_t)yp, W, H, x, y, inner);
        if (k.size == 3) tmp[o] = maxr(lum[o] / (t > 0.f ? t : 1.f), 0.f);
        else tmp[o] = inner ? lum[o] / maxr(t, 0.00001f) : 1.f;          // std::max(val, 0.00001f)
    }
}

// gaussianBlur(tmp, tmpI, sigma, nullptr, GAUSS_MULT) fused with check_stop (L189-199) of the same pixel
__global__ void __launch_bounds__(256) k_rld_mult(RldK k, const float* __restrict__ tmp, float* __restrict__ tmpI, const float* __restrict__ lum, float* __restrict__ out,
                                                  const unsigned char* __restrict__ imp, const float* __restrict__ blend, float amount, size_t yp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        bool inner;
        const float t = rld_conv(k, tmp + o, (ptrdiff_t)yp, W, H, x, y, inner);
        float v = tmpI[o];
        if (inner) { v = v * t; tmpI[o] = v; }
        const float cur = out[o];
        if (cur != cur) {
            const float l = lum[o];
            const float delta = l * 0.2f;
            if (fabsf(v - l) > delta) {
                float res;
                if (v != v) res = l;
                else { const float bb = imp[o] ? 0.f : blend[o] * amount; res = bb * maxr(v, 0.0f) + (1.f - bb) * l; }
                out[o] = res;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_rld_final(float* __restrict__ YY, const float* __restrict__ out, const float* __restrict__ tmpI, const float* __restrict__ lum,
                                                   const unsigned char* __restrict__ imp, const float* __restrict__ blend, float amount, size_t yp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        float l = out[o];
        if (l != l) {
            const float v = tmpI[o], lu = lum[o];
            if (v != v) l = lu;
            else { const float bb = imp[o] ? 0.f : blend[o] * amount; l = bb * maxr(v, 0.0f) + (1.f - bb) * lu; }
        }
        YY[o] = maxr(l - 1000.f, 0.f);
    }
}

// YY = intp(CornerBoostMask(x, y), YY2, YY), ipsharpen.cc L313-338, L762-771
__global__ void __launch_bounds__(256) k_rld_corner_mix(float* __restrict__ YY, const float* __restrict__ YY2, size_t yp, int W, int H,
                                                        int ox, int oy, int w2, int h2, float r2, float sg)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t o = (size_t)y * yp + x;
        const int xx = x + ox - w2, yy = y + oy - h2;
        const float distance = sqrtf((float)(xx * xx + yy * yy));
        const float d = maxr(distance - r2, 0.f);
        const float e = sleef::xexpf_scalar((-(d * d)) / sg);
        const float m = 1.f - maxr(0.f, minr(e, 1.f));
        YY[o] = m * YY2[o] + (1.f - m) * YY[o];
    }
}

// multiply(rgb, YY, Y), rt_algo.cc L958-975
__global__ void __launch_bounds__(256) k_rld_multiply(float* __restrict__ R, float* __restrict__ G, float* __restrict__ B, size_t ip,
                                                      const float* __restrict__ YY, const float* __restrict__ Y, size_t yp, int W, int H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const size_t i = (size_t)y * ip + x, o = (size_t)y * yp + x;
        const float den = Y[o];
        if (den > 0.f) {
            const float f = YY[o] / den;
            R[i] *= f; G[i] *= f; B[i] *= f;
        }
    }
}

}  // namespace

static int rld_kernel(double sigma, RldK* k)
{   // the coefficient set gaussianBlurImpl would use for GAUSS_DIV / GAUSS_MULT at this sigma (gauss.cc L1444-1511, L52-92); 0.25 <= sigma <= 1.15
    if (sigma < 0.6) {
        double c0 = 1.0, c1 = std::exp(-0.5 * ((1.0 / sigma) * (1.0 / sigma))), c2 = std::exp(-((1.0 / sigma) * (1.0 / sigma)));
        const double sum = c0 + 4.0 * (c1 + c2);
        c0 /= sum; c1 /= sum; c2 /= sum;
        double b1 = std::exp(-1.0 / (2.0 * sigma * sigma));
        const double bsum = 2.0 * b1 + 1.0;
        b1 /= bsum;
        k->size = 3; k->c3[0] = (float)c0; k->c3[1] = (float)c1; k->c3[2] = (float)c2; k->c3[3] = (float)(1.0 / bsum); k->c3[4] = (float)b1;
        return 0;
    }
    const int half = sigma <= 0.84 ? 2 : 3;
    const double lim = half == 2 ? (3.0 * 0.84) * (3.0 * 0.84) : (3.0 * 1.15) * (3.0 * 1.15);
    const float sg = (float)sigma;
    const double temp = -2.f * (sg * sg);
    float kk[7][7], sum = 0.f;
    for (int i = -half; i <= half; ++i)
        for (int j = -half; j <= half; ++j) {
            if ((i * i + j * j) <= lim) { kk[i + half][j + half] = (float)std::exp((i * i + j * j) / temp); sum += kk[i + half][j + half]; }
            else kk[i + half][j + half] = 0.f;
        }
    for (int i = 0; i <= 2 * half; ++i) for (int j = 0; j <= 2 * half; ++j) kk[i][j] /= sum;
    if (half == 2) { k->size = 5; k->c[0] = kk[0][1]; k->c[1] = kk[0][2]; k->c[2] = kk[1][1]; k->c[3] = kk[1][2]; k->c[4] = kk[2][2]; }
    else { k->size = 7; k->c[0] = kk[0][2]; k->c[1] = kk[0][3]; k->c[2] = kk[1][1]; k->c[3] = kk[1][2]; k->c[4] = kk[1][3]; k->c[5] = kk[2][2]; k->c[6] = kk[2][3]; k->c[7] = kk[3][3]; }
    return 0;
}

// deconvsharpening(YY = copy of Y, blend, impulse, sigma, amount) into the plane YY
static int rld_deconv(art_hp_ctx* ctx, const float* Y, const float* blend, const unsigned char* imp, double sigma, float amount,
                      float* YY, float* lum, float* tmp, float* tmpI, float* out, size_t yp, int W, int H)
{
    cudaStream_t st = ctx->stream;
    const dim3 blk(256), grid((W + 255) / 256, std::min(H, 148 * 8));
    if (amount <= 0 || sigma < 0.2f) {       // returns at once (L146-155): the copy of Y stays as it is
        ART_CUDA(ctx, cudaMemcpyAsync(YY, Y, yp * (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return ART_HP_OK;
    }
    RldK k{};
    k_rld_init<<<grid, blk, 0, st>>>(Y, lum, tmpI, out, yp, W, H);
    if (sigma > 1.15) {        // gauss.cc L1490-1511: recursive GAUSS_DIV / GAUSS_MULT, then check_stop alone (k.size == 0)
        for (int it = 0; it < 20; ++it) {
            int rc;
            if ((rc = art_gauss_divmult_dev(ctx, tmpI, yp, tmp, yp, lum, yp, W, H, sigma, 2))) return rc;
            if ((rc = art_gauss_divmult_dev(ctx, tmp, yp, tmpI, yp, nullptr, 0, W, H, sigma, 1))) return rc;
            art_prof_begin(ctx, "k_rld_check");
            k_rld_mult<<<grid, blk, 0, st>>>(k, tmp, tmpI, lum, out, imp, blend, amount, yp, W, H);
            art_prof_end(ctx);
        }
        ctx->launches += 20;
    } else {
        rld_kernel(sigma, &k);
        art_prof_begin(ctx, "k_rld_iterations");
        for (int it = 0; it < 20; ++it) {
            k_rld_div<<<grid, blk, 0, st>>>(k, tmpI, lum, tmp, yp, W, H);
            k_rld_mult<<<grid, blk, 0, st>>>(k, tmp, tmpI, lum, out, imp, blend, amount, yp, W, H);
        }
        art_prof_end(ctx);
        ctx->launches += 40;
    }
    art_prof_begin(ctx, "k_rld_final");
    k_rld_final<<<grid, blk, 0, st>>>(YY, out, tmpI, lum, imp, blend, amount, yp, W, H);
    art_prof_end(ctx);
    ctx->launches += 2;
    return ART_HP_OK;
}

static int art_rld_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_sharpen_params* p, const double* ws9)
{
    const double scale = p->scale > 0 ? p->scale : 1.0;
    const double sigma = p->deconvradius / scale;
    const float amount = p->deconvamount / 100.f;
    const float delta = (float)(p->deconvCornerBoost / scale);
    const bool boost = delta > 0.01f;
    auto unsupported = [&](double sg) { return amount > 0 && !(sg < 0.2f) && (sg >= 25.0 || sg < 0.25 || (sg > 1.15 && (W < 4 || H < 4))); };
    if (unsupported(sigma) || (boost && unsupported(sigma + delta)))
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "rld is on the hot path for 0.25 <= sigma < 25 (3x3 / 5x5 / 7x7 and recursive GAUSS_DIV / GAUSS_MULT forms; the recursive ones need W, H >= 4), got %.3f", boost ? sigma + delta : sigma);
    if (!boost && (amount <= 0 || sigma < 0.2f)) return ART_HP_OK;           // deconvsharpening returns at once: multiply() then scales by exactly 1
    cudaStream_t st = ctx->stream;
    int rc;
    const size_t yp = round_up((size_t)W, 32), pl = yp * (size_t)H;
    float* planes = nullptr;
    if ((rc = art_pool_alloc(ctx, (8 * pl + pl / 4 + 64) * sizeof(float), (void**)&planes))) return rc;
    float *Y = planes, *blend = Y + pl, *lum = blend + pl, *tmp = lum + pl, *tmpI = tmp + pl, *out = tmpI + pl, *YY = out + pl, *YY2 = YY + pl;
    unsigned char* imp = reinterpret_cast<unsigned char*>(YY2 + pl);
    const dim3 blk(256), grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_rld_copy_lum");
    k_rld_copy_lum<<<grid, blk, 0, st>>>(r, g, b, ip, Y, yp, W, H, (float)ws9[3], (float)ws9[4], (float)ws9[5]);
    art_prof_end(ctx);
    const float s_scale = (float)std::sqrt(scale);
    if (p->contrast == 0.0) {
        k_usm_fill<<<grid, blk, 0, st>>>(blend, yp, W, H, 1.f);
    } else {
        k_usm_contrast<<<grid, blk, 0, st>>>(Y, yp, blend, yp, W, H, (float)(p->contrast / 100.f), s_scale, 1.f, 0.0625f / 327.68f * 1.f);
        if ((rc = art_gauss_dev(ctx, blend, yp, blend, yp, W, H, (double)(2.f / s_scale)))) { art_pool_free(ctx, planes); return rc; }
    }
    // markImpulse(W, H, Y, impulse, 2.f): lpf = gaussianBlur(Y, max(2, thresh - 1)) into tmp
    if ((rc = art_gauss_dev(ctx, Y, yp, tmp, yp, W, H, 2.0))) { art_pool_free(ctx, planes); return rc; }
    art_prof_begin(ctx, "k_rld_impulse");
    k_rld_impulse<<<grid, blk, 0, st>>>(Y, tmp, yp, imp, W, H, 3.5f / 24.0f);
    art_prof_end(ctx);
    ctx->launches += 4;
    if ((rc = rld_deconv(ctx, Y, blend, imp, sigma, amount, YY, lum, tmp, tmpI, out, yp, W, H))) { art_pool_free(ctx, planes); return rc; }
    if (boost) {        // doSharpening L757-771
        if ((rc = rld_deconv(ctx, Y, blend, imp, sigma + delta, amount, YY2, lum, tmp, tmpI, out, yp, W, H))) { art_pool_free(ctx, planes); return rc; }
        const int fw = p->full_width > 0 ? p->full_width : W, fh = p->full_height > 0 ? p->full_height : H;
        const int w2 = fw / 2, h2 = fh / 2;
        const float radius = (float)std::max(w2, h2);
        const float lat = float(p->deconvCornerLatitude) / 150.f;
        const float lim = lat < 0.f ? 0.f : (lat > 1.f ? 1.f : lat);
        const float r2 = (radius - radius * lim) / 2.f;
        const float sg = 2.f * ((radius * 0.3f) * (radius * 0.3f));
        art_prof_begin(ctx, "k_rld_corner_mix");
        k_rld_corner_mix<<<grid, blk, 0, st>>>(YY, YY2, yp, W, H, p->offset_x, p->offset_y, w2, h2, r2, sg);
        art_prof_end(ctx);
        ctx->launches++;
    }
    art_prof_begin(ctx, "k_rld_multiply");
    k_rld_multiply<<<grid, blk, 0, st>>>(r, g, b, ip, YY, Y, yp, W, H);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    art_pool_free(ctx, planes);
    return ART_HP_OK;
}

// the BL_BEGIN / BL_OPERn arguments of bilateral05 .. bilateral25 (bilateral2.h L151-484): LUT scale, half width, quarter kernel
struct BlKernel { int scale, half; int q[36]; };
static const BlKernel BL_KERNELS[21] = {
    {318, 1, {1, 7, 7, 55}},
    {768, 1, {1, 4, 4, 16}},
    {366, 2, {0, 0, 1, 0, 8, 21, 1, 21, 59}},
    {753, 2, {0, 0, 1, 0, 5, 10, 1, 10, 23}},
    {595, 2, {0, 1, 2, 1, 6, 12, 2, 12, 22}},
    {910, 2, {0, 1, 2, 1, 4, 7, 2, 7, 12}},
    {209, 3, {0, 0, 1, 1, 0, 2, 5, 8, 1, 5, 18, 27, 1, 8, 27, 41}},
    {322, 3, {0, 0, 1, 1, 0, 1, 4, 6, 1, 4, 11, 16, 1, 6, 16, 23}},
    {336, 3, {0, 0, 1, 1, 0, 2, 4, 6, 1, 4, 11, 14, 1, 6, 14, 19}},
    {195, 3, {0, 1, 2, 3, 1, 4, 8, 10, 2, 8, 17, 21, 3, 10, 21, 28}},
    {132, 4, {0, 0, 0, 1, 1, 0, 1, 2, 4, 5, 0, 2, 6, 12, 14, 1, 4, 12, 22, 28, 1, 5, 14, 28, 35}},
    {180, 4, {0, 0, 0, 1, 1, 0, 1, 2, 3, 4, 0, 2, 5, 9, 10, 1, 3, 9, 15, 19, 1, 4, 10, 19, 23}},
    {195, 4, {0, 0, 1, 1, 1, 0, 1, 2, 3, 4, 1, 2, 5, 8, 9, 1, 3, 8, 13, 16, 1, 4, 9, 16, 19}},
    {151, 4, {0, 0, 1, 2, 2, 0, 1, 3, 5, 5, 1, 3, 6, 10, 12, 2, 5, 10, 16, 19, 2, 5, 12, 19, 22}},
    {151, 4, {0, 0, 1, 2, 2, 0, 1, 3, 4, 5, 1, 3, 5, 8, 9, 2, 4, 8, 12, 14, 2, 5, 9, 14, 16}},
    {116, 5, {0, 0, 0, 1, 1, 1, 0, 0, 1, 2, 3, 3, 0, 1, 2, 4, 7, 7, 1, 2, 4, 8, 12, 14, 1, 3, 7, 12, 18, 20, 1, 3, 7, 14, 20, 23}},
    {127, 5, {0, 0, 0, 1, 1, 1, 0, 0, 1, 2, 3, 3, 0, 1, 2, 4, 6, 7, 1, 2, 4, 8, 11, 12, 1, 3, 6, 11, 15, 17, 1, 3, 7, 12, 17, 19}},
    {109, 5, {0, 0, 0, 1, 1, 2, 0, 1, 2, 3, 3, 4, 1, 2, 3, 5, 7, 8, 1, 3, 5, 9, 12, 13, 1, 3, 7, 12, 16, 18, 2, 4, 8, 13, 18, 20}},
    {132, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 3, 5, 6, 7, 1, 2, 5, 7, 10, 11, 1, 3, 6, 10, 13, 14, 1, 3, 7, 11, 14, 16}},
    {156, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 3, 4, 5, 6, 1, 2, 4, 6, 8, 9, 1, 3, 5, 8, 10, 11, 1, 3, 6, 9, 11, 12}},
    {173, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 2, 4, 5, 5, 1, 2, 4, 5, 7, 7, 1, 3, 5, 7, 9, 9, 1, 3, 5, 7, 9, 10}},
};
// kernel k serves sigma < BL_LIMITS[k] (the dispatcher, L487-547), the last one everything above
static const double BL_LIMITS[20] = {0.55, 0.65, 0.75, 0.85, 0.95, 1.05, 1.15, 1.25, 1.35, 1.45, 1.55, 1.65, 1.75, 1.85, 1.95, 2.05, 2.15, 2.25, 2.35, 2.45};

// bilateral<float, float>(src, dst, buffer, W, H, sigma, sens): src -> dst, planes of pitch yp
static int usm_bilateral(art_hp_ctx* ctx, const float* src, float* dst, size_t yp, int W, int H, double sigma, int sens)
{
    cudaStream_t st = ctx->stream;
    if (sigma < 0.45) {         // L490-497
        ART_CUDA(ctx, cudaMemcpyAsync(dst, src, yp * (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return ART_HP_OK;
    }
    int kx = 0;
    while (kx < 20 && !(sigma < BL_LIMITS[kx])) kx++;
    const BlKernel& K = BL_KERNELS[kx];
    int rc = art_reserve(ctx, ctx->d_bl_lut, 0x20000 * sizeof(float));
    if (rc) return rc;
    if (ctx->bl_lut_scale != K.scale || ctx->bl_lut_sens != sens) {     // BL_BEGIN, L42-45: host libm exp in double, as the reference
        std::vector<float> ec(0x20000);
        const double scale = K.scale, s = sens;
        for (int i = 0; i < 0x20000; i++) ec[i] = (float)(std::exp(-(double)(i - 0x10000) * (double)(i - 0x10000) / (2.0 * s * s)) * scale);
        ART_CUDA(ctx, cudaMemcpyAsync(ctx->d_bl_lut.p, ec.data(), 0x20000 * sizeof(float), cudaMemcpyHostToDevice, st));
        ART_CUDA(ctx, cudaStreamSynchronize(st));       // pageable source: keep it alive until the copy has run
        ctx->bl_lut_scale = K.scale; ctx->bl_lut_sens = sens;
    }
    BlK k{};
    k.half = K.half;
    for (int i = 0; i < (K.half + 1) * (K.half + 1); ++i) k.q[i] = (float)K.q[i];
    const dim3 blk(256), grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_usm_bilateral");
    k_usm_bilateral<<<grid, blk, 0, st>>>(k, src, dst, yp, W, H, (const float*)ctx->d_bl_lut.p);
    art_prof_end(ctx);
    ctx->launches++;
    return ART_HP_OK;
}

int art_usm_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_sharpen_params* p, const double* ws9)
{
    if (p->amount < 1 || W < 8 || H < 8) return ART_HP_OK;      // doSharpening L716-718
    if (p->method == 1) return art_rld_dev(ctx, r, g, b, ip, W, H, p, ws9);
    if (p->method != 0) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "sharpening method %d (psf) is not on the hot path", p->method);
    if (p->edgesonly && p->edges_tolerance < 1) return ctx->fail(ART_HP_ERR_INVALID, "edges_tolerance must be >= 1, got %d", p->edges_tolerance);
    if (p->edgesonly && !(p->edges_radius >= 0)) return ctx->fail(ART_HP_ERR_INVALID, "negative edges_radius");
    cudaStream_t st = ctx->stream;
    int rc;
    if (!ctx->usm_tables_ready) {
        if ((rc = art_reserve(ctx, ctx->d_usm_tables, 2 * 65536 * sizeof(float)))) return rc;
        k_usm_tables<<<256, 256, 0, st>>>((float*)ctx->d_usm_tables.p);
        ctx->launches++;
        ctx->usm_tables_ready = true;
    }
    const float* glut = (const float*)ctx->d_usm_tables.p;
    const size_t yp = round_up((size_t)W, 32);
    float* planes = nullptr;
    if ((rc = art_pool_alloc(ctx, (p->edgesonly ? 5 : 4) * yp * (size_t)H * sizeof(float), (void**)&planes))) return rc;
    float *Y = planes, *YY = Y + yp * H, *b2 = YY + yp * H, *blend = b2 + yp * H;
    float* base = p->edgesonly ? blend + yp * H : YY;           // b3 of unsharp_mask (L239-262)
    const dim3 blk(256), grid((W + 255) / 256, std::min(H, 148 * 8));

    art_prof_begin(ctx, "k_usm_lum");
    k_usm_lum<<<grid, blk, 0, st>>>(r, g, b, ip, Y, YY, yp, W, H, (float)ws9[3], (float)ws9[4], (float)ws9[5], glut);
    art_prof_end(ctx);
    ctx->launches++;

    const double scale = p->scale > 0 ? p->scale : 1.0;
    const float s_scale = (float)std::sqrt(scale);
    // contrastThreshold = pow_F(contrast / 100.f, 1.2f) * s_scale (L727) is zero exactly when contrast is: exp(1.2 * log(0)) = 0
    if (p->contrast == 0.0) {     // rt_algo.cc L417-422
        art_prof_begin(ctx, "k_usm_fill");
        k_usm_fill<<<grid, blk, 0, st>>>(blend, yp, W, H, 1.f);
        art_prof_end(ctx);
        ctx->launches++;
    } else {
        art_prof_begin(ctx, "k_usm_contrast");
        k_usm_contrast<<<grid, blk, 0, st>>>(Y, yp, blend, yp, W, H, (float)(p->contrast / 100.f), s_scale, 1.f, 0.0625f / 327.68f * 1.f);
        art_prof_end(ctx);
        ctx->launches++;
        if ((rc = art_gauss_dev(ctx, blend, yp, blend, yp, W, H, (double)(2.f / s_scale)))) { art_pool_free(ctx, planes); return rc; }
    }
    if (p->edgesonly && (rc = usm_bilateral(ctx, YY, base, yp, W, H, p->edges_radius / scale, p->edges_tolerance))) { art_pool_free(ctx, planes); return rc; }
    if ((rc = art_gauss_dev(ctx, base, yp, b2, yp, W, H, p->radius / scale))) { art_pool_free(ctx, planes); return rc; }

    const Thr thr{(double)p->threshold[0], (double)p->threshold[1], (double)p->threshold[2], (double)p->threshold[3]};
    art_prof_begin(ctx, "k_usm_apply");
    if (p->halocontrol) k_usm_apply_halo<<<grid, blk, 0, st>>>(r, g, b, ip, Y, YY, base, b2, blend, yp, W, H, p->amount, p->halocontrol_amount, thr, glut + 65536);
    else k_usm_apply<<<grid, blk, 0, st>>>(r, g, b, ip, Y, YY, base, b2, blend, yp, W, H, p->amount, thr, glut + 65536);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    art_pool_free(ctx, planes);
    return ART_HP_OK;
}
