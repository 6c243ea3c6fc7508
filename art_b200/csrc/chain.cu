// The per-pixel colour / curve chain of ImProcFunctions::process (reference rtengine/improcfun.cc L567-641) fused into
// one pass over the three planes: 12 B read + 12 B written per pixel where the reference makes one full pass per stage
// plus two colour-space conversions (~240 B/px).
//
//   STAGE_1  exposure            ipexposure.cc expcomp L29-73
//   STAGE_3  saturationVibrance  ipsaturation.cc L44-83
//            toneCurve           iptonecurve.cc L553-716 for basecurve LINEAR, contrast 0, white point 1, one curve in mode
//                                STD or FILMLIKE: filmlike_clip (L214-231) + StandardToneCurve / AdobeToneCurve::Apply
//                                (curves.h L360-368, L425-472) over the host-built lutToneCurve
//            rgbCurves           iprgbcurves.cc L63-146 over the three host-built LUTs
//            labAdjustments      iplabadjustments.cc L185-344: Imagefloat::setMode(LAB) (imagefloat.cc L841-878), the L / a / b
//                                LUT loop (L252-283), and the setMode(RGB) the next stage triggers (imagefloat.cc L949-972)
//
// The reference runs every row as 4-pixel SSE2 groups plus a scalar tail of W % 4 pixels, and the two code paths round
// differently (LUT interpolation, Lab2XYZ's Y branch, a group falling back to scalar Lab code when one of its pixels is
// out of range).  One thread owns one group (or the tail of its row) and follows the same rules: bit-identical results.
// Curves and LUTs stay host-built (curves.cc) and are passed by pointer, as SURVEY.md 8(b) prescribes.
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>

namespace {

struct ChainArgs {
    float *r, *g, *b; size_t pitch; int W, H;
    int do_exp; float exp_scale, black;
    int do_sat, vib_on; float saturation, vibrance, noise, wy0, wy1, wy2;
    int tc_mode; const float* tc_lut; float Lmax;
    const float *rc, *gc, *bc;
    int do_lab; const float *lc, *ac, *bcl; float chroma; const float *cachef, *cachefy;
    float ws[9], iws[9];
};

__device__ __forceinline__ float maxr(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float minr(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float vmaxf_(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float vminf_(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }

constexpr int CLIP_BELOW = 1, CLIP_ABOVE = 2;
__device__ __forceinline__ float lut_s(const float* __restrict__ data, int size, int clip, float index)
{   // LUT.h L437-459
    int idx = (int)index;
    if (index < 0.f || !(index == index)) {
        if (clip & CLIP_BELOW) return data[0];
        idx = 0;
    } else if (index > (float)(size - 2)) {
        if (clip & CLIP_ABOVE) return data[size - 1];
        idx = size - 2;
    }
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
__device__ __forceinline__ float lut_v(const float* __restrict__ data, int size, float index)
{   // LUT.h L349-377
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}

__device__ __forceinline__ float pow_F(float a, float b) { return sleef::xexpf_scalar(b * sleef::xlogf_scalar(a)); }

__device__ __forceinline__ float apply_vibrance(float x, float vib, float noise)
{   // ipsaturation.cc L29-38
    const float ax = fabsf(x / 65535.f);
    if (ax > noise) {
        const float sgn = (float)((0.f < x) - (x < 0.f));
        return sgn * pow_F(ax, vib) * 65535.f;
    }
    return x;
}

__device__ __forceinline__ void clip_tone(float& r, float& g, float& b, float L)
{   // color.cc L6650-6658
    const float r_ = r > L ? L : r;
    const float b_ = b > L ? L : b;
    const float g_ = b_ + ((r_ - b_) * (g - b) / (r - b));
    r = r_; g = g_; b = b_;
}
__device__ __forceinline__ void filmlike_clip(float& r, float& g, float& b, float L)
{   // color.cc L6662-6688
    if (r >= g) {
        if (g > b) clip_tone(r, g, b, L);
        else if (b > r) clip_tone(b, r, g, L);
        else if (b > g) clip_tone(r, b, g, L);
        else { r = r > L ? L : r; g = g > L ? L : g; b = g; }
    } else {
        if (r >= b) clip_tone(g, r, b, L);
        else if (b > g) clip_tone(b, g, r, L);
        else clip_tone(g, b, r, L);
    }
}
__device__ __forceinline__ void set_lut_val(const float* __restrict__ lut, float& v) { v = lut_s(lut, 65536, CLIP_BELOW | CLIP_ABOVE, maxr(v, 0.f)); }
__device__ __forceinline__ void rgb_tone(const float* __restrict__ lut, float& r, float& g, float& b)
{   // curves.h L462-472
    const float rold = r, gold = g, bold = b;
    set_lut_val(lut, r);
    set_lut_val(lut, b);
    g = b + ((r - b) * (gold - bold) / (rold - bold));
}

constexpr float D50X = 0.9642f, D50Z = 0.8249f;
__device__ __forceinline__ float xyz2lab_f(const float* __restrict__ cachef, float f)
{   // Color::computeXYZ2Lab, color.cc L1247-1259
    const double kappa = 24389.0 / 27.0;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * ((kappa * (double)f / 65535.f + 16.0) / 116.0));
    else if (f > 65535.f) return 327.68f * sleef::xcbrtf_scalar(f / 65535.f);
    return lut_s(cachef, 65536, CLIP_BELOW, f);
}
__device__ __forceinline__ float xyz2lab_fy(const float* __restrict__ cachefy, float f)
{   // Color::computeXYZ2LabY, color.cc L1262-1274
    const double kappa = 24389.0 / 27.0;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * (kappa * (double)f / 65535.f));
    else if (f > 65535.f) return 327.68f * (116.f * sleef::xcbrtf_scalar(f / 65535.f) - 16.f);
    return lut_s(cachefy, 65536, CLIP_BELOW, f);
}
__device__ __forceinline__ float f2xyz(float f)
{   // color.h L767-770
    const float epsilonExpInv3f = (float)(6.0 / 29.0), kappaInvf = (float)(27.0 / 24389.0);
    return (f > epsilonExpInv3f) ? f * f * f : (116.f * f - 16.f) * kappaInvf;
}

// everything before the Lab stage, one pixel; VEC = the pixel sits in a 4-wide SSE2 group of the reference's row loops
template <bool VEC>
__device__ __forceinline__ void rgb_stages(const ChainArgs& a, float& r, float& g, float& b)
{
    if (a.do_exp) {
        const float tr = r * a.exp_scale - a.black, tg = g * a.exp_scale - a.black, tb = b * a.exp_scale - a.black;
        r = VEC ? vmaxf_(tr, 0.f) : maxr(tr, 0.f);
        g = VEC ? vmaxf_(tg, 0.f) : maxr(tg, 0.f);
        b = VEC ? vmaxf_(tb, 0.f) : maxr(tb, 0.f);
    }
    if (a.do_sat) {
        const float l = r * a.wy0 + g * a.wy1 + b * a.wy2;      // rgbLuminance over the float TMatrix (iccstore.h L38)
        float rl = r - l, gl = g - l, bl = b - l;
        if (a.vib_on) { rl = apply_vibrance(rl, a.vibrance, a.noise); gl = apply_vibrance(gl, a.vibrance, a.noise); bl = apply_vibrance(bl, a.vibrance, a.noise); }
        r = maxr(l + a.saturation * rl, a.noise);
        g = maxr(l + a.saturation * gl, a.noise);
        b = maxr(l + a.saturation * bl, a.noise);
    }
    if (a.tc_mode >= 0) {
        filmlike_clip(r, g, b, a.Lmax);
        if (a.tc_mode == 0) { set_lut_val(a.tc_lut, r); set_lut_val(a.tc_lut, g); set_lut_val(a.tc_lut, b); }
        else {
            r = maxr(0.f, minr(r, a.Lmax)); g = maxr(0.f, minr(g, a.Lmax)); b = maxr(0.f, minr(b, a.Lmax));
            if (r >= g) {
                if (g > b) rgb_tone(a.tc_lut, r, g, b);
                else if (b > r) rgb_tone(a.tc_lut, b, r, g);
                else if (b > g) rgb_tone(a.tc_lut, r, b, g);
                else { set_lut_val(a.tc_lut, r); set_lut_val(a.tc_lut, g); b = g; }
            } else {
                if (r >= b) rgb_tone(a.tc_lut, g, r, b);
                else if (b > g) rgb_tone(a.tc_lut, b, g, r);
                else rgb_tone(a.tc_lut, g, b, r);
            }
        }
    }
    if (a.rc) r = VEC ? lut_v(a.rc, 65536, r) : lut_s(a.rc, 65536, 0, r);
    if (a.gc) g = VEC ? lut_v(a.gc, 65536, g) : lut_s(a.gc, 65536, 0, g);
    if (a.bc) b = VEC ? lut_v(a.bc, 65536, b) : lut_s(a.bc, 65536, 0, b);
}

// lab_adjustments' loop and Lab -> RGB for one pixel already in Lab (L, a, b)
template <bool VEC>
__device__ __forceinline__ void lab_tail(const ChainArgs& A, float L, float a, float bb, float& r, float& g, float& b)
{
    if (VEC) {
        L = lut_v(A.lc, 32770, L);
        a = (lut_v(A.ac, 65536, a + 32768.f) - 32768.f) * A.chroma;
        bb = (lut_v(A.bcl, 65536, bb + 32768.f) - 32768.f) * A.chroma;
    } else {
        L = lut_s(A.lc, 32770, 0, L);
        a = (lut_s(A.ac, 65536, CLIP_BELOW | CLIP_ABOVE, a + 32768.f) - 32768.f) * A.chroma;
        bb = (lut_s(A.bcl, 65536, CLIP_BELOW | CLIP_ABOVE, bb + 32768.f) - 32768.f) * A.chroma;
    }
    const float c1By116 = (float)(1.0 / 116.0), c16By116 = (float)(16.0 / 116.0);
    const double kappa = 24389.0 / 27.0;
    float X, Y, Z;
    const float LL = L / 327.68f, aa = a / 327.68f, b2 = bb / 327.68f;
    const float fy = c1By116 * LL + c16By116;
    const float fx = 0.002f * aa + fy;
    const float fz = fy - (0.005f * b2);
    X = 65535.f * f2xyz(fx) * D50X;
    Z = 65535.f * f2xyz(fz) * D50Z;
    if (VEC) {      // Lab2XYZ(vfloat ...), color.cc L1228-1245
        const float res1 = fy * fy * fy;
        const float res2 = LL / (float)kappa;
        Y = (LL > 8.f) ? res1 : res2;
        Y *= 65535.f;
    } else {        // Lab2XYZ(float ...), L1203-1213
        Y = ((double)LL > 8.0) ? 65535.0f * fy * fy * fy : (float)((double)(65535.0f * LL) / kappa);
    }
    r = A.iws[0] * X + A.iws[1] * Y + A.iws[2] * Z;
    g = A.iws[3] * X + A.iws[4] * Y + A.iws[5] * Z;
    b = A.iws[6] * X + A.iws[7] * Y + A.iws[8] * Z;
}

__global__ void __launch_bounds__(256) k_chain(ChainArgs A)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int x0 = gx * 4;
    if (x0 >= A.W) return;
    A.noise = pow_F(2.f, -16.f);       // ipsaturation.cc L52, evaluated with the same sleef steps
    for (int y = blockIdx.y; y < A.H; y += gridDim.y) {
        const size_t row = (size_t)y * A.pitch;
        if (x0 + 4 <= A.W) {
            float r[4], g[4], b[4];
            const float4 vr = *reinterpret_cast<const float4*>(A.r + row + x0), vg = *reinterpret_cast<const float4*>(A.g + row + x0),
                         vb = *reinterpret_cast<const float4*>(A.b + row + x0);
            r[0] = vr.x; r[1] = vr.y; r[2] = vr.z; r[3] = vr.w;
            g[0] = vg.x; g[1] = vg.y; g[2] = vg.z; g[3] = vg.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) rgb_stages<true>(A, r[k], g[k], b[k]);
            if (A.do_lab) {
                float X[4], Y[4], Z[4];
                bool slow = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // rgb2lab(vfloat ...): rgbxyz + XYZ2Lab, color.cc L841-846, L1401-1437
                    X[k] = A.ws[0] * r[k] + A.ws[1] * g[k] + A.ws[2] * b[k];
                    Y[k] = A.ws[3] * r[k] + A.ws[4] * g[k] + A.ws[5] * b[k];
                    Z[k] = A.ws[6] * r[k] + A.ws[7] * g[k] + A.ws[8] * b[k];
                    X[k] = X[k] / D50X;
                    Z[k] = Z[k] / D50Z;
                    const float mx = vmaxf_(X[k], vmaxf_(Y[k], Z[k])), mn = vminf_(X[k], vminf_(Y[k], Z[k]));
                    slow = slow || mx > 65535.f || mn < 0.f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float L, a, bb;
                    if (slow) {
                        const float fx = xyz2lab_f(A.cachef, X[k]), fy = xyz2lab_f(A.cachef, Y[k]), fz = xyz2lab_f(A.cachef, Z[k]);
                        L = xyz2lab_fy(A.cachefy, Y[k]);
                        a = 500.f * (fx - fy);
                        bb = 200.f * (fy - fz);
                    } else {
                        const float fx = lut_v(A.cachef, 65536, X[k]), fy = lut_v(A.cachef, 65536, Y[k]), fz = lut_v(A.cachef, 65536, Z[k]);
                        L = lut_v(A.cachefy, 65536, Y[k]);
                        a = 500.f * (fx - fy);
                        bb = 200.f * (fy - fz);
                    }
                    lab_tail<true>(A, L, a, bb, r[k], g[k], b[k]);
                }
            }
            *reinterpret_cast<float4*>(A.r + row + x0) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(A.g + row + x0) = make_float4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(A.b + row + x0) = make_float4(b[0], b[1], b[2], b[3]);
        } else {
            for (int x = x0; x < A.W; ++x) {      // the scalar tail of the row
                float r = A.r[row + x], g = A.g[row + x], b = A.b[row + x];
                rgb_stages<false>(A, r, g, b);
                if (A.do_lab) {
                    const float Xs = A.ws[0] * r + A.ws[1] * g + A.ws[2] * b;
                    const float Ys = A.ws[3] * r + A.ws[4] * g + A.ws[5] * b;
                    const float Zs = A.ws[6] * r + A.ws[7] * g + A.ws[8] * b;
                    const float xd = Xs / D50X, zd = Zs / D50Z;
                    const float fx = xyz2lab_f(A.cachef, xd), fy = xyz2lab_f(A.cachef, Ys), fz = xyz2lab_f(A.cachef, zd);
                    const float L = xyz2lab_fy(A.cachefy, Ys);
                    lab_tail<false>(A, L, 500.0f * (fx - fy), 200.0f * (fy - fz), r, g, b);
                }
                A.r[row + x] = r; A.g[row + x] = g; A.b[row + x] = b;
            }
        }
    }
}

}  // namespace

// LUT slots in the context's device / pinned staging: 0 tone curve, 1-3 rgb curves, 4 L curve, 5-6 a / b curves, 7-8 cachef / cachefy
constexpr size_t LUT_SLOT = 65536 + 64;

int art_chain_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const art_hp_chain_params* p)
{
    if ((pitch & 3) || (reinterpret_cast<uintptr_t>(r) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
        return ctx->fail(ART_HP_ERR_INVALID, "planes must be 16-byte aligned with a pitch that is a multiple of 4 floats");
    cudaStream_t st = ctx->stream;
    int rc = art_reserve(ctx, ctx->d_chain, 9 * LUT_SLOT * sizeof(float));
    if (rc) return rc;
    float* d = (float*)ctx->d_chain.p;
    if (!ctx->h_chain) {
        ART_CUDA(ctx, cudaMallocHost(&ctx->h_chain, 9 * LUT_SLOT * sizeof(float)));
        ART_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
    } else {
        ART_CUDA(ctx, cudaEventSynchronize(ctx->ev_chain));        // the previous call's uploads left the staging buffer
    }
    float* h = (float*)ctx->h_chain;
    if (!ctx->chain_cache_ready) {      // Color::cachef / cachefy, color.cc L205-233 (host libm cbrt, as in the reference)
        float* cf = h + 7 * LUT_SLOT; float* cfy = h + 8 * LUT_SLOT;
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
        const int epsmaxint = (int)(MAXVALF * eps);
        int i = 0;
        for (; i <= epsmaxint; i++) { cf[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); cfy[i] = (float)(327.68 * (kappa * i / MAXVALF)); }
        for (; i < 65536; i++) { cf[i] = (float)(327.68 * std::cbrt((double)i / MAXVALF)); cfy[i] = (float)(327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0)); }
        cf[65536] = cf[65535]; cfy[65536] = cfy[65535];
        ART_CUDA(ctx, cudaMemcpyAsync(d + 7 * LUT_SLOT, cf, 2 * LUT_SLOT * sizeof(float), cudaMemcpyHostToDevice, st));
        ctx->chain_cache_ready = true;
    }
    auto up = [&](int slot, const float* src, int n) -> const float* {
        if (!src) return nullptr;
        memcpy(h + slot * LUT_SLOT, src, sizeof(float) * n);
        h[slot * LUT_SLOT + n] = src[n - 1];       // LUT<T> allocates s + 3 entries; data[size] is touched (times a zero weight) at the top index
        cudaMemcpyAsync(d + slot * LUT_SLOT, h + slot * LUT_SLOT, sizeof(float) * (n + 1), cudaMemcpyHostToDevice, st);
        return d + slot * LUT_SLOT;
    };
    ChainArgs a{};
    a.r = r; a.g = g; a.b = b; a.pitch = pitch; a.W = W; a.H = H;
    a.do_exp = p->exposure_enabled; a.exp_scale = p->exp_scale; a.black = p->black;
    a.do_sat = p->saturation_enabled && (p->saturation || p->vibrance);
    a.vib_on = p->vibrance != 0;
    a.saturation = 1.f + p->saturation / 100.f;
    a.vibrance = 1.f - p->vibrance / 1000.f;
    a.tc_mode = p->tonecurve_lut ? p->tonecurve_mode : -1;
    a.tc_lut = up(0, p->tonecurve_lut, 65536);
    a.Lmax = 65535.f * 1.f;
    a.rc = up(1, p->rcurve, 65536); a.gc = up(2, p->gcurve, 65536); a.bc = up(3, p->bcurve, 65536);
    a.do_lab = p->lab_enabled;
    if (a.do_lab) {
        a.lc = up(4, p->lab_lcurve, 32770); a.ac = up(5, p->lab_acurve, 65536); a.bcl = up(6, p->lab_bcurve, 65536);
        a.chroma = p->lab_chroma;
        a.cachef = d + 7 * LUT_SLOT; a.cachefy = d + 8 * LUT_SLOT;
    }
    if ((a.do_sat || a.do_lab) && !p->ws) return ctx->fail(ART_HP_ERR_INVALID, "ws is required by saturation and Lab stages");
    if (a.do_lab && (!p->iws || !p->lab_lcurve || !p->lab_acurve || !p->lab_bcurve)) return ctx->fail(ART_HP_ERR_INVALID, "Lab stage needs iws and the three curves");
    if (a.tc_mode > 1) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "tone curve mode %d", a.tc_mode);
    if (p->ws) { a.wy0 = (float)p->ws[3]; a.wy1 = (float)p->ws[4]; a.wy2 = (float)p->ws[5]; for (int i = 0; i < 9; ++i) a.ws[i] = (float)p->ws[i]; }
    if (p->iws) for (int i = 0; i < 9; ++i) a.iws[i] = (float)p->iws[i];
    ART_CUDA(ctx, cudaEventRecord(ctx->ev_chain, st));
    const dim3 blk(64, 1), grid(((W + 3) / 4 + 63) / 64, std::min(H, 148 * 16));
    art_prof_begin(ctx, "k_chain");
    k_chain<<<grid, blk, 0, st>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
