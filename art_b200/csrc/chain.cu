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
    int tc_mode; const float* tc_lut; float Lmax, whitecoeff;
    unsigned* hist;
    const float* sl;
    const art_hp_curve_stage* stages; int nstages;           // device copies (poly arrays in device memory)
    const float *pq, *pq_inv, *satlut, *hues;                 // jzazbz_pq_, jzazbz_pq_inv_ (color.cc L322-326), satcurve_lut's table, ApplyState's hue constants
    float to_out[9], to_work[9];
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


// ---- the reference's default tone curve: NeutralToneCurve::BatchApply (curves.cc L891-1037) and apply_satcurve (iptonecurve.cc L398-441) ----
__device__ __forceinline__ float powf_cr(float a, float b) { return (float)pow((double)a, (double)b); }   // glibc powf is correctly rounded but for rare ties
__device__ __forceinline__ float pq_fn(float X)
{   // color.cc L67-74
    X = X < 1e-10f ? 1e-10f : X;
    const float XX = powf_cr(X * 1e-4f, 0.1593017578125f);
    return powf_cr((0.8359375f + 18.8515625f * XX) / (1 + 18.6875f * XX), 134.034375f);
}
__device__ __forceinline__ float pq_inv_fn(float X)
{   // color.cc L77-84
    X = X < 1e-10f ? 1e-10f : X;
    const float XX = powf_cr(X, 7.460772656268214e-03f);
    return 1e4f * powf_cr((0.8359375f - XX) / (18.6875f * XX - 18.8515625f), 6.277394636015326f);
}
__device__ __forceinline__ void mat3(const float* m, float a, float b, float c, float& x, float& y, float& z)
{
    x = m[0] * a + m[1] * b + m[2] * c;
    y = m[3] * a + m[4] * b + m[5] * c;
    z = m[6] * a + m[7] * b + m[8] * c;
}
__device__ __forceinline__ void xyz2jzazbz(const float* __restrict__ pq, float X, float Y, float Z, float& Jz, float& az, float& bz)
{   // color.cc L6706-6722, XYZ_D50_to_D65 L37-49
    const float x = 0.9555766f * X + -0.0230393f * Y + 0.0631636f * Z;
    const float y = -0.0282895f * X + 1.0099416f * Y + 0.0210077f * Z;
    const float z = 0.0122982f * X + -0.0204830f * Y + 1.3299098f * Z;
    const float l = 0.674207838f * x + 0.382799340f * y - 0.047570458f * z;
    const float m = 0.149284160f * x + 0.739628340f * y + 0.083327300f * z;
    const float s = 0.070941080f * x + 0.174768000f * y + 0.670970020f * z;
    const float Lp = (l >= 0.f && l <= 1.f) ? lut_s(pq, 65536, 0, l * 65535.f) : pq_fn(l);
    const float Mp = (m >= 0.f && m <= 1.f) ? lut_s(pq, 65536, 0, m * 65535.f) : pq_fn(m);
    const float Sp = (s >= 0.f && s <= 1.f) ? lut_s(pq, 65536, 0, s * 65535.f) : pq_fn(s);
    const float Iz = 0.5f * (Lp + Mp);
    az = 3.524000f * Lp - 4.066708f * Mp + 0.542708f * Sp;
    bz = 0.199076f * Lp + 1.096799f * Mp - 1.295875f * Sp;
    Jz = (0.44f * Iz) / (1.f - 0.56f * Iz) - 1.6295499532821566e-11f;
}
__device__ __forceinline__ void jzazbz2xyz(const float* __restrict__ pqi, float Jz, float az, float bz, float& X, float& Y, float& Z)
{   // color.cc L6724-6742, XYZ_D65_to_D50 L52-64
    Jz = Jz + 1.6295499532821566e-11f;
    const float Iz = Jz / (0.44f + 0.56f * Jz);
    const float l = Iz + 1.386050432715393e-1f * az + 5.804731615611869e-2f * bz;
    const float m = Iz - 1.386050432715393e-1f * az - 5.804731615611891e-2f * bz;
    const float s = Iz - 9.601924202631895e-2f * az - 8.118918960560390e-1f * bz;
    const float L = (l >= 0.f && l <= 1.f) ? lut_s(pqi, 65536, 0, l * 65535.f) : pq_inv_fn(l);
    const float M = (m >= 0.f && m <= 1.f) ? lut_s(pqi, 65536, 0, m * 65535.f) : pq_inv_fn(m);
    const float S = (s >= 0.f && s <= 1.f) ? lut_s(pqi, 65536, 0, s * 65535.f) : pq_inv_fn(s);
    const float x = +1.661373055774069e+00f * L - 9.145230923250668e-01f * M + 2.313620767186147e-01f * S;
    const float y = -3.250758740427037e-01f * L + 1.571847038366936e+00f * M - 2.182538318672940e-01f * S;
    const float z = -9.098281098284756e-02f * L - 3.127282905230740e-01f * M + 1.522766561305260e+00f * S;
    X = 1.0478112f * x + 0.0228866f * y + -0.0501270f * z;
    Y = 0.0295424f * x + 0.9904844f * y + -0.0170491f * z;
    Z = -0.0092345f * x + 0.0150436f * y + 0.7521316f * z;
}
__device__ __forceinline__ void rgb2jzczhz(const float* __restrict__ pq, const float* ws, float R, float G, float B, float& Jz, float& cz, float& hz)
{   // color.h L1791-1796; jzazbz2jzch = yuv2hsl(bz, az, h, c), color.cc L6691-6695
    float X, Y, Z, az, bz;
    mat3(ws, R, G, B, X, Y, Z);
    xyz2jzazbz(pq, X, Y, Z, Jz, az, bz);
    cz = sqrtf(bz * bz + az * az);
    hz = sleef::xatan2f(bz, az);
}
__device__ __forceinline__ void jzczhz2rgb(const float* __restrict__ pqi, const float* iws, float Jz, float cz, float hz, float& R, float& G, float& B)
{   // color.h L1799-1804; jzch2jzazbz = hsl2yuv(h, c, bz, az), color.cc L6698-6703
    float sn, cs, X, Y, Z;
    sleef::xsincosf(hz, sn, cs);
    jzazbz2xyz(pqi, Jz, cz * cs, cz * sn, X, Y, Z);
    mat3(iws, X, Y, Z, R, G, B);
}
// Curve::getVal of the composed curve above the LUT, stage by stage (see art_hp_curve_stage)
__device__ __noinline__ double curve_eval(const art_hp_curve_stage* __restrict__ st, int n, double t)
{
    for (int s = 0; s < n; ++s) {
        if (st[s].kind == 1) {
            const double* __restrict__ px = st[s].poly_x; const int np = st[s].n;
            int lo = 0, hi = np;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (px[mid] < t) lo = mid + 1; else hi = mid; }
            if (lo == np) { t = st[s].poly_y[np - 1]; continue; }
            int d = lo;
            if (lo + 1 < np && t - px[lo] > px[lo + 1] - t) ++d;
            const double v = st[s].poly_y[d];
            t = v < 0.0 ? 0.0 : v;
        } else if (st[s].kind == 2) {
            const double w = st[s].w;
            double x = w < t ? w : t;
            x = x < 0.0 ? 0.0 : x;
            const double p = pow(x / w, st[s].a);
            t = log(p * (st[s].b - 1.0) + 1.0) / log(st[s].b) * w;
        }
    }
    return t;
}
// curves::setLutVal, curves.h L224-231
__device__ __forceinline__ void set_lut_val_c(const ChainArgs& a, float& v)
{
    if (v <= 65535.f || !a.nstages) v = lut_s(a.tc_lut, 65536, CLIP_BELOW | CLIP_ABOVE, maxr(v, 0.f));
    else v = (float)(curve_eval(a.stages, a.nstages, (double)(v / 65535.f)) * 65535.f);
}
__device__ __forceinline__ void rgb_tone(const ChainArgs& a, float& r, float& g, float& b)
{   // curves.h L462-472
    const float rold = r, gold = g, bold = b;
    set_lut_val_c(a, r);
    set_lut_val_c(a, b);
    g = b + ((r - b) * (gold - bold) / (rold - bold));
}
__device__ __forceinline__ float lim_f(float v, float lo, float hi) { const float m = hi < v ? hi : v; return lo < m ? m : lo; }      // rt_math.h LIM
// WeightedStdToneCurve::Triangle / Apply, curves.h L499-562
__device__ __forceinline__ float triangle(float a, float a1, float b, float whitept)
{
    if (a != b) {
        const float a2 = a1 - a;
        return b < a ? b + a2 * b / a : b + a2 * (whitept - b) / (whitept - a);
    }
    return a1;
}
__device__ __forceinline__ void weighted_std_tone(const ChainArgs& A, float& ir, float& ig, float& ib)
{
    const float w = A.Lmax;
    const float r = lim_f(ir, 0.f, w), g = lim_f(ig, 0.f, w), b = lim_f(ib, 0.f, w);
    float r1 = r; set_lut_val_c(A, r1);
    const float g1 = triangle(r, r1, g, w), b1 = triangle(r, r1, b, w);
    float g2 = g; set_lut_val_c(A, g2);
    const float r2 = triangle(g, g2, r, w), b2 = triangle(g, g2, b, w);
    float b3 = b; set_lut_val_c(A, b3);
    const float r3 = triangle(b, b3, r, w), g3 = triangle(b, b3, g, w);
    ir = lim_f(r1 * 0.50f + r2 * 0.25f + r3 * 0.25f, 0.f, w);
    ig = lim_f(g1 * 0.25f + g2 * 0.50f + g3 * 0.25f, 0.f, w);
    ib = lim_f(b1 * 0.25f + b2 * 0.25f + b3 * 0.50f, 0.f, w);
}
// SatAndValueBlendingToneCurve::Apply, curves.h L634-668, with Color::rgb2hsvtc / hsv2rgbdcp (color.h L423-506); reads the LUT directly
__device__ __forceinline__ void sat_value_tone(const ChainArgs& A, float& ir, float& ig, float& ib)
{
    float r = lim_f(ir, 0.f, 65535.f), g = lim_f(ig, 0.f, 65535.f), b = lim_f(ib, 0.f, 65535.f);
    const float lum = (r + g + b) / 3.f;
    const float newLum = lut_s(A.tc_lut, 65536, CLIP_BELOW | CLIP_ABOVE, lum);
    if (newLum == lum) return;
    float h, s;
    const float mn0 = g < r ? g : r, var_Min = b < mn0 ? b : mn0;
    const float mx0 = r < g ? g : r, var_Max = mx0 < b ? b : mx0;
    const float del_Max = var_Max - var_Min;
    const float v = var_Max / 65535.f;
    if (del_Max < 0.00001f) { h = 0.f; s = 0.f; }
    else {
        s = del_Max / var_Max;
        if (r == var_Max) h = (g < b ? 6.f : 0.f) + (g - b) / del_Max;
        else if (g == var_Max) h = 2.f + (b - r) / del_Max;
        else h = 4.f + (r - g) / del_Max;
    }
    float dV;
    if (newLum > lum) { const float coef = (newLum - lum) / (65535.f - lum); dV = (1.f - v) * coef; s *= 1.f - coef; }
    else { const float coef = (newLum - lum) / lum; dV = v * coef; }
    float vv = v + dV;
    const int sector = (int)h;
    const float f = h - sector;
    vv *= 65535.f;
    const float vs = vv * s, p = vv - vs, q = vv - f * vs, t = p + vv - q;
    switch (sector) {
    case 1: r = q; g = vv; b = p; break;
    case 2: r = p; g = vv; b = t; break;
    case 3: r = p; g = q; b = vv; break;
    case 4: r = t; g = p; b = vv; break;
    case 5: r = vv; g = p; b = q; break;
    default: r = vv; g = t; b = p;
    }
    ir = r; ig = g; ib = b;
}
// LuminanceToneCurve::Apply, curves.h L474-496; the luminance row of the float TMatrix (apply_tc, iptonecurve.cc L73-84)
__device__ __forceinline__ void luminance_tone(const ChainArgs& A, float& ir, float& ig, float& ib)
{
    const float w = A.Lmax;
    const float r = lim_f(ir, 0.f, w), g = lim_f(ig, 0.f, w), b = lim_f(ib, 0.f, w);
    float currLuminance = r * A.wy0 + g * A.wy1 + b * A.wy2;
    float newLuminance = currLuminance;
    set_lut_val_c(A, newLuminance);
    currLuminance = currLuminance == 0.f ? 0.00001f : currLuminance;
    const float coef = newLuminance / currLuminance;
    ir = lim_f(r * coef, 0.f, w); ig = lim_f(g * coef, 0.f, w); ib = lim_f(b * coef, 0.f, w);
}
__device__ __forceinline__ float lim01(float a) { const float m = 1.f < a ? 1.f : a; return 0.f < m ? m : 0.f; }
__device__ __forceinline__ float gauss_hue(float x, float b, float c) { return sleef::xexpf_scalar(-((x - b) * (x - b)) / (2 * (c * c))); }

__device__ __forceinline__ void neutral_tone(const ChainArgs& a, float& R, float& G, float& B)
{
    const float th[3] = {0.85f, 0.75f, 0.95f};
    const float dl[3] = {1.1f, 1.2f, 1.5f};
    const float PI_F_180 = (float)(3.14159265358979323846 / 180.0);
    const float whitept = 65535.f * a.whitecoeff;
    float rgb[3], jl, jc, jh;
    rgb[0] = R / 65535.f; rgb[0] = rgb[0] < 0.f ? 0.f : rgb[0];
    rgb[1] = G / 65535.f; rgb[1] = rgb[1] < 0.f ? 0.f : rgb[1];
    rgb[2] = B / 65535.f; rgb[2] = rgb[2] < 0.f ? 0.f : rgb[2];
    rgb2jzczhz(a.pq, a.ws, rgb[0], rgb[1], rgb[2], jl, jc, jh);
    const float ilum = jl;
    float hue = jh;
    const float iY = (rgb[0] + rgb[1] + rgb[2]) / 3.f;
    {
        float x = 0.f, y = 0.f, z = 0.f;
        x += a.to_out[0] * rgb[0]; x += a.to_out[1] * rgb[1]; x += a.to_out[2] * rgb[2];
        y += a.to_out[3] * rgb[0]; y += a.to_out[4] * rgb[1]; y += a.to_out[5] * rgb[2];
        z += a.to_out[6] * rgb[0]; z += a.to_out[7] * rgb[1]; z += a.to_out[8] * rgb[2];
        rgb[0] = x; rgb[1] = y; rgb[2] = z;
    }
    float ac = rgb[0] < rgb[1] ? rgb[1] : rgb[0];
    ac = ac < rgb[2] ? rgb[2] : ac;
    const float aac = fabsf(ac);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = ac != 0.f ? (ac - rgb[k]) / aac : 0.f;
        const float s = (1.f - th[k]) / sqrtf(dl[k] - 1.f);
        const float cd = d < th[k] ? d : s * sqrtf(d - th[k] + (s * s) / 4.0f) - s * sqrtf((s * s) / 4.0f) + th[k];
        rgb[k] = ac - cd * aac;
    }
    {
        float x = 0.f, y = 0.f, z = 0.f;
        x += a.to_work[0] * rgb[0]; x += a.to_work[1] * rgb[1]; x += a.to_work[2] * rgb[2];
        y += a.to_work[3] * rgb[0]; y += a.to_work[4] * rgb[1]; y += a.to_work[5] * rgb[2];
        z += a.to_work[6] * rgb[0]; z += a.to_work[7] * rgb[1]; z += a.to_work[8] * rgb[2];
        rgb[0] = x; rgb[1] = y; rgb[2] = z;
    }
    const float oY = (rgb[0] + rgb[1] + rgb[2]) / 3.f;
    if (oY > 0.f) {
        const float f = iY / oY;
        rgb[0] *= f; rgb[1] *= f; rgb[2] *= f;
        filmlike_clip(rgb[0], rgb[1], rgb[2], whitept);       // Lmax = ToneCurve::whitept = 65535 * whitecoeff on [0, 1] data, as the reference has it
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float nt = rgb[j] * 65535.f;
        set_lut_val_c(a, nt);
        rgb[j] = nt / 65535.f;
    }
    rgb2jzczhz(a.pq, a.ws, rgb[0], rgb[1], rgb[2], jl, jc, jh);
    const float rhue = a.hues[0], bhue = a.hues[1], yhue = a.hues[2], rrange = a.hues[3], brange = a.hues[4], yrange = a.hues[5];
    float hue_shift = 15.f * PI_F_180 * gauss_hue(hue, rhue, rrange);
    hue_shift += -5.f * PI_F_180 * gauss_hue(hue, bhue, brange);
    hue_shift *= lim01((rgb[0] + rgb[1] + rgb[2]) / (3.f * a.whitecoeff));
    hue += hue_shift;
    float sat = jc;
    {
        const float olum = jl;
        float ccf = ilum > 1e-5f ? (1.f - (lim01((olum / ilum) - 1.f) * 0.2f)) : 1.f;
        ccf = lim01(ccf + 0.5f * gauss_hue(hue, yhue, yrange));
        sat *= ccf;
    }
    jzczhz2rgb(a.pq_inv, a.iws, jl, sat, hue, rgb[0], rgb[1], rgb[2]);
    float v;
    v = rgb[0] * 65535.f; v = whitept < v ? whitept : v; R = 0.f < v ? v : 0.f;
    v = rgb[1] * 65535.f; v = whitept < v ? whitept : v; G = 0.f < v ? v : 0.f;
    v = rgb[2] * 65535.f; v = whitept < v ? whitept : v; B = 0.f < v ? v : 0.f;
}
__device__ __forceinline__ void sat_curve(const ChainArgs& a, float& R, float& G, float& B)
{
    float X, Y, Z, Jz, az, bz;
    mat3(a.ws, R / 65535.f, G / 65535.f, B / 65535.f, X, Y, Z);
    xyz2jzazbz(a.pq, X, Y, Z, Jz, az, bz);
    float cz = sqrtf(bz * bz + az * az);
    const float hz = sleef::xatan2f(bz, az);
    cz *= lut_s(a.satlut, 65536, CLIP_BELOW, Y * 65535.f);
    float r, g, b;
    jzczhz2rgb(a.pq_inv, a.iws, Jz, cz, hz, r, g, b);
    R = r * 65535.f; G = g * 65535.f; B = b * 65535.f;
}
// ApplyState's hue constants (curves.cc L877-887), computed once with the same device arithmetic
__global__ void k_tone_hues(const float* __restrict__ pq, float* __restrict__ out)
{
    const float rec2020[9] = {0.6734241f, 0.1656411f, 0.1251286f, 0.2790177f, 0.6753402f, 0.0456377f, -0.0019300f, 0.0299784f, 0.7973330f};
    float j, c, rh, bh, yh, oh;
    rgb2jzczhz(pq, rec2020, 1.f, 0.f, 0.f, j, c, rh);
    rgb2jzczhz(pq, rec2020, 0.f, 0.f, 1.f, j, c, bh);
    rgb2jzczhz(pq, rec2020, 1.f, 1.f, 0.f, j, c, yh);
    rgb2jzczhz(pq, rec2020, 1.f, 0.5f, 0.f, j, c, oh);
    out[0] = rh; out[1] = bh; out[2] = yh;
    out[3] = fabsf(oh - rh); out[4] = out[3]; out[5] = fabsf(oh - yh) * 0.8f;
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
    if (a.tc_mode == 2) {
        neutral_tone(a, r, g, b);
    } else if (a.tc_mode >= 0) {
        filmlike_clip(r, g, b, a.Lmax);
        if (a.tc_mode == 0) { set_lut_val_c(a, r); set_lut_val_c(a, g); set_lut_val_c(a, b); }
        else if (a.tc_mode == 3) weighted_std_tone(a, r, g, b);
        else if (a.tc_mode == 4) sat_value_tone(a, r, g, b);
        else if (a.tc_mode == 5) luminance_tone(a, r, g, b);
        else {
            r = maxr(0.f, minr(r, a.Lmax)); g = maxr(0.f, minr(g, a.Lmax)); b = maxr(0.f, minr(b, a.Lmax));
            if (r >= g) {
                if (g > b) rgb_tone(a, r, g, b);
                else if (b > r) rgb_tone(a, b, r, g);
                else if (b > g) rgb_tone(a, r, b, g);
                else { set_lut_val_c(a, r); set_lut_val_c(a, g); b = g; }
            } else {
                if (r >= b) rgb_tone(a, g, r, b);
                else if (b > g) rgb_tone(a, b, g, r);
                else rgb_tone(a, g, b, r);
            }
        }
    }
    if (a.satlut) sat_curve(a, r, g, b);
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

// ImProcFunctions::softLight's apply lambda (ipsoftlight.cc L57-66): LUTf f(65536) clips below and above; scalar loop in the reference
__device__ __forceinline__ float soft_light(const float* __restrict__ f, float x) { return x <= 65535.f ? lut_s(f, 65536, CLIP_BELOW | CLIP_ABOVE, x) : x; }

// (int)f as x86's cvttss2si gives it: NaN and out-of-range values become INT_MIN (CUDA's conversion saturates and maps NaN to 0)
__device__ __forceinline__ int cvtt_x86(float v) { return (v >= -2147483648.f && v < 2147483648.f) ? (int)v : (int)0x80000000; }

// HIST: labAdjustments' hist16 (iplabadjustments.cc L307-334: hist16thr[(int)L]++ over the Lab L plane, LUT<T>::operator[](int) clamps the
// index, LUT.h L308-311) of the frame as it stands after the stages before the Lab stage; nothing is stored
template <bool HIST>
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
                    if (HIST) atomicAdd(A.hist + min(max(cvtt_x86(L), 0), 65535), 1u);
                    else lab_tail<true>(A, L, a, bb, r[k], g[k], b[k]);
                }
            }
            if (HIST) continue;
            if (A.sl) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { r[k] = soft_light(A.sl, r[k]); g[k] = soft_light(A.sl, g[k]); b[k] = soft_light(A.sl, b[k]); }
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
                    if (HIST) { atomicAdd(A.hist + min(max(cvtt_x86(L), 0), 65535), 1u); continue; }
                    lab_tail<false>(A, L, 500.0f * (fx - fy), 200.0f * (fy - fz), r, g, b);
                }
                if (A.sl) { r = soft_light(A.sl, r); g = soft_light(A.sl, g); b = soft_light(A.sl, b); }
                A.r[row + x] = r; A.g[row + x] = g; A.b[row + x] = b;
            }
        }
    }
}

}  // namespace

// LUT slots in the context's device / pinned staging: 0 tone curve, 1-3 rgb curves, 4 L curve, 5-6 a / b curves, 7-8 cachef / cachefy,
// 9-10 jzazbz_pq_ / jzazbz_pq_inv_, 11 the saturation curve's table, 12 ApplyState's hue constants (device-computed), 13 softLight's table
constexpr size_t LUT_SLOT = 65536 + 64;
constexpr int N_SLOTS = 14;
constexpr int STAGE_SLOTS = 4;     // pinned staging only: the curve stages above the LUT when they fit (1 MB), else a pageable copy

static int chain_run(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const art_hp_chain_params* p, unsigned* d_hist);

int art_chain_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const art_hp_chain_params* p)
{
    return chain_run(ctx, W, H, r, g, b, pitch, p, nullptr);
}

// labAdjustments' hist16 of the frame as the Lab stage would see it (the stages before it applied on the fly, the planes untouched)
int art_chain_lab_hist_dev(art_hp_ctx* ctx, int W, int H, const float* r, const float* g, const float* b, size_t pitch, const art_hp_chain_params* p, unsigned* d_hist)
{
    if (!p->ws) return ctx->fail(ART_HP_ERR_INVALID, "the Lab histogram needs ws");
    ART_CUDA(ctx, cudaMemsetAsync(d_hist, 0, 65536 * sizeof(unsigned), ctx->stream));
    return chain_run(ctx, W, H, const_cast<float*>(r), const_cast<float*>(g), const_cast<float*>(b), pitch, p, d_hist);
}

static int chain_run(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const art_hp_chain_params* p, unsigned* d_hist)
{
    // every check comes before the first upload: a rejected call leaves the staging buffer and its event untouched
    if ((pitch & 3) || (reinterpret_cast<uintptr_t>(r) & 15) || (reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
        return ctx->fail(ART_HP_ERR_INVALID, "planes must be 16-byte aligned with a pitch that is a multiple of 4 floats");
    const int do_sat = p->saturation_enabled && (p->saturation || p->vibrance);
    const int tc_mode = p->tonecurve_lut ? p->tonecurve_mode : -1;
    const bool jz = tc_mode == 2 || p->satcurve_lut;
    if ((do_sat || p->lab_enabled || jz) && !p->ws) return ctx->fail(ART_HP_ERR_INVALID, "ws is required by the saturation, NEUTRAL tone curve and Lab stages");
    if (jz && !p->iws) return ctx->fail(ART_HP_ERR_INVALID, "the NEUTRAL tone curve and the saturation curve need iws");
    if (!d_hist && p->lab_enabled && (!p->iws || !p->lab_lcurve || !p->lab_acurve || !p->lab_bcurve)) return ctx->fail(ART_HP_ERR_INVALID, "Lab stage needs iws and the three curves");
    if (tc_mode > 5) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "tone curve mode %d (PERCEPTUAL is not built)", tc_mode);
    if (tc_mode == 5 && !p->ws) return ctx->fail(ART_HP_ERR_INVALID, "the LUMINANCE tone curve needs ws");
    const float whitecoeff = p->tonecurve_whitept > 0.f ? p->tonecurve_whitept : 1.f;
    const int nstages = (tc_mode >= 0 && p->tonecurve_stages) ? p->tonecurve_nstages : 0;
    if (nstages < 0 || nstages > 4) return ctx->fail(ART_HP_ERR_INVALID, "tonecurve_nstages %d (0..4)", nstages);
    for (int i = 0; i < nstages; ++i) {
        const art_hp_curve_stage& st = p->tonecurve_stages[i];
        if (st.kind < 0 || st.kind > 2) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "curve stage kind %d", st.kind);
        if (st.kind == 1 && (st.n < 1 || !st.poly_x || !st.poly_y)) return ctx->fail(ART_HP_ERR_INVALID, "curve stage %d: empty polyline", i);
        if (st.kind == 2 && !(st.w > 0.0 && st.b > 0.0 && st.b != 1.0)) return ctx->fail(ART_HP_ERR_INVALID, "curve stage %d: contrast curve parameters", i);
    }
    cudaStream_t st = ctx->stream;
    int rc = art_reserve(ctx, ctx->d_chain, N_SLOTS * LUT_SLOT * sizeof(float));
    if (rc) return rc;
    float* d = (float*)ctx->d_chain.p;
    if (!ctx->h_chain) {
        ART_CUDA(ctx, cudaMallocHost(&ctx->h_chain, (N_SLOTS + STAGE_SLOTS) * LUT_SLOT * sizeof(float)));
        ART_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
    } else {
        ART_CUDA(ctx, cudaEventSynchronize(ctx->ev_chain));        // the previous call's uploads left the staging buffer
    }
    float* h = (float*)ctx->h_chain;
    cudaError_t up_err = cudaSuccess;
    if (!ctx->chain_cache_ready) {      // Color::cachef / cachefy, color.cc L205-233 (host libm cbrt, as in the reference); jzazbz_pq_ / _inv_, L322-326 (host powf)
        float* cf = h + 7 * LUT_SLOT; float* cfy = h + 8 * LUT_SLOT; float* pq = h + 9 * LUT_SLOT; float* pqi = h + 10 * LUT_SLOT;
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
        const int epsmaxint = (int)(MAXVALF * eps);
        int i = 0;
        for (; i <= epsmaxint; i++) { cf[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); cfy[i] = (float)(327.68 * (kappa * i / MAXVALF)); }
        for (; i < 65536; i++) { cf[i] = (float)(327.68 * std::cbrt((double)i / MAXVALF)); cfy[i] = (float)(327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0)); }
        cf[65536] = cf[65535]; cfy[65536] = cfy[65535];
        for (i = 0; i < 65536; ++i) {
            float X = (float)i / 65535.f;
            X = std::max(X, 1e-10f);
            const float XX = std::pow(X * 1e-4f, 0.1593017578125f);
            pq[i] = std::pow((0.8359375f + 18.8515625f * XX) / (1 + 18.6875f * XX), 134.034375f);
            const float XI = std::pow(X, 7.460772656268214e-03f);
            pqi[i] = 1e4f * std::pow((0.8359375f - XI) / (18.6875f * XI - 18.8515625f), 6.277394636015326f);
        }
        pq[65536] = pq[65535]; pqi[65536] = pqi[65535];
        up_err = cudaMemcpyAsync(d + 7 * LUT_SLOT, cf, 4 * LUT_SLOT * sizeof(float), cudaMemcpyHostToDevice, st);
        if (up_err == cudaSuccess) {
            k_tone_hues<<<1, 1, 0, st>>>(d + 9 * LUT_SLOT, d + 12 * LUT_SLOT);
            ctx->launches++;
            ctx->chain_cache_ready = true;
        }
    }
    auto up = [&](int slot, const float* src, int n) -> const float* {
        if (!src) return nullptr;
        memcpy(h + slot * LUT_SLOT, src, sizeof(float) * n);
        h[slot * LUT_SLOT + n] = src[n - 1];       // LUT<T> allocates s + 3 entries; data[size] is touched (times a zero weight) at the top index
        const cudaError_t e = cudaMemcpyAsync(d + slot * LUT_SLOT, h + slot * LUT_SLOT, sizeof(float) * (n + 1), cudaMemcpyHostToDevice, st);
        if (up_err == cudaSuccess) up_err = e;
        return d + slot * LUT_SLOT;
    };
    ChainArgs a{};
    a.r = r; a.g = g; a.b = b; a.pitch = pitch; a.W = W; a.H = H;
    a.do_exp = p->exposure_enabled; a.exp_scale = p->exp_scale; a.black = p->black;
    a.do_sat = do_sat;
    a.vib_on = p->vibrance != 0;
    a.saturation = 1.f + p->saturation / 100.f;
    a.vibrance = 1.f - p->vibrance / 1000.f;
    a.tc_mode = tc_mode;
    a.tc_lut = up(0, p->tonecurve_lut, 65536);
    a.whitecoeff = whitecoeff;
    a.Lmax = 65535.f * whitecoeff;
    a.rc = up(1, p->rcurve, 65536); a.gc = up(2, p->gcurve, 65536); a.bc = up(3, p->bcurve, 65536);
    a.do_lab = p->lab_enabled || d_hist;
    a.hist = d_hist;
    if (a.do_lab) {
        if (!d_hist) { a.lc = up(4, p->lab_lcurve, 32770); a.ac = up(5, p->lab_acurve, 65536); a.bcl = up(6, p->lab_bcurve, 65536); }
        a.chroma = p->lab_chroma;
        a.cachef = d + 7 * LUT_SLOT; a.cachefy = d + 8 * LUT_SLOT;
    }
    a.pq = d + 9 * LUT_SLOT; a.pq_inv = d + 10 * LUT_SLOT; a.hues = d + 12 * LUT_SLOT;
    a.satlut = up(11, p->satcurve_lut, 65536);
    a.sl = d_hist ? nullptr : up(13, p->softlight_lut, 65536);
    static const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) { a.to_out[i] = p->neutral_to_out ? p->neutral_to_out[i] : ident[i]; a.to_work[i] = p->neutral_to_work ? p->neutral_to_work[i] : ident[i]; }
    if (p->ws) { a.wy0 = (float)p->ws[3]; a.wy1 = (float)p->ws[4]; a.wy2 = (float)p->ws[5]; for (int i = 0; i < 9; ++i) a.ws[i] = (float)p->ws[i]; }
    if (p->iws) for (int i = 0; i < 9; ++i) a.iws[i] = (float)p->iws[i];
    ART_CUDA(ctx, cudaEventRecord(ctx->ev_chain, st));        // recorded on every path that queued an upload
    if (up_err != cudaSuccess) return ctx->fail(ART_HP_ERR_CUDA, "curve upload failed: %s", cudaGetErrorString(up_err));
    if (nstages) {
        // the stages above the LUT: descriptors + polylines in one device block.  Only samples above 65535 read them; when a
        // polyline ends at or below 1 (white point 1) every such sample lies beyond it and the last point alone decides
        size_t doubles = 0;
        int keep[4];
        for (int i = 0; i < nstages; ++i) {
            const art_hp_curve_stage& s = p->tonecurve_stages[i];
            keep[i] = s.kind != 1 ? 0 : (i == 0 && s.poly_x[s.n - 1] <= 1.0) ? 1 : s.n;
            doubles += 2 * (size_t)keep[i];
        }
        const size_t hdr = round_up(4 * sizeof(art_hp_curve_stage), 256);
        if ((rc = art_reserve(ctx, ctx->d_chain_stages, hdr + doubles * sizeof(double)))) return rc;
        const size_t total = hdr + doubles * sizeof(double);
        char* hbase;
        if (total <= STAGE_SLOTS * LUT_SLOT * sizeof(float)) hbase = reinterpret_cast<char*>(h + N_SLOTS * LUT_SLOT);     // pinned, guarded by ev_chain (re-recorded below)
        else { ctx->h_chain_stages.assign(total, 0); hbase = ctx->h_chain_stages.data(); }       // pageable: cudaMemcpyAsync stages it before returning
        memset(hbase, 0, hdr);
        art_hp_curve_stage* hd = reinterpret_cast<art_hp_curve_stage*>(hbase);
        double* hp = reinterpret_cast<double*>(hbase + hdr);
        double* dp = reinterpret_cast<double*>((char*)ctx->d_chain_stages.p + hdr);
        size_t off = 0;
        for (int i = 0; i < nstages; ++i) {
            hd[i] = p->tonecurve_stages[i];
            if (hd[i].kind == 1) {
                const int n = keep[i], skip = hd[i].n - n;
                memcpy(hp + off, hd[i].poly_x + skip, n * sizeof(double)); hd[i].poly_x = dp + off; off += n;
                memcpy(hp + off, p->tonecurve_stages[i].poly_y + skip, n * sizeof(double)); hd[i].poly_y = dp + off; off += n;
                hd[i].n = n;
            }
        }
        ART_CUDA(ctx, cudaMemcpyAsync(ctx->d_chain_stages.p, hbase, total, cudaMemcpyHostToDevice, st));
        ART_CUDA(ctx, cudaEventRecord(ctx->ev_chain, st));
        a.stages = reinterpret_cast<const art_hp_curve_stage*>(ctx->d_chain_stages.p);
        a.nstages = nstages;
    }
    const dim3 blk(64, 1), grid(((W + 3) / 4 + 63) / 64, std::min(H, 148 * 16));
    art_prof_begin(ctx, d_hist ? "k_chain_hist" : "k_chain");
    if (d_hist) k_chain<true><<<grid, blk, 0, st>>>(a); else k_chain<false><<<grid, blk, 0, st>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
