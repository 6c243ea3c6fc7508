/*
 * oracle/chain_port.c -- plain-C restatement of the per-pixel colour / curve chain of ImProcFunctions::process.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Follows reference rtengine/: ipexposure.cc expcomp L29-73; ipsaturation.cc apply_vibrance L29-38, saturationVibrance
 * L44-83; iptonecurve.cc filmlike_clip L214-231, apply L34-45; curves.h setLutVal L223-230, StandardToneCurve::Apply
 * L360-368, AdobeToneCurve::Apply / RGBTone L425-472; iprgbcurves.cc rgbCurves pixel loop L113-146; iplabadjustments.cc
 * lab_adjustments pixel loop L252-283; imagefloat.cc rgb_to_lab L841-878, lab_to_rgb L949-972; color.cc rgbxyz / xyz2rgb
 * L833-894, Lab2XYZ L1203-1245, computeXYZ2Lab(Y) L1247-1275, XYZ2Lab L1382-1437, filmlike_clip L6650-6688; color.h f2xyz
 * L767-781, rgbLuminance L203-207; LUT.h operator[](float) L437-459 and operator[](vfloat) L349-377.
 *
 * The reference processes every row in groups of four pixels with SSE2 and the remaining W % 4 pixels with scalar code;
 * the two paths differ (LUT interpolation a*hi + (1-a)*lo versus lo + (hi-lo)*a, Lab2XYZ's Y branch, a whole group
 * taking the scalar Lab route when one of its four pixels is out of range).  A pixel x is in a group iff
 * (x & ~3) + 4 <= W.  Pinned bit-exact against the reference code compiled in place (tests/test_oracle_chain.py).
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

static inline float maxr(float a, float b) { return a < b ? b : a; }        /* rt_math.h max */
static inline float minr(float a, float b) { return b < a ? b : a; }        /* rt_math.h min */
static inline float vmaxf_(float a, float b) { return a > b ? a : b; }      /* _mm_max_ps(a, b) */
static inline float vminf_(float a, float b) { return a < b ? a : b; }      /* _mm_min_ps(a, b) */
static inline float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }
static inline int in_group(int x, int W) { return (x & ~3) + 4 <= W; }

/* ---- LUT.h ---- */
#define CLIP_BELOW 1
#define CLIP_ABOVE 2
static inline float lut_s(const float* data, int size, int clip, float index)
{
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
static inline float lut_v(const float* data, int size, float index)
{
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}

/* ---- expcomp ---- */
int artoracle_chain_expcomp(float* R, float* G, float* B, int W, int H, float exp_scale, float black)
{
    float* ch[3] = {R, G, B};
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 3; ++c) {
                float* v = ch[c] + (size_t)y * W + x;
                const float t = *v * exp_scale - black;
                *v = in_group(x, W) ? vmaxf_(t, 0.f) : maxr(t, 0.f);
            }
    return 0;
}

/* ---- channelMixer (ipchmixer.cc L152-232), the per-pixel loop: m = RR RG RB / GR GG GB / BR BG BB (RGB_MATRIX: the per-mille sliders / 1000.f;
 * PRIMARIES_CHROMA: get_mixer_matrix's result).  SSE2 groups clamp with _mm_max_ps (NaN -> 0), the row tail with rt_math.h max (NaN stays). ---- */
int artoracle_chmixer(float* R, float* G, float* B, int W, int H, const float* m)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t o = (size_t)y * W + x;
            const float r = R[o], g = G[o], b = B[o];
            const float rmix = (r * m[0] + g * m[1] + b * m[2]);
            const float gmix = (r * m[3] + g * m[4] + b * m[5]);
            const float bmix = (r * m[6] + g * m[7] + b * m[8]);
            if (in_group(x, W)) { R[o] = vmaxf_(rmix, 0.f); G[o] = vmaxf_(gmix, 0.f); B[o] = vmaxf_(bmix, 0.f); }
            else { R[o] = maxr(rmix, 0.f); G[o] = maxr(gmix, 0.f); B[o] = maxr(bmix, 0.f); }
        }
    return 0;
}

/* ---- saturationVibrance ---- */
static inline float lum_d(float r, float g, float b, const double* ws)
{   /* Color::rgbLuminance(r, g, b, TMatrix), color.h L203-207; TMatrix is const float (*)[3] (iccstore.h L38): float arithmetic */
    const float w0 = (float)ws[3], w1 = (float)ws[4], w2 = (float)ws[5];
    return r * w0 + g * w1 + b * w2;
}
static inline float apply_vibrance(float x, float vib, float noise)
{
    const float ax = fabsf(x / 65535.f);
    if (ax > noise) {
        const float sgn = (float)((0.f < x) - (x < 0.f));
        return sgn * pow_F_scalar(ax, vib) * 65535.f;
    }
    return x;
}
int artoracle_chain_saturation(float* R, float* G, float* B, int W, int H, int sat, int vibr, const double* ws)
{
    const float saturation = 1.f + sat / 100.f;
    const float vibrance = 1.f - vibr / 1000.f;
    const float noise = pow_F_scalar(2.f, -16.f);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const float r = R[i], g = G[i], b = B[i];
        const float l = lum_d(r, g, b, ws);
        float rl = r - l, gl = g - l, bl = b - l;
        if (vibr) { rl = apply_vibrance(rl, vibrance, noise); gl = apply_vibrance(gl, vibrance, noise); bl = apply_vibrance(bl, vibrance, noise); }
        R[i] = maxr(l + saturation * rl, noise);
        G[i] = maxr(l + saturation * gl, noise);
        B[i] = maxr(l + saturation * bl, noise);
    }
    return 0;
}

/* ---- tone curve: filmlike_clip, then StandardToneCurve (mode 0) or AdobeToneCurve (mode 1) ---- */
static inline void clip_tone(float* r, float* g, float* b, float L)
{
    const float r_ = *r > L ? L : *r;
    const float b_ = *b > L ? L : *b;
    const float g_ = b_ + ((r_ - b_) * (*g - *b) / (*r - *b));
    *r = r_; *g = g_; *b = b_;
}
static void filmlike_clip(float* r, float* g, float* b, float L)
{
    if (*r >= *g) {
        if (*g > *b) clip_tone(r, g, b, L);
        else if (*b > *r) clip_tone(b, r, g, L);
        else if (*b > *g) clip_tone(r, b, g, L);
        else { *r = *r > L ? L : *r; *g = *g > L ? L : *g; *b = *g; }
    } else {
        if (*r >= *b) clip_tone(g, r, b, L);
        else if (*b > *g) clip_tone(b, g, r, L);
        else clip_tone(g, b, r, L);
    }
}
static inline void set_lut_val(const float* lut, float* val)
{   /* curves.h L223-230 with `curve == nullptr` (the LUT serves every sample, clipped above); the Curve::getVal branch is tone_port.c's */
    *val = lut_s(lut, 65536, CLIP_BELOW | CLIP_ABOVE, maxr(*val, 0.f));
}
static inline void rgb_tone(const float* lut, float* r, float* g, float* b)
{
    const float rold = *r, gold = *g, bold = *b;
    set_lut_val(lut, r);
    set_lut_val(lut, b);
    *g = *b + ((*r - *b) * (gold - bold) / (rold - bold));
}
int artoracle_chain_tonecurve(float* R, float* G, float* B, int W, int H, int mode, const float* lut, float whitept)
{
    const float Lmax = 65535.f * whitept;
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        filmlike_clip(&R[i], &G[i], &B[i], Lmax);
        if (mode == 0) {
            set_lut_val(lut, &R[i]); set_lut_val(lut, &G[i]); set_lut_val(lut, &B[i]);
        } else {
            float r = maxr(0.f, minr(R[i], Lmax)), g = maxr(0.f, minr(G[i], Lmax)), b = maxr(0.f, minr(B[i], Lmax));
            if (r >= g) {
                if (g > b) rgb_tone(lut, &r, &g, &b);
                else if (b > r) rgb_tone(lut, &b, &r, &g);
                else if (b > g) rgb_tone(lut, &r, &b, &g);
                else { set_lut_val(lut, &r); set_lut_val(lut, &g); b = g; }
            } else {
                if (r >= b) rgb_tone(lut, &g, &r, &b);
                else if (b > g) rgb_tone(lut, &b, &g, &r);
                else rgb_tone(lut, &g, &b, &r);
            }
            R[i] = r; G[i] = g; B[i] = b;
        }
    }
    return 0;
}

/* ---- the other per-pixel tone-curve classes apply_tc dispatches to (iptonecurve.cc L48-85), after the same filmlike_clip pass:
 * mode 3 = WeightedStdToneCurve::Apply (curves.h L499-562), 4 = SatAndValueBlendingToneCurve::Apply (L634-668, Color::rgb2hsvtc /
 * hsv2rgbdcp, color.h L423-506), 5 = LuminanceToneCurve::Apply (L474-496; ws = the float TMatrix) ---- */
static inline float lim_f(float v, float lo, float hi) { const float m = hi < v ? hi : v; return lo < m ? m : lo; }   /* LIM = max(low, min(val, high)) with std::min / max operand order */
static inline float triangle(float a, float a1, float b, float whitept)
{
    if (a != b) {
        const float a2 = a1 - a;
        return b < a ? b + a2 * b / a : b + a2 * (whitept - b) / (whitept - a);
    }
    return a1;
}
static inline float min3f(float a, float b, float c) { const float m = b < a ? b : a; return c < m ? c : m; }   /* rt_math.h min(a, b, c) = min(min(a, b), min(c)) */
static inline float max3f(float a, float b, float c) { const float m = a < b ? b : a; return m < c ? c : m; }
int artoracle_chain_tonecurve_ex(float* R, float* G, float* B, int W, int H, int mode, const float* lut, float whitecoeff, const double* wsd)
{
    const float whitept = 65535.f * whitecoeff;
    float wy[3] = {0.f, 0.f, 0.f};
    if (mode < 3 || mode > 5) return -1;
    if (mode == 5) { if (!wsd) return -1; for (int i = 0; i < 3; ++i) wy[i] = (float)wsd[3 + i]; }
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        filmlike_clip(&R[i], &G[i], &B[i], whitept);
        if (mode == 3) {
            float r = lim_f(R[i], 0.f, whitept), g = lim_f(G[i], 0.f, whitept), b = lim_f(B[i], 0.f, whitept);
            float r1 = r; set_lut_val(lut, &r1);
            const float g1 = triangle(r, r1, g, whitept), b1 = triangle(r, r1, b, whitept);
            float g2 = g; set_lut_val(lut, &g2);
            const float r2 = triangle(g, g2, r, whitept), b2 = triangle(g, g2, b, whitept);
            float b3 = b; set_lut_val(lut, &b3);
            const float r3 = triangle(b, b3, r, whitept), g3 = triangle(b, b3, g, whitept);
            R[i] = lim_f(r1 * 0.50f + r2 * 0.25f + r3 * 0.25f, 0.f, whitept);
            G[i] = lim_f(g1 * 0.25f + g2 * 0.50f + g3 * 0.25f, 0.f, whitept);
            B[i] = lim_f(b1 * 0.25f + b2 * 0.25f + b3 * 0.50f, 0.f, whitept);
        } else if (mode == 4) {
            float r = lim_f(R[i], 0.f, 65535.f), g = lim_f(G[i], 0.f, 65535.f), b = lim_f(B[i], 0.f, 65535.f);      /* CLIP */
            const float lum = (r + g + b) / 3.f;
            const float newLum = lut_s(lut, 65536, CLIP_BELOW | CLIP_ABOVE, lum);
            if (newLum == lum) continue;            /* the reference returns before writing back: the pixel keeps its unclipped values */
            float h, s, v;
            {
                const float var_Min = min3f(r, g, b), var_Max = max3f(r, g, b), del_Max = var_Max - var_Min;
                v = var_Max / 65535.f;
                if (del_Max < 0.00001f) { h = 0.f; s = 0.f; }
                else {
                    s = del_Max / var_Max;
                    if (r == var_Max) h = (g < b ? 6.f : 0.f) + (g - b) / del_Max;
                    else if (g == var_Max) h = 2.f + (b - r) / del_Max;
                    else h = 4.f + (r - g) / del_Max;
                }
            }
            float dV;
            if (newLum > lum) { const float coef = (newLum - lum) / (65535.f - lum); dV = (1.f - v) * coef; s *= 1.f - coef; }
            else { const float coef = (newLum - lum) / lum; dV = v * coef; }
            {
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
            }
            R[i] = r; G[i] = g; B[i] = b;
        } else {
            float r = lim_f(R[i], 0.f, whitept), g = lim_f(G[i], 0.f, whitept), b = lim_f(B[i], 0.f, whitept);
            float currLuminance = r * wy[0] + g * wy[1] + b * wy[2];
            float newLuminance = currLuminance;
            set_lut_val(lut, &newLuminance);
            currLuminance = currLuminance == 0.f ? 0.00001f : currLuminance;
            const float coef = newLuminance / currLuminance;
            R[i] = lim_f(r * coef, 0.f, whitept); G[i] = lim_f(g * coef, 0.f, whitept); B[i] = lim_f(b * coef, 0.f, whitept);
        }
    }
    return 0;
}

/* ---- rgbCurves: LUTs built with flags 0 (no clipping) ---- */
int artoracle_chain_rgbcurves(float* R, float* G, float* B, int W, int H, const float* rc, const float* gc, const float* bc)
{
    float* ch[3] = {R, G, B};
    const float* cv[3] = {rc, gc, bc};
    for (int c = 0; c < 3; ++c) {
        if (!cv[c]) continue;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float* v = ch[c] + (size_t)y * W + x;
                *v = in_group(x, W) ? lut_v(cv[c], 65536, *v) : lut_s(cv[c], 65536, 0, *v);
            }
    }
    return 0;
}

/* ---- softLight: the apply lambda of ImProcFunctions::softLight (ipsoftlight.cc L57-78) over its host-built table f (LUTf(65536): clips below and above) ---- */
int artoracle_chain_softlight(float* R, float* G, float* B, int W, int H, const float* f)
{
    float* ch[3] = {R, G, B};
    for (int c = 0; c < 3; ++c)
        for (size_t i = 0; i < (size_t)W * H; ++i) {
            const float x = ch[c][i];
            ch[c][i] = x <= 65535.f ? lut_s(f, 65536, CLIP_BELOW | CLIP_ABOVE, x) : x;
        }
    return 0;
}

/* ---- blackAndWhite: the two pixel loops of ImProcFunctions::blackAndWhite (ipbw.cc L283-312 mixer with optional gamma tables, L343-362 colour cast in
 * YUV) with Imagefloat::setMode(YUV) before the second and the next stage's setMode(RGB) after it (imagefloat.cc L700-725, L779-803).  Tables are
 * LUTf(65536): the SSE2 groups read them with the vector rule, the row tails with the scalar one ---- */
int artoracle_bw(float* R, float* G, float* B, int W, int H, const double* ws9, float bwr, float bwg, float bwb, float kcorec,
                 const float* gr, const float* gg, const float* gb, const float* ul, const float* vl)
{
    const float w0 = (float)ws9[3], w1 = (float)ws9[4], w2 = (float)ws9[5];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const int vec = in_group(x, W);
            float r = R[i], g = G[i], b = B[i];
            if (gr) {
                r = vec ? lut_v(gr, 65536, r) : lut_s(gr, 65536, CLIP_BELOW | CLIP_ABOVE, r);
                g = vec ? lut_v(gg, 65536, g) : lut_s(gg, 65536, CLIP_BELOW | CLIP_ABOVE, g);
                b = vec ? lut_v(gb, 65536, b) : lut_s(gb, 65536, CLIP_BELOW | CLIP_ABOVE, b);
            }
            const float bw = ((bwr * r + bwg * g + bwb * b) * kcorec);
            r = g = b = bw;
            if (ul) {
                const float Y = r * w0 + g * w1 + b * w2;
                float u = Y - b, v = r - Y;
                u += vec ? lut_v(ul, 65536, Y) : lut_s(ul, 65536, CLIP_BELOW | CLIP_ABOVE, Y);
                v += vec ? lut_v(vl, 65536, Y) : lut_s(vl, 65536, CLIP_BELOW | CLIP_ABOVE, Y);
                b = Y - u; r = v + Y;
                g = (Y - r * w0 - b * w2) / w1;
            }
            R[i] = r; G[i] = g; B[i] = b;
        }
    return 0;
}

/* ---- proPhotoBlue (improcfun.cc L312-357), the step ImProcFunctions::process runs after toneEqualizer when the working profile is ProPhoto (L585-587):
 * a pixel with r == 0 or g == 0 and no negative channel loses 1 % of its HSV saturation; Color::rgb2hsv computes in double (color.cc L586-622),
 * Color::hsv2rgb in float (L654-694).  The SSE2 group test only skips groups without such a pixel: the rule is per pixel ---- */
int artoracle_prophoto_blue(float* R, float* G, float* B, int W, int H)
{
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const float r = R[i], g = G[i], b = B[i];
        const float mn0 = g < r ? g : r, mn = b < mn0 ? b : mn0;       /* rtengine::min(r, g, b) */
        if (!((r == 0.0f || g == 0.0f) && mn >= 0.f)) continue;
        const double var_R = r / 65535.0, var_G = g / 65535.0, var_B = b / 65535.0;
        const double m0 = var_G < var_R ? var_G : var_R, var_Min = var_B < m0 ? var_B : m0;
        const double x0 = var_R < var_G ? var_G : var_R, var_Max = x0 < var_B ? var_B : x0;
        const double del_Max = var_Max - var_Min;
        float h = 0.f, s, v = (float)var_Max;
        if (del_Max < 0.00001 && del_Max > -0.00001) s = 0.f;
        else {
            s = (float)(del_Max / (var_Max == 0.0 ? 1.0 : var_Max));
            if (var_R == var_Max) h = (float)((var_G - var_B) / del_Max);
            else if (var_G == var_Max) h = (float)(2.0 + (var_B - var_R) / del_Max);
            else if (var_B == var_Max) h = (float)(4.0 + (var_R - var_G) / del_Max);
            h /= 6.f;
            if (h < 0.f) h += 1.f;
            if (h > 1.f) h -= 1.f;
        }
        s *= 0.99f;
        const float h1 = h * 6.f;
        const int k = (int)h1;
        const float f = h1 - k;
        const float p = v * (1.f - s), q = v * (1.f - s * f), t = v * (1.f - s * (1.f - f));
        float r1, g1, b1;
        if (k == 1) { r1 = q; g1 = v; b1 = p; }
        else if (k == 2) { r1 = p; g1 = v; b1 = t; }
        else if (k == 3) { r1 = p; g1 = q; b1 = v; }
        else if (k == 4) { r1 = t; g1 = p; b1 = v; }
        else if (k == 5) { r1 = v; g1 = p; b1 = q; }
        else { r1 = v; g1 = t; b1 = p; }
        R[i] = r1 * 65535.0f; G[i] = g1 * 65535.0f; B[i] = b1 * 65535.0f;
    }
    return 0;
}

/* ---- Lab ---- */
static float g_cachef[65536], g_cachefy[65536];
static int g_cache_ready = 0;
static void cache_init(void)
{   /* color.cc L205-233 */
    if (g_cache_ready) return;
    const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
    const int epsmaxint = (int)(MAXVALF * eps);
    int i = 0;
    for (; i <= epsmaxint; i++) { g_cachef[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); g_cachefy[i] = (float)(327.68 * (kappa * i / MAXVALF)); }
    for (; i < 65536; i++) { g_cachef[i] = (float)(327.68 * cbrt((double)i / MAXVALF)); g_cachefy[i] = (float)(327.68 * (116.0 * cbrt((double)i / MAXVALF) - 16.0)); }
    g_cache_ready = 1;
}
#define D50X 0.9642f
#define D50Z 0.8249f
static inline float xyz2lab_f(float f)
{
    const double kappa = 24389.0 / 27.0;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * ((kappa * f / 65535.f + 16.0) / 116.0));
    else if (f > 65535.f) return 327.68f * xcbrtf_scalar(f / 65535.f);
    return lut_s(g_cachef, 65536, CLIP_BELOW, f);
}
static inline float xyz2lab_fy(float f)
{
    const double kappa = 24389.0 / 27.0;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * (kappa * f / 65535.f));
    else if (f > 65535.f) return 327.68f * (116.f * xcbrtf_scalar(f / 65535.f) - 16.f);
    return lut_s(g_cachefy, 65536, CLIP_BELOW, f);
}
static inline void xyz2lab_scalar(float x, float y, float z, float* L, float* a, float* b)
{
    const float fx = xyz2lab_f(x), fy = xyz2lab_f(y), fz = xyz2lab_f(z);
    *L = xyz2lab_fy(y);
    *a = 500.0f * (fx - fy);
    *b = 200.0f * (fy - fz);
}
static inline float f2xyz_f(float f)
{
    const float epsilonExpInv3f = (float)(6.0 / 29.0), kappaInvf = (float)(27.0 / 24389.0);
    return (f > epsilonExpInv3f) ? f * f * f : (116.f * f - 16.f) * kappaInvf;
}

/* one row of Imagefloat::rgb_to_lab: r <- a, g <- L, b <- b */
static void row_rgb_to_lab(float* r, float* g, float* b, int W, const float ws[9])
{
    int x = 0;
    for (; x < W - 3; x += 4) {
        float X[4], Y[4], Z[4];
        int slow = 0;
        for (int k = 0; k < 4; ++k) {
            const float rv = r[x + k], gv = g[x + k], bv = b[x + k];
            X[k] = ws[0] * rv + ws[1] * gv + ws[2] * bv;
            Y[k] = ws[3] * rv + ws[4] * gv + ws[5] * bv;
            Z[k] = ws[6] * rv + ws[7] * gv + ws[8] * bv;
            X[k] = X[k] / D50X;
            Z[k] = Z[k] / D50Z;
            const float mx = vmaxf_(X[k], vmaxf_(Y[k], Z[k])), mn = vminf_(X[k], vminf_(Y[k], Z[k]));
            if (mx > 65535.f || mn < 0.f) slow = 1;
        }
        for (int k = 0; k < 4; ++k) {
            float L, a, bb;
            if (slow) xyz2lab_scalar(X[k], Y[k], Z[k], &L, &a, &bb);
            else {
                const float fx = lut_v(g_cachef, 65536, X[k]), fy = lut_v(g_cachef, 65536, Y[k]), fz = lut_v(g_cachef, 65536, Z[k]);
                L = lut_v(g_cachefy, 65536, Y[k]);
                a = 500.f * (fx - fy);
                bb = 200.f * (fy - fz);
            }
            g[x + k] = L; r[x + k] = a; b[x + k] = bb;
        }
    }
    for (; x < W; ++x) {
        const float rv = r[x], gv = g[x], bv = b[x];
        const float X = ws[0] * rv + ws[1] * gv + ws[2] * bv;
        const float Y = ws[3] * rv + ws[4] * gv + ws[5] * bv;
        const float Z = ws[6] * rv + ws[7] * gv + ws[8] * bv;
        float L, a, bb;
        xyz2lab_scalar(X / D50X, Y, Z / D50Z, &L, &a, &bb);
        g[x] = L; r[x] = a; b[x] = bb;
    }
}

static void row_lab_to_rgb(float* r, float* g, float* b, int W, const float iws[9])
{
    const float c1By116 = (float)(1.0 / 116.0), c16By116 = (float)(16.0 / 116.0);
    const double kappa = 24389.0 / 27.0;
    for (int x = 0; x < W; ++x) {
        const float L = g[x], a = r[x], bb = b[x];
        float X, Y, Z;
        if (in_group(x, W)) {       /* Lab2XYZ(vfloat...) */
            const float Lq = L / 327.68f, aq = a / 327.68f, bq = bb / 327.68f;
            const float fy = c1By116 * Lq + c16By116;
            const float fx = 0.002f * aq + fy;
            const float fz = fy - (0.005f * bq);
            X = 65535.f * f2xyz_f(fx) * D50X;
            Z = 65535.f * f2xyz_f(fz) * D50Z;
            const float res1 = fy * fy * fy;
            const float res2 = Lq / (float)kappa;
            Y = (Lq > (float)8.0) ? res1 : res2;
            Y *= 65535.f;
        } else {                    /* Lab2XYZ(float...) */
            const float LL = L / 327.68f, aa = a / 327.68f, b2 = bb / 327.68f;
            const float fy = (c1By116 * LL) + c16By116;
            const float fx = (0.002f * aa) + fy;
            const float fz = fy - (0.005f * b2);
            X = 65535.0f * f2xyz_f(fx) * D50X;
            Z = 65535.0f * f2xyz_f(fz) * D50Z;
            Y = ((double)LL > 8.0) ? 65535.0f * fy * fy * fy : (float)(65535.0f * LL / kappa);
        }
        r[x] = iws[0] * X + iws[1] * Y + iws[2] * Z;
        g[x] = iws[3] * X + iws[4] * Y + iws[5] * Z;
        b[x] = iws[6] * X + iws[7] * Y + iws[8] * Z;
    }
}

int artoracle_chain_rgb2lab(float* R, float* G, float* B, int W, int H, const double* wsd, const double* iwsd, int back)
{
    cache_init();
    float ws[9], iws[9];
    for (int i = 0; i < 9; ++i) { ws[i] = (float)wsd[i]; iws[i] = (float)(iwsd ? iwsd[i] : 0.0); }
    for (int y = 0; y < H; ++y) {
        if (back) row_lab_to_rgb(R + (size_t)y * W, G + (size_t)y * W, B + (size_t)y * W, W, iws);
        else row_rgb_to_lab(R + (size_t)y * W, G + (size_t)y * W, B + (size_t)y * W, W, ws);
    }
    return 0;
}

/* labAdjustments: setMode(LAB), lab_adjustments' loop, setMode(RGB) */
int artoracle_chain_lab(float* R, float* G, float* B, int W, int H, const float* lc, const float* ac, const float* bc, float chroma,
                        const double* wsd, const double* iwsd)
{
    cache_init();
    float ws[9], iws[9];
    for (int i = 0; i < 9; ++i) { ws[i] = (float)wsd[i]; iws[i] = (float)iwsd[i]; }
    for (int y = 0; y < H; ++y) {
        float *r = R + (size_t)y * W, *g = G + (size_t)y * W, *b = B + (size_t)y * W;
        row_rgb_to_lab(r, g, b, W, ws);
        for (int x = 0; x < W; ++x) {
            if (in_group(x, W)) {
                g[x] = lut_v(lc, 32770, g[x]);
                r[x] = (lut_v(ac, 65536, r[x] + 32768.f) - 32768.f) * chroma;
                b[x] = (lut_v(bc, 65536, b[x] + 32768.f) - 32768.f) * chroma;
            } else {
                g[x] = lut_s(lc, 32770, 0, g[x]);
                r[x] = (lut_s(ac, 65536, CLIP_BELOW | CLIP_ABOVE, r[x] + 32768.f) - 32768.f) * chroma;
                b[x] = (lut_s(bc, 65536, CLIP_BELOW | CLIP_ABOVE, b[x] + 32768.f) - 32768.f) * chroma;
            }
        }
        row_lab_to_rgb(r, g, b, W, iws);
    }
    return 0;
}
