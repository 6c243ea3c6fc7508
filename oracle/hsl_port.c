/*
 * oracle/hsl_port.c -- CPU restatement of the reference's HSL equalizer.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates ImProcFunctions::hslEqualizer (reference rtengine/iphsl.cc L29-221), STAGE_1 of ImProcFunctions::process
 * (improcfun.cc L580-584), together with the two mode changes around it: Imagefloat::setMode(YUV) on entry (rgb_to_yuv,
 * imagefloat.cc L700-725) and the setMode(RGB) the next stage performs on the YUV image it leaves (yuv_to_rgb, L779-803).
 *   Y, u, v scaled to [0, 1]; (u, v) -> (h, s) by Color::yuv2hsl (color.cc L6691-6695: sqrt, sleef xatan2f)
 *   per curve (S, L, H in this order): mask = FlatCurve::getVal(hue01(h)) (flatcurves.cc L339-365, double arithmetic over the
 *   host-built polyline), guidedFilter(Y, mask, mask, radius, eps) with automatic subsampling (guidedfilter.cc), then the
 *   per-pixel update through tolin() = sign * LIM01(xlog2lin(|2 (m - 0.5)|, base)) (sleef.h L1309-1313)
 *   (h, s) -> (u, v) by Color::hsl2yuv (xsincosf), scale back by 65535.
 * The curves come in as the polylines the FlatCurve constructor builds (poly_x, poly_y, dyByDx), n = 0 for an identity curve;
 * `coeff` is the fixed local FlatCurve of L119-123.
 * Pinned bit-exact against the reference's own function body compiled in place (oracle/_ref, shim_tone.cc) in
 * tests/test_oracle_hsl.py.  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

int artoracle_guided_filter(const float* guide, const float* src, float* dst, long stride, int W, int H, int r, float epsilon, int subsampling);
float artoracle_xatan2f(float y, float x);                  /* tone_port.c: sleef.h L1155-1188 */
void artoracle_xsincosf(float d, float* sn, float* cs);     /* tone_port.c: sleefsseavx.h L1051-1101 */

typedef struct { int n; const double *px, *py, *dy; } flat_t;

static double flat_getval(const flat_t* c, double t)
{   /* FCT_MinMaxCPoints, flatcurves.cc L344-365 */
    if (t < c->px[0]) t += 1.0;
    unsigned k_lo = 0, k_hi = (unsigned)c->n - 1;
    while (k_hi > 1 + k_lo) {
        const unsigned k = (k_hi + k_lo) / 2;
        if (c->px[k] > t) k_hi = k; else k_lo = k;
    }
    return c->py[k_lo] + (t - c->px[k_lo]) * c->dy[k_lo];
}
static inline float lim01f(float a) { const float m = 1.f < a ? 1.f : a; return 0.f < m ? m : 0.f; }      /* max(T(0), min(a, T(1))) */
static inline float sgnf(float a) { return (float)((0.f < a) - (a < 0.f)); }
static inline float xlog2lin_(float x, float base) { return (pow_F_scalar(base, x) - 1.f) / (base - 1.f); }
static inline float hue01(float h)
{
    const float pi2 = 2.f * (float)3.14159265358979323846;
    const float v = h / pi2;
    if (v < 0.f) return 1.f + v;
    else if (v > 1.f) return v - 1.f;
    return v;
}
static inline float tolin(float y, float base)
{
    const float v = (y - 0.5f) * 2.f;
    return sgnf(v) * lim01f(xlog2lin_(fabsf(v), base));
}

/* smoothing = params->hsl.smoothing; scale = ImProcFunctions::scale.  Curves: hc / sc / lc / coeff as (n, px, py, dy) */
int artoracle_hsl_equalizer(float* R, float* G, float* B, int W, int H, const double* ws9,
                            int nh, const double* hx, const double* hy, const double* hd,
                            int ns, const double* sx, const double* sy, const double* sd,
                            int nl, const double* lx, const double* ly, const double* ld,
                            int nc, const double* cx, const double* cy, const double* cd,
                            int smoothing, double scale)
{
    const size_t n = (size_t)W * H;
    const float w0 = (float)ws9[3], w1 = (float)ws9[4], w2 = (float)ws9[5];
    const flat_t hcurve = {nh, hx, hy, hd}, scurve = {ns, sx, sy, sd}, lcurve = {nl, lx, ly, ld}, coeff = {nc, cx, cy, cd};
    const float down = 1.f / 65535.f;
    const float PI_F = (float)3.14159265358979323846;
    float* mask = (float*)malloc(sizeof(float) * n);
    if (!mask) return 1;
    int rc = 0;
    /* setMode(YUV): g = Y, b = u, r = v; normalizeFloatTo1; yuv2hsl: r = h, b = s */
    for (size_t k = 0; k < n; ++k) {
        const float Y = R[k] * w0 + G[k] * w1 + B[k] * w2;
        float u = Y - B[k], v = R[k] - Y;
        G[k] = Y * down; u *= down; v *= down;
        B[k] = sqrtf(u * u + v * v);
        R[k] = artoracle_xatan2f(u, v);
    }
    const float smooth = powf(10.f, lim01f(smoothing / 10.f)) - 1.f;
    if (ns) {
        for (size_t k = 0; k < n; ++k) mask[k] = (float)flat_getval(&scurve, hue01(R[k]));
        const int radius = (int)(4 / scale * smooth + 0.5);
        if (radius > 0) rc = artoracle_guided_filter(G, mask, mask, W, W, H, radius, 0.001f, 0);
        for (size_t k = 0; k < n && !rc; ++k) {
            const float f = tolin(mask[k], 2.f);
            const float s = (float)(1.f + (f < 0 ? flat_getval(&coeff, B[k]) : 1.f - flat_getval(&coeff, B[k])));
            B[k] *= 1.f + sgnf(f) * pow_F_scalar(lim01f(fabsf(f)), s);
        }
    }
    if (nl && !rc) {
        for (size_t k = 0; k < n; ++k) mask[k] = (float)flat_getval(&lcurve, hue01(R[k]));
        const int radius = (int)(25 / scale * smooth + 0.5);
        if (radius > 0) rc = artoracle_guided_filter(G, mask, mask, W, W, H, radius, 0.0001f, 0);
        for (size_t k = 0; k < n && !rc; ++k) G[k] *= 1.f + tolin(mask[k], 10.f);
    }
    if (nh && !rc) {
        for (size_t k = 0; k < n; ++k) mask[k] = (float)flat_getval(&hcurve, hue01(R[k]));
        const int radius = (int)(4 / scale * smooth + 0.5);
        if (radius > 0) rc = artoracle_guided_filter(G, mask, mask, W, W, H, radius, 0.001f, 0);
        for (size_t k = 0; k < n && !rc; ++k) R[k] += tolin(mask[k], 32.f) * PI_F;
    }
    /* hsl2yuv; normalizeFloatTo65535; then the next stage's setMode(RGB): yuv2rgb (color.h L790-796) */
    for (size_t k = 0; k < n && !rc; ++k) {
        float sn, cs;
        artoracle_xsincosf(R[k], &sn, &cs);
        const float u = B[k] * sn * 65535.f, v = B[k] * cs * 65535.f, Y = G[k] * 65535.f;
        const float b = Y - u, r = v + Y;
        B[k] = b; R[k] = r;
        G[k] = (Y - r * w0 - b * w2) / w1;
    }
    free(mask);
    return rc;
}
