/*
 * oracle/toneeq_port.c -- CPU restatement of the reference's tone equalizer.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates ImProcFunctions::toneEqualizer (reference rtengine/iptoneequalizer.cc L343-371) and tone_eq() (L68-338) without the
 * colour-map preview branch (lcms2, PREVIEW pipeline only):
 *   frame * gain, gain = 1 / 65535 * 2^-pivot; Y = LIM(rgbLuminance, 1e-5, 32)
 *   regularization > 0: guidedFilterLog(10, Y, radius 5 / scale + 0.5, eps 0.014) (guidedfilter.cc L243-269: the guide is the log image itself)
 *   regularization > 1: Y2 = Y; Y = 2^(round(5 LIM(log2 max(Y, 1e-9), -16, 6)) / 5); guidedFilter(Y2, Y, Y, 350 / scale, 0.004); and for
 *                       reg = 5 - min(regularization, 4) > 1 once more with radius (reg - 1) and eps / 100
 *   correction(y) = sum_c gauss(center_c, LIM(log2 max(y, 0), -14, 4)) factor_c / w_sum over twelve 2-EV bands, read from a 65536-entry
 *   table for Y <= 1 and evaluated directly above it -- per 4-pixel SSE2 group in the vector loop (one pixel above 1 sends the whole group
 *   through sleef's VECTOR xlogf / xexpf), per pixel in the scalar row tail; RGB *= correction; frame * (1 / gain).
 * Pinned bit-exact against the reference's own tone_eq compiled in place (oracle/_ref, shim_tone.cc) in tests/test_oracle_toneeq.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

int artoracle_guided_filter(const float* guide, const float* src, float* dst, long stride, int W, int H, int r, float epsilon, int subsampling);

static inline float maxr(float a, float b) { return a < b ? b : a; }        /* rt_math.h max */
static inline float minr(float a, float b) { return b < a ? b : a; }
static inline float lim_f(float v, float lo, float hi) { return maxr(lo, minr(v, hi)); }
static inline float vmaxf_(float a, float b) { return a > b ? a : b; }      /* _mm_max_ps(a, b) */
static inline float vminf_(float a, float b) { return a < b ? a : b; }
static inline float vclampf_(float v, float lo, float hi) { return vmaxf_(vminf_(hi, v), lo); }
static inline float lut_s(const float* data, int size, float index)
{   /* LUT.h L437-459, clip below and above */
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) return data[size - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
static inline float lut_v(const float* data, int size, float index)
{   /* LUT.h L349-377 */
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}
static inline float log2_(float x) { return xlogf_scalar(x) / xlogf_scalar(2.f); }
static inline float exp2_(float x) { return pow_F_scalar(2.f, x); }
static inline float gauss_(float b, float x) { return xexpf_scalar(-((x - b) * (x - b)) / 4.0f); }
static inline float xlog2lin_(float x, float base) { return (pow_F_scalar(base, x) - 1.f) / (base - 1.f); }

static const float centers[12] = {-16.0f, -14.0f, -12.0f, -10.0f, -8.0f, -6.0f, -4.0f, -2.0f, 0.0f, 2.0f, 4.0f, 6.0f};

static float conv(int v, float lo, float hi) { const float f = v < 0 ? lo : hi; return exp2_((float)v / 100.f * f); }

static float process_pixel(float y, const float* factors, float w_sum)
{
    const float luma = lim_f(log2_(maxr(y, 0.f)), -14.f, 4.f);
    float correction = 0.0f;
    for (int c = 0; c < 12; ++c) correction += gauss_(centers[c], luma) * factors[c];
    return correction / w_sum;
}
static float vprocess_pixel(float y, const float* factors, float w_sum)
{
    const float luma = vminf_(vmaxf_(xlogf_vector(vmaxf_(y, 0.f)) / xlogf_scalar(2.f), -14.f), 4.f);
    float correction = 0.f;
    for (int c = 0; c < 12; ++c) { const float d = luma - centers[c]; correction += xexpf_vector(-(d * d) / 4.f) * factors[c]; }
    return correction / w_sum;
}

int artoracle_tone_equalizer(float* R, float* G, float* B, int W, int H, const double* ws9, const int* bands, int regularization, double pivot, double scale)
{
    const size_t n = (size_t)W * H;
    const float w0 = (float)ws9[3], w1 = (float)ws9[4], w2 = (float)ws9[5];
    const float gain = (float)(1.f / 65535.f * pow(2.f, -pivot));
    float* Y = (float*)malloc(sizeof(float) * n * 2);
    float* lut = (float*)malloc(sizeof(float) * 65536);
    if (!Y || !lut) { free(Y); free(lut); return 1; }
    float* Y2 = Y + n;
    int rc = 0;
    const float factors[12] = {conv(bands[0], 2.f, 3.f), conv(bands[0], 2.f, 3.f), conv(bands[0], 2.f, 3.f), conv(bands[0], 2.f, 3.f), conv(bands[0], 2.f, 3.f),
                               conv(bands[1], 2.f, 3.f), conv(bands[2], 2.5f, 2.5f), conv(bands[3], 3.f, 2.f), conv(bands[4], 3.f, 2.f), conv(bands[4], 3.f, 2.f),
                               conv(bands[4], 3.f, 2.f), conv(bands[4], 3.f, 2.f)};
    for (size_t k = 0; k < n; ++k) {
        R[k] *= gain; G[k] *= gain; B[k] *= gain;
        Y[k] = lim_f(R[k] * w0 + G[k] * w1 + B[k] * w2, 1e-5f, 32.f);
    }
    const int detail = regularization > 0 ? 5 : 0;
    int radius = (int)((float)detail / scale + 0.5f);
    float epsilon = 0.01f + 0.002f * (float)(detail - 3 > 0 ? detail - 3 : 0);
    if (radius > 0) {
        for (size_t k = 0; k < n; ++k) Y[k] = xlin2log_scalar(maxr(Y[k], 0.f), 10.f);
        rc = artoracle_guided_filter(Y, Y, Y, W, W, H, radius, epsilon, 0);
        for (size_t k = 0; k < n; ++k) Y[k] = xlog2lin_(maxr(Y[k], 0.f), 10.f);
    }
    if (regularization > 1 && !rc) {
        for (size_t k = 0; k < n; ++k) {
            const float l = lim_f(log2_(maxr(Y[k], 1e-9f)), centers[0], centers[11]);      /* std::max(Y, 1e-9f) = Y < 1e-9f ? 1e-9f : Y */
            const float ll = roundf(l * 5.f) / 5.f;
            Y2[k] = Y[k];
            Y[k] = exp2_(ll);
        }
        radius = (int)(350.f / scale);
        epsilon = 0.004f;
        rc = artoracle_guided_filter(Y2, Y, Y, W, W, H, radius, epsilon, 0);
        const int reg = 5 - (regularization < 4 ? regularization : 4);
        if (reg > 1 && !rc) rc = artoracle_guided_filter(Y2, Y, Y, W, W, H, radius * (reg - 1), epsilon / 100, 0);
    }
    float w_sum = 0.f;
    for (int i = 0; i < 12; ++i) w_sum += gauss_(centers[i], 0.f);
    for (int i = 0; i < 65536; ++i) lut[i] = process_pixel((float)i / 65535.f, factors, w_sum);
    for (int y = 0; y < H && !rc; ++y) {
        float *r = R + (size_t)y * W, *g = G + (size_t)y * W, *b = B + (size_t)y * W;
        const float* cy = Y + (size_t)y * W;
        int x = 0;
        for (; x < W - 3; x += 4) {
            const int any = cy[x] > 1.f || cy[x + 1] > 1.f || cy[x + 2] > 1.f || cy[x + 3] > 1.f;      /* _mm_movemask_ps(cY > 1) */
            for (int k = 0; k < 4; ++k) {
                const float corr = any ? vprocess_pixel(cy[x + k], factors, w_sum) : lut_v(lut, 65536, cy[x + k] * 65535.f);
                r[x + k] *= corr; g[x + k] *= corr; b[x + k] *= corr;
            }
        }
        for (; x < W; ++x) {
            const float corr = cy[x] > 1.f ? process_pixel(cy[x], factors, w_sum) : lut_s(lut, 65536, cy[x] * 65535.f);
            r[x] *= corr; g[x] *= corr; b[x] *= corr;
        }
    }
    const float back = 1.f / gain;
    for (size_t k = 0; k < n; ++k) { R[k] *= back; G[k] *= back; B[k] *= back; }
    free(Y); free(lut);
    return rc;
}
