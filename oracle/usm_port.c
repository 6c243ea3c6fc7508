/*
 * oracle/usm_port.c -- plain-C restatement of the reference's unsharp-mask sharpening.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Follows reference rtengine/: ipsharpen.cc ImProcFunctions::doSharpening L711-790 ("usm" route), apply_gamma L46-78,
 * sharpenHaloCtrl L80-141, unsharp_mask L232-312 (edgesonly == false); rt_algo.cc calcBlendFactor L47-63, buildBlendMask
 * L315-496 (autoContrast == false), get_luminance L942-956, multiply L958-975; procparams.h Threshold<T>::multiply
 * L445-503; color.h rgbLuminance L203-207; LUT.h operator[](float) L437-459; gauss.cc through oracle/gauss_port.c.
 *
 * buildBlendMask runs rows as 4-pixel SSE2 groups starting at column 2 (`for (i = 2; i < W - 5; i += 4)`) with the vector
 * xexpf and finishes each row with the scalar xexpf; both are restated.  Pinned bit-exact against the reference functions
 * compiled in place (tests/test_oracle_usm.py).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

int artoracle_gauss(const float* src, long ss, float* dst, long ds, int W, int H, double sigma);

static inline float maxr(float a, float b) { return a < b ? b : a; }
static inline float minr(float a, float b) { return b < a ? b : a; }

static inline float lut_clip(const float* data, int size, float index)
{   /* LUT.h L437-459 with LUT_CLIP_BELOW | LUT_CLIP_ABOVE */
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) return data[size - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}

/* apply_gamma<reverse>, ipsharpen.cc L46-78, pivot 1 */
static void apply_gamma(float* Y, size_t n, float gamma, int reverse)
{
    const float pivot = 1.f;
    if (!reverse) gamma = 1.f / gamma;
    float* glut = (float*)malloc(sizeof(float) * 65536);
    glut[0] = 0;
    const float d = 65535.f * pivot;
    for (int i = 1; i < 65536; ++i) {
        glut[i] = pow_F_scalar((float)i / d, gamma) * pivot;
        glut[i] *= 65535.f;
    }
    for (size_t k = 0; k < n; ++k) {
        float l = Y[k];
        if (l >= 0.f && l < 65536.f) l = lut_clip(glut, 65536, l);
        else {
            l = pow_F_scalar(maxr(l / d, 1e-18f), gamma) * pivot;
            l *= 65535.f;
        }
        Y[k] = l;
    }
    free(glut);
}

/* Threshold<int>(bottom_left, top_left, bottom_right, top_right, false)::multiply<float, float, float>, procparams.h L476-502 */
static inline float threshold_multiply(const int* t, float x, float y_max)
{
    const double val = x;
    const double bl = t[0], tl = t[1], br = t[2], tr = t[3];
    if (val == br && br == tr) return y_max;
    if (val >= br) return 0;
    if (val > tr) return (float)(y_max * (1.0 - (val - tr) / (br - tr)));
    if (val >= tl) return y_max;
    if (val > bl) return (float)(y_max * (val - bl) / (tl - bl));
    return 0;
}

/* buildBlendMask(luminance, blend, W, H, contrastThreshold, amount = 1, autoContrast = false, blur_radius, luminance_factor = 1) */
int artoracle_blend_mask(const float* lum, float* blend, int W, int H, float contrastThreshold, float amount, float blur_radius)
{
    if (contrastThreshold == 0.f) {
        for (size_t k = 0; k < (size_t)W * H; ++k) blend[k] = amount;
        return 0;
    }
    const float scale = 0.0625f / 327.68f * 1.f;
#define L(j, i) lum[(size_t)(j) * W + (i)]
#define BL(j, i) blend[(size_t)(j) * W + (i)]
    for (int j = 2; j < H - 2; ++j) {
        int i = 2;
        for (; i < W - 5; i += 4)
            for (int k = i; k < i + 4; ++k) {
                const float a = L(j, k + 1) - L(j, k - 1), b = L(j + 1, k) - L(j - 1, k), c = L(j, k + 2) - L(j, k - 2), d = L(j + 2, k) - L(j - 2, k);
                const float contrast = sqrtf(a * a + b * b + c * c + d * d) * scale;
                BL(j, k) = amount * (1.f / (1.f + xexpf_vector(16.f - 16.f * contrast / contrastThreshold)));
            }
        for (; i < W - 2; ++i) {
            const float a = L(j, i + 1) - L(j, i - 1), b = L(j + 1, i) - L(j - 1, i), c = L(j, i + 2) - L(j, i - 2), d = L(j + 2, i) - L(j - 2, i);
            const float contrast = sqrtf(a * a + b * b + c * c + d * d) * scale;
            BL(j, i) = amount * (1.f / (1.f + xexpf_scalar(16.f - 16.f * contrast / contrastThreshold)));
        }
    }
    for (int j = 0; j < 2; ++j) for (int i = 2; i < W - 2; ++i) BL(j, i) = BL(2, i);
    for (int j = H - 2; j < H; ++j) for (int i = 2; i < W - 2; ++i) BL(j, i) = BL(H - 3, i);
    for (int j = 0; j < H; ++j) {
        BL(j, 0) = BL(j, 1) = BL(j, 2);
        BL(j, W - 2) = BL(j, W - 1) = BL(j, W - 3);
    }
#undef L
#undef BL
    /* gaussianBlur(blend, blend, W, H, blur_radius) under MXCSR.FTZ: every value is in [1.1e-7, 1], no subnormal can arise */
    return artoracle_gauss(blend, W, blend, W, W, H, (double)blur_radius);
}

/* thr = Threshold<int> {bottom_left, top_left, bottom_right, top_right}; blend_out (optional) receives the blend mask */
int artoracle_usm(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                  const int* thr, int halocontrol, int halocontrol_amount, float* blend_out)
{
    if (amount < 1 || W < 8 || H < 8) return 0;
    const size_t n = (size_t)W * H;
    const float w0 = (float)wsd[3], w1 = (float)wsd[4], w2 = (float)wsd[5];
    float* Y = (float*)malloc(sizeof(float) * n);
    float* YY = (float*)malloc(sizeof(float) * n);
    float* b2 = (float*)malloc(sizeof(float) * n);
    float* blend = (float*)malloc(sizeof(float) * n);
    for (size_t k = 0; k < n; ++k) Y[k] = R[k] * w0 + G[k] * w1 + B[k] * w2;
    const float s_scale = (float)sqrt(scale);
    float contrast = pow_F_scalar((float)(contrast_p / 100.f), 1.2f) * s_scale;
    int rc = artoracle_blend_mask(Y, blend, W, H, contrast, 1.f, 2.f / s_scale);
    if (blend_out) memcpy(blend_out, blend, sizeof(float) * n);
    memcpy(YY, Y, sizeof(float) * n);
    /* unsharp_mask */
    apply_gamma(YY, n, 3.f, 0);
    if (!rc) rc = artoracle_gauss(YY, W, b2, W, W, H, radius / scale);
    if (!halocontrol) {
        for (size_t k = 0; k < n; ++k) {
            const float diff = YY[k] - b2[k];
            const float delta = threshold_multiply(thr, minr(fabsf(diff), 2000.f), amount * diff * 0.01f);
            YY[k] = blend[k] * (YY[k] + delta) + (1.f - blend[k]) * YY[k];
        }
    } else {        /* sharpenHaloCtrl(Y, b2, labCopy, blend, ...), L80-141: base is a copy of Y taken before the loop (L289-299) */
        const float scl = (100.f - halocontrol_amount) * 0.01f;
        const float sharpFac = amount * 0.01f;
        float* nL = (float*)malloc(sizeof(float) * n);
        memcpy(nL, YY, sizeof(float) * n);
#define NL(i, j) nL[(size_t)(i) * W + (j)]
        for (int i = 2; i < H - 2; i++) {
            float max1 = 0, max2 = 0, min1 = 0, min2 = 0;
            for (int j = 2; j < W - 2; j++) {
                const float np1 = 2.f * (NL(i - 2, j) + NL(i - 2, j + 1) + NL(i - 2, j + 2) + NL(i - 1, j) + NL(i - 1, j + 1) + NL(i - 1, j + 2) + NL(i, j) + NL(i, j + 1) + NL(i, j + 2)) / 27.f + NL(i - 1, j + 1) / 3.f;
                const float np2 = 2.f * (NL(i - 1, j) + NL(i - 1, j + 1) + NL(i - 1, j + 2) + NL(i, j) + NL(i, j + 1) + NL(i, j + 2) + NL(i + 1, j) + NL(i + 1, j + 1) + NL(i + 1, j + 2)) / 27.f + NL(i, j + 1) / 3.f;
                const float np3 = 2.f * (NL(i, j) + NL(i, j + 1) + NL(i, j + 2) + NL(i + 1, j) + NL(i + 1, j + 1) + NL(i + 1, j + 2) + NL(i + 2, j) + NL(i + 2, j + 1) + NL(i + 2, j + 2)) / 27.f + NL(i + 1, j + 1) / 3.f;
                const float maxn = maxr(maxr(np1, np2), np3), minn = minr(minr(np1, np2), np3);       /* rt_math.h variadic max / min */
                float max_ = maxr(maxr(max1, max2), maxn), min_ = minr(minr(min1, min2), minn);
                max1 = max2; max2 = maxn; min1 = min2; min2 = minn;
                const float labL = YY[(size_t)i * W + j];
                if (max_ < labL) max_ = labL;
                if (min_ > labL) min_ = labL;
                const float diff = NL(i, j) - b2[(size_t)i * W + j];
                const float delta = threshold_multiply(thr, minr(fabsf(diff), 2000.f), sharpFac * diff);
                float newL = labL + delta;
                if (newL > max_) newL = max_ + (newL - max_) * scl;
                else if (newL < min_) newL = min_ - (min_ - newL) * scl;
                const float bl = blend[(size_t)i * W + j];
                YY[(size_t)i * W + j] = bl * newL + (1.f - bl) * YY[(size_t)i * W + j];
            }
        }
#undef NL
        free(nL);
    }
    apply_gamma(YY, n, 3.f, 1);
    /* multiply(rgb, YY, Y) */
    for (size_t k = 0; k < n; ++k)
        if (Y[k] > 0.f) {
            const float f = YY[k] / Y[k];
            R[k] *= f; G[k] *= f; B[k] *= f;
        }
    free(Y); free(YY); free(b2); free(blend);
    return rc;
}
