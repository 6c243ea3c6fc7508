/*
 * oracle/smoothing_port.c -- CPU restatement of the reference's chroma smoothing after RGB_denoise.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates denoise::denoiseGuidedSmoothing (reference rtengine/ipsmoothing.cc L875-897), which ImProcFunctions::denoise runs when
 * smoothingEnabled (ipdenoise.cc L1171-1172; guidedChromaRadius defaults to 3): the frame is scaled to [0, 1]
 * (Imagefloat::normalizeFloatTo1, imagefloat.cc L396-432), guided_smoothing(..., Channel::C, radius, 0.001, scale) (L334-409) filters
 * each of R, G, B in a log encoding (guidedFilterLog, guidedfilter.cc L243-263: xlin2log base 10, guidedFilter with automatic
 * subsampling, xlog2lin) guided by the log luminance, then keeps the INPUT luminance and takes the filtered chroma scaled by
 * Y_in / Y_filtered (Color::rgb2yuv / yuv2rgb, color.h L783-796), and the frame is scaled back by 65535.
 * Pinned bit-exact against the reference's own functions compiled in place (oracle/_ref) in tests/test_oracle_smoothing.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

int artoracle_guided_filter(const float* guide, const float* src, float* dst, long stride, int W, int H, int r, float epsilon, int subsampling);

static inline float max0(float v) { return v < 0.f ? 0.f : v; }        /* rtengine::max(v, 0.f) = v < 0 ? 0 : v */
static inline float xlog2lin_(float x, float base) { return (pow_F_scalar(base, x) - 1.f) / (base - 1.f); }

int artoracle_denoise_guided_smoothing(float* R, float* G, float* B, int W, int H, const double* ws9, int guidedChromaRadius, double scale)
{
    if (guidedChromaRadius == 0) return 0;
    const size_t n = (size_t)W * H;
    const float w0 = (float)ws9[3], w1 = (float)ws9[4], w2 = (float)ws9[5];
    float* ch[3] = {R, G, B};
    const float down = 1.f / 65535.f;
    for (int c = 0; c < 3; ++c) for (size_t k = 0; k < n; ++k) ch[c][k] *= down;
    const int r = (int)round(guidedChromaRadius / scale) > 0 ? (int)round(guidedChromaRadius / scale) : 0;
    int rc = 0;
    if (r > 0) {
        float* in = (float*)malloc(sizeof(float) * n * 4);
        if (!in) return 1;
        float *iR = in, *iG = in + n, *iB = in + 2 * n, *guide = in + 3 * n;
        memcpy(iR, R, sizeof(float) * n); memcpy(iG, G, sizeof(float) * n); memcpy(iB, B, sizeof(float) * n);
        for (size_t k = 0; k < n; ++k) {
            const float l = R[k] * w0 + G[k] * w1 + B[k] * w2;
            guide[k] = xlin2log_scalar(max0(l), 10.f);
        }
        for (int c = 0; c < 3 && !rc; ++c) {       /* guidedFilterLog(guide, 10.f, chan, r, epsilon) */
            for (size_t k = 0; k < n; ++k) ch[c][k] = xlin2log_scalar(max0(ch[c][k]), 10.f);
            rc = artoracle_guided_filter(guide, ch[c], ch[c], W, W, H, r, 0.001f, 0);
            for (size_t k = 0; k < n; ++k) ch[c][k] = xlog2lin_(max0(ch[c][k]), 10.f);
        }
        for (size_t k = 0; k < n && !rc; ++k) {
            const float iY = iR[k] * w0 + iG[k] * w1 + iB[k] * w2;
            float oY = R[k] * w0 + G[k] * w1 + B[k] * w2;
            float ou = oY - B[k], ov = R[k] - oY;
            const float bump = oY > 1e-5f ? iY / oY : 1.f;
            ou *= bump; ov *= bump; oY = iY;
            const float b = oY - ou, rr = ov + oY;
            B[k] = b; R[k] = rr;
            G[k] = (oY - rr * w0 - b * w2) / w1;
        }
        free(in);
    }
    for (int c = 0; c < 3; ++c) for (size_t k = 0; k < n; ++k) ch[c][k] *= 65535.f;
    return rc;
}
