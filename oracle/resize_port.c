/*
 * oracle/resize_port.c -- CPU restatement of the reference's Lanczos resampler.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates ImProcFunctions::Lanczos (reference rtengine/ipresize.cc L38-207) without the Lab round trip around it
 * (src->setMode(LAB) / dst->setMode(mode), imagefloat.cc -- the per-pixel colour conversions are the colour chain's): a = 3 lobes,
 * support = int(2 a / min(scale, 1)) + 1 taps, weights a sin(pi x) sin(pi x / a) / (pi x)^2 through sleef's xsinf, normalised per
 * output row / column; every output row is interpolated vertically into a source-width line (the SSE2 4-column groups and the
 * scalar tail accumulate in the same order), then horizontally.
 * Pinned bit-exact against the reference's own function compiled in place (oracle/_ref) in tests/test_oracle_resize.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static float xsinf_(float d)
{   /* sleef.h L993-1016; xrintf is cvtss2si (round to nearest even) on an SSE2 build */
    int q = (int)lrintf(d * (float)0.31830988618379067154);
    float u, s;
    d = q * (-0.78515625f * 4) + d;
    d = q * (-0.00024127960205078125f * 4) + d;
    d = q * (-6.3329935073852539062e-07f * 4) + d;
    d = q * (-4.9604681473525147339e-10f * 4) + d;
    s = d * d;
    if ((q & 1) != 0) d = -d;
    u = 2.6083159809786593541503e-06f;
    u = u * s + -0.0001981069071916863322258f;
    u = u * s + 0.00833307858556509017944336f;
    u = u * s + -0.166666597127914428710938f;
    u = s * (u * d) + d;
    return u;
}

static float lanc(float x, float a)
{   /* L38-48 */
    if (x * x < 1e-6f) return 1.0f;
    else if (x * x > a * a) return 0.0f;
    else {
        x = (float)3.14159265358979323846 * x;
        return a * xsinf_(x) * xsinf_(x / a) / (x * x);
    }
}

/* three planes (the reference's g / r / b = L / a / b slots are treated alike), contiguous, sW x sH -> dW x dH */
int artoracle_lanczos(const float* s0, const float* s1, const float* s2, int sW, int sH, float* d0, float* d1, float* d2, int dW, int dH, float scale)
{
    if (sW < 1 || sH < 1 || dW < 1 || dH < 1 || !(scale > 0.f)) return 1;
    const float delta = 1.0f / scale;
    const float a = 3.0f;
    const float sc = scale < 1.0f ? scale : 1.0f;
    const int support = (int)(2.0f * a / sc) + 1;
    float* wwh = (float*)calloc((size_t)support * dW, sizeof(float));
    int* jj0 = (int*)malloc(sizeof(int) * dW); int* jj1 = (int*)malloc(sizeof(int) * dW);
    for (int j = 0; j < dW; j++) {      /* L82-109 */
        const float x0 = ((float)j + 0.5f) * delta - 0.5f;
        float* w = wwh + (size_t)j * support;
        float ws = 0.0f;
        int lo = (int)floorf(x0 - a / sc) + 1; if (lo < 0) lo = 0;
        int hi = (int)floorf(x0 + a / sc) + 1; if (hi > sW) hi = sW;
        jj0[j] = lo; jj1[j] = hi;
        for (int jj = lo; jj < hi; jj++) {
            const int k = jj - lo;
            const float z = sc * (x0 - (float)jj);
            w[k] = lanc(z, a);
            ws += w[k];
        }
        for (int k = 0; k < support; k++) w[k] /= ws;
    }
    const float* S[3] = {s0, s1, s2};
    float* Dp[3] = {d0, d1, d2};
#pragma omp parallel
    {
        float* line = (float*)malloc(sizeof(float) * (size_t)sW * 3);
        float* w = (float*)calloc((size_t)support, sizeof(float));
#pragma omp for
        for (int i = 0; i < dH; i++) {  /* L130-199 */
            const float y0 = ((float)i + 0.5f) * delta - 0.5f;
            float ws = 0.0f;
            int ii0 = (int)floorf(y0 - a / sc) + 1; if (ii0 < 0) ii0 = 0;
            int ii1 = (int)floorf(y0 + a / sc) + 1; if (ii1 > sH) ii1 = sH;
            for (int ii = ii0; ii < ii1; ii++) {
                const int k = ii - ii0;
                const float z = sc * (y0 - (float)ii);
                w[k] = lanc(z, a);
                ws += w[k];
            }
            for (int k = 0; k < support; k++) w[k] /= ws;
            for (int c = 0; c < 3; ++c) {
                float* l = line + (size_t)c * sW;
                for (int j = 0; j < sW; j++) {
                    float v = 0.0f;
                    for (int ii = ii0; ii < ii1; ii++) v += w[ii - ii0] * S[c][(size_t)ii * sW + j];
                    l[j] = v;
                }
                for (int j = 0; j < dW; j++) {
                    const float* wh = wwh + (size_t)support * j;
                    float v = 0.0f;
                    for (int jj = jj0[j]; jj < jj1[j]; jj++) v += wh[jj - jj0[j]] * l[jj];
                    Dp[c][(size_t)i * dW + j] = v;
                }
            }
        }
        free(line); free(w);
    }
    free(wwh); free(jj0); free(jj1);
    return 0;
}
