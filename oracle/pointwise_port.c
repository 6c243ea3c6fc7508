/*
 * oracle/pointwise_port.c -- CPU restatement of the per-pixel stages.  TEST INFRASTRUCTURE ONLY.
 *
 * artoracle_scale_convert: RawImageSource::getImage's gain/clip at skip == 1 (reference
 * rtengine/rawimagesource.cc L943-1025: `rtot *= rm; if (doClip) rtot = CLIP(rtot)`, CLIP = rt_math.h
 * L97-101) followed by the matrix branch of RawImageSource::colorSpaceConversion_ (L3197-3211:
 * float = double coefficient * float sample summed left to right in double).
 * Pinned against oracle/_ref (the reference's own CLIP and its own matrix loop) in
 * tests/test_oracle_pointwise.py: bit-exact.
 */
#include <stddef.h>

static inline float clip65535_(float a)
{   /* LIM(a, 0.f, 65535.f) = max(0, min(a, 65535)), rt_math.h L84-101 */
    const float m = a < 65535.f ? a : 65535.f;
    return 0.f < m ? m : 0.f;
}

int artoracle_scale_convert(int W, int H, float* r, float* g, float* b, long stride,
                            const float* mul, int doClip, const double* mat)
{
#pragma omp parallel for
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const size_t k = (size_t)i * stride + j;
            float x = r[k] * mul[0], y = g[k] * mul[1], z = b[k] * mul[2];
            if (doClip) { x = clip65535_(x); y = clip65535_(y); z = clip65535_(z); }
            if (mat) {
                const float nx = (float)(mat[0] * x + mat[1] * y + mat[2] * z);
                const float ny = (float)(mat[3] * x + mat[4] * y + mat[5] * z);
                const float nz = (float)(mat[6] * x + mat[7] * y + mat[8] * z);
                x = nx; y = ny; z = nz;
            }
            r[k] = x; g[k] = y; b[k] = z;
        }
    return 0;
}
