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

/* RawImageSource::scaleColors, Bayer branch (rawimagesource.cc L2731-2772, dynamicRowNoiseFilter off):
 * val = max(0.f, raw - cblacksom[c4]) * scale_mul[c4]; chmax[c] = max(chmax[c], val). */
static inline unsigned fc_pw_(unsigned filters, int row, int col)
{
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}
int artoracle_scale_colors_bayer(int W, int H, unsigned filters, float* raw, long stride,
                                 const float* cblacksom, const float* scale_mul, float* chmax)
{
    float mx[3] = {0.f, 0.f, 0.f};
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const int c = fc_pw_(filters, row, col);
            const int c4 = (c == 1 && !(row & 1)) ? 3 : c;
            const float d = raw[(size_t)row * stride + col] - cblacksom[c4];
            const float val = (0.f < d ? d : 0.f) * scale_mul[c4];
            raw[(size_t)row * stride + col] = val;
            mx[c] = mx[c] < val ? val : mx[c];
        }
    chmax[0] = mx[0]; chmax[1] = mx[1]; chmax[2] = mx[2];
    return 0;
}
