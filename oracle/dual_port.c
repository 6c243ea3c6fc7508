/*
 * oracle/dual_port.c -- CPU restatement of the reference's dual demosaic, "<first demosaicer> + bilinear" forms, manual contrast.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates the part of RawImageSource::dual_demosaic_RT (reference rtengine/dual_demosaic_RT.cc L39-152) that follows the first
 * demosaicer (AMaZE / RCD, restated in amaze_port.c / rcd_port.c) for Method::AMAZEBILINEAR / RCDBILINEAR  (autoContrast: usm_port.c's artoracle_auto_contrast_threshold):
 *   Color::RGB2L          (color.cc L1343-1380)  L of the demosaiced frame through the sRGB -> XYZ row of L95-99; SSE2 groups of four
 *                         read cachefy with the vector LUT form (LUT.h L349-377) unless one of the four Y is outside [0, 65535],
 *                         then (and in the row tail) each goes through computeXYZ2LabY
 *   buildBlendMask        (rt_algo.cc L405-494; usm_port.c) with contrast / 100, amount 1, blur radius 2
 *   bayer_bilinear_demosaic(blend, ...) (bayer_bilinear_demosaic.cc L33-75): every interior pixel becomes
 *                         intp(blend, first demosaicer, bilinear) -- the border row / column pair keeps the first demosaicer's values
 * Pinned bit-exact against the reference's own functions compiled in place (oracle/_ref) in tests/test_oracle_dual.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

const float* artoracle_cachef(int which);
float artoracle_xyz2laby(float f);
int artoracle_blend_mask(const float* lum, float* blend, int W, int H, float contrastThreshold, float amount, float blur_radius);
float artoracle_auto_contrast_threshold(const float* lum, int W, int H, float contrastThreshold, float luminance_factor);

static inline unsigned fc_(unsigned filters, int row, int col) { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
static inline float vclampf_(float v, float lo, float hi) { const float m = v > lo ? v : lo; return m < hi ? m : hi; }     /* minps(maxps(v, lo), hi): NaN -> lo */
static inline float lut_v(const float* data, int size, float index)
{
    const int idx = (int)vclampf_(index, 0.f, (float)(size - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclampf_(index, 0.f, (float)(size - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;
}

/* Color::RGB2L over a plane; wp = the 3x3 of dual_demosaic_RT.cc L95-99 (float constants) */
int artoracle_rgb2l(const float* R, const float* G, const float* B, float* L, int W, int H, const float* wp9)
{
    const float* cachefy = artoracle_cachef(1);
    const float w0 = wp9[3], w1 = wp9[4], w2 = wp9[5];
    for (int y = 0; y < H; ++y) {
        const float *r = R + (size_t)y * W, *g = G + (size_t)y * W, *b = B + (size_t)y * W;
        float* l = L + (size_t)y * W;
        int i = 0;
        for (; i < W - 3; i += 4) {
            float yv[4];
            int slow = 0;
            for (int k = 0; k < 4; ++k) {
                yv[k] = w0 * r[i + k] + w1 * g[i + k] + w2 * b[i + k];
                if (yv[k] > 65535.f || yv[k] < 0.f) slow = 1;
            }
            for (int k = 0; k < 4; ++k) l[i + k] = slow ? artoracle_xyz2laby(yv[k]) : lut_v(cachefy, 65536, yv[k]);
        }
        for (; i < W; ++i) l[i] = artoracle_xyz2laby(w0 * r[i] + w1 * g[i] + w2 * b[i]);
    }
    return 0;
}

/* bayer_bilinear_demosaic(blend, rawData, red, green, blue) */
int artoracle_bilinear_blend(const float* raw, const float* blend, int W, int H, unsigned filters, float* red, float* green, float* blue)
{
#define RAW(i, j) raw[(size_t)(i) * W + (j)]
#define INTP(a, b, c) ((a) * (b) + (1.f - (a)) * (c))
    for (int i = 1; i < H - 1; ++i) {
        float *n1 = red, *n2 = blue;
        if (fc_(filters, i, 0) == 2 || fc_(filters, i, 1) == 2) { float* t = n1; n1 = n2; n2 = t; }
        for (int j = 2 - (fc_(filters, i, 1) & 1); j < W - 2; j += 2) {
            const size_t o = (size_t)i * W + j;
            const float b0 = blend[o], b1 = blend[o + 1];
            green[o] = INTP(b0, green[o], RAW(i, j));
            n1[o] = INTP(b0, n1[o], (RAW(i, j - 1) + RAW(i, j + 1)) * 0.5f);
            n2[o] = INTP(b0, n2[o], (RAW(i - 1, j) + RAW(i + 1, j)) * 0.5f);
            green[o + 1] = INTP(b1, green[o + 1], ((RAW(i - 1, j + 1) + RAW(i, j)) + (RAW(i, j + 2) + RAW(i + 1, j + 1))) * 0.25f);
            n1[o + 1] = INTP(b1, n1[o + 1], RAW(i, j + 1));
            n2[o + 1] = INTP(b1, n2[o + 1], ((RAW(i - 1, j) + RAW(i - 1, j + 2)) + (RAW(i + 1, j) + RAW(i + 1, j + 2))) * 0.25f);
        }
    }
#undef RAW
#undef INTP
    return 0;
}

/* dual_demosaic_RT after the first demosaicer: red / green / blue hold its result on entry.  *contrast in percent, in / out (L108-112):
 * with auto_contrast the threshold comes from the frame (usm_port.c: artoracle_auto_contrast_threshold; needs W, H >= 80). */
int artoracle_dual_bilinear_ex(const float* raw, int W, int H, unsigned filters, float* red, float* green, float* blue, double* contrast, int auto_contrast, float* blend_out);
int artoracle_dual_bilinear(const float* raw, int W, int H, unsigned filters, float* red, float* green, float* blue, double contrast, float* blend_out)
{
    return artoracle_dual_bilinear_ex(raw, W, H, filters, red, green, blue, &contrast, 0, blend_out);
}
int artoracle_dual_bilinear_ex(const float* raw, int W, int H, unsigned filters, float* red, float* green, float* blue, double* contrast, int auto_contrast, float* blend_out)
{
    if (auto_contrast && (W < 80 || H < 80)) return 1;
    static const float xyz_rgb[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
    const size_t n = (size_t)W * H;
    float* L = (float*)malloc(sizeof(float) * n);
    float* blend = (float*)malloc(sizeof(float) * n);
    if (!L || !blend) { free(L); free(blend); return 1; }
    artoracle_rgb2l(red, green, blue, L, W, H, xyz_rgb);
    float contrastf = *contrast / 100.0;
    if (auto_contrast) contrastf = artoracle_auto_contrast_threshold(L, W, H, contrastf, 1.f);
    int rc = artoracle_blend_mask(L, blend, W, H, contrastf, 1.f, 2.f);
    *contrast = contrastf * 100.f;
    if (!rc) rc = artoracle_bilinear_blend(raw, blend, W, H, filters, red, green, blue);
    if (blend_out) memcpy(blend_out, blend, sizeof(float) * n);
    free(L); free(blend);
    return rc;
}

/* fast_xtrans_interpolate_blend(blend, rawData, red, green, blue), xtrans_demosaic.cc L1033-1092: the dual demosaic's flat-region
 * demosaicer on X-Trans: 3x3 weighted sums per colour, mixed with the first demosaicer's planes by intp(blend, first, fast); the
 * 8-pixel border keeps the first demosaicer's values */
int artoracle_xtrans_fast_blend(int W, int H, const int* xtrans36, const float* raw, const float* blend, float* red, float* green, float* blue)
{
    static const float weight[3][3] = {{0.25f, 0.5f, 0.25f}, {0.5f, 0.f, 0.5f}, {0.25f, 0.5f, 0.25f}};
#define FCOL(r, c) xtrans36[((r) % 6) * 6 + ((c) % 6)]
#define INTP(a, b, c) ((a) * (b) + (1.f - (a)) * (c))
    for (int row = 8; row < H - 8; ++row)
        for (int col = 8; col < W - 8; ++col) {
            float sum[3] = {0.f, 0.f, 0.f};
            for (int v = -1; v <= 1; v++)
                for (int h = -1; h <= 1; h++) sum[FCOL(row + v, col + h)] += raw[(size_t)(row + v) * W + (col + h)] * weight[v + 1][h + 1];
            const size_t o = (size_t)row * W + col;
            const float bl = blend[o], x = raw[o];
            switch (FCOL(row, col)) {
            case 0:
                red[o] = INTP(bl, red[o], x); green[o] = INTP(bl, green[o], sum[1] * 0.5f); blue[o] = INTP(bl, blue[o], sum[2]);
                break;
            case 1:
                green[o] = INTP(bl, green[o], x);
                if (FCOL(row, col - 1) == FCOL(row, col + 1)) { red[o] = INTP(bl, red[o], sum[0]); blue[o] = INTP(bl, blue[o], sum[2]); }
                else { red[o] = INTP(bl, red[o], sum[0] * 1.3333333f); blue[o] = INTP(bl, blue[o], sum[2] * 1.3333333f); }
                break;
            case 2:
                red[o] = INTP(bl, red[o], sum[0]); green[o] = INTP(bl, green[o], sum[1] * 0.5f); blue[o] = INTP(bl, blue[o], x);
                break;
            }
        }
#undef FCOL
#undef INTP
    return 0;
}

/* dual_demosaic_RT for Method::AMAZEVNG4 / RCDVNG4 (L128-147) after the first demosaicer: the flat-region planes come from VNG4 (vng4_port.c) */
int artoracle_vng4(int W, int H, unsigned prefilters, const float* raw, float* red, float* green, float* blue);
int artoracle_dual_vng4(const float* raw, int W, int H, unsigned prefilters, float* red, float* green, float* blue, double* contrast, int auto_contrast, float* blend_out)
{
    static const float xyz_rgb[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
    if (auto_contrast && (W < 80 || H < 80)) return 1;
    const size_t n = (size_t)W * H;
    float* L = (float*)malloc(sizeof(float) * n * 5);
    if (!L) return 1;
    float *blend = L + n, *tr = blend + n, *tg = tr + n, *tb = tg + n;
    artoracle_rgb2l(red, green, blue, L, W, H, xyz_rgb);
    float contrastf = *contrast / 100.0;
    if (auto_contrast) contrastf = artoracle_auto_contrast_threshold(L, W, H, contrastf, 1.f);
    int rc = artoracle_blend_mask(L, blend, W, H, contrastf, 1.f, 2.f);
    *contrast = contrastf * 100.f;
    if (!rc) rc = artoracle_vng4(W, H, prefilters, raw, tr, tg, tb);
    if (!rc)
        for (size_t k = 0; k < n; ++k) {
            red[k] = blend[k] * red[k] + (1.f - blend[k]) * tr[k];
            green[k] = blend[k] * green[k] + (1.f - blend[k]) * tg[k];
            blue[k] = blend[k] * blue[k] + (1.f - blend[k]) * tb[k];
        }
    if (blend_out) memcpy(blend_out, blend, sizeof(float) * n);
    free(L);
    return rc;
}
