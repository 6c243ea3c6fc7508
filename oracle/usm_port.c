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
int artoracle_gauss_iir(float* src, long ss, float* dst, long ds, const float* divb, long vs, int W, int H, double sigma, int type);

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


/* ---------------------------------------------------------------------------------------------------------------------------
 * buildBlendMask's automatic contrast threshold (rt_algo.cc L317-414: tileAverage L65-86, tileVariance L88-110, calcContrastThreshold
 * L112-170).  Pass 0 looks for the flattest 80x80 tile on an 80-pixel grid, pass 1 for the flattest 40x40 tile on a 10-pixel grid and
 * then around it pixel by pixel; the threshold is the smallest c / 100 for which the blend factors of that tile sum to <= 1 % of it.
 * The SSE2 loops keep four lane sums (columns 0..3 mod 4 of each group) beside the scalar tail's sum; vhadd adds lanes (0 + 2) + (1 + 3).
 * ------------------------------------------------------------------------------------------------------------------------- */
static float tile_average(const float* d, int W, int tileY, int tileX, int ts)
{
    float avg = 0.f, v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = tileY; y < tileY + ts; ++y) {
        int x = tileX;
        for (; x < tileX + ts - 3; x += 4) for (int k = 0; k < 4; ++k) v[k] += d[(size_t)y * W + x + k];
        for (; x < tileX + ts; ++x) avg += d[(size_t)y * W + x];
    }
    avg += (v[0] + v[2]) + (v[1] + v[3]);
    return avg / (float)(ts * ts);
}
static float tile_variance(const float* d, int W, int tileY, int tileX, int ts, float avg)
{
    float var = 0.f, v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int y = tileY; y < tileY + ts; ++y) {
        int x = tileX;
        for (; x < tileX + ts - 3; x += 4) for (int k = 0; k < 4; ++k) { const float t = d[(size_t)y * W + x + k] - avg; v[k] += t * t; }
        for (; x < tileX + ts; ++x) { const float t = d[(size_t)y * W + x] - avg; var += t * t; }
    }
    var += (v[0] + v[2]) + (v[1] + v[3]);
    return var / ((float)(ts * ts) * avg);
}
static float calc_contrast_threshold(const float* lum, int W, int tileY, int tileX, int ts, float factor)
{
    const float scale = 0.0625f / 327.68f * factor;
    const int n = ts - 4;
    float* bl = (float*)malloc(sizeof(float) * (size_t)n * n);
#define L(j, i) lum[(size_t)(j) * W + (i)]
    for (int j = tileY + 2; j < tileY + ts - 2; ++j)
        for (int i = tileX + 2; i < tileX + ts - 2; ++i) {
            const float a = L(j, i + 1) - L(j, i - 1), b = L(j + 1, i) - L(j - 1, i), c = L(j, i + 2) - L(j, i - 2), d = L(j + 2, i) - L(j - 2, i);
            bl[(size_t)(j - tileY - 2) * n + (i - tileX - 2)] = sqrtf(a * a + b * b + c * c + d * d) * scale;
        }
#undef L
    const float limit = (float)(n * n) / 100.f;
    int c;
    for (c = 1; c < 100; ++c) {
        const float thr = c / 100.f;
        float sum = 0.f, v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < n; ++j) {
            int i = 0;
            for (; i < ts - 7; i += 4) for (int k = 0; k < 4; ++k) v[k] += 1.f / (1.f + xexpf_vector(16.f - 16.f * bl[(size_t)j * n + i + k] / thr));
            for (; i < n; ++i) sum += 1.f / (1.f + xexpf_scalar(16.f - 16.f * bl[(size_t)j * n + i] / thr));
        }
        sum += (v[0] + v[2]) + (v[1] + v[3]);
        if (sum <= limit) break;
    }
    free(bl);
    return c / 100.f;
}
static float tile_score(const float* lum, int W, int tileY, int tileX, int ts, float minLum, float maxLum)
{
    const float avg = tile_average(lum, W, tileY, tileX, ts);
    if (avg < minLum || avg > maxLum) return INFINITY;
    const float v = tile_variance(lum, W, tileY, tileX, ts, avg);
    return v < 0.5f ? INFINITY : v;
}
/* the autoContrast block of buildBlendMask: returns the threshold (the caller's value when neither pass assigns one); needs W, H >= 80 */
float artoracle_auto_contrast_threshold(const float* lum, int W, int H, float contrastThreshold, float luminance_factor)
{
    const float minLum = 2000.f / luminance_factor, maxLum = 20000.f / luminance_factor;
    for (int pass = 0; pass < 2; ++pass) {
        const int ts = 80 / (pass + 1);
        const int skip = pass == 0 ? ts : ts / 4;
        const int ntw = W / skip - 3 * pass, nth = H / skip - 3 * pass;
        float minvar = INFINITY;
        int minI = 0, minJ = 0;
        for (int i = 0; i < nth; ++i)
            for (int j = 0; j < ntw; ++j) {
                const float v = tile_score(lum, W, i * skip, j * skip, ts, minLum, maxLum);
                if (v < minvar) { minvar = v; minI = i; minJ = j; }
            }
        if (minvar <= 1.f || pass == 1) {
            const int minY = skip * minI, minX = skip * minJ;
            if (pass == 0) return calc_contrast_threshold(lum, W, minY, minX, ts, luminance_factor);
            const int y0 = minY - skip > 0 ? minY - skip : 0, x0 = minX - skip > 0 ? minX - skip : 0;
            const int y1 = minY + skip < H - ts ? minY + skip : H - ts, x1 = minX + skip < W - ts ? minX + skip : W - ts;
            float mv = INFINITY;
            int mi = 0, mj = 0;
            for (int i = 0; i < y1 - y0 + 1; ++i)
                for (int j = 0; j < x1 - x0 + 1; ++j) {
                    const float v = tile_score(lum, W, y0 + i, x0 + j, ts, minLum, maxLum);
                    if (v < mv) { mv = v; mi = i; mj = j; }
                }
            contrastThreshold = mv <= 8.f ? calc_contrast_threshold(lum, W, y0 + mi, x0 + mj, ts, luminance_factor) : 0.f;
        }
    }
    return contrastThreshold;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * bilateral<float, float>(src, dst, buffer, W, H, sigma, sens), bilateral2.h L38-547: 21 fixed integer kernels (3x3 .. 11x11, one per
 * 0.1 step of sigma) weighted by a range LUT ec[d + 65536] = exp(-d^2 / (2 sens^2)) * scale, read with LUTf's interpolating
 * float index; each sum runs row-major over src[i - a][j - b], a, b = -h .. h, in float; the h-pixel border is copied.
 * Rows: {LUT scale, half width h, quarter kernel (h + 1) x (h + 1), row-major, corner first} -- the BL_BEGIN / BL_OPERn arguments.
 * ------------------------------------------------------------------------------------------------------------------------- */
typedef struct { int scale, half; int q[36]; } bl_kernel;
static const bl_kernel BL_KERNELS[21] = {
    {318, 1, {1, 7, 7, 55}},                                                                                  /* sigma 0.5, L151-164 */
    {768, 1, {1, 4, 4, 16}},
    {366, 2, {0, 0, 1, 0, 8, 21, 1, 21, 59}},
    {753, 2, {0, 0, 1, 0, 5, 10, 1, 10, 23}},
    {595, 2, {0, 1, 2, 1, 6, 12, 2, 12, 22}},
    {910, 2, {0, 1, 2, 1, 4, 7, 2, 7, 12}},                                                                   /* 1.0 */
    {209, 3, {0, 0, 1, 1, 0, 2, 5, 8, 1, 5, 18, 27, 1, 8, 27, 41}},
    {322, 3, {0, 0, 1, 1, 0, 1, 4, 6, 1, 4, 11, 16, 1, 6, 16, 23}},
    {336, 3, {0, 0, 1, 1, 0, 2, 4, 6, 1, 4, 11, 14, 1, 6, 14, 19}},
    {195, 3, {0, 1, 2, 3, 1, 4, 8, 10, 2, 8, 17, 21, 3, 10, 21, 28}},
    {132, 4, {0, 0, 0, 1, 1, 0, 1, 2, 4, 5, 0, 2, 6, 12, 14, 1, 4, 12, 22, 28, 1, 5, 14, 28, 35}},            /* 1.5 */
    {180, 4, {0, 0, 0, 1, 1, 0, 1, 2, 3, 4, 0, 2, 5, 9, 10, 1, 3, 9, 15, 19, 1, 4, 10, 19, 23}},
    {195, 4, {0, 0, 1, 1, 1, 0, 1, 2, 3, 4, 1, 2, 5, 8, 9, 1, 3, 8, 13, 16, 1, 4, 9, 16, 19}},
    {151, 4, {0, 0, 1, 2, 2, 0, 1, 3, 5, 5, 1, 3, 6, 10, 12, 2, 5, 10, 16, 19, 2, 5, 12, 19, 22}},
    {151, 4, {0, 0, 1, 2, 2, 0, 1, 3, 4, 5, 1, 3, 5, 8, 9, 2, 4, 8, 12, 14, 2, 5, 9, 14, 16}},
    {116, 5, {0, 0, 0, 1, 1, 1, 0, 0, 1, 2, 3, 3, 0, 1, 2, 4, 7, 7, 1, 2, 4, 8, 12, 14, 1, 3, 7, 12, 18, 20, 1, 3, 7, 14, 20, 23}},   /* 2.0 */
    {127, 5, {0, 0, 0, 1, 1, 1, 0, 0, 1, 2, 3, 3, 0, 1, 2, 4, 6, 7, 1, 2, 4, 8, 11, 12, 1, 3, 6, 11, 15, 17, 1, 3, 7, 12, 17, 19}},
    {109, 5, {0, 0, 0, 1, 1, 2, 0, 1, 2, 3, 3, 4, 1, 2, 3, 5, 7, 8, 1, 3, 5, 9, 12, 13, 1, 3, 7, 12, 16, 18, 2, 4, 8, 13, 18, 20}},
    {132, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 3, 5, 6, 7, 1, 2, 5, 7, 10, 11, 1, 3, 6, 10, 13, 14, 1, 3, 7, 11, 14, 16}},
    {156, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 3, 4, 5, 6, 1, 2, 4, 6, 8, 9, 1, 3, 5, 8, 10, 11, 1, 3, 6, 9, 11, 12}},
    {173, 5, {0, 0, 1, 1, 1, 1, 0, 1, 1, 2, 3, 3, 1, 1, 2, 4, 5, 5, 1, 2, 4, 5, 7, 7, 1, 3, 5, 7, 9, 9, 1, 3, 5, 7, 9, 10}},           /* 2.5 */
};
/* the dispatcher's thresholds (L487-547): kernel k serves sigma < BL_LIMITS[k], the last one everything above */
static const double BL_LIMITS[20] = {0.55, 0.65, 0.75, 0.85, 0.95, 1.05, 1.15, 1.25, 1.35, 1.45, 1.55, 1.65, 1.75, 1.85, 1.95, 2.05, 2.15, 2.25, 2.35, 2.45};

static inline float bl_lut(const float* ec, float index)
{   /* LUTf::operator[](float), LUT.h L437-459: 0x20000 entries, clipped both ways */
    if (index < 0.f || !(index == index)) return ec[0];
    if (index > 131070.f) return ec[131071];
    const int idx = (int)index;
    const float diff = index - (float)idx;
    const float p1 = ec[idx];
    const float p2 = ec[idx + 1] - p1;
    return p1 + p2 * diff;
}

int artoracle_bilateral(const float* src, float* dst, int W, int H, double sigma, double sens)
{
    if (sigma < 0.45) { memcpy(dst, src, sizeof(float) * (size_t)W * H); return 0; }
    int kx = 0;
    while (kx < 20 && !(sigma < BL_LIMITS[kx])) kx++;
    const bl_kernel* K = &BL_KERNELS[kx];
    const int h = K->half, hw = h + 1;
    float* ec = (float*)malloc(sizeof(float) * 0x20000);
    if (!ec) return 1;
    const double scale = K->scale;
    for (int i = 0; i < 0x20000; i++) ec[i] = (exp(-(double)(i - 0x10000) * (double)(i - 0x10000) / (2.0 * sens * sens)) * scale);
#pragma omp parallel for
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            const size_t o = (size_t)i * W + j;
            if (i < h || j < h || i >= H - h || j >= W - h) { dst[o] = src[o]; continue; }
            const float c = src[o];
            float v = 0.f, den = 0.f;
            int first = 1;
            for (int a = -h; a <= h; a++)
                for (int b = -h; b <= h; b++) {
                    const float coef = (float)K->q[(h - abs(a)) * hw + (h - abs(b))];
                    const float s = src[(size_t)(i - a) * W + (j - b)];
                    const float e = bl_lut(ec, s - c + 65536.0f);
                    const float tv = coef * (s * e), td = coef * e;
                    if (first) { v = tv; den = td; first = 0; }
                    else { v = v + tv; den = den + td; }
                }
            dst[o] = v / den;
        }
    free(ec);
    return 0;
}

/* thr = Threshold<int> {bottom_left, top_left, bottom_right, top_right}; blend_out (optional) receives the blend mask */
int artoracle_usm_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                     const int* thr, int halocontrol, int halocontrol_amount, float* blend_out, int edgesonly, double edges_radius, int edges_tolerance);
int artoracle_usm(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                  const int* thr, int halocontrol, int halocontrol_amount, float* blend_out)
{
    return artoracle_usm_ex(R, G, B, W, H, wsd, scale, contrast_p, radius, amount, thr, halocontrol, halocontrol_amount, blend_out, 0, 1.9, 1800);
}
/* edgesonly (unsharp_mask L239-262): the difference image is taken on a bilateral-filtered copy (base = b3), the blur of that copy */
int artoracle_usm_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                     const int* thr, int halocontrol, int halocontrol_amount, float* blend_out, int edgesonly, double edges_radius, int edges_tolerance)
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
    float* base = YY;
    if (edgesonly) {
        base = (float*)malloc(sizeof(float) * n);
        if (!rc) rc = artoracle_bilateral(YY, base, W, H, edges_radius / scale, (double)edges_tolerance);
    }
    if (!rc) rc = artoracle_gauss(base, W, b2, W, W, H, radius / scale);
    if (!halocontrol) {
        for (size_t k = 0; k < n; ++k) {
            const float diff = base[k] - b2[k];
            const float delta = threshold_multiply(thr, minr(fabsf(diff), 2000.f), amount * diff * 0.01f);
            YY[k] = blend[k] * (YY[k] + delta) + (1.f - blend[k]) * YY[k];
        }
    } else {        /* sharpenHaloCtrl(Y, b2, labCopy, blend, ...), L80-141: base is a copy of Y taken before the loop (L289-299) */
        const float scl = (100.f - halocontrol_amount) * 0.01f;
        const float sharpFac = amount * 0.01f;
        float* nL = (float*)malloc(sizeof(float) * n);
        memcpy(nL, base, sizeof(float) * n);        /* edgesonly passes b3 itself (L302) */
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
    if (edgesonly) free(base);
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

/* ---------------------------------------------------------------------------------------------------------------------------
 * "rld" route of doSharpening (ipsharpen.cc L747-771, no corner boost): markImpulse (rt_algo.cc L497-591), deconvsharpening
 * (ipsharpen.cc L144-230) over gaussianBlur's GAUSS_DIV / GAUSS_MULT forms: gauss3x3div / mult (gauss.cc L177-274), gauss5x5div / mult (L331-378,
 * L415-443), gauss7x7div / mult (L276-329, L380-413), kernels L52-92, dispatch L1444-1511; above sigma 1.15 the recursive forms
 * (artoracle_gauss_iir in gauss_port.c).
 * ------------------------------------------------------------------------------------------------------------------------- */
#define S_(i, j) src[(size_t)(i) * W + (j)]
#define D_(i, j) dst[(size_t)(i) * W + (j)]
#define V_(i, j) divb[(size_t)(i) * W + (j)]

static void kernel5(float sigma, float k[5][5])
{   /* compute5x5kernel, L73-92 */
    const double temp = -2.f * (sigma * sigma);
    float sum = 0.f;
    for (int i = -2; i <= 2; ++i)
        for (int j = -2; j <= 2; ++j) {
            if ((i * i + j * j) <= (3.0 * 0.84) * (3.0 * 0.84)) { k[i + 2][j + 2] = (float)exp((i * i + j * j) / temp); sum += k[i + 2][j + 2]; }
            else k[i + 2][j + 2] = 0.f;
        }
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) k[i][j] /= sum;
}
static void kernel7(float sigma, float k[7][7])
{   /* compute7x7kernel, L52-71 */
    const double temp = -2.f * (sigma * sigma);
    float sum = 0.f;
    for (int i = -3; i <= 3; ++i)
        for (int j = -3; j <= 3; ++j) {
            if ((i * i + j * j) <= (3.0 * 1.15) * (3.0 * 1.15)) { k[i + 3][j + 3] = (float)exp((i * i + j * j) / temp); sum += k[i + 3][j + 3]; }
            else k[i + 3][j + 3] = 0.f;
        }
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) k[i][j] /= sum;
}
static inline float conv5(const float* src, int W, int i, int j, const float k[5][5])
{
    const float c21 = k[0][1], c20 = k[0][2], c11 = k[1][1], c10 = k[1][2], c00 = k[2][2];
    return c21 * (S_(i - 2, j - 1) + S_(i - 2, j + 1) + S_(i - 1, j - 2) + S_(i - 1, j + 2) + S_(i + 1, j - 2) + S_(i + 1, j + 2) + S_(i + 2, j - 1) + S_(i + 2, j + 1)) +
           c20 * (S_(i - 2, j) + S_(i, j - 2) + S_(i, j + 2) + S_(i + 2, j)) +
           c11 * (S_(i - 1, j - 1) + S_(i - 1, j + 1) + S_(i + 1, j - 1) + S_(i + 1, j + 1)) +
           c10 * (S_(i - 1, j) + S_(i, j - 1) + S_(i, j + 1) + S_(i + 1, j)) +
           c00 * S_(i, j);
}
static inline float conv7(const float* src, int W, int i, int j, const float k[7][7])
{   /* note `src[i - 2][j + 1] * c21` inside the c21 group: as written in the reference (L302, L402) */
    const float c31 = k[0][2], c30 = k[0][3], c22 = k[1][1], c21 = k[1][2], c20 = k[1][3], c11 = k[2][2], c10 = k[2][3], c00 = k[3][3];
    return c31 * (S_(i - 3, j - 1) + S_(i - 3, j + 1) + S_(i - 1, j - 3) + S_(i - 1, j + 3) + S_(i + 1, j - 3) + S_(i + 1, j + 3) + S_(i + 3, j - 1) + S_(i + 3, j + 1)) +
           c30 * (S_(i - 3, j) + S_(i, j - 3) + S_(i, j + 3) + S_(i + 3, j)) +
           c22 * (S_(i - 2, j - 2) + S_(i - 2, j + 2) + S_(i + 2, j - 2) + S_(i + 2, j + 2)) +
           c21 * (S_(i - 2, j - 1) + S_(i - 2, j + 1) * c21 + S_(i - 1, j - 2) + S_(i - 1, j + 2) + S_(i + 1, j - 2) + S_(i + 1, j + 2) + S_(i + 2, j - 1) + S_(i + 2, j + 1)) +
           c20 * (S_(i - 2, j) + S_(i, j - 2) + S_(i, j + 2) + S_(i + 2, j)) +
           c11 * (S_(i - 1, j - 1) + S_(i - 1, j + 1) + S_(i + 1, j - 1) + S_(i + 1, j + 1)) +
           c10 * (S_(i - 1, j) + S_(i, j - 1) + S_(i, j + 1) + S_(i + 1, j)) +
           c00 * S_(i, j);
}
/* the 3x3 forms: value of the (bordered) 3x3 blur at (i, j) */
static inline float conv3(const float* src, int W, int H, int i, int j, float c0, float c1, float c2, float b0, float b1)
{
    const int top = (i == 0 || i == H - 1), side = (j == 0 || j == W - 1);
    if (top && side) return S_(i, j);
    if (top) return b1 * (S_(i, j - 1) + S_(i, j + 1)) + b0 * S_(i, j);
    if (side) return b1 * (S_(i - 1, j) + S_(i + 1, j)) + b0 * S_(i, j);
    return c2 * (S_(i - 1, j - 1) + S_(i - 1, j + 1) + S_(i + 1, j - 1) + S_(i + 1, j + 1)) + c1 * (S_(i - 1, j) + S_(i, j - 1) + S_(i, j + 1) + S_(i + 1, j)) + c0 * S_(i, j);
}
/* gaussianBlur(src, dst, W, H, sigma, nullptr, GAUSS_DIV, divb) / (..., GAUSS_MULT), src != dst; 1 = unsupported sigma (>= 25) */
static int gauss_divmult(const float* src, float* dst, const float* divb, int W, int H, double sigma, int mult)
{
    if (sigma < 0.25) {        /* GAUSS_SKIP: plain copy whatever the type (L1436-1443) */
        memcpy(dst, src, sizeof(float) * (size_t)W * H);
        return 0;
    }
    if (sigma < 0.6) {
        double c0 = 1.0, c1 = exp(-0.5 * ((1.0 / sigma) * (1.0 / sigma))), c2 = exp(-((1.0 / sigma) * (1.0 / sigma)));
        const double sum = c0 + 4.0 * (c1 + c2);
        c0 /= sum; c1 /= sum; c2 /= sum;
        double b1 = exp(-1.0 / (2.0 * sigma * sigma));
        const double bsum = 2.0 * b1 + 1.0;
        b1 /= bsum;
        const double b0 = 1.0 / bsum;
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const float t = conv3(src, W, H, i, j, (float)c0, (float)c1, (float)c2, (float)b0, (float)b1);
                if (mult) D_(i, j) *= t;
                else D_(i, j) = maxr(V_(i, j) / (t > 0.f ? t : 1.f), 0.f);
            }
        return 0;
    }
    if (sigma <= 0.84) {
        float k[5][5];
        kernel5((float)sigma, k);
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const int inner = i >= 2 && i < H - 2 && j >= 2 && j < W - 2;
                if (mult) { if (inner) D_(i, j) *= conv5(src, W, i, j, k); }
                else D_(i, j) = inner ? V_(i, j) / maxr(conv5(src, W, i, j, k), 0.00001f) : 1.f;
            }
        return 0;
    }
    if (sigma <= 1.15) {
        float k[7][7];
        kernel7((float)sigma, k);
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < W; ++j) {
                const int inner = i >= 3 && i < H - 3 && j >= 3 && j < W - 3;
                if (mult) { if (inner) D_(i, j) *= conv7(src, W, i, j, k); }
                else D_(i, j) = inner ? V_(i, j) / maxr(conv7(src, W, i, j, k), 0.00001f) : 1.f;
            }
        return 0;
    }
    /* recursive forms (gauss.cc L1490-1511); GAUSS_MULT blurs its source plane in place on the way */
    return artoracle_gauss_iir((float*)src, W, dst, W, divb, W, W, H, sigma, mult ? 1 : 2);
}
#undef S_
#undef D_
#undef V_

/* markImpulse(width, height, src, impulse, thresh), rt_algo.cc L497-591: the SSE2 and scalar forms agree (same summation order per pixel;
 * the sum contains hpfabs itself, so it cannot round below it and the sign test equals the comparison) */
int artoracle_mark_impulse(const float* src, unsigned char* impulse, int W, int H, float thresh)
{
    float* lpf = (float*)malloc(sizeof(float) * (size_t)W * H);
    int rc = artoracle_gauss(src, W, lpf, W, W, H, (double)maxr(2.f, thresh - 1.f));
    const float impthr = maxr(1.f, 5.5f - thresh);
    const float impthrDiv24 = impthr / 24.0f;
    for (int i = 0; i < H && !rc; i++)
        for (int j = 0; j < W; j++) {
            const float hpfabs = fabsf(src[(size_t)i * W + j] - lpf[(size_t)i * W + j]);
            float hfnbrave = 0;
            for (int i1 = i - 2 > 0 ? i - 2 : 0; i1 <= (i + 2 < H - 1 ? i + 2 : H - 1); i1++)
                for (int j1 = j - 2 > 0 ? j - 2 : 0; j1 <= (j + 2 < W - 1 ? j + 2 : W - 1); j1++)
                    hfnbrave += fabsf(src[(size_t)i1 * W + j1] - lpf[(size_t)i1 * W + j1]);
            impulse[(size_t)i * W + j] = (hpfabs > ((hfnbrave - hpfabs) * impthrDiv24));
        }
    free(lpf);
    return rc;
}

/* deconvsharpening(luminance, blend, impulse, W, H, sigma, amount), ipsharpen.cc L144-230 */
static int deconv(float* lum, const float* blend, const unsigned char* impulse, int W, int H, double sigma, float amount)
{
    if (amount <= 0) return 0;
    if (sigma < 0.2f) return 0;
    if (sigma >= 25.0 || (sigma > 1.15 && (W < 4 || H < 4))) return 1;
    const size_t n = (size_t)W * H;
    const int maxiter = 20;
    const float delta_factor = 0.2f, offset = 1000.f;
    float* tmp = (float*)malloc(sizeof(float) * n); float* tmpI = (float*)malloc(sizeof(float) * n); float* out = (float*)malloc(sizeof(float) * n);
    for (size_t k = 0; k < n; ++k) { lum[k] += offset; tmpI[k] = maxr(lum[k], 0.f); out[k] = NAN; }
#define GET_OUTPUT(k) ((tmpI[k] != tmpI[k]) ? lum[k] : ({ const float b_ = impulse[k] ? 0.f : blend[k] * amount; b_ * maxr(tmpI[k], 0.0f) + (1.f - b_) * lum[k]; }))
    for (int it = 0; it < maxiter; it++) {
        gauss_divmult(tmpI, tmp, lum, W, H, sigma, 0);
        gauss_divmult(tmp, tmpI, NULL, W, H, sigma, 1);
        for (size_t k = 0; k < n; ++k)
            if (out[k] != out[k]) {
                const float l = lum[k];
                const float delta = l * delta_factor;
                if (fabsf(tmpI[k] - l) > delta) out[k] = GET_OUTPUT(k);
            }
    }
    for (size_t k = 0; k < n; ++k) {
        float l = out[k];
        if (l != l) l = GET_OUTPUT(k);
        lum[k] = maxr(l - offset, 0.f);
    }
#undef GET_OUTPUT
    free(tmp); free(tmpI); free(out);
    return 0;
}

/* CornerBoostMask, ipsharpen.cc L313-338 */
static float corner_mask(int x, int y, int ox, int oy, int width, int height, int latitude)
{
    const int w2 = width / 2, h2 = height / 2;
    const float radius = (float)(w2 > h2 ? w2 : h2);
    const float lat = (float)latitude / 150.f;
    const float lim = maxr(0.f, minr(lat, 1.f));                                  /* LIM01 */
    const float r2 = (radius - radius * lim) / 2.f;
    const float sg = 2.f * ((radius * 0.3f) * (radius * 0.3f));
    const int xx = x + ox - w2, yy = y + oy - h2;
    const float distance = sqrtf((float)(xx * xx + yy * yy));
    const float d = maxr(distance - r2, 0.f);
    const float e = xexpf_scalar((-(d * d)) / sg);
    return 1.f - maxr(0.f, minr(e, 1.f));
}

int artoracle_rld_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double deconvradius,
                     int deconvamount, float* impulse_out, double deconvCornerBoost, int deconvCornerLatitude, int offset_x, int offset_y,
                     int full_width, int full_height)
{
    if (W < 8 || H < 8) return 0;
    const size_t n = (size_t)W * H;
    const float w0 = (float)wsd[3], w1 = (float)wsd[4], w2 = (float)wsd[5];
    float* Y = (float*)malloc(sizeof(float) * n); float* YY = (float*)malloc(sizeof(float) * n); float* blend = (float*)malloc(sizeof(float) * n);
    unsigned char* impulse = (unsigned char*)malloc(n);
    for (size_t k = 0; k < n; ++k) Y[k] = R[k] * w0 + G[k] * w1 + B[k] * w2;
    const float s_scale = (float)sqrt(scale);
    const float contrast = pow_F_scalar((float)(contrast_p / 100.f), 1.2f) * s_scale;
    int rc = artoracle_blend_mask(Y, blend, W, H, contrast, 1.f, 2.f / s_scale);
    if (!rc) rc = artoracle_mark_impulse(Y, impulse, W, H, 2.f);
    if (impulse_out) for (size_t k = 0; k < n; ++k) impulse_out[k] = impulse[k];
    memcpy(YY, Y, sizeof(float) * n);
    const double sigma = deconvradius / scale;
    const float amount = deconvamount / 100.f;
    const float delta = (float)(deconvCornerBoost / scale);
    if (delta > 0.01f) {       /* doSharpening L757-771 */
        float* YY2 = (float*)malloc(sizeof(float) * n);
        memcpy(YY2, Y, sizeof(float) * n);
        if (!rc) rc = deconv(YY, blend, impulse, W, H, sigma, amount);
        if (!rc) rc = deconv(YY2, blend, impulse, W, H, sigma + delta, amount);
        const int fw = full_width > 0 ? full_width : W, fh = full_height > 0 ? full_height : H;
        for (int y = 0; y < H && !rc; ++y)
            for (int x = 0; x < W; ++x) {
                const float m = corner_mask(x, y, offset_x, offset_y, fw, fh, deconvCornerLatitude);
                const size_t k = (size_t)y * W + x;
                YY[k] = m * YY2[k] + (1.f - m) * YY[k];
            }
        free(YY2);
    } else if (!rc) rc = deconv(YY, blend, impulse, W, H, sigma, amount);
    for (size_t k = 0; k < n && !rc; ++k)
        if (Y[k] > 0.f) {
            const float f = YY[k] / Y[k];
            R[k] *= f; G[k] *= f; B[k] *= f;
        }
    free(Y); free(YY); free(blend); free(impulse);
    return rc;
}

int artoracle_rld(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double deconvradius,
                  int deconvamount, float* impulse_out)
{
    return artoracle_rld_ex(R, G, B, W, H, wsd, scale, contrast_p, deconvradius, deconvamount, impulse_out, 0.0, 25, 0, 0, 0, 0);
}
