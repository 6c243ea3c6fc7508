/*
 * oracle/nlmeans_port.c -- CPU restatement of the reference's NL-means smoothing.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates (reference) rtengine/nlmeans.cc NLMeans L50-280 (SSE2 build: 4-wide groups use the vector LUT lookup of
 * rtengine/LUT.h L349-377, the trailing samples of each tile row the scalar lookup L437-459; the tile loop runs with
 * the MXCSR flush-to-zero bit set, L158-159), rtengine/FTblockDN.cc laplacian L1366-1403 and detail_mask L1408-1476
 * (BlurType::GAUSS and BOX), rtengine/rescale.h rescaleBilinear L27-77.
 * Pinned bit-exact against those functions compiled in place (oracle/_ref) in tests/test_oracle_nlmeans.py.
 * Compile with -ffp-contract=off.
 */
#include <stdlib.h>
#include <xmmintrin.h>
#include "sleef_port.h"

int artoracle_gauss(const float* src, long ss, float* dst, long ds, int W, int H, double sigma);
int artoracle_boxblur(const float* src, long ss, float* dst, long ds, int W, int H, int radius);

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int ilim(int v, int lo, int hi) { return imax(lo, imin(v, hi)); }
static inline float fmaxr(float a, float b) { return a < b ? b : a; }       /* std::max(a, b) */
static inline float fminr(float a, float b) { return b < a ? b : a; }       /* std::min(a, b) */
static inline float flim(float v, float lo, float hi) { return fmaxr(lo, fminr(v, hi)); }   /* rt_math.h LIM */

/* rescale.h L27-77 */
static void rescale_bilinear(const float* src, int Ws, int Hs, float* dst, int Wd, int Hd)
{
    const float col_scale = (float)Ws / (float)Wd;
    const float row_scale = (float)Hs / (float)Hd;
    for (int y = 0; y < Hd; ++y) {
        const float ymrs = y * row_scale;
        for (int x = 0; x < Wd; ++x) {
            const float fx = x * col_scale, fy = ymrs;
            const int xi = imin((int)fx, Ws - 1), yi = imin((int)fy, Hs - 1);
            const float xf = fx - xi, yf = fy - yi;
            const int xi1 = imin(xi + 1, Ws - 1), yi1 = imin(yi + 1, Hs - 1);
            const float bl = src[(size_t)yi * Ws + xi], br = src[(size_t)yi * Ws + xi1];
            const float tl = src[(size_t)yi1 * Ws + xi], tr = src[(size_t)yi1 * Ws + xi1];
            const float b = xf * br + (1.f - xf) * bl;
            const float t = xf * tr + (1.f - xf) * tl;
            dst[(size_t)y * Wd + x] = yf * t + (1.f - yf) * b;
        }
    }
}

/* FTblockDN.cc L1366-1403 */
static void laplacian(const float* src, float* dst, int W, int H, float threshold, float ceiling, float factor)
{
    const float f = factor / ceiling;
#define G(y, x) fmaxr(src[(size_t)(y) * W + (x)], 0.f)
    for (int y = 0; y < H; ++y) {
        const int n = (y - 1 < 0) ? y + 1 : y - 1, s = (y + 1 >= H) ? y - 1 : y + 1;
        for (int x = 0; x < W; ++x) {
            const int w = (x - 1 < 0) ? x + 1 : x - 1, e = (x + 1 >= W) ? x - 1 : x + 1;
            const float v = -8.f * G(y, x) + G(n, x) + G(s, x) + G(y, w) + G(y, e) + G(n, w) + G(n, e) + G(s, w) + G(s, e);
            dst[(size_t)y * W + x] = flim(fabsf(v) - threshold, 0.f, ceiling) * f;
        }
    }
#undef G
}

/* FTblockDN.cc L1408-1476.  blur_type: 0 off, 1 box, 2 gauss */
int artoracle_detail_mask(const float* src, float* mask, int W, int H, float scaling, float threshold, float ceiling, float factor,
                          int blur_type, float blur)
{
    if (W < 8 || H < 8) {
        for (size_t i = 0; i < (size_t)W * H; ++i) mask[i] = 1.f;
        return 0;
    }
    const int W4 = W / 4, H4 = H / 4;
    float* L2 = (float*)malloc(sizeof(float) * (size_t)W4 * H4);
    float* m2 = (float*)malloc(sizeof(float) * (size_t)W4 * H4);
    if (!L2 || !m2) { free(L2); free(m2); return 1; }
    rescale_bilinear(src, W, H, L2, W4, H4);
    for (size_t i = 0; i < (size_t)W4 * H4; ++i) L2[i] = xlin2log_scalar(L2[i] / scaling, 50.f);
    laplacian(L2, m2, W4, H4, threshold / scaling, ceiling / scaling, factor);
    rescale_bilinear(m2, W4, H4, mask, W, H);
    const float thr = 1.f - factor;
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        const float x = flim(mask[i] + thr, 0.f, 1.f);
        mask[i] = xlin2log_scalar(pow_F_scalar(x, 2.23f), 101.f);
    }
    free(L2); free(m2);
    if (blur_type == 2) return artoracle_gauss(mask, W, mask, W, W, H, (double)blur);
    if (blur_type == 1 && (int)blur > 0)
        for (int i = 0; i < 3; ++i) { int rc = artoracle_boxblur(mask, W, mask, W, W, H, (int)blur); if (rc) return rc; }
    return 0;
}

/* LUT.h: LUTf(8192), default clip flags (below and above) */
#define LUTSZ 8192
static inline float lut_scalar(const float* data, float index)
{   /* L437-459 */
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(LUTSZ - 2)) return data[LUTSZ - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
static inline float vclamp(float v, float lo, float hi) { return fminr(fmaxr(v, lo), hi); }    /* NaN -> lo, not reached */
static inline float lut_vector(const float* data, float index)
{   /* L349-377 */
    const int idx = (int)vclamp(index, 0.f, (float)(LUTSZ - 2));
    const float lower = data[idx], upper = data[idx + 1];
    const float diff = vclamp(index, 0.f, (float)(LUTSZ - 1)) - (float)idx;
    return diff * upper + (1.f - diff) * lower;          /* vintpf */
}

/* nlmeans.cc L50-280 */
int artoracle_nlmeans(float* img, int W, int H, float normcoeff, int strength, int detail_thresh, float scale)
{
    if (!strength) return 0;
    const int search_radius = (int)ceilf(5.f / scale);
    const int patch_radius = (int)ceilf(2.f / scale);
    const float ph = powf((float)strength / 100.f, 0.9f) / 10.f / scale;
    const float h2 = ph * ph;
    const float amount = flim((float)detail_thresh / 100.f, 0.f, 0.99f);
    float* mask = (float*)malloc(sizeof(float) * (size_t)W * H);
    if (!mask) return 1;
    int rc = artoracle_detail_mask(img, mask, W, H, normcoeff, 1e-3f * normcoeff, normcoeff, amount, 2, 2.f / scale);
    if (rc) { free(mask); return rc; }

    const int border = search_radius + patch_radius;
    const int WW = W + border * 2, HH = H + border * 2;
    const float factor = normcoeff;
    float* src = (float*)malloc(sizeof(float) * (size_t)WW * HH);
    if (!src) { free(mask); return 1; }
    for (int y = 0; y < HH; ++y) {
        const int yy = y <= border ? 0 : y >= H ? H - 1 : y - border;          /* sic, L102-109 */
        for (int x = 0; x < WW; ++x) {
            const int xx = x <= border ? 0 : x >= W ? W - 1 : x - border;
            src[(size_t)y * WW + x] = img[(size_t)yy * W + xx] / factor;
        }
    }
    memset(img, 0, sizeof(float) * (size_t)W * H);
    float* dst = img;

    const float lutfactor = 100.f / (float)(LUTSZ - 1);
    float* explut = (float*)malloc(sizeof(float) * LUTSZ);
    for (int i = 0; i < LUTSZ; ++i) explut[i] = xexpf_scalar(-((float)i * lutfactor));
    for (size_t i = 0; i < (size_t)W * H; ++i) mask[i] = (1.f / (mask[i] * h2)) / lutfactor;

    const int tile_size = 150;
    const int step = tile_size - 2 * border;
    const int ntiles_x = (int)ceilf((float)WW / step);
    const int ntiles_y = (int)ceilf((float)HH / step);
    const int ntiles = ntiles_x * ntiles_y;
    int failed = 0;

#pragma omp parallel
    {
        const unsigned old = _MM_GET_FLUSH_ZERO_MODE();
        _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
        float* St = (float*)malloc(sizeof(float) * tile_size * tile_size);
        float* SW = (float*)malloc(sizeof(float) * tile_size * tile_size);
        if (!St || !SW) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp barrier
#pragma omp for schedule(dynamic, 2)
        for (int tile = 0; tile < ntiles; ++tile) {
            if (failed) continue;
            const int tile_y = tile / ntiles_x, tile_x = tile % ntiles_x;
            const int start_y = tile_y * step, end_y = imin(start_y + tile_size, HH), TH = end_y - start_y;
            const int start_x = tile_x * step, end_x = imin(start_x + tile_size, WW), TW = end_x - start_x;
#define SRC(y, x) src[(size_t)(y) * WW + (x)]
#define YC(y) ilim((y) + start_y, 0, HH - 1)
#define XC(x) ilim((x) + start_x, 0, WW - 1)
#define ST(y, x) St[(y) * TW + (x)]
            memset(SW, 0, sizeof(float) * (size_t)TW * TH);
            for (int ty = -search_radius; ty <= search_radius; ++ty) {
                for (int tx = -search_radius; tx <= search_radius; ++tx) {
                    ST(0, 0) = 0.f;
                    for (int xx = 1; xx < TW; ++xx) { float d = SRC(YC(0), XC(xx)) - SRC(YC(ty), XC(xx + tx)); ST(0, xx) = ST(0, xx - 1) + d * d; }
                    for (int yy = 1; yy < TH; ++yy) { float d = SRC(YC(yy), XC(0)) - SRC(YC(yy + ty), XC(tx)); ST(yy, 0) = ST(yy - 1, 0) + d * d; }
                    for (int yy = 1; yy < TH; ++yy)
                        for (int xx = 1; xx < TW; ++xx) {
                            float d = SRC(YC(yy), XC(xx)) - SRC(YC(yy + ty), XC(xx + tx));
                            ST(yy, xx) = (ST(yy, xx - 1) + ST(yy - 1, xx)) - (ST(yy - 1, xx - 1) - d * d);
                        }
                    for (int yy = start_y + border; yy < end_y - border; ++yy) {
                        const int y = yy - border;
                        for (int xx = start_x + border; xx < end_x - border; ++xx) {
                            const int vec = ((xx - (start_x + border)) & ~3) + (start_x + border) < end_x - border - 3;
                            const int x = xx - border, sx = xx + tx, sy = yy + ty, sty = yy - start_y, stx = xx - start_x;
                            float dist2 = ST(sty + patch_radius, stx + patch_radius) + ST(sty - patch_radius, stx - patch_radius)
                                          - ST(sty + patch_radius, stx - patch_radius) - ST(sty - patch_radius, stx + patch_radius);
                            dist2 = fmaxr(dist2, 0.f);
                            const float d = dist2 * mask[(size_t)y * W + x];
                            const float weight = vec ? lut_vector(explut, d) : lut_scalar(explut, d);
                            SW[(y - start_y) * TW + (x - start_x)] += weight;
                            const float Yv = weight * SRC(sy, sx);
                            dst[(size_t)y * W + x] += Yv;
                        }
                    }
                }
            }
            for (int yy = start_y + border; yy < end_y - border; ++yy) {
                const int y = yy - border;
                for (int xx = start_x + border; xx < end_x - border; ++xx) {
                    const int x = xx - border;
                    const float Yv = dst[(size_t)y * W + x];
                    const float f = 1e-5f + SW[(y - start_y) * TW + (x - start_x)];
                    dst[(size_t)y * W + x] = (Yv / f) * factor;
                }
            }
        }
        free(St); free(SW);
        _MM_SET_FLUSH_ZERO_MODE(old);
    }
    free(src); free(mask); free(explut);
    return failed;
}
