/*
 * oracle/fattal_port.c -- plain-C restatement of the reference's Fattal tone mapping.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; nothing under art_b200/ does.
 *
 * Follows reference rtengine/tmo_fattal02.cc: ToneMapFattal02 L1053-1215, tmo_fattal02 L421-681, downSample L157-177,
 * gaussianBlur L179-247, createGaussianPyramids L249-281, calculateGradients L285-320, upSample L324-341,
 * calculateFiMatrix L359-417, solve_pde_fft L869-950 with transform_ev2normal / transform_normal2ev L731-810 and
 * get_lambda L813-823, find_fast_dim L1014-1050; rtengine/FTblockDN.cc do_median_denoise L87-421 (one iteration, the
 * upper-bound form tone mapping uses); rtengine/rescale.h rescaleBilinear L27-77, rescaleNearest L80-106;
 * rtengine/color.h rgbLuminance L203-207.
 *
 * The two 2-D REDFT00 transforms are FFTW calls in the reference (fftw3f, absent from this image): they go through
 * oracle/dct_standin.h, a double-precision restatement of the published definition.  PARITY UNPINNED at that boundary;
 * everything else is pinned bit-exact against the reference functions compiled in place (tests/test_oracle_fattal.py).
 * pow() in calculateFiMatrix is the platform libm's powf, as it is for the reference.
 * Compile with -ffp-contract=off.
 */
#define _POSIX_C_SOURCE 200112L
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"
#include "dct_standin.h"

static inline float fmaxr(float a, float b) { return a < b ? b : a; }       /* std::max(a, b) */
static inline int imin(int a, int b) { return b < a ? b : a; }
static inline int imax(int a, int b) { return a < b ? b : a; }

/* ---- rescale.h ---- */
static void fat_rescale_bilinear(const float* src, int Ws, int Hs, float* dst, int Wd, int Hd)
{
    const float col_scale = (float)Ws / (float)Wd;
    const float row_scale = (float)Hs / (float)Hd;
#pragma omp parallel for
    for (int y = 0; y < Hd; ++y) {
        const float fy = y * row_scale;
        for (int x = 0; x < Wd; ++x) {
            const float fx = x * col_scale;
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

static void fat_rescale_nearest(const float* src, int sW, int sH, float* dst, int dW, int dH)
{
#pragma omp parallel for
    for (int y = 0; y < dH; ++y) {
        const int sy = y * sH / dH;
        for (int x = 0; x < dW; ++x) dst[(size_t)y * dW + x] = src[(size_t)sy * sW + x * sW / dW];
    }
}

/* ---- FTblockDN.cc do_median_denoise, iterations == 1 ---- */
static int cmpf(const void* a, const void* b) { const float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }
static float median_of(float* v, int n) { qsort(v, n, sizeof(float), cmpf); return v[n / 2]; }

int artoracle_median_denoise(const float* src, float* dst, float upperBound, int useUpper, int W, int H, int type)
{
    static const int border_of[6] = {1, 1, 2, 2, 3, 4};
    if (type < 0 || type > 5) return 1;
    const int border = border_of[type];
    float* out = (float*)malloc(sizeof(float) * (size_t)W * H);
    memcpy(out, src, sizeof(float) * (size_t)W * H);            /* borders and unfiltered samples are copies */
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = border; i < H - border; ++i) {
        float pp[81];
        for (int j = border; j < W - border; ++j) {
            if (useUpper && !(src[(size_t)i * W + j] <= upperBound)) continue;
            int n = 0;
            for (int ii = -border; ii <= border; ++ii)
                for (int jj = -border; jj <= border; ++jj) {
                    const int d = abs(ii) + abs(jj);
                    int take = 1;
                    if (type == 0) take = d <= 1;               /* 3x3 soft: the plus */
                    else if (type == 2) take = d <= 2;          /* 5x5 soft: the 13-point diamond */
                    if (take) pp[n++] = src[(size_t)(i + ii) * W + j + jj];
                }
            out[(size_t)i * W + j] = median_of(pp, n);
        }
    }
    memcpy(dst, out, sizeof(float) * (size_t)W * H);
    free(out);
    return 0;
}

/* ---- tmo_fattal02.cc ---- */
static void fat_downsample(const float* A, int aw, float* B, int width, int height)
{
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            float p = A[(size_t)(2 * y) * aw + 2 * x];
            p += A[(size_t)(2 * y) * aw + 2 * x + 1];
            p += A[(size_t)(2 * y + 1) * aw + 2 * x];
            p += A[(size_t)(2 * y + 1) * aw + 2 * x + 1];
            B[(size_t)y * width + x] = p * 0.25f;
        }
}

static void fat_gaussian_blur(const float* I, float* L, int width, int height)
{
    if (width < 3 || height < 3) {
        if (I != L) memcpy(L, I, sizeof(float) * (size_t)width * height);
        return;
    }
    float* T = (float*)malloc(sizeof(float) * (size_t)width * height);
    for (int y = 0; y < height; ++y) {
        const float* r = I + (size_t)y * width;
        float* t = T + (size_t)y * width;
        for (int x = 1; x < width - 1; ++x) {
            float v = 2.f * r[x];
            v += r[x - 1];
            v += r[x + 1];
            t[x] = v * 0.25f;
        }
        t[0] = (3.f * r[0] + r[1]) * 0.25f;
        t[width - 1] = (3.f * r[width - 1] + r[width - 2]) * 0.25f;
    }
    for (int x = 0; x < width; ++x) {
        for (int y = 1; y < height - 1; ++y) {
            float v = 2.f * T[(size_t)y * width + x];
            v += T[(size_t)(y - 1) * width + x];
            v += T[(size_t)(y + 1) * width + x];
            L[(size_t)y * width + x] = v * 0.25f;
        }
        L[x] = (3.f * T[x] + T[(size_t)width + x]) * 0.25f;
        L[(size_t)(height - 1) * width + x] = (3.f * T[(size_t)(height - 1) * width + x] + T[(size_t)(height - 2) * width + x]) * 0.25f;
    }
    free(T);
}

#define NLEVELS 7

static float fat_gradients(const float* Hm, float* G, int width, int height, int k)
{
    const float divider = (float)pow(2.0, k + 1);
    double avg = 0.0;
    for (int y = 0; y < height; ++y) {
        const int n = (y == 0 ? 0 : y - 1), s = (y + 1 == height ? y : y + 1);
        for (int x = 0; x < width; ++x) {
            const int w = (x == 0 ? 0 : x - 1), e = (x + 1 == width ? x : x + 1);
            const float gx = Hm[(size_t)y * width + w] - Hm[(size_t)y * width + e];
            const float gy = Hm[(size_t)s * width + x] - Hm[(size_t)n * width + x];
            const float g = sqrtf(gx * gx + gy * gy) / divider;
            G[(size_t)y * width + x] = g;
            avg += g;
        }
    }
    return (float)(avg / (width * height));
}

static void fat_upsample(const float* A, int aw, int ah, float* B, int width, int height)
{
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            int ax = (int)(x * 0.5f), ay = (int)(y * 0.5f);
            ax = ax < aw ? ax : aw - 1;
            ay = ay < ah ? ay : ah - 1;
            B[(size_t)y * width + x] = A[(size_t)ay * aw + ax];
        }
}

/* the solver: F (n0 rows x n1 cols) -> U, buf scratch; solve_pde_fft L869-950 */
static void fat_solve_pde(float* F, float* U, float* buf, int width, int height)
{
    /* transform_normal2ev(F, buf) */
    artdct_redft00_2d(height, width, F, buf);
    const float factor = 1.0f / ((height - 1) * (width - 1));
    for (size_t i = 0; i < (size_t)width * height; ++i) buf[i] *= factor;
    for (int x = 0; x < width; ++x) { buf[x] *= 0.5f; buf[(size_t)(height - 1) * width + x] *= 0.5f; }
    for (int y = 0; y < height; ++y) { buf[(size_t)y * width] *= 0.5f; buf[(size_t)y * width + width - 1] *= 0.5f; }
    /* eigenvalues, get_lambda */
    double* l1 = (double*)malloc(sizeof(double) * height);
    double* l2 = (double*)malloc(sizeof(double) * width);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < height; ++i) { const double s = sin((double)i / (2 * (height - 1)) * pi); l1[i] = -4.0 * (s * s); }
    for (int i = 0; i < width; ++i) { const double s = sin((double)i / (2 * (width - 1)) * pi); l2[i] = -4.0 * (s * s); }
#pragma omp parallel for
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) buf[(size_t)y * width + x] = (float)(buf[(size_t)y * width + x] / (l1[y] + l2[x]));
    buf[0] = 0.f;
    free(l1); free(l2);
    /* transform_ev2normal(buf, U) */
    for (int y = 1; y < height - 1; ++y)
        for (int x = 1; x < width - 1; ++x) buf[(size_t)y * width + x] *= 0.25f;
    for (int x = 1; x < width - 1; ++x) { buf[x] *= 0.5f; buf[(size_t)(height - 1) * width + x] *= 0.5f; }
    for (int y = 1; y < height - 1; ++y) { buf[(size_t)y * width] *= 0.5f; buf[(size_t)y * width + width - 1] *= 0.5f; }
    artdct_redft00_2d(height, width, buf, U);
}

/* tmo_fattal02 L421-681: Y (width x height, contiguous) -> L; Y and L may be the same array, as at the call site */
int artoracle_tmo_fattal02(int width, int height, const float* Y, float* L, float alfa, float beta, float noise, int detail_level)
{
    if (detail_level < 0) detail_level = 0;
    if (detail_level > 3) detail_level = 3;
    const int fullwidth = width, fullheight = height;
    float* H = (float*)malloc(sizeof(float) * (size_t)width * height);
    const float eps = 1e-4f;
#pragma omp parallel for
    for (int i = 0; i < height; ++i) {
        int j = 0;
        for (; j < width - 3; j += 4)
            for (int q = 0; q < 4; ++q) H[(size_t)i * width + j + q] = xlogf_vector(Y[(size_t)i * width + j + q] + eps);
        for (; j < width; ++j) H[(size_t)i * width + j] = xlogf_scalar(Y[(size_t)i * width + j] + eps);
    }
    float* fullH = NULL;
    const int dim = imax(width, height);
    if (dim > 1920) {
        const float s = 1920.f / (float)dim;
        const int w = (int)((float)(size_t)width * s), h = (int)((float)(size_t)height * s);
        float* HH = (float*)malloc(sizeof(float) * (size_t)w * h);
        fat_rescale_bilinear(H, width, height, HH, w, h);
        fullH = H; H = HH; width = w; height = h;
    }
    /* pyramids */
    float* pyr[NLEVELS]; int pw[NLEVELS], ph[NLEVELS];
    pyr[0] = H; pw[0] = width; ph[0] = height;
    {
        int w = width, h = height;
        float* Lb = (float*)malloc(sizeof(float) * (size_t)w * h);
        fat_gaussian_blur(pyr[0], Lb, w, h);
        for (int k = 1; k < NLEVELS; ++k) {
            if (w > 2 && h > 2) {
                const int aw = w;
                w /= 2; h /= 2;
                pyr[k] = (float*)malloc(sizeof(float) * (size_t)w * h);
                fat_downsample(Lb, aw, pyr[k], w, h);
            } else {
                pyr[k] = (float*)malloc(sizeof(float) * (size_t)w * h);
                memcpy(pyr[k], Lb, sizeof(float) * (size_t)w * h);
            }
            pw[k] = w; ph[k] = h;
            if (k < NLEVELS - 1) {
                free(Lb);
                Lb = (float*)malloc(sizeof(float) * (size_t)w * h);
                fat_gaussian_blur(pyr[k], Lb, w, h);
            }
        }
        free(Lb);
    }
    float* grad[NLEVELS]; float avg[NLEVELS];
    for (int k = 0; k < NLEVELS; ++k) {
        grad[k] = (float*)malloc(sizeof(float) * (size_t)pw[k] * ph[k]);
        avg[k] = fat_gradients(pyr[k], grad[k], pw[k], ph[k], k);
        if (k != 0) free(pyr[k]);
    }
    /* calculateFiMatrix */
    float* FI = (float*)malloc(sizeof(float) * (size_t)width * height);
    {
        float* fi[NLEVELS];
        fi[NLEVELS - 1] = (float*)malloc(sizeof(float) * (size_t)pw[NLEVELS - 1] * ph[NLEVELS - 1]);
        for (size_t i = 0; i < (size_t)pw[NLEVELS - 1] * ph[NLEVELS - 1]; ++i) fi[NLEVELS - 1][i] = 1.0f;
        for (int k = NLEVELS - 1; k >= 0; --k) {
            const int w = pw[k], h = ph[k];
            if ((k >= detail_level || k == NLEVELS - 1) && beta != 1.f) {
                const float a = alfa * avg[k];
                for (size_t i = 0; i < (size_t)w * h; ++i) {
                    const float g = (grad[k][i] < 1e-4f) ? (float)1e-4 : grad[k][i];
                    const float value = powf((g + noise) / a, beta - 1.0f);
                    fi[k][i] *= value;
                }
            }
            if (k > 1) fi[k - 1] = (float*)malloc(sizeof(float) * (size_t)pw[k - 1] * ph[k - 1]);
            else fi[0] = FI;
            if (k > 0) {
                fat_upsample(fi[k], w, h, fi[k - 1], pw[k - 1], ph[k - 1]);
                fat_gaussian_blur(fi[k - 1], fi[k - 1], pw[k - 1], ph[k - 1]);
            }
        }
        for (int k = 1; k < NLEVELS; ++k) free(fi[k]);
    }
    for (int k = 0; k < NLEVELS; ++k) free(grad[k]);
    if (fullH) {
        free(H);
        H = fullH;
        float* FI2 = (float*)malloc(sizeof(float) * (size_t)fullwidth * fullheight);
        fat_rescale_bilinear(FI, width, height, FI2, fullwidth, fullheight);
        free(FI); FI = FI2;
        width = fullwidth; height = fullheight;
    }
    /* attenuated gradients, L doubles as Gy */
    float* Gx = (float*)malloc(sizeof(float) * (size_t)width * height);
    float* Gy = L;
#pragma omp parallel for
    for (int y = 0; y < height; ++y) {
        const int yp1 = (y + 1 >= height ? height - 2 : y + 1);
        for (int x = 0; x < width; ++x) {
            const int xp1 = (x + 1 >= width ? width - 2 : x + 1);
            const size_t c = (size_t)y * width + x;
            Gx[c] = (float)((H[(size_t)y * width + xp1] - H[c]) * 0.5 * (FI[(size_t)y * width + xp1] + FI[c]));
            Gy[c] = (float)((H[(size_t)yp1 * width + x] - H[c]) * 0.5 * (FI[(size_t)yp1 * width + x] + FI[c]));
        }
    }
    free(H);
    /* divergence into FI */
#pragma omp parallel for
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const size_t c = (size_t)y * width + x;
            float v = Gx[c] + Gy[c];
            if (x > 0) v -= Gx[c - 1];
            if (y > 0) v -= Gy[c - width];
            if (x == 0) v += Gx[c];
            if (y == 0) v += Gy[c];
            FI[c] = v;
        }
    fat_solve_pde(FI, L, Gx, width, height);
    free(Gx); free(FI);
#pragma omp parallel for
    for (int i = 0; i < height; ++i) {
        int j = 0;
        for (; j < width - 3; j += 4)
            for (int q = 0; q < 4; ++q) L[(size_t)i * width + j + q] = xexpf_vector(L[(size_t)i * width + j + q]);
        for (; j < width; ++j) L[(size_t)i * width + j] = xexpf_scalar(L[(size_t)i * width + j]);
    }
    return 0;
}

/* find_fast_dim L1014-1050 (round_up_pow2 L998-1012) */
int artoracle_find_fast_dim(int dim)
{
    unsigned v = (unsigned)dim;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    const int d1 = (int)v;
    const int d[12] = {d1 / 128 * 65, d1 / 64 * 33, d1 / 512 * 273, d1 / 16 * 9, d1 / 8 * 5, d1 / 16 * 11,
                       d1 / 128 * 91, d1 / 4 * 3, d1 / 64 * 49, d1 / 16 * 13, d1 / 8 * 7, d1};
    for (int i = 0; i < 12; ++i) if (d[i] >= dim) return d[i];
    return dim;
}

static inline float fat_luminance(float r, float g, float b, const double* ws)
{
    /* Color::rgbLuminance(r, g, b, TMatrix), color.h L203-207; TMatrix is const float (*)[3] (iccstore.h L38): float arithmetic */
    const float w0 = (float)ws[3], w1 = (float)ws[4], w2 = (float)ws[5];
    return r * w0 + g * w1 + b * w2;
}

/* ToneMapFattal02 L1053-1215; planes contiguous W x H, in place; ws = working-space matrix, row-major 3x3 */
int artoracle_fattal(float* R, float* G, float* B, int w, int h, int threshold, int amount, int satcontrol, const double* ws)
{
    const int detail_level = 3;
    float alpha = 1.f;
    if (threshold < 0) alpha += (threshold * 0.9f) / 100.f;
    else if (threshold > 0) alpha += threshold / 100.f;
    const float beta = 1.f - (amount * 0.3f) / 100.f;
    if (alpha <= 0 || beta <= 0) return 0;
    const size_t n = (size_t)w * h;
    float* Yr = (float*)malloc(sizeof(float) * n);
    const float epsilon = 1e-4f, luminance_noise_floor = 65.535f, min_luminance = 1.f;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) Yr[i] = fmaxr(fat_luminance(R[i], G[i], B[i], ws), min_luminance);
    const int w2 = artoracle_find_fast_dim(w) + 1, h2 = artoracle_find_fast_dim(h) + 1;
    float* L = (float*)malloc(sizeof(float) * (size_t)w2 * h2);
    {
        const float r = (float)imax(w, h) / 1920.f;
        const int med = r >= 3 ? 4 : r >= 2 ? 3 : r >= 1 ? 2 : 1;
        artoracle_median_denoise(Yr, Yr, luminance_noise_floor, 1, w, h, med);
    }
    const float noise = alpha * 0.01f;
    fat_rescale_nearest(Yr, w, h, L, w2, h2);
    artoracle_tmo_fattal02(w2, h2, L, L, alpha, beta, noise, detail_level);
    const float hr = (float)h2 / (float)h, wr = (float)w2 / (float)w;
    float scale = 65535.f, offset = 0.f;
    {
        float ratio; int ww, hh;
        if (w >= h) { ratio = 200.f / w; ww = 200; hh = (int)(ratio * h); }
        else { ratio = 200.f / h; hh = 200; ww = (int)(ratio * w); }
        const int sz = ww * hh, idx = sz / 2;
        const int oidx = imax(1, imin((int)(sz * 0.05f + 0.5f), sz - 1));
        float* tmp = (float*)malloc(sizeof(float) * (size_t)(sz > 0 ? sz : 1));
        fat_rescale_nearest(Yr, w, h, tmp, ww, hh);
        qsort(tmp, sz, sizeof(float), cmpf);
        const float oldMedian = tmp[idx];
        float old_min = 0.f;
        for (int i = 0; i <= oidx; ++i) old_min += tmp[i];
        old_min /= oidx;
        fat_rescale_nearest(L, w2, h2, tmp, ww, hh);
        qsort(tmp, sz, sizeof(float), cmpf);
        const float newMedian = tmp[idx];
        scale = (oldMedian == 0.f || newMedian == 0.f) ? 65535.f : (oldMedian / newMedian);
        float new_min = 0.f;
        for (int i = 0; i <= oidx; ++i) new_min += tmp[i];
        new_min /= oidx;
        offset = old_min - new_min;
        free(tmp);
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int y = 0; y < h; ++y) {
        const int yy = imin((int)(y * hr + 1), h2 - 1);
        for (int x = 0; x < w; ++x) {
            const int xx = imin((int)(x * wr + 1), w2 - 1);
            const size_t c = (size_t)y * w + x;
            const float Y = fmaxr(Yr[c], epsilon);
            const float l = fmaxr(L[(size_t)yy * w2 + xx], epsilon) * (scale / Y);
            float r = R[c], g = G[c], b = B[c], s = 1.f;
            if (l > 1.f) {
                r = fmaxr(r * l - offset, r);
                g = fmaxr(g * l - offset, g);
                b = fmaxr(b * l - offset, b);
                if (satcontrol) s = pow_F_scalar(1.f / l, 0.3f);
            } else {
                r *= l; g *= l; b *= l;
                if (satcontrol) s = pow_F_scalar(l, 0.3f);
            }
            if (satcontrol && s != 1.f) {
                const float ll = fat_luminance(r, g, b, ws);
                const float rl = r - ll, gl = g - ll, bl = b - ll;
                r = ll + s * rl; g = ll + s * gl; b = ll + s * bl;
            }
            R[c] = r; G[c] = g; B[c] = b;
        }
    }
    free(Yr); free(L);
    return 0;
}

/* exported for tests of the stand-in and of the GPU transform */
void artoracle_redft00_2d(int n0, int n1, const float* in, float* out) { artdct_redft00_2d(n0, n1, in, out); }
