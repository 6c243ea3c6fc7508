#define _POSIX_C_SOURCE 200112L
/*
 * oracle/denoise_port.c -- CPU restatement of denoise::RGB_denoise for the path ART actually takes.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates (reference) rtengine/FTblockDN.cc RGB_denoise L1638-2689 with kall = 0, isRAW = true, colorSpace RGB,
 * aggressive = false (QUALITY_STANDARD), chrominanceMethod MANUAL, no luminance noise curve, optional chrominance
 * noise curve (the driver ipdenoise.cc L1140-1150 always sets one), Tile_calc L442-478 (always one tile), and the
 * pieces it calls: Color::gammaf2lut (color.cc L1128-1170, SSE2 build), Color::gammaf (color.h L1202-1205),
 * rgb2yuv / yuv2rgb (color.h L782-796), rgbxyz (color.cc L833-838), XYZ2Lab (color.cc L1247-1274, L1382-1399),
 * Noise_residualAB L607-635, detail_recovery L1479-1635 run by one thread (its overlap-add races otherwise),
 * RGBtile_denoise L494-525, RGBoutput_tile_row L531-558, boxabsblur (boxblur.h L745-888).
 * The block DCT is FFTW's in the reference; here it is oracle/dct_standin.h (parity unpinned at that boundary).
 * Pinned against the reference function compiled in place over the same stand-in (oracle/_ref) in
 * tests/test_oracle_denoise.py: bit-exact.
 * Compile with -ffp-contract=off.
 */
#include <stdio.h>
#include "sleef_port.h"
#include "dct_standin.h"

void* artoracle_wavelet_new(const float* src, int W, int H, int maxlvl, int subsamp);
int artoracle_wavelet_maxlevel(void* p);
int artoracle_wavelet_level_W(void* p, int l);
int artoracle_wavelet_level_H(void* p, int l);
float* artoracle_wavelet_band(void* p, int l, int dir);
void artoracle_wavelet_reconstruct(void* p, float* dst, float blend);
void artoracle_wavelet_delete(void* p);
float artoracle_madrgb(const float* data, int n);
int artoracle_wavelet_denoise_L(void* wL, const float* noisevarlum, const float* madL, double scale);
int artoracle_wavelet_denoise_AB(void* wL, void* wab, const float* noisevarchrom, const float* madL, float noisevar_ab,
                                 int useNoiseCCurve, int autoch, double scale);
int artoracle_detail_mask(const float* src, float* mask, int W, int H, float scaling, float threshold, float ceiling, float factor,
                          int blur_type, float blur);

#define TS 64
#define OFFSET 25
#define BLKRAD 1
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline float sqrf(float x) { return x * x; }

/* LUTf(size, LUT_CLIP_BELOW)::operator[](float), LUT.h L437-459 */
static inline float lut_clip_below(const float* data, int size, float index)
{
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) idx = size - 2;
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
/* LUTf(size) with both clips (the NoiseCurve's LUT) */
static inline float lut_clip_both(const float* data, int size, float index)
{
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) return data[size - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}

/* Color::gammaf2lut, SSE2 build */
void artoracle_gammaf2lut(float* lut, float gamma, float start, float slope, float divisor, float factor)
{
    const float gammav = 1.f / gamma;
    const float slopev = (slope / divisor) * factor;
    const float divisorv = xlogf_scalar(divisor);
    const float comparev = start * divisor;
    const int border = (int)(start * divisor);
    const int border1 = border - (border & 3);
    const int border2 = border1 + 4;
    int i = 0;
    for (; i < border1; ++i) lut[i] = (float)i * slopev;
    for (; i < border2; ++i) {
        const float iv = (float)i;
        const float r0 = iv * slopev;
        const float r1 = xexpf_vector((xlogf_vector(iv) - divisorv) * gammav) * factor;
        lut[i] = iv <= comparev ? r0 : r1;
    }
    for (; i < 65536; ++i) lut[i] = xexpf_nocheck((xlogf_nocheck((float)i) - divisorv) * gammav) * factor;
}
static inline float gammaf_(float x, float gamma, float start, float slope)
{
    return x <= start ? x * slope : xexpf_scalar(xlogf_scalar(x) / gamma);
}

/* Color::cachef / cachefy (color.cc L205-233) and computeXYZ2Lab / computeXYZ2LabY (L1247-1274) */
static float* g_cachef = NULL; static float* g_cachefy = NULL;
static void init_cachef(void)
{
    if (g_cachef) return;
    g_cachef = (float*)malloc(sizeof(float) * 65536); g_cachefy = (float*)malloc(sizeof(float) * 65536);
    const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
    const int epsmaxint = (int)(MAXVALF * eps);
    int i = 0;
    for (; i <= epsmaxint; i++) { g_cachef[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); g_cachefy[i] = (float)(327.68 * (kappa * i / MAXVALF)); }
    for (; i < 65536; i++) { g_cachef[i] = (float)(327.68 * cbrt((double)i / MAXVALF)); g_cachefy[i] = (float)(327.68 * (116.0 * cbrt((double)i / MAXVALF) - 16.0)); }
}
const float* artoracle_cachef(int which) { init_cachef(); return which ? g_cachefy : g_cachef; }
static float computeXYZ2Lab(float f)
{
    const double kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * ((kappa * f / MAXVALF + 16.0) / 116.0));
    else if (f > 65535.f) return 327.68f * xcbrtf_scalar(f / 65535.f);
    return lut_clip_below(g_cachef, 65536, f);
}
static float computeXYZ2LabY(float f)
{
    const double kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * (kappa * f / MAXVALF));
    else if (f > 65535.f) return 327.68f * (116.f * xcbrtf_scalar(f / 65535.f) - 16.f);
    return lut_clip_below(g_cachefy, 65536, f);
}

/* boxabsblur(src, dst, 3, 3, 64, 64, temp), boxblur.h L745-888 (W % 4 == 0: all columns in the vector class) */
static void boxabsblur64(const float* src, float* dst, int rad, float* temp)
{
    const int W = TS, H = TS;
    for (int row = 0; row < H; row++) {
        int len = rad + 1;
        float tempval = fabsf(src[row * W + 0]);
        for (int j = 1; j <= rad; j++) tempval += fabsf(src[row * W + j]);
        tempval /= len;
        temp[row * W + 0] = tempval;
        for (int col = 1; col <= rad; col++) {
            tempval = (tempval * len + fabsf(src[row * W + col + rad])) / (len + 1);
            temp[row * W + col] = tempval;
            len++;
        }
        const float rlen = 1.f / (float)len;
        for (int col = rad + 1; col < W - rad; col++) {
            tempval = tempval + ((float)(fabsf(src[row * W + col + rad]) - fabsf(src[row * W + col - rad - 1]))) * rlen;
            temp[row * W + col] = tempval;
        }
        for (int col = W - rad; col < W; col++) {
            tempval = (tempval * len - fabsf(src[row * W + col - rad - 1])) / (len - 1);
            temp[row * W + col] = tempval;
            len--;
        }
    }
    for (int col = 0; col < W; col++) {
        float len = (float)(rad + 1);
        float t = temp[col];
        for (int i = 1; i <= rad; i++) t = t + temp[i * W + col];
        t = t / len;
        dst[col] = t;
        for (int row = 1; row <= rad; row++) {
            const float lp1 = len + 1.f;
            t = (t * len + temp[(row + rad) * W + col]) / lp1;
            dst[row * W + col] = t;
            len = lp1;
        }
        const float rlen = 1.f / len;
        for (int row = rad + 1; row < H - rad; row++) {
            t = t + (temp[(row + rad) * W + col] - temp[(row - rad - 1) * W + col]) * rlen;
            dst[row * W + col] = t;
        }
        for (int row = H - rad; row < H; row++) {
            const float lm1 = len - 1.f;
            t = (t * len - temp[(row - rad - 1) * W + col]) / lm1;
            dst[row * W + col] = t;
            len = lm1;
        }
    }
}

static float compute_detail(float d)
{   /* L1481-1485 */
    const float a = (float)(((100. - d) * (100. - d)) + 50. * (100. - d)) * TS * 0.5f;
    return a * a;
}

/* tilemask_in / tilemask_out, L1833-1849 */
void artoracle_tilemasks(float* tin, float* tout)
{
    const float epsilon = 0.001f / (TS * TS);
    const int border = imax(2, TS / 16);
    const double RT_PI = 3.14159265358979323846;
    for (int i = 0; i < TS; ++i) {
        const float i1 = (float)abs((i > TS / 2 ? i - TS + 1 : i));
        const float vmask = (i1 < border ? (float)(sin((RT_PI * i1) / (2 * border)) * sin((RT_PI * i1) / (2 * border))) : 1.0f);
        const float vmask2 = (i1 < 2 * border ? (float)(sin((RT_PI * i1) / (2 * border)) * sin((RT_PI * i1) / (2 * border))) : 1.0f);
        for (int j = 0; j < TS; ++j) {
            const float j1 = (float)abs((j > TS / 2 ? j - TS + 1 : j));
            const double sj = sin((RT_PI * j1) / (2 * border));
            tin[i * TS + j] = (float)((vmask * (j1 < border ? sj * sj : (double)1.0f)) + epsilon);
            tout[i * TS + j] = (float)((vmask2 * (j1 < 2 * border ? sj * sj : (double)1.0f)) + epsilon);
        }
    }
}

/* detail_recovery, L1479-1635, one thread.  L: wavelet-denoised luma (in/out), Lin: luma before reconstruction */
int artoracle_detail_recovery(float* L, const float* Lin, int width, int height, float params_Ldetail, int detail_thresh,
                              const float* tin, const float* tout, double scale)
{
    const float detail_hi = compute_detail(params_Ldetail);
    const float detail_lo = compute_detail(0.f);
    const int numblox_W = (int)ceil(((float)width) / OFFSET) + 2 * BLKRAD;
    const int numblox_H = (int)ceil(((float)height) / OFFSET) + 2 * BLKRAD;
    float* Ldetail = (float*)calloc((size_t)width * height, sizeof(float));
    float* totwt = (float*)calloc((size_t)width * height, sizeof(float));
    float* mask = NULL;
    if (detail_thresh > 0) {
        mask = (float*)malloc(sizeof(float) * (size_t)width * height);
        float amount = (float)detail_thresh / 100.f;
        amount = amount < 0.f ? 0.f : amount > 1.f ? 1.f : amount;
        int rc = artoracle_detail_mask(L, mask, width, height, 65535.f, 25.f, 10000.f, amount, 2, (float)(25.f / scale));
        if (rc) return rc;
    }
    float* Lblox = (float*)fftwf_malloc(sizeof(float) * (size_t)numblox_W * TS * TS);
    float* fLblox = (float*)fftwf_malloc(sizeof(float) * (size_t)numblox_W * TS * TS);
    float* detail_factor = (float*)malloc(sizeof(float) * (size_t)numblox_W * TS * TS);
    float* pBuf = (float*)malloc(sizeof(float) * (size_t)(width + TS + 2 * BLKRAD * OFFSET));
    float nbrwt[TS * TS], blurbuffer[TS * TS];
    const int nfwd[2] = {TS, TS};
    const fftw_r2r_kind fwdkind[2] = {FFTW_REDFT10, FFTW_REDFT10}, bwdkind[2] = {FFTW_REDFT01, FFTW_REDFT01};
    fftwf_plan pf = fftwf_plan_many_r2r(2, nfwd, numblox_W, Lblox, NULL, 1, TS * TS, fLblox, NULL, 1, TS * TS, fwdkind, 0);
    fftwf_plan pb = fftwf_plan_many_r2r(2, nfwd, numblox_W, fLblox, NULL, 1, TS * TS, Lblox, NULL, 1, TS * TS, bwdkind, 0);
    const int blur_rad = imax(1, (int)(3 / scale));
    const float DCTnorm = 1.0f / (4 * TS * TS);

    for (int vblk = 0; vblk < numblox_H; ++vblk) {
        const int top = (vblk - BLKRAD) * OFFSET;
        float* datarow = pBuf + BLKRAD * OFFSET;
        for (int i = 0; i < TS; ++i) {
            const int row = top + i;
            int rr = row;
            if (row < 0) rr = imin(-row, height - 1);
            else if (row >= height) rr = imax(0, 2 * height - 2 - row);
            for (int j = 0; j < width; ++j) datarow[j] = (Lin[(size_t)rr * width + j] - L[(size_t)rr * width + j]);
            for (int j = -BLKRAD * OFFSET; j < 0; ++j) datarow[j] = datarow[imin(-j, width - 1)];
            for (int j = width; j < width + TS + BLKRAD * OFFSET; ++j) datarow[j] = datarow[imax(0, 2 * width - 2 - j)];
            for (int hblk = 0; hblk < numblox_W; ++hblk) {
                const int left = (hblk - BLKRAD) * OFFSET;
                const int indx = hblk * TS;
                if (top + i >= 0 && top + i < height) {
                    int j;
                    for (j = 0; j < imin((-left), TS); ++j) {
                        Lblox[(indx + i) * TS + j] = tin[i * TS + j] * datarow[left + j];
                        detail_factor[(indx + i) * TS + j] = detail_lo;
                    }
                    for (; j < imin(TS, width - left); ++j) {
                        Lblox[(indx + i) * TS + j] = tin[i * TS + j] * datarow[left + j];
                        totwt[(size_t)(top + i) * width + left + j] += tin[i * TS + j] * tout[i * TS + j];
                        detail_factor[(indx + i) * TS + j] = detail_thresh > 0 ? compute_detail(params_Ldetail * mask[(size_t)(top + i) * width + left + j]) : detail_hi;
                    }
                    for (; j < TS; ++j) {
                        Lblox[(indx + i) * TS + j] = tin[i * TS + j] * datarow[left + j];
                        detail_factor[(indx + i) * TS + j] = detail_lo;
                    }
                } else {
                    for (int j = 0; j < TS; ++j) {
                        Lblox[(indx + i) * TS + j] = tin[i * TS + j] * datarow[left + j];
                        detail_factor[(indx + i) * TS + j] = detail_lo;
                    }
                }
            }
        }
        fftwf_execute_r2r(pf, Lblox, fLblox);
        for (int hblk = 0; hblk < numblox_W; ++hblk) {        /* RGBtile_denoise */
            float* blk = fLblox + (size_t)hblk * TS * TS;
            boxabsblur64(blk, nbrwt, blur_rad, blurbuffer);
            for (int n = 0; n < TS * TS; ++n)
                blk[n] = blk[n] * (1.0f - xexpf_vector(-(nbrwt[n] * nbrwt[n]) / detail_factor[(size_t)hblk * TS * TS + n]));
        }
        fftwf_execute_r2r(pb, fLblox, fLblox);
        {   /* RGBoutput_tile_row */
            const int nbw = (int)ceil(((float)width) / OFFSET);
            const int imin_ = imax(0, -top);
            const int bottom = imin(top + TS, height);
            const int imax_ = bottom - top;
            for (int i = imin_; i < imax_; ++i)
                for (int hblk = 0; hblk < nbw; ++hblk) {
                    const int left = (hblk - BLKRAD) * OFFSET;
                    const int right = imin(left + TS, width);
                    const int jmin = imax(0, -left);
                    const int jmax = right - left;
                    const int indx = hblk * TS;
                    for (int j = jmin; j < jmax; ++j)
                        Ldetail[(size_t)(top + i) * width + left + j] += tout[i * TS + j] * fLblox[(indx + i) * TS + j] * DCTnorm;
                }
        }
    }
    for (size_t i = 0; i < (size_t)width * height; ++i) L[i] += Ldetail[i] / totwt[i];
    fftwf_destroy_plan(pf); fftwf_destroy_plan(pb);
    fftwf_free(Lblox); fftwf_free(fLblox); free(detail_factor); free(pBuf); free(Ldetail); free(totwt); free(mask);
    return 0;
}

/*
 * p: luminance, luminanceDetail, luminanceDetailThreshold, chrominance, chrominanceRedGreen, chrominanceBlueYellow, gamma, scale
 * wp: working space matrix (row-major 3x3 double).  ccurve: the NoiseCurve's 501-entry LUT or NULL; calclum: 3 planes of
 * ((H+1)/2) x ((W+1)/2), needed when ccurve is in use.  out2: nresi, highresi.
 */
int artoracle_wavelet_denoise_AB_bishrink(void* wL, void* wab, const float* noisevarchrom, const float* madL, float noisevar_ab,
                                          int useNoiseCCurve, int autoch, double scale);
int artoracle_rgb_denoise_ex(float* r, float* g, float* b, int W, int H, const double* p, const double* wp,
                             const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* out2, int aggressive);
int artoracle_rgb_denoise_ex2(float* r, float* g, float* b, int W, int H, const double* p, const double* wp, const double* wp_inverse,
                              const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* out2, int aggressive, int lab_mode);
int artoracle_rgb_denoise(float* r, float* g, float* b, int W, int H, const double* p, const double* wp,
                          const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* out2)
{
    return artoracle_rgb_denoise_ex(r, g, b, W, H, p, wp, ccurve, ccurve_sum, cl_r, cl_g, cl_b, out2, 0);
}
/* aggressive = DenoiseParams::aggressive: nrQuality QUALITY_HIGH (L1671-1672): two more wavelet levels (L2260-2262), BiShrink for the
 * chroma channels and the luminance FOLLOWED by the standard shrinkage (L2339-2349, L2376-2386, L2412-2421; WaveletDenoiseAll_BiShrinkL
 * computes exactly what WaveletDenoiseAllL does, so the luminance is simply shrunk twice), qhighFactor 1 / 0.9 */
int artoracle_rgb_denoise_ex(float* r, float* g, float* b, int W, int H, const double* p, const double* wp,
                             const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* out2, int aggressive)
{
    return artoracle_rgb_denoise_ex2(r, g, b, W, H, p, wp, NULL, ccurve, ccurve_sum, cl_r, cl_g, cl_b, out2, aggressive, 0);
}
/* Color::denoiseGammaTab / denoiseIGammaTab (color.cc L188-189, L278-292; gamma55 / igamma55 color.h L1155-1169): LUTf(65536, 0) */
static float *g_dn_gamma = NULL, *g_dn_igamma = NULL;
static void init_dn_tabs(void)
{
    if (g_dn_gamma) return;
    float* a = (float*)malloc(sizeof(float) * 65536); float* b = (float*)malloc(sizeof(float) * 65536);
    for (int i = 0; i < 65536; i++) {
        const double x = i / 65535.0;
        a[i] = (float)(65535.0 * (x <= 0.013189 ? x * 10.0 : 1.593503 * exp(log(x) / 5.5) - 0.593503));
        b[i] = (float)(65535.0 * (x <= 0.131889 ? x / 10.0 : exp(log((x + 0.593503) / 1.593503) * 5.5)));
    }
    g_dn_gamma = a; g_dn_igamma = b;
}
const float* artoracle_denoise_gamma_tab(int inverse) { init_dn_tabs(); return inverse ? g_dn_igamma : g_dn_gamma; }
static inline float lut_noclip(const float* data, int size, float index)
{   /* LUT.h L437-459 with flags 0: extrapolates at both ends */
    int idx = (int)index;
    if (index < 0.f || !(index == index)) idx = 0;
    else if (index > (float)(size - 2)) idx = size - 2;
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
static inline float f2xyz_(float f)
{   /* color.h L767-770 */
    const float epsilonExpInv3f = (float)(6.0 / 29.0), kappaInvf = (float)(27.0 / 24389.0);
    return (f > epsilonExpInv3f) ? f * f * f : (116.f * f - 16.f) * kappaInvf;
}
/* lab_mode = DenoiseParams::ColorSpace::LAB (L1996): the samples go through denoiseIGammaTab before the gamma curve and through Color::rgb2lab
 * instead of rgb2yuv (L2093-2118); on the way out Color::lab2rgb with the inverse working-space matrix and denoiseGammaTab (L2519-2538) */
int artoracle_rgb_denoise_ex2(float* r, float* g, float* b, int W, int H, const double* p, const double* wp, const double* wp_inverse,
                              const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* out2, int aggressive, int lab_mode)
{
    if (lab_mode && !wp_inverse) return 1;
    float wpinv[3][3] = {{0}};
    if (wp_inverse) for (int i = 0; i < 9; ++i) wpinv[i / 3][i % 3] = (float)wp_inverse[i];
    if (lab_mode) init_dn_tabs();
    const double luminance = p[0], luminanceDetail = p[1], chrominance = p[3], chromRG = p[4], chromBY = p[5], scale = p[7];
    const int detail_thresh = (int)p[2];
    if (luminance == 0 && chrominance == 0 && !ccurve) return 0;
    init_cachef();
    const int useNoiseCCurve = (ccurve && ccurve_sum > 5.f);
    const float noiseluma = (float)luminance;
    const float noisevarL = (float)(((noiseluma / 125.0) * (1.0 + noiseluma / 25.0)) * ((noiseluma / 125.0) * (1.0 + noiseluma / 25.0)));
    const int denoiseLuminance = (noisevarL > 0.00001f);
    float wpi[3][3];
    for (int i = 0; i < 9; ++i) wpi[i / 3][i % 3] = (float)wp[i];
    const size_t n = (size_t)W * H;
    const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
    float* ccalc = NULL;
    if (useNoiseCCurve) {        /* L1706-1770 */
        ccalc = (float*)malloc(sizeof(float) * (size_t)w2 * h2);
        const float cn100Precalc = sqrf(1.f + 1.f * (4.f * lut_clip_both(ccurve, 501, 100.f / 60.f)));
        for (size_t i = 0; i < (size_t)w2 * h2; ++i) {
            const float RL = cl_r[i], GL = cl_g[i], BL = cl_b[i];
            const float XL = ((wpi[0][0] * RL + wpi[0][1] * GL + wpi[0][2] * BL));
            const float YL = ((wpi[1][0] * RL + wpi[1][1] * GL + wpi[1][2] * BL));
            const float ZL = ((wpi[2][0] * RL + wpi[2][1] * GL + wpi[2][2] * BL));
            const float x = XL / 0.9642f, z = ZL / 0.8249f, y = YL;
            const float fx = computeXYZ2Lab(x), fy = computeXYZ2Lab(y), fz = computeXYZ2Lab(z);
            const float AA = (500.0f * (fx - fy)), BB = (200.0f * (fy - fz));
            const float cN = sqrtf(sqrf(AA) + sqrf(BB));
            ccalc[i] = cN > 100 ? sqrf(1.f + 1.f * (4.f * lut_clip_both(ccurve, 501, cN / 60.f))) : cn100Precalc;
        }
        (void)computeXYZ2LabY;
    }
    if (!(luminance != 0 || chrominance != 0)) { free(ccalc); return 0; }

    const float gam = (float)p[6];
    const float gamthresh = 0.001f;
    float* gamcurve = (float*)malloc(sizeof(float) * 65536);
    float* igamcurve = (float*)malloc(sizeof(float) * 65536);
    const float gamslope = (float)(exp(log((double)gamthresh) / gam) / gamthresh);
    artoracle_gammaf2lut(gamcurve, gam, gamthresh, gamslope, 65535.f, 65535.f);
    const float igam = 1.f / gam;
    const float igamthresh = gamthresh * gamslope;
    const float igamslope = 1.f / gamslope;
    artoracle_gammaf2lut(igamcurve, igam, igamthresh, igamslope, 65535.f, 65535.f);
    const float gain = powf(2.0f, (float)0.0);
    const float params_Ldetail = fminf((float)luminanceDetail, 99.9f);
    float tin[TS * TS], tout[TS * TS];
    if (denoiseLuminance) artoracle_tilemasks(tin, tout);

    const int width = W, height = H, width2 = (width + 1) / 2;
    float interm_med = (float)chrominance / 10.0;
    float intermred = chromRG > 0. ? (float)(chromRG / 10.) : (float)((float)chromRG / 7.0);
    float intermblue = chromBY > 0. ? (float)(chromBY / 10.) : (float)((float)chromBY / 7.0);
    float realred = interm_med + intermred;
    if (realred <= 0.f) realred = 0.001f;
    float realblue = interm_med + intermblue;
    if (realblue <= 0.f) realblue = 0.001f;
    const float noisevarab_r = sqrf(realred), noisevarab_b = sqrf(realblue);
    const float maxNoiseVarab = fmaxf(noisevarab_b, noisevarab_r);

    float* Lp = (float*)malloc(sizeof(float) * n), * ap = (float*)malloc(sizeof(float) * n), * bp = (float*)malloc(sizeof(float) * n);
    float* noisevarlum = (float*)malloc(sizeof(float) * (size_t)h2 * w2);
    float* noisevarchrom = (float*)malloc(sizeof(float) * (size_t)h2 * w2);
#define APPLY_GAMMA(v) ((gam > 1.f && (v) > 0.f) ? ((v) < 65535.f ? lut_clip_below(gamcurve, 65536, (v)) : (gammaf_((v) / 65535.f, gam, gamthresh, gamslope) * 65535.f)) : (v))
#define APPLY_IGAMMA(v) ((gam > 1.f && (v) > 0.f) ? ((v) < 65536.f ? lut_clip_below(igamcurve, 65536, (v)) : (gammaf_((v) / 65535.f, igam, igamthresh, igamslope) * 65535.f)) : (v))
    for (int i = 0; i < height; ++i)
        for (int j = 0; j < width; ++j) {
            float X = gain * r[(size_t)i * W + j], Y = gain * g[(size_t)i * W + j], Z = gain * b[(size_t)i * W + j];
            if (lab_mode) { X = lut_noclip(g_dn_igamma, 65536, X); Y = lut_noclip(g_dn_igamma, 65536, Y); Z = lut_noclip(g_dn_igamma, 65536, Z); }
            X = APPLY_GAMMA(X); Y = APPLY_GAMMA(Y); Z = APPLY_GAMMA(Z);
            float l, u, v;
            if (lab_mode) {     /* Color::rgb2lab(X, Y, Z, l, v, u, wpi): rgbxyz + XYZ2Lab (color.cc L833-838, L1382-1399) */
                const float x = (wpi[0][0] * X + wpi[0][1] * Y + wpi[0][2] * Z), y = (wpi[1][0] * X + wpi[1][1] * Y + wpi[1][2] * Z), z = (wpi[2][0] * X + wpi[2][1] * Y + wpi[2][2] * Z);
                const float fx = computeXYZ2Lab(x / 0.9642f), fy = computeXYZ2Lab(y), fz = computeXYZ2Lab(z / 0.8249f);
                l = computeXYZ2LabY(y); v = 500.0f * (fx - fy); u = 200.0f * (fy - fz);
            } else {
                l = X * wpi[1][0] + Y * wpi[1][1] + Z * wpi[1][2];      /* rgb2yuv */
                u = l - Z; v = X - l;
            }
            Lp[(size_t)i * W + j] = l; ap[(size_t)i * W + j] = v; bp[(size_t)i * W + j] = u;
            if (((i | j) & 1) == 0) {
                noisevarlum[(i >> 1) * width2 + (j >> 1)] = noisevarL;
                noisevarchrom[(i >> 1) * width2 + (j >> 1)] = useNoiseCCurve ? maxNoiseVarab * ccalc[(size_t)(i >> 1) * w2 + (j >> 1)] : 1.f;
            }
        }

    /* wavelet levels, L2246-2293 */
    int levwav = 5;
    const float maxreal = fmaxf(realred, realblue);
    if (maxreal < 8.f) levwav = 5; else if (maxreal < 10.f) levwav = 6; else if (maxreal < 15.f) levwav = 7; else levwav = 8;
    if (aggressive) levwav += 2;
    if (levwav > 8) levwav = 8;
    levwav = imax(5, (int)(levwav - ceil(log(scale))));
    const int minsizetile = imin(W, H);
    int maxlev2 = 8;
    if (minsizetile < 256) maxlev2 = 7;
    if (minsizetile < 128) maxlev2 = 6;
    if (minsizetile < 64) maxlev2 = 5;
    if (minsizetile < 32) maxlev2 = 4;
    if (minsizetile < 16) maxlev2 = 3;
    levwav = imin(maxlev2, levwav);

    void* Ldecomp = artoracle_wavelet_new(Lp, W, H, levwav, 1);
    float madL[8][3];
    memset(madL, 0, sizeof madL);
    const int maxlvl = artoracle_wavelet_maxlevel(Ldecomp);
    for (int lvl = 0; lvl < maxlvl; ++lvl)
        for (int dir = 1; dir < 4; ++dir) {
            const float m = artoracle_madrgb(artoracle_wavelet_band(Ldecomp, lvl, dir), artoracle_wavelet_level_W(Ldecomp, lvl) * artoracle_wavelet_level_H(Ldecomp, lvl));
            madL[lvl][dir - 1] = m * m;
        }
    float chresid = 0.f, chmaxresid = 0.f, chresidtemp, chmaxresidtemp;
    float* chan[2] = {ap, bp};
    const float nv[2] = {noisevarab_r, noisevarab_b};
    float resid2[2], max2[2];
    for (int c = 0; c < 2; ++c) {
        void* dec = artoracle_wavelet_new(chan[c], W, H, levwav, 1);
        /* QUALITY_HIGH runs BiShrink and THEN the standard shrinkage (L2339-2349) */
        if (aggressive) artoracle_wavelet_denoise_AB_bishrink(Ldecomp, dec, noisevarchrom, &madL[0][0], nv[c], useNoiseCCurve, 0, scale);
        artoracle_wavelet_denoise_AB(Ldecomp, dec, noisevarchrom, &madL[0][0], nv[c], useNoiseCCurve, 0, scale);
        float resid = 0.f, maxresid = 0.f;      /* Noise_residualAB */
        const int ml = artoracle_wavelet_maxlevel(dec);
        for (int lvl = 0; lvl < ml; ++lvl)
            for (int dir = 1; dir < 4; ++dir) {
                const float m = artoracle_madrgb(artoracle_wavelet_band(dec, lvl, dir), artoracle_wavelet_level_W(dec, lvl) * artoracle_wavelet_level_H(dec, lvl));
                const float madC = m * m;
                resid += madC;
                if (madC > maxresid) maxresid = madC;
            }
        resid2[c] = resid; max2[c] = maxresid;
        artoracle_wavelet_reconstruct(dec, chan[c], 1.f);
        artoracle_wavelet_delete(dec);
    }
    chresidtemp = resid2[0]; chmaxresidtemp = max2[0];
    chresid = resid2[1]; chmaxresid = max2[1];
    chresid += chresidtemp; chmaxresid += chmaxresidtemp;
    chresid = sqrtf(chresid / (6 * (levwav)));
    if (out2) { out2[1] = chresid + 0.66f * (sqrtf(chmaxresid) - chresid); out2[0] = chresid; }

    float* Lin = NULL;
    if (denoiseLuminance) {
        /* QUALITY_HIGH: WaveletDenoiseAll_BiShrinkL (the same computation as WaveletDenoiseAllL) and then WaveletDenoiseAllL again, L2412-2421 */
        if (aggressive) artoracle_wavelet_denoise_L(Ldecomp, noisevarlum, &madL[0][0], scale);
        artoracle_wavelet_denoise_L(Ldecomp, noisevarlum, &madL[0][0], scale);
        Lin = (float*)malloc(sizeof(float) * n);
        memcpy(Lin, Lp, sizeof(float) * n);
        artoracle_wavelet_reconstruct(Ldecomp, Lp, 1.f);
    }
    artoracle_wavelet_delete(Ldecomp);
    if (denoiseLuminance) {
        int rc = artoracle_detail_recovery(Lp, Lin, W, H, params_Ldetail, detail_thresh, tin, tout, scale);
        if (rc) return rc;
    }
    const float newGain = 1.f / gain;
    const float qhighFactor = aggressive ? 1.f / (float)0.9 : 1.0f;
    for (size_t i = 0; i < n; ++i) {
        const float c_h = sqrtf(sqrf(ap[i]) + sqrf(bp[i]));
        if (c_h > 3000.f) {
            ap[i] *= 1.f + qhighFactor * realred / 100.f;
            bp[i] *= 1.f + qhighFactor * realblue / 100.f;
        }
        /* yuv2rgb(L, b, a): Y, u = b, v = a */
        const float Yv = Lp[i], u = bp[i], v = ap[i];
        float X, Y, Z;
        if (lab_mode) {     /* Color::lab2rgb(L, a, b, X, Y, Z, wpi_inverse): Lab2XYZ (color.cc L1203-1214) + xyz2rgb (L880-885) */
            const float c1By116 = (float)(1.0 / 116.0), c16By116 = (float)(16.0 / 116.0);
            const double kappa = 24389.0 / 27.0;
            const float LL = Yv / 327.68f, aa = v / 327.68f, bb = u / 327.68f;
            const float fy = (c1By116 * LL) + c16By116;
            const float fx = (0.002f * aa) + fy;
            const float fz = fy - (0.005f * bb);
            const float x = 65535.0f * f2xyz_(fx) * 0.9642f;
            const float z = 65535.0f * f2xyz_(fz) * 0.8249f;
            const float y = ((double)LL > 8.0) ? 65535.0f * fy * fy * fy : (float)(65535.0f * LL / kappa);
            X = (wpinv[0][0] * x + wpinv[0][1] * y + wpinv[0][2] * z);
            Y = (wpinv[1][0] * x + wpinv[1][1] * y + wpinv[1][2] * z);
            Z = (wpinv[2][0] * x + wpinv[2][1] * y + wpinv[2][2] * z);
        } else {
            Z = Yv - u;
            X = v + Yv;
            Y = (Yv - X * wpi[1][0] - Z * wpi[1][2]) / wpi[1][1];
        }
        X = APPLY_IGAMMA(X); Y = APPLY_IGAMMA(Y); Z = APPLY_IGAMMA(Z);
        if (lab_mode) { X = lut_noclip(g_dn_gamma, 65536, X); Y = lut_noclip(g_dn_gamma, 65536, Y); Z = lut_noclip(g_dn_gamma, 65536, Z); }
        r[i] = newGain * X; g[i] = newGain * Y; b[i] = newGain * Z;
    }
    free(Lp); free(ap); free(bp); free(Lin); free(noisevarlum); free(noisevarchrom); free(gamcurve); free(igamcurve); free(ccalc);
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * Automatic chroma estimator (DenoiseParams::ChrominanceMethod::AUTOMATIC, the reference's default): RGB_denoise_info
 * (ipdenoise.cc L227-669) over one crop, WaveletDenoiseAll_info / ShrinkAll_info (FTblockDN.cc L1227-1364), calcautodn_info
 * (ipdenoise.cc L66-206) and the nine-crop combination of ImProcFunctions::denoiseComputeParams (L964-1075).
 * ------------------------------------------------------------------------------------------------------------------------- */
static inline float mulsignf_(float x, float y)
{
    union { float f; uint32_t u; } a, b;
    a.f = x; b.f = y;
    a.u ^= (b.u & 0x80000000u);
    return a.f;
}
static float atan2kf_(float y, float x)
{   /* sleef.h L1155-1177; the SSE form (sleefsseavx.h L1168-1197) evaluates the same products and sums */
    float s, t, u, q = 0.f;
    if (x < 0) { x = -x; q = -2.f; }
    if (y > x) { t = x; x = y; y = -t; q += 1.f; }
    s = y / x;
    t = s * s;
    u = 0.00282363896258175373077393f;
    u = u * t + -0.0159569028764963150024414f;
    u = u * t + 0.0425049886107444763183594f;
    u = u * t + -0.0748900920152664184570312f;
    u = u * t + 0.106347933411598205566406f;
    u = u * t + -0.142027363181114196777344f;
    u = u * t + 0.199926957488059997558594f;
    u = u * t + -0.333331018686294555664062f;
    t = u * t;
    t = t * s + s;
    return q * (float)(3.14159265358979323846 / 2.0) + t;
}
static float xatan2f_(float y, float x)
{   /* sleef.h L1179-1188 */
    const float PI_F = (float)3.14159265358979323846;
    float r = atan2kf_(fabsf(y), x);
    r = mulsignf_(r, x);
    if (isinf(x) || x == 0) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .5f)) : 0);
    if (isinf(y)) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .25f)) : 0);
    if (y == 0) r = (copysignf(1.f, x) == -1 ? PI_F : 0);
    return (x != x) || (y != y) ? NAN : mulsignf_(r, y);
}

/* src: the crop (W x H); calc: the crop's even pixels converted to the working space (provicalc, (H+1)/2 x (W+1)/2).
 * out[15] = chaut, Nb, redaut, blueaut, maxredaut, maxblueaut, minredaut, minblueaut, chromina, sigma, lumema, sigma_L, redyel, skinc, nsknc
 * (entries the reference leaves untouched keep the caller's values).  Returns 2 when the crop is too small for levwav wavelet levels. */
int artoracle_denoise_info(const float* r, const float* g, const float* b, int W, int H, const float* cr, const float* cg, const float* cb,
                           double gamma, int aggressive, double scale, double expcomp, const double* wp9, float* out)
{
    init_cachef();
    const int wid = (W + 1) / 2, hei = (H + 1) / 2;
    float wp[3][3];
    for (int i = 0; i < 9; ++i) wp[i / 3][i % 3] = (float)wp9[i];
    const size_t n2 = (size_t)wid * hei, n = (size_t)W * H;
    float* lumcalc = (float*)malloc(sizeof(float) * n2); float* acalc = (float*)malloc(sizeof(float) * n2); float* bcalc = (float*)malloc(sizeof(float) * n2);
    for (size_t k = 0; k < n2; ++k) {       /* L271-287: Color::rgbxyz + Color::XYZ2Lab */
        const float RL = cr[k], GL = cg[k], BL = cb[k];
        const float XL = ((wp[0][0] * RL + wp[0][1] * GL + wp[0][2] * BL));
        const float YL = ((wp[1][0] * RL + wp[1][1] * GL + wp[1][2] * BL));
        const float ZL = ((wp[2][0] * RL + wp[2][1] * GL + wp[2][2] * BL));
        const float x = XL / 0.9642f, z = ZL / 0.8249f, y = YL;
        const float fx = computeXYZ2Lab(x), fy = computeXYZ2Lab(y), fz = computeXYZ2Lab(z);
        lumcalc[k] = computeXYZ2LabY(y);
        acalc[k] = (500.0f * (fx - fy));
        bcalc[k] = (200.0f * (fy - fz));
    }
    /* RGB_denoise_infoGamCurve, L209-225 (isRAW) */
    const float gam = (float)gamma;
    const float gamthresh = 0.001f;
    const float gamslope = (float)(exp(log((double)gamthresh) / gam) / gamthresh);
    float* gamcurve = (float*)malloc(sizeof(float) * 65536);
    artoracle_gammaf2lut(gamcurve, gam, gamthresh, gamslope, 65535.f, 32768.f);
    const float gain = powf(2.0f, (float)expcomp);      /* L293 */

    /* the whole crop is one tile (Tile_calc with kall == 0, L311); isRAW branch L390-474 */
    float* nvlum = (float*)malloc(sizeof(float) * n2); float* nvchrom = (float*)malloc(sizeof(float) * n2); float* nvhue = (float*)malloc(sizeof(float) * n2);
    for (size_t k = 0; k < n2; ++k) {
        const float aN = acalc[k], bN = bcalc[k];
        float cN = sqrtf(sqrf(aN) + sqrf(bN));
        nvhue[k] = xatan2f_(bN, aN);
        if (cN < 100.f) cN = 100.f;
        nvchrom[k] = cN;
        float Llum = lumcalc[k];
        Llum = Llum < 2.f ? 2.f : Llum;
        Llum = Llum > 32768.f ? 32768.f : Llum;
        nvlum[k] = Llum;
    }
    float* la = (float*)malloc(sizeof(float) * n); float* lb = (float*)malloc(sizeof(float) * n);
    for (size_t k = 0; k < n; ++k) {
        float X = gain * r[k], Y = gain * g[k], Z = gain * b[k];
        X = X < 65535.f ? lut_noclip(gamcurve, 65536, X) : (gammaf_(X / 65535.f, gam, gamthresh, gamslope) * 32768.f);
        Y = Y < 65535.f ? lut_noclip(gamcurve, 65536, Y) : (gammaf_(Y / 65535.f, gam, gamthresh, gamslope) * 32768.f);
        Z = Z < 65535.f ? lut_noclip(gamcurve, 65536, Z) : (gammaf_(Z / 65535.f, gam, gamthresh, gamslope) * 32768.f);
        const float l = X * wp[1][0] + Y * wp[1][1] + Z * wp[1][2];      /* Color::rgb2yuv */
        la[k] = X - l;          /* labdn->a = v */
        lb[k] = l - Z;          /* labdn->b = u */
    }
    const int schoice = aggressive ? 2 : 0;
    const int levwav = imax(2, (int)(5 - ceil(log(scale))));
    void* adec = artoracle_wavelet_new(la, W, H, levwav, 1);
    void* bdec = artoracle_wavelet_new(lb, W, H, levwav, 1);
    int rc = 0;
    if (artoracle_wavelet_maxlevel(adec) < levwav) rc = 2;
    /* WaveletDenoiseAll_info -> ShrinkAll_info, FTblockDN.cc L1227-1364 */
    float chau = 0.f, chred = 0.f, chblue = 0.f, maxchred = 0.f, maxchblue = 0.f, minchred = 100000000.f, minchblue = 100000000.f;
    int nb = 0;
    for (int lvl = 0; lvl < levwav && !rc; ++lvl) {
        const int W_ab = artoracle_wavelet_level_W(adec, lvl), H_ab = artoracle_wavelet_level_H(adec, lvl);
        if (lvl == 1) {
            float chro = 0.f, dev = 0.f, devL = 0.f, lume = 0.f, red_yel = 0.f, skin_c = 0.f;
            int nc = 0, nL = 0, nry = 0, nsk = 0;
            for (int i = 0; i < H_ab; ++i)
                for (int j = 0; j < W_ab; ++j) {
                    const float c = nvchrom[(size_t)i * wid + j], h = nvhue[(size_t)i * wid + j], l = nvlum[(size_t)i * wid + j];
                    chro += c;
                    ++nc;
                    dev += sqrf(c - (chro / nc));
                    if (h > -0.8f && h < 2.0f && c > 10000.f) { red_yel += c; ++nry; }
                    if (h > 0.f && h < 1.6f && c < 10000.f) { skin_c += c; ++nsk; }
                    lume += l;
                    ++nL;
                    devL += sqrf(l - (lume / nL));
                }
            if (nc > 0) { out[8] = chro / nc; out[9] = sqrtf(dev / nc); out[14] = (float)nsk / (float)nc; }
            else out[14] = (float)nsk;
            if (nL > 0) { out[10] = lume / nL; out[11] = sqrtf(devL / nL); }
            if (nry > 0) out[12] = red_yel / nry;
            if (nsk > 0) out[13] = skin_c / nsk;
        }
        const float reduc = (schoice == 2) ? (float)0.9 : 1.f;
        for (int dir = 1; dir < 4; ++dir) {
            const float ma = artoracle_madrgb(artoracle_wavelet_band(adec, lvl, dir), W_ab * H_ab);
            const float mada = ma * ma;
            chred += mada;
            if (mada > maxchred) maxchred = mada;
            if (mada < minchred) minchred = mada;
            out[4] = sqrtf(reduc * maxchred);
            out[6] = sqrtf(reduc * minchred);
            const float mb = artoracle_madrgb(artoracle_wavelet_band(bdec, lvl, dir), W_ab * H_ab);
            const float madb = mb * mb;
            chblue += madb;
            if (madb > maxchblue) maxchblue = madb;
            if (madb < minchblue) minchblue = madb;
            out[5] = sqrtf(reduc * maxchblue);
            out[7] = sqrtf(reduc * minchblue);
            chau += (mada + madb);
            ++nb;
            out[0] = sqrtf(reduc * chau / (nb + nb));
            out[2] = sqrtf(reduc * chred / nb);
            out[3] = sqrtf(reduc * chblue / nb);
            out[1] = (float)nb;
        }
    }
    artoracle_wavelet_delete(adec); artoracle_wavelet_delete(bdec);
    free(lumcalc); free(acalc); free(bcalc); free(gamcurve); free(nvlum); free(nvchrom); free(nvhue); free(la); free(lb);
    return rc;
}

/* calcautodn_info, ipdenoise.cc L66-206 */
static void calcautodn_info_(int aggressive, float* chaut, float* delta, int Nb, int levaut, float maxmax, float lumema, float chromina, int mode, int lissage,
                             float redyel, float skinc, float nsknc)
{
    float reducdelta = 1.f;
    if (aggressive) reducdelta = (float)0.9;
    float c = (*chaut * Nb - maxmax) / (Nb - 1);
    if ((redyel > 5000.f || skinc > 1000.f) && nsknc < 0.4f && chromina > 3000.f) c *= 0.45f;
    else if ((redyel > 12000.f || skinc > 1200.f) && nsknc < 0.3f && chromina > 3000.f) c *= 0.3f;
    if (mode == 0 || mode == 2) {
        if (chromina > 10000.f) c *= 0.7f; else if (chromina > 6000.f) c *= 0.9f; else if (chromina < 3000.f) c *= 1.2f; else if (chromina < 2000.f) c *= 1.5f;
        if (lumema < 2500.f) c *= 1.3f; else if (lumema < 5000.f) c *= 1.2f; else if (lumema > 20000.f) c *= 0.9f;
    } else if (mode == 1) {
        if (chromina > 10000.f) c *= 0.8f; else if (chromina > 6000.f) c *= 0.9f; else if (chromina < 3000.f) c *= 1.5f; else if (chromina < 2000.f) c *= 2.2f;
        if (lumema < 2500.f) c *= 1.2f; else if (lumema < 5000.f) c *= 1.1f; else if (lumema > 20000.f) c *= 0.9f;
    }
    if (levaut == 0 && c > 300.f) c = 0.714286f * c + 85.71428f;
    float d = maxmax - c;
    d *= reducdelta;
    if (lissage == 1 || lissage == 2) {
        if (c < 200.f && d < 200.f) d *= 0.95f; else if (c < 200.f && d < 400.f) d *= 0.5f; else if (c < 200.f && d >= 400.f) d = 200.f;
        else if (c < 400.f && d < 400.f) d *= 0.4f; else if (c < 400.f && d >= 400.f) d = 120.f;
        else if (c < 550.f) d *= 0.15f; else if (c < 650.f) d *= 0.1f; else d *= 0.07f;
        if (mode == 0 || mode == 2) { if (chromina < 6000.f) d *= 1.4f; if (lumema < 5000.f) d *= 1.4f; }
        else if (mode == 1) { if (chromina < 6000.f) d *= 1.2f; if (lumema < 5000.f) d *= 1.2f; }
    }
    if (lissage == 0) {
        if (c < 200.f && d < 200.f) d *= 0.95f; else if (c < 200.f && d < 400.f) d *= 0.7f; else if (c < 200.f && d >= 400.f) d = 280.f;
        else if (c < 400.f && d < 400.f) d *= 0.6f; else if (c < 400.f && d >= 400.f) d = 200.f;
        else if (c < 550.f) d *= 0.3f; else if (c < 650.f) d *= 0.2f; else d *= 0.15f;
        if (mode == 0 || mode == 2) { if (chromina < 6000.f) d *= 1.4f; if (lumema < 5000.f) d *= 1.4f; }
        else if (mode == 1) { if (chromina < 6000.f) d *= 1.2f; if (lumema < 5000.f) d *= 1.2f; }
    }
    *chaut = c; *delta = d;
}

/* The nine-crop combination of denoiseComputeParams (L964-1075).  stats: 9 x 15 floats in artoracle_denoise_info's layout, crop k = hcr * 3 + wcr.
 * out3 = store.chrominance, store.chrominanceRedGreen, store.chrominanceBlueYellow (before chrominanceAutoFactor). */
int artoracle_denoise_auto_params(const float* stats, int isRAW, int aggressive, float* out3)
{
    const float autoNR = 10, autoNRmax = 40, lowdenoise = 1.f, adjustr = 1.f;
    const int levaut = 0, mode = 1, lissage = 0;
    float ch_M[9], max_r[9], max_b[9], min_r[9], min_b[9], delta[9];
    float Max_R[9] = {0}, Max_B[9] = {0}, Min_R[9], Min_B[9];
    const float multip = isRAW ? 1.f : 2.f;
    for (int k = 0; k < 9; ++k) {
        const float* s = stats + 15 * k;
        ch_M[k] = 1.0f * s[0]; max_r[k] = 1.0f * s[4]; max_b[k] = 1.0f * s[5]; min_r[k] = 1.0f * s[6]; min_b[k] = 1.0f * s[7];
        const float maxmax = max_r[k] > max_b[k] ? max_r[k] : max_b[k];       /* rtengine::max(a, b) */
        calcautodn_info_(aggressive, &ch_M[k], &delta[k], (int)s[1], levaut, maxmax, s[10], s[8], mode, lissage, s[12], s[13], s[14]);
    }
    for (int k = 0; k < 9; ++k) {
        if (max_r[k] > max_b[k]) {
            Max_R[k] = (delta[k]) / ((autoNRmax * multip * adjustr * lowdenoise) / 2.f);
            Min_B[k] = -(ch_M[k] - min_b[k]) / (autoNRmax * multip * adjustr * lowdenoise);
            Max_B[k] = 0.f; Min_R[k] = 0.f;
        } else {
            Max_B[k] = (delta[k]) / ((autoNRmax * multip * adjustr * lowdenoise) / 2.f);
            Min_R[k] = -(ch_M[k] - min_r[k]) / (autoNRmax * multip * adjustr * lowdenoise);
            Min_B[k] = 0.f; Max_R[k] = 0.f;
        }
    }
    float chM = 0.f, MaxR = 0.f, MaxB = 0.f, MinR = 100000000000.f, MinB = 100000000000.f, maxr, maxb;
    float MaxRMoy = 0.f, MaxBMoy = 0.f, MinRMoy = 0.f, MinBMoy = 0.f;
    for (int k = 0; k < 9; ++k) {
        chM += ch_M[k]; MaxBMoy += Max_B[k]; MaxRMoy += Max_R[k]; MinRMoy += Min_R[k]; MinBMoy += Min_B[k];
        if (Max_R[k] > MaxR) MaxR = Max_R[k];
        if (Max_B[k] > MaxB) MaxB = Max_B[k];
        if (Min_R[k] < MinR) MinR = Min_R[k];
        if (Min_B[k] < MinB) MinB = Min_B[k];
    }
    chM /= 9; MaxBMoy /= 9; MaxRMoy /= 9; MinBMoy /= 9; MinRMoy /= 9;
    if (MaxR > MaxB) { maxr = MaxRMoy + (MaxR - MaxRMoy) * 0.66f; maxb = MinBMoy + (MinB - MinBMoy) * 0.66f; }
    else { maxb = MaxBMoy + (MaxB - MaxBMoy) * 0.66f; maxr = MinRMoy + (MinR - MinRMoy) * 0.66f; }
    out3[0] = chM / (autoNR * multip * adjustr);
    out3[1] = maxr;
    out3[2] = maxb;
    return 0;
}

/* Color::computeXYZ2LabY for other ports (color.cc L1262-1274) */
float artoracle_xyz2laby(float f) { init_cachef(); return computeXYZ2LabY(f); }

/* the block DCT stand-in alone, for cross-checks against an independent implementation (tests/test_oracle_denoise.py):
 * kind 0 = REDFT10 x REDFT10 (FTblockDN.cc L1604), 1 = REDFT01 x REDFT01 (L1614) on one n x n block */
int artoracle_block_dct(int n, int kind, const float* in, float* out)
{
    const int nn[2] = {n, n};
    const fftw_r2r_kind k[2] = {kind ? FFTW_REDFT01 : FFTW_REDFT10, kind ? FFTW_REDFT01 : FFTW_REDFT10};
    fftwf_plan p = fftwf_plan_many_r2r(2, nn, 1, NULL, NULL, 1, n * n, NULL, NULL, 1, n * n, k, FFTW_ESTIMATE);
    if (!p) return 1;
    artdct_block(p, in, out);
    fftwf_destroy_plan(p);
    return 0;
}
