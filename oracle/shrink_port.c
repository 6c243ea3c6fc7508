/*
 * oracle/shrink_port.c -- CPU restatement of FTblockDN's wavelet shrinkage.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates (reference rtengine/FTblockDN.cc) MadRgb L569-603, ShrinkAllL L638-726, ShrinkAllAB L729-839,
 * WaveletDenoiseAllL L1111-1167 (edge == 0, vari == nullptr) and WaveletDenoiseAllAB L1170-1221, the SSE2
 * build: coefficients in the 4-wide vector loops use the vector xexpf (rtengine/sleefsseavx.h L1326-1345) and
 * the vector expression association; the n % 4 trailing coefficients use the scalar xexpf (rtengine/sleef.h
 * L1247-1266) and the scalar association.  The averaging filter is the flat boxblur (rtengine/boxblur.h
 * L558-742) with its three column classes (8/4-wide vector columns multiply by 1/len, the W % 4 trailing
 * columns divide and start from a sum of quotients).
 * Pinned bit-exact against the reference functions compiled in place (oracle/_ref) in tests/test_oracle_shrink.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void* artoracle_wavelet_new(const float* src, int W, int H, int maxlvl, int subsamp);
int artoracle_wavelet_maxlevel(void* p);
int artoracle_wavelet_level_W(void* p, int l);
int artoracle_wavelet_level_H(void* p, int l);
float* artoracle_wavelet_band(void* p, int l, int dir);

#include "sleef_port.h"

/* boxblur(T* src, A* dst, A* buffer, radx, rady, W, H), boxblur.h L558-742 (radx == rady == rad >= 1) */
static void boxblur_flat(const float* src, float* dst, float* temp, int rad, int W, int H)
{
    for (int row = H - 1; row >= 0; row--) {
        int len = rad + 1;
        float t = src[row * W];
        for (int j = 1; j <= rad; j++) t += src[row * W + j];
        t = t / len;
        temp[row * W] = t;
        for (int col = 1; col <= rad; col++) {
            t = (t * len + src[row * W + col + rad]) / (len + 1);
            temp[row * W + col] = t;
            len++;
        }
        const float reclen = 1.f / len;
        for (int col = rad + 1; col < W - rad; col++) {
            t = t + ((float)(src[row * W + col + rad] - src[row * W + col - rad - 1])) * reclen;
            temp[row * W + col] = t;
        }
        for (int col = W - rad; col < W; col++) {
            t = (t * len - src[row * W + col - rad - 1]) / (len - 1);
            temp[row * W + col] = t;
            len--;
        }
    }
    const int nvec = W - (W % 4);          /* 8-wide then 4-wide vector columns (L614-688) */
    for (int col = 0; col < nvec; col++) {
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
    for (int col = nvec; col < W; col++) {  /* scalar trailing columns (L690-710) */
        int len = rad + 1;
        dst[col] = temp[col] / len;
        for (int i = 1; i <= rad; i++) dst[col] += temp[i * W + col] / len;
        for (int row = 1; row <= rad; row++) {
            dst[row * W + col] = (dst[(row - 1) * W + col] * len + temp[(row + rad) * W + col]) / (len + 1);
            len++;
        }
        for (int row = rad + 1; row < H - rad; row++)
            dst[row * W + col] = dst[(row - 1) * W + col] + (temp[(row + rad) * W + col] - temp[(row - rad - 1) * W + col]) / len;
        for (int row = H - rad; row < H; row++) {
            dst[row * W + col] = (dst[(row - 1) * W + col] * len - temp[(row - rad - 1) * W + col]) / (len - 1);
            len--;
        }
    }
}

float artoracle_madrgb(const float* data, int n)
{   /* L569-603 */
    if (n <= 1) return 0;
    int* histo = (int*)calloc(65536, sizeof(int));
    for (int i = 0; i < n; ++i) {
        int v = abs((int)data[i]);
        histo[v < 65535 ? v : 65535]++;
    }
    int median = 0, count = 0;
    while (count < n / 2) { count += histo[median]; ++median; }
    const int count_ = count - histo[median - 1];
    free(histo);
    return (((median - 1) + (n / 2 - count_) / ((float)(count - count_))) / 0.6745);
}

static int blur_radius(int level, double scale) { const int r = (int)((level + 2) / scale); return r > 1 ? r : 1; }

static void shrink_L(void* wL, float** buffer, int level, int dir, const float* noisevarlum, const float* madL, double scale)
{
    const float eps = 0.01f;
    float *sf = buffer[0], *sfd = buffer[1], *blur = buffer[2];
    const int W = artoracle_wavelet_level_W(wL, level), H = artoracle_wavelet_level_H(wL, level), n = W * H;
    float* c = artoracle_wavelet_band(wL, level, dir);
    const float mad_L = madL[dir - 1];
    const float levelFactor = mad_L * 5.f / (float)(level + 1);
    int i;
    for (i = 0; i < n - 3; i += 4)
        for (int l = 0; l < 4; ++l) {
            const float mad = noisevarlum[i + l] * levelFactor;
            const float mag = c[i + l] * c[i + l];
            sf[i + l] = mag / (mag + mad * xexpf_vector(-mag / (9.0f * mad)) + eps);
        }
    for (; i < n; ++i) {
        const float mag = c[i] * c[i];
        sf[i] = mag / (mag + levelFactor * noisevarlum[i] * xexpf_scalar(-mag / (9 * levelFactor * noisevarlum[i])) + eps);
    }
    boxblur_flat(sf, sfd, blur, blur_radius(level, scale), W, H);
    for (i = 0; i < n - 3; i += 4)
        for (int l = 0; l < 4; ++l) {
            const float s = sf[i + l], d = sfd[i + l];
            c[i + l] = c[i + l] * (d * d + s * s) / (d + s + eps);
        }
    for (; i < n; ++i) {
        const float s = sf[i];
        c[i] *= (sfd[i] * sfd[i] + s * s) / (sfd[i] + s + eps);
    }
}

static void shrink_AB(void* wL, void* wab, float** buffer, int level, int dir, const float* noisevarchrom, float noisevar_ab,
                      int useNoiseCCurve, int autoch, const float* madL, double scale)
{
    const float eps = 0.01f;
    if (autoch && noisevar_ab <= 0.001f) noisevar_ab = 0.02f;
    float *sf = buffer[0], *sfd = buffer[1], *blur = buffer[2];
    const int W = artoracle_wavelet_level_W(wab, level), H = artoracle_wavelet_level_H(wab, level), n = W * H;
    const float* cL = artoracle_wavelet_band(wL, level, dir);
    float* cab = artoracle_wavelet_band(wab, level, dir);
    const float mad_L = madL[dir - 1];
    float madab = artoracle_madrgb(cab, n);
    madab = madab * madab;
    if (!(noisevar_ab > 0.001f)) return;
    madab = useNoiseCCurve ? madab : madab * noisevar_ab;
    const float rmadLm9 = 1.f / (mad_L * 9.f);
    int i;
    for (i = 0; i < n - 3; i += 4)
        for (int l = 0; l < 4; ++l) {
            const float mad_ab = noisevarchrom[i + l] * madab;
            float mag_L = cL[i + l];
            const float mag_ab = cab[i + l] * cab[i + l];
            mag_L = (mag_L * mag_L) * rmadLm9;
            sf[i + l] = 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
        }
    for (; i < n; ++i) {
        const float mag_L = cL[i] * cL[i], mag_ab = cab[i] * cab[i];
        sf[i] = (1.f - xexpf_scalar(-(mag_ab / (noisevarchrom[i] * madab)) - (mag_L / (9.f * mad_L))));
    }
    boxblur_flat(sf, sfd, blur, blur_radius(level, scale), W, H);
    for (i = 0; i < n - 3; i += 4)
        for (int l = 0; l < 4; ++l) {
            const float s = sf[i + l], d = sfd[i + l];
            cab[i + l] = cab[i + l] * (d * d + s * s) / (d + s + eps);
        }
    for (; i < n; ++i) {
        const float s = sf[i];
        cab[i] *= (sfd[i] * sfd[i] + s * s) / (sfd[i] + s + eps);
    }
}

static float** shrink_buffers(void* w, int maxlvl)
{
    int mw = 0, mh = 0;
    for (int l = 0; l < maxlvl; ++l) {
        if (artoracle_wavelet_level_W(w, l) > mw) mw = artoracle_wavelet_level_W(w, l);
        if (artoracle_wavelet_level_H(w, l) > mh) mh = artoracle_wavelet_level_H(w, l);
    }
    float** b = (float**)malloc(3 * sizeof(float*));
    for (int k = 0; k < 3; ++k) b[k] = (float*)malloc(sizeof(float) * ((size_t)mw * mh + 128));
    return b;
}

int artoracle_wavelet_denoise_L(void* wL, const float* noisevarlum, const float* madL /*[8][3]*/, double scale)
{
    int maxlvl = artoracle_wavelet_maxlevel(wL);
    if (maxlvl > 5) maxlvl = 5;                      /* L1115 */
    float** b = shrink_buffers(wL, maxlvl);
    for (int lvl = 0; lvl < maxlvl; ++lvl)
        for (int dir = 1; dir < 4; ++dir) shrink_L(wL, b, lvl, dir, noisevarlum, madL + 3 * lvl, scale);
    for (int k = 0; k < 3; ++k) free(b[k]);
    free(b);
    return 0;
}

/* WaveletDenoiseAll_BiShrinkAB, L976-1108 (QUALITY_HIGH = DenoiseParams::aggressive): the MADs of every band first, then the
 * coarsest level through ShrinkAllAB (with those MADs) and every finer level through the "simple" shrinkage, which has no local
 * averaging and squares (1 - exp(.)); note SQR(noisevar_ab) where ShrinkAllAB has noisevar_ab. */
int artoracle_wavelet_denoise_AB_bishrink(void* wL, void* wab, const float* noisevarchrom, const float* madL, float noisevar_ab,
                                          int useNoiseCCurve, int autoch, double scale)
{
    const int maxlvl = artoracle_wavelet_maxlevel(wL);
    if (autoch && noisevar_ab <= 0.001f) noisevar_ab = 0.02f;
    float** b = shrink_buffers(wL, maxlvl);
    for (int lvl = maxlvl - 1; lvl >= 0; lvl--)
        for (int dir = 1; dir < 4; ++dir) {
            if (lvl == maxlvl - 1) {       /* ShrinkAllAB recomputes nothing else: the band's MAD is the same before and after the others shrink */
                shrink_AB(wL, wab, b, lvl, dir, noisevarchrom, noisevar_ab, useNoiseCCurve, autoch, madL + 3 * lvl, scale);
                continue;
            }
            const int W = artoracle_wavelet_level_W(wab, lvl), H = artoracle_wavelet_level_H(wab, lvl), n = W * H;
            const float* cL = artoracle_wavelet_band(wL, lvl, dir);
            float* cab = artoracle_wavelet_band(wab, lvl, dir);
            float madab = artoracle_madrgb(cab, n);
            madab = madab * madab;
            const float mad_Lr = madL[3 * lvl + dir - 1];
            const float mad_abr = useNoiseCCurve ? noisevar_ab * madab : (noisevar_ab * noisevar_ab) * madab;
            if (!(noisevar_ab > 0.001f)) continue;
            const float rmad_Lm9 = 1.f / (mad_Lr * 9.f);
            int i;
            for (i = 0; i < n - 3; i += 4)
                for (int l = 0; l < 4; ++l) {
                    const float mad_ab = noisevarchrom[i + l] * mad_abr;
                    const float t = cab[i + l];
                    float mag_L = cL[i + l];
                    const float mag_ab = t * t;
                    mag_L = (mag_L * mag_L) * rmad_Lm9;
                    const float f = 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
                    cab[i + l] = t * (f * f);
                }
            for (; i < n; ++i) {
                const float mag_L = cL[i] * cL[i], mag_ab = cab[i] * cab[i];
                const float f = 1.f - xexpf_scalar(-(mag_ab / (noisevarchrom[i] * mad_abr)) - (mag_L / (9.f * mad_Lr)));
                cab[i] *= f * f;
            }
        }
    for (int k = 0; k < 3; ++k) free(b[k]);
    free(b);
    return 0;
}

int artoracle_wavelet_denoise_AB(void* wL, void* wab, const float* noisevarchrom, const float* madL, float noisevar_ab,
                                 int useNoiseCCurve, int autoch, double scale)
{
    const int maxlvl = artoracle_wavelet_maxlevel(wL);
    float** b = shrink_buffers(wL, maxlvl);
    for (int lvl = 0; lvl < maxlvl; ++lvl)
        for (int dir = 1; dir < 4; ++dir) shrink_AB(wL, wab, b, lvl, dir, noisevarchrom, noisevar_ab, useNoiseCCurve, autoch, madL + 3 * lvl, scale);
    for (int k = 0; k < 3; ++k) free(b[k]);
    free(b);
    return 0;
}
