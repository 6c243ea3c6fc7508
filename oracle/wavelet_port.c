/*
 * oracle/wavelet_port.c -- CPU restatement of the reference's wavelet decomposition / reconstruction.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates rtengine::wavelet_decomposition (reference rtengine/cplx_wavelet_dec.h L97-270) and
 * rtengine::wavelet_level (rtengine/cplx_wavelet_level.h L77-114 geometry, L205-763 filters) for Daub4Len == 6
 * (cplx_wavelet_filter_coeffs.h Daub4_anal, offset 2) and skipcrop == 1:
 *   level l is decimated ("subsamp_out") iff bit l of `subsampling` is set: 6-tap FIR, clamped borders, output
 *   (w+1)/2 x (h+1)/2; otherwise an undecimated Haar pair with tap spacing `skip` (L82-104);
 *   synthesis mirrors it (Haar: 0.5 * (lo + hi + lo[-skip] - hi[-skip]); FIR: polyphase with shift 3, the
 *   vertical stage blends into the destination: dst = dst * (1 - blend) + blend * 4 * tot, L546).
 * Accumulation order follows the reference loops exactly.  Pinned bit-exact against the reference headers
 * compiled unmodified (oracle/_ref) in tests/test_oracle_wavelet.py.  Compile with -ffp-contract=off.
 */
#include <stdlib.h>
#include <string.h>

#define TAPS 6
#define OFFS 2
#define MAXLEV 10
static const float ANAL[2][TAPS] = {
    {0.f, 0.f, 0.34150635f, 0.59150635f, 0.15849365f, -0.091506351f},
    {-0.091506351f, -0.15849365f, 0.59150635f, -0.34150635f, 0.f, 0.f}};

typedef struct {
    int w, h, w2, h2, skip, sub;
    float* band[4];          /* [1..3] owned (one block), [0] unused */
} wlevel;

typedef struct {
    int nlev, W, H, subsamp;
    wlevel lev[MAXLEV];
    float* coeff0;           /* lowpass of the last level */
} wdec;

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static void decompose_level(wlevel* L, const float* src, float* dst)
{
    const int w = L->w, h = L->h, w2 = L->w2, skip = L->skip;
    float* tmpLo = (float*)malloc(sizeof(float) * 2 * (size_t)w);
    float* tmpHi = tmpLo + w;
    if (L->sub) {
        for (int row = 0; row < h; row += 2) {
            for (int k = 0; k < w; k++) {           /* AnalysisFilterSubsampVertical, L331-398 */
                float lo = 0.f, hi = 0.f;
                for (int j = 0; j < TAPS; j++) {
                    const float s = src[(size_t)imax(0, imin(row + skip * (OFFS - j), h - 1)) * w + k];
                    lo += ANAL[0][j] * s;
                    hi += ANAL[1][j] * s;
                }
                tmpLo[k] = lo; tmpHi[k] = hi;
            }
            for (int pass = 0; pass < 2; pass++) {  /* AnalysisFilterSubsampHorizontal, L301-329 */
                const float* t = pass ? tmpHi : tmpLo;
                float* dLo = pass ? L->band[2] : dst;
                float* dHi = pass ? L->band[3] : L->band[1];
                for (int i = 0; i < w; i += 2) {
                    float lo = 0.f, hi = 0.f;
                    for (int j = 0; j < TAPS; j++) {
                        const float s = t[imax(0, imin(i + skip * (OFFS - j), w - 1))];
                        lo += ANAL[0][j] * s;
                        hi += ANAL[1][j] * s;
                    }
                    dLo[(size_t)(row / 2) * w2 + i / 2] = lo;
                    dHi[(size_t)(row / 2) * w2 + i / 2] = hi;
                }
            }
        }
    } else {
        for (int row = 0; row < h; row++) {
            int have = 0;
            if (row < h - skip) {                   /* AnalysisFilterHaarVertical, L223-239 */
                for (int j = 0; j < w; j++) {
                    tmpLo[j] = 0.25f * (src[(size_t)row * w + j] + src[(size_t)(row + skip) * w + j]);
                    tmpHi[j] = 0.25f * (src[(size_t)row * w + j] - src[(size_t)(row + skip) * w + j]);
                }
                have = 1;
            } else if (row >= imax(h - skip, skip)) {
                for (int j = 0; j < w; j++) {
                    tmpLo[j] = 0.25f * (src[(size_t)row * w + j] + src[(size_t)(row - skip) * w + j]);
                    tmpHi[j] = 0.25f * (src[(size_t)row * w + j] - src[(size_t)(row - skip) * w + j]);
                }
                have = 1;
            }
            (void)have;
            for (int pass = 0; pass < 2; pass++) {  /* AnalysisFilterHaarHorizontal, L205-221 */
                const float* t = pass ? tmpHi : tmpLo;
                float* dLo = pass ? L->band[2] : dst;
                float* dHi = pass ? L->band[3] : L->band[1];
                for (int i = 0; i < w - skip; i++) {
                    dLo[(size_t)row * w + i] = t[i] + t[i + skip];
                    dHi[(size_t)row * w + i] = t[i] - t[i + skip];
                }
                for (int i = imax(w - skip, skip); i < w; i++) {
                    dLo[(size_t)row * w + i] = t[i] + t[i - skip];
                    dHi[(size_t)row * w + i] = t[i] - t[i - skip];
                }
            }
        }
    }
    free(tmpLo);
}

void* artoracle_wavelet_new(const float* src, int W, int H, int maxlvl, int subsamp)
{
    if (maxlvl < 1 || maxlvl > MAXLEV) return NULL;
    wdec* d = (wdec*)calloc(1, sizeof(wdec));
    d->W = W; d->H = H; d->subsamp = subsamp; d->nlev = maxlvl;
    float* buffer[2];
    const size_t nb = (size_t)(W / 2 + 1) * (H / 2 + 1);
    /* the reference sizes its ping-pong buffers for a decimated level 0 (dec.h L159-175); an undecimated level 0
     * needs the full frame */
    const size_t nfull = (subsamp & 1) ? nb : (size_t)W * H;
    buffer[0] = (float*)calloc(nfull, sizeof(float));
    buffer[1] = (float*)calloc(nfull, sizeof(float));
    int bi = 0, w = W, h = H;
    for (int l = 0; l < maxlvl; l++) {
        wlevel* L = &d->lev[l];
        L->w = w; L->h = h;
        L->sub = (subsamp >> l) & 1;
        if (subsamp) {                               /* level.h L82-104 (skipcrop == 1) */
            L->skip = 1;
            for (int n = 0; n < l; n++) L->skip *= 2 - ((subsamp >> n) & 1);
        } else L->skip = 1 << l;
        L->w2 = L->sub ? (w + 1) / 2 : w;
        L->h2 = L->sub ? (h + 1) / 2 : h;
        float* blk = (float*)calloc(3 * (size_t)L->w2 * L->h2, sizeof(float));
        for (int j = 1; j < 4; j++) L->band[j] = blk + (size_t)L->w2 * L->h2 * (j - 1);
        if (l == 0) decompose_level(L, src, buffer[bi ^ 1]);
        else { bi ^= 1; decompose_level(L, buffer[bi], buffer[bi ^ 1]); }
        w = L->w2; h = L->h2;
    }
    d->coeff0 = buffer[bi ^ 1];
    free(buffer[bi]);
    return d;
}

int artoracle_wavelet_maxlevel(void* p) { return ((wdec*)p)->nlev; }
int artoracle_wavelet_level_W(void* p, int l) { return ((wdec*)p)->lev[l].w2; }
int artoracle_wavelet_level_H(void* p, int l) { return ((wdec*)p)->lev[l].h2; }
int artoracle_wavelet_level_stride(void* p, int l) { return ((wdec*)p)->lev[l].skip; }
float* artoracle_wavelet_band(void* p, int l, int dir) { wdec* d = (wdec*)p; return dir == 0 ? d->coeff0 : d->lev[l].band[dir]; }

static void reconstruct_level(wlevel* L, float* tmpLo, float* tmpHi, const float* src, float* dst, float blend)
{
    const int skip = L->skip;
    if (!L->sub) {
        const int w = L->w, h = L->h;
        for (int pass = 0; pass < 2; pass++) {      /* SynthesisFilterHaarHorizontal, L244-264; hi pair first (L743-744) */
            const float* lo = pass ? src : L->band[2];
            const float* hi = pass ? L->band[1] : L->band[3];
            float* o = pass ? tmpLo : tmpHi;
            for (int k = 0; k < h; k++) {
                for (int i = 0; i < skip; i++) o[(size_t)k * w + i] = lo[(size_t)k * w + i] + hi[(size_t)k * w + i];
                for (int i = skip; i < w; i++)
                    o[(size_t)k * w + i] = 0.5f * (lo[(size_t)k * w + i] + hi[(size_t)k * w + i] + lo[(size_t)k * w + i - skip] - hi[(size_t)k * w + i - skip]);
            }
        }
        for (int i = 0; i < skip; i++)              /* SynthesisFilterHaarVertical, L266-298 */
            for (int j = 0; j < w; j++) dst[(size_t)w * i + j] = tmpLo[(size_t)i * w + j] + tmpHi[(size_t)i * w + j];
        for (int i = skip; i < h; i++)
            for (int j = 0; j < w; j++)
                dst[(size_t)w * i + j] = 0.5f * (tmpLo[(size_t)i * w + j] + tmpHi[(size_t)i * w + j] + tmpLo[(size_t)(i - skip) * w + j] - tmpHi[(size_t)(i - skip) * w + j]);
        return;
    }
    /* decimated level: synthesis filters are the analysis filters reversed (dec.h L112-113) */
    float fLo[TAPS], fHi[TAPS];
    for (int i = 0; i < TAPS; i++) { fLo[i] = ANAL[0][TAPS - 1 - i]; fHi[i] = ANAL[1][TAPS - 1 - i]; }
    const int srcw = L->w2, dstw = L->w, srch = L->h2, dsth = L->h;
    const int shift = skip * (TAPS - OFFS - 1);
    for (int pass = 0; pass < 2; pass++) {          /* SynthesisFilterSubsampHorizontal, L447-515; hi pair first */
        const float* lo = pass ? src : L->band[2];
        const float* hi = pass ? L->band[1] : L->band[3];
        float* o = pass ? tmpLo : tmpHi;
        for (int k = 0; k < srch; k++)
            for (int i = 0; i < dstw; i++) {
                float tot = 0.f;
                const int i_src = (i + shift) / 2, begin = (i + shift) % 2;
                for (int j = begin, l = 0; j < TAPS; j += 2, l += skip) {
                    const int arg = imax(0, imin(i_src - l, srcw - 1));
                    tot += ((fLo[j] * lo[(size_t)k * srcw + arg] + fHi[j] * hi[(size_t)k * srcw + arg]));
                }
                o[(size_t)k * dstw + i] = tot;
            }
    }
    const float srcFactor = 1.f - blend;            /* SynthesisFilterSubsampVertical, L518-597 */
    for (int i = 0; i < dsth; i++) {
        const int i_src = (i + shift) / 2, begin = (i + shift) % 2;
        for (int k = 0; k < dstw; k++) {
            float tot = 0.f;
            for (int j = begin, l = 0; j < TAPS; j += 2, l += skip) {
                const size_t arg = (size_t)imax(0, imin(i_src - l, srch - 1)) * dstw + k;
                tot += ((fLo[j] * tmpLo[arg] + fHi[j] * tmpHi[arg]));
            }
            dst[(size_t)dstw * i + k] = dst[(size_t)dstw * i + k] * srcFactor + blend * 4.f * tot;
        }
    }
}

/* consumes the decomposition (the reference deletes the levels as it goes, dec.h L220-268) */
void artoracle_wavelet_reconstruct(void* p, float* dst, float blend)
{
    wdec* d = (wdec*)p;
    for (int l = d->nlev - 1; l > 0; l--) {
        wlevel* L = &d->lev[l];
        float* tmpHi = (float*)malloc(sizeof(float) * (size_t)L->w * L->h);
        /* tmpLo = wavcoeffs[2] of the level (dec.h L219): band[2] is consumed by the first horizontal call first */
        float* tmpLo = (float*)malloc(sizeof(float) * (size_t)L->w * L->h);
        reconstruct_level(L, tmpLo, tmpHi, d->coeff0, d->coeff0, 1.f);
        free(tmpLo); free(tmpHi);
    }
    wlevel* L = &d->lev[0];
    float* tmpLo = (float*)malloc(sizeof(float) * (size_t)L->w * L->h2 + 16);
    float* tmpHi = (float*)malloc(sizeof(float) * (size_t)L->w * L->h2 + 16);
    if (L->sub) reconstruct_level(L, tmpLo, tmpHi, d->coeff0, dst, blend);
    else { free(tmpLo); free(tmpHi); tmpLo = (float*)malloc(sizeof(float) * (size_t)L->w * L->h); tmpHi = (float*)malloc(sizeof(float) * (size_t)L->w * L->h); reconstruct_level(L, tmpLo, tmpHi, d->coeff0, dst, blend); }
    free(tmpLo); free(tmpHi);
}

void artoracle_wavelet_delete(void* p)
{
    wdec* d = (wdec*)p;
    if (!d) return;
    for (int l = 0; l < d->nlev; l++) free(d->lev[l].band[1]);
    free(d->coeff0);
    free(d);
}
