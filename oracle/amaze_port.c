/*
 * oracle/amaze_port.c -- CPU restatement of the reference's AMaZE demosaic (SSE2 code path).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Restates RawImageSource::amaze_demosaic_RT (reference rtengine/amaze_demosaic_RT.cc L41-1595),
 * the `#ifdef __SSE2__` branches -- what every x86-64 build of the reference executes -- as scalar
 * C that walks the same 4-lane groups.  The scalar (#else) branches of the reference are NOT
 * numerically equivalent to its SSE2 branches and are not what we follow.
 *
 * Exactness notes (all verified against oracle/_ref/libartref_det.so, tests/test_oracle_amaze.py):
 *  - fixed tile grid: 160x160 tiles at stride 128 from (-16,-16) (L182-183), 16 px mirrored pad;
 *  - three in-place passes: hcd/vcd variance select (L536-581; hcd has the SSE "lanes 0,1 see the
 *    previous vector's new values, lanes 2,3 see old values" semantics, vcd is a true recurrence down
 *    rows of equal parity), hvwt (L958-974) and pmwt (L1213-1223) are recurrences down rows;
 *  - the per-thread scratch is ONE block whose sub-buffers alias each other (L124-174); some passes
 *    read cells that were last written under another name (e.g. Dgrb[1] over the second half of
 *    vcdalt, Dgrb2 over dgintv, rbm over vcd, pmwt over delhvsqsum).  We keep the same layout so those
 *    reads see the same bytes.  The scratch is zeroed at the start of every tile (the deterministic
 *    variant; the stock code leaves the previous tile's bytes, which makes a handful of pixels depend
 *    on the OpenMP schedule);
 *  - every vector loop also processes the lanes that overrun its nominal column range; we reproduce
 *    the loop bounds literally so the same cells get written.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TS 160
#define TSH 80
enum { v1 = TS, v2 = 2 * TS, v3 = 3 * TS, p1 = -TS + 1, p2 = -2 * TS + 2, p3 = -3 * TS + 3, m1 = TS + 1, m2 = 2 * TS + 2, m3 = 3 * TS + 3 };

/* scratch layout, bytes from the 64-byte aligned base (L129-174; cldf = 2 -> 128-byte gaps) */
#define FULL ((size_t)TS * TS * 4)
#define HALF ((size_t)TS * TSH * 4)
#define GAP 128
#define OFF_RGBGREEN 0
#define OFF_DELHVSQSUM (OFF_RGBGREEN + FULL + GAP)
#define OFF_DIRWTS0 (OFF_DELHVSQSUM + FULL + GAP)
#define OFF_DIRWTS1 (OFF_DIRWTS0 + FULL + GAP)
#define OFF_VCD (OFF_DIRWTS1 + FULL + GAP)
#define OFF_HCD (OFF_VCD + FULL + GAP)
#define OFF_VCDALT (OFF_HCD + FULL + GAP)
#define OFF_HCDALT (OFF_VCDALT + FULL + GAP)
#define OFF_CDDIFFSQ (OFF_HCDALT + FULL + GAP)
#define OFF_HVWT (OFF_CDDIFFSQ + FULL + 2 * GAP)
#define OFF_DGINTV (OFF_HVWT + HALF + GAP)          /* = Dgrb2 */
#define OFF_DGINTH (OFF_DGINTV + FULL + GAP)
#define OFF_DGRBSQ1M (OFF_DGINTH + FULL + GAP)
#define OFF_DGRBSQ1P (OFF_DGRBSQ1M + HALF + GAP)
#define OFF_CFA (OFF_DGRBSQ1P + HALF + GAP)
#define OFF_NYQUIST (OFF_CFA + FULL + GAP)
#define OFF_NYQUTEST (OFF_NYQUIST + (size_t)TS * TSH + GAP)
#define SCRATCH_BYTES (14 * (size_t)4 * TS * TS + (size_t)TS * TSH + 18 * GAP)

static inline unsigned fc_(unsigned filters, int row, int col)
{   /* rawimage.h L186-189 */
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}
static inline float sq(float x) { return x * x; }
static inline float vminf_(float a, float b) { return a < b ? a : b; }          /* _mm_min_ps */
static inline float vmaxf_(float a, float b) { return a > b ? a : b; }          /* _mm_max_ps */
static inline float vintpf_(float a, float b, float c) { return a * b + (1.f - a) * c; }  /* sleefsseavx.h L1435 */
static inline float vmedian_(float a, float b, float c) { return vmaxf_(vminf_(a, b), vminf_(c, vmaxf_(a, b))); } /* median.h L63-67 */
static inline float smedian_(float a, float b, float c)
{   /* median.h L54-57: std::max(std::min(a,b), std::min(c, std::max(a,b))) */
    const float mn = b < a ? b : a, mx = a < b ? b : a;
    const float t = mx < c ? mx : c;
    return mn < t ? t : mn;
}
static inline float stdmaxf_(float a, float b) { return a < b ? b : a; }
static inline int sat8_(int x) { return x > 127 ? 127 : (x < -128 ? -128 : x); }

typedef struct {
    char* base;
    float *rgbgreen, *delhvsqsum, *dirwts0, *dirwts1, *vcd, *hcd, *vcdalt, *hcdalt, *cddiffsq, *hvwt;
    float *Dgrb0, *Dgrb1, *delp, *delm, *rbint, *Dgrb2, *dgintv, *dginth, *Dgrbsq1m, *Dgrbsq1p, *cfa, *pmwt, *rbm, *rbp, *nyqutest;
    unsigned char *nyquist, *nyquist2;
} amz_t;

static void amz_bind(amz_t* a, char* data)
{
    a->base = data;
    a->rgbgreen = (float*)(data + OFF_RGBGREEN);
    a->delhvsqsum = (float*)(data + OFF_DELHVSQSUM);
    a->dirwts0 = (float*)(data + OFF_DIRWTS0);
    a->dirwts1 = (float*)(data + OFF_DIRWTS1);
    a->vcd = (float*)(data + OFF_VCD);
    a->hcd = (float*)(data + OFF_HCD);
    a->vcdalt = (float*)(data + OFF_VCDALT);
    a->hcdalt = (float*)(data + OFF_HCDALT);
    a->cddiffsq = (float*)(data + OFF_CDDIFFSQ);
    a->hvwt = (float*)(data + OFF_HVWT);
    a->Dgrb0 = a->vcdalt;                                   /* L148 */
    a->Dgrb1 = a->vcdalt + TS * TSH;
    a->delp = a->cddiffsq;                                  /* L150 */
    a->delm = (float*)((char*)a->delp + HALF + GAP);        /* L152 */
    a->rbint = a->delm;                                     /* L154 */
    a->dgintv = (float*)(data + OFF_DGINTV);
    a->Dgrb2 = a->dgintv;                                   /* L156-158: {h,v} pairs */
    a->dginth = (float*)(data + OFF_DGINTH);
    a->Dgrbsq1m = (float*)(data + OFF_DGRBSQ1M);
    a->Dgrbsq1p = (float*)(data + OFF_DGRBSQ1P);
    a->cfa = (float*)(data + OFF_CFA);
    a->pmwt = a->delhvsqsum;                                /* L167 */
    a->rbm = a->vcd;                                        /* L169 */
    a->rbp = (float*)((char*)a->rbm + HALF + GAP);          /* L170 */
    a->nyquist = (unsigned char*)(data + OFF_NYQUIST);
    a->nyquist2 = (unsigned char*)a->cddiffsq;              /* L173 */
    a->nyqutest = (float*)(data + OFF_NYQUTEST);
}

typedef struct {
    int W, H;
    unsigned filters;
    const float* raw; long rs;
    float *R, *G, *B; long os;
    float clip_pt, clip_pt8;
    int ex, ey;
} amz_job;

/* stop_after: 0 = run everything; otherwise return after the stage with that number (numbering of the
 * CUDA passes in art_b200/csrc/amaze.cu, used by tests to localise a divergence pass by pass). */
static void amaze_tile(const amz_job* J, amz_t* S, int top, int left, int stop_after)
{
    const int W = J->W, H = J->H;
    const unsigned F = J->filters;
    const float eps = 1e-5f, epssq = 1e-10f, arthresh = 0.75f, nyqthresh = 0.5f;
    const float gaussodd[4] = {0.14659727707323927f, 0.103592713382435f, 0.0732036125103057f, 0.0365543548389495f};
    const float gaussgrad[6] = {nyqthresh * 0.07384411893421103f, nyqthresh * 0.06207511968171489f, nyqthresh * 0.0521818194747806f,
                                nyqthresh * 0.03687419286733595f, nyqthresh * 0.03099732204057846f, nyqthresh * 0.018413194161458882f};
    const float gausseven[2] = {0.13719494435797422f, 0.05640252782101291f};
    const float gquinc[4] = {0.169917f, 0.108947f, 0.069855f, 0.0287182f};
    const float clip_pt = J->clip_pt, clip_pt8 = J->clip_pt8;
    float *cfa = S->cfa, *rgbgreen = S->rgbgreen, *dirwts0 = S->dirwts0, *dirwts1 = S->dirwts1, *delhvsqsum = S->delhvsqsum;
    float *vcd = S->vcd, *hcd = S->hcd, *vcdalt = S->vcdalt, *hcdalt = S->hcdalt, *cddiffsq = S->cddiffsq, *hvwt = S->hvwt;
    float *dgintv = S->dgintv, *dginth = S->dginth, *nyqutest = S->nyqutest;
    unsigned char *nyquist = S->nyquist, *nyquist2 = S->nyquist2;

    memset(S->base, 0, SCRATCH_BYTES);      /* deterministic variant (see header) */

    const int bottom = (top + TS < H + 16) ? top + TS : H + 16;     /* L186-192 */
    const int right = (left + TS < W + 16) ? left + TS : W + 16;
    const int rr1 = bottom - top, cc1 = right - left;
    const int rrmin = top < 0 ? 16 : 0, ccmin = left < 0 ? 16 : 0;   /* L195-198 */
    const int rrmax = bottom > H ? H - top : rr1, ccmax = right > W ? W - left : cc1;

    /* ---- tile fill with mirrored borders (L206-334), the nine loops in the reference's order.
     * Kept on purpose: (1) in the image-corner blocks the top rows / left columns are mirrored about
     * row/column 16 (rawData[winy + 32 - rr][...], L307, L323, L331), not about 0 as on the edges;
     * (2) when the image bottom (right) edge falls inside the last 16 rows (columns) of a full-size
     * tile, the 16 border rows (columns) run past row (column) 159: rows >= 160 land in the bytes that
     * follow cfa / rgbgreen in the scratch block (the nyquist flags / delhvsqsum), columns >= 160 wrap
     * into the next row.  The flat indexing below reproduces both. */
#define PUT(idx, v) do { const float v_ = (v) / 65535.f; cfa[idx] = v_; rgbgreen[idx] = v_; } while (0)
#define RAW(r, c) J->raw[(long)(r) * J->rs + (c)]
    if (rrmin > 0)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = ccmin, row = 32 - rr + top; cc < ccmax; cc++) PUT(rr * TS + cc, RAW(row, cc + left));
    for (int rr = rrmin; rr < rrmax; rr++)
        for (int cc = ccmin; cc < ccmax; cc++) PUT(rr * TS + cc, RAW(rr + top, cc + left));
    if (rrmax < rr1)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = ccmin; cc < ccmax; cc++) PUT((rrmax + rr) * TS + cc, RAW(H - rr - 2, left + cc));
    if (ccmin > 0)
        for (int rr = rrmin; rr < rrmax; rr++)
            for (int cc = 0; cc < 16; cc++) PUT(rr * TS + cc, RAW(rr + top, 32 - cc + left));
    if (ccmax < cc1)
        for (int rr = rrmin; rr < rrmax; rr++)
            for (int cc = 0; cc < 16; cc++) PUT(rr * TS + ccmax + cc, RAW(top + rr, W - cc - 2));
    if (rrmin > 0 && ccmin > 0)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = 0; cc < 16; cc++) PUT(rr * TS + cc, RAW(32 - rr, 32 - cc));
    if (rrmax < rr1 && ccmax < cc1)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = 0; cc < 16; cc++) PUT((rrmax + rr) * TS + ccmax + cc, RAW(H - rr - 2, W - cc - 2));
    if (rrmin > 0 && ccmax < cc1)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = 0; cc < 16; cc++) PUT(rr * TS + ccmax + cc, RAW(32 - rr, W - cc - 2));
    if (rrmax < rr1 && ccmin > 0)
        for (int rr = 0; rr < 16; rr++)
            for (int cc = 0; cc < 16; cc++) PUT((rrmax + rr) * TS + cc, RAW(H - rr - 2, 32 - cc));
#undef PUT
#undef RAW

    if (stop_after == 1) return;
    /* ---- horizontal and vertical gradients (L342-350) */
    for (int rr = 2; rr < rr1 - 2; rr++)
        for (int indx = rr * TS; indx < rr * TS + cc1; indx += 4)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + l;
                const float delh = fabsf(cfa[i + 1] - cfa[i - 1]);
                const float delv = fabsf(cfa[i + v1] - cfa[i - v1]);
                dirwts1[i] = eps + fabsf(cfa[i + 2] - cfa[i]) + fabsf(cfa[i] - cfa[i - 2]) + delh;
                dirwts0[i] = eps + fabsf(cfa[i + v2] - cfa[i]) + fabsf(cfa[i] - cfa[i - v2]) + delv;
                delhvsqsum[i] = sq(delh) + sq(delv);
            }

    if (stop_after == 2) return;
    /* ---- interpolate vertical and horizontal colour differences (L380-431) */
    for (int rr = 4; rr < rr1 - 4; rr++)
        for (int indx = rr * TS + 4; indx < rr * TS + cc1 - 7; indx += 4)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + l;
                const float sgn = (fc_(F, rr, 4 + l) & 1) ? -1.f : 1.f;   /* sgnv: +1 at R/B sites, -1 at G sites */
                const float cfav = cfa[i];
                const float cru = cfa[i - v1] * (dirwts0[i - v2] + dirwts0[i]) / (dirwts0[i - v2] * (eps + cfav) + dirwts0[i] * (eps + cfa[i - v2]));
                const float crd = cfa[i + v1] * (dirwts0[i + v2] + dirwts0[i]) / (dirwts0[i + v2] * (eps + cfav) + dirwts0[i] * (eps + cfa[i + v2]));
                const float crl = cfa[i - 1] * (dirwts1[i - 2] + dirwts1[i]) / (dirwts1[i - 2] * (eps + cfav) + dirwts1[i] * (eps + cfa[i - 2]));
                const float crr = cfa[i + 1] * (dirwts1[i + 2] + dirwts1[i]) / (dirwts1[i + 2] * (eps + cfav) + dirwts1[i] * (eps + cfa[i + 2]));
                const float guha = cfa[i - v1] + 0.5f * (cfav - cfa[i - v2]);
                const float gdha = cfa[i + v1] + 0.5f * (cfav - cfa[i + v2]);
                const float glha = cfa[i - 1] + 0.5f * (cfav - cfa[i - 2]);
                const float grha = cfa[i + 1] + 0.5f * (cfav - cfa[i + 2]);
                float guar = fabsf(1.f - cru) < arthresh ? cfav * cru : guha;
                float gdar = fabsf(1.f - crd) < arthresh ? cfav * crd : gdha;
                float glar = fabsf(1.f - crl) < arthresh ? cfav * crl : glha;
                float grar = fabsf(1.f - crr) < arthresh ? cfav * crr : grha;
                const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
                const float vwt = dirwts0[i - v1] / (dirwts0[i + v1] + dirwts0[i - v1]);
                const float Ginthha = vintpf_(hwt, grha, glha);
                const float Gintvha = vintpf_(vwt, gdha, guha);
                const float hcdaltv = sgn * (Ginthha - cfav);
                const float vcdaltv = sgn * (Gintvha - cfav);
                hcdalt[i] = hcdaltv;
                vcdalt[i] = vcdaltv;
                const int clip = (cfav > clip_pt8) || (Gintvha > clip_pt8) || (Ginthha > clip_pt8);
                if (clip) { guar = guha; gdar = gdha; glar = glha; grar = grha; }
                vcd[i] = clip ? vcdaltv : sgn * (vintpf_(vwt, gdar, guar) - cfav);
                hcd[i] = clip ? hcdaltv : sgn * (vintpf_(hwt, grar, glar) - cfav);
                dgintv[i] = vminf_(sq(guha - gdha), sq(guar - gdar));
                dginth[i] = vminf_(sq(glha - grha), sq(glar - grar));
            }

    if (stop_after == 3) return;
    /* ---- variance select + saturation bounds, in place (L536-581) */
    for (int rr = 4; rr < rr1 - 4; rr++)
        for (int indx = rr * TS + 4; indx < rr * TS + cc1 - 4; indx += 4) {
            float nh[4], nv[4], nd[4];
            for (int l = 0; l < 4; ++l) {       /* all loads of the vector happen before its stores */
                const int i = indx + l;
                const float sgn = (fc_(F, rr, 4 + l) & 1) ? -1.f : 1.f, nsgn = -sgn, sgn3 = sgn + sgn + sgn;
                float hcdv = hcd[i];
                const float hcdvar = sq(hcd[i - 2] - hcdv) + sq(hcd[i - 2] - hcd[i + 2]) + sq(hcdv - hcd[i + 2]);
                const float hcdaltv = hcdalt[i];
                const float hcdaltvar = sq(hcdalt[i - 2] - hcdaltv) + sq(hcdalt[i - 2] - hcdalt[i + 2]) + sq(hcdaltv - hcdalt[i + 2]);
                float vcdv = vcd[i];
                const float vcdvar = sq(vcd[i - v2] - vcdv) + sq(vcd[i - v2] - vcd[i + v2]) + sq(vcdv - vcd[i + v2]);
                const float vcdaltv = vcdalt[i];
                const float vcdaltvar = sq(vcdalt[i - v2] - vcdaltv) + sq(vcdalt[i - v2] - vcdalt[i + v2]) + sq(vcdaltv - vcdalt[i + v2]);
                hcdv = hcdaltvar < hcdvar ? hcdaltv : hcdv;
                vcdv = vcdaltvar < vcdvar ? vcdaltv : vcdv;
                const float c = cfa[i];
                {
                    const float Ginth = sgn * hcdv + c;
                    const float temp2 = sgn3 * hcdv;
                    const float hwt = 1.f + temp2 / (eps + Ginth + c);
                    const int hmask = nsgn * hcdv > 0.f;
                    const float old = hcdv;
                    const float temp = nsgn * (c - vmedian_(Ginth, cfa[i - 1], cfa[i + 1]));
                    hcdv = (temp2 < -(c + Ginth)) ? temp : vintpf_(hwt, hcdv, temp);
                    hcdv = hmask ? hcdv : old;
                    hcdv = (Ginth > clip_pt) ? temp : hcdv;
                }
                {
                    const float Gintv = sgn * vcdv + c;
                    const float temp2 = sgn3 * vcdv;
                    const float vwt = 1.f + temp2 / (eps + Gintv + c);
                    const int vmask = nsgn * vcdv > 0.f;
                    const float old = vcdv;
                    const float temp = nsgn * (c - vmedian_(Gintv, cfa[i - v1], cfa[i + v1]));
                    vcdv = (temp2 < -(c + Gintv)) ? temp : vintpf_(vwt, vcdv, temp);
                    vcdv = vmask ? vcdv : old;
                    vcdv = (Gintv > clip_pt) ? temp : vcdv;
                }
                nh[l] = hcdv; nv[l] = vcdv; nd[l] = sq(vcdv - hcdv);
            }
            for (int l = 0; l < 4; ++l) { hcd[indx + l] = nh[l]; vcd[indx + l] = nv[l]; cddiffsq[indx + l] = nd[l]; }
        }

    if (stop_after == 5) return;
    /* ---- hvwt (L681-723): 4 R/B sites (stride 2) per vector */
    for (int rr = 6; rr < rr1 - 6; rr++)
        for (int indx = rr * TS + 6 + (fc_(F, rr, 2) & 1); indx < rr * TS + cc1 - 6; indx += 8)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l;
                float temp = vcd[i];
                const float uave = temp + vcd[i - v1] + vcd[i - v2] + vcd[i - v3];
                const float dave = temp + vcd[i + v1] + vcd[i + v2] + vcd[i + v3];
                float Dgrbvvaru = sq(temp - uave) + sq(vcd[i - v1] - uave) + sq(vcd[i - v2] - uave) + sq(vcd[i - v3] - uave);
                float Dgrbvvard = sq(temp - dave) + sq(vcd[i + v1] - dave) + sq(vcd[i + v2] - dave) + sq(vcd[i + v3] - dave);
                const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);       /* vadivapb */
                const float vwt = dirwts0[i - v1] / (dirwts0[i - v1] + dirwts0[i + v1]);
                temp = hcd[i];
                const float lave = temp + (hcd[i - 3] + hcd[i - 2]) + hcd[i - 1];           /* vaddc2vfu(hcd[indx-3]) */
                const float rave = temp + (hcd[i + 1] + hcd[i + 2]) + hcd[i + 3];
                float Dgrbhvarl = sq(temp - lave) + sq(hcd[i - 1] - lave) + sq(hcd[i - 2] - lave) + sq(hcd[i - 3] - lave);
                float Dgrbhvarr = sq(temp - rave) + sq(hcd[i + 1] - rave) + sq(hcd[i + 2] - rave) + sq(hcd[i + 3] - rave);
                const float vcdvar = epssq + vintpf_(vwt, Dgrbvvard, Dgrbvvaru);
                const float hcdvar = epssq + vintpf_(hwt, Dgrbhvarr, Dgrbhvarl);
                Dgrbvvaru = dgintv[i - v1] + dgintv[i - v2];
                Dgrbvvard = dgintv[i + v1] + dgintv[i + v2];
                Dgrbhvarl = dginth[i - 2] + dginth[i - 1];                                   /* vaddc2vfu(dginth[indx-2]) */
                Dgrbhvarr = dginth[i + 1] + dginth[i + 2];
                const float vcdvar1 = epssq + dgintv[i] + vintpf_(vwt, Dgrbvvard, Dgrbvvaru);
                const float hcdvar1 = epssq + dginth[i] + vintpf_(hwt, Dgrbhvarr, Dgrbhvarl);
                const float varwt = hcdvar / (vcdvar + hcdvar);
                const float diffwt = hcdvar1 / (vcdvar1 + hcdvar1);
                const int dec = ((0.5f - varwt) * (0.5f - diffwt) > 0.f) && (fabsf(0.5f - diffwt) < fabsf(0.5f - varwt));
                hvwt[i >> 1] = dec ? varwt : diffwt;
            }

    if (stop_after == 6) return;
    /* ---- nyquist test value (L789-844): vector part then scalar remainder (different association!) */
    for (int rr = 6; rr < rr1 - 6; rr++) {
        int cc = 6 + (fc_(F, rr, 2) & 1);
        int indx = rr * TS + cc;
        for (; cc < cc1 - 7; cc += 8, indx += 8)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l;
                nyqutest[i >> 1] =
                    (gaussodd[0] * cddiffsq[i] +
                     gaussodd[1] * (cddiffsq[i - m1] + cddiffsq[i + p1] + cddiffsq[i - p1] + cddiffsq[i + m1]) +
                     gaussodd[2] * (cddiffsq[i - v2] + cddiffsq[i - 2] + cddiffsq[i + 2] + cddiffsq[i + v2]) +
                     gaussodd[3] * (cddiffsq[i - m2] + cddiffsq[i + p2] + cddiffsq[i - p2] + cddiffsq[i + m2])) -
                    (gaussgrad[0] * delhvsqsum[i] +
                     gaussgrad[1] * (delhvsqsum[i - v1] + delhvsqsum[i - 1] + delhvsqsum[i + 1] + delhvsqsum[i + v1]) +
                     gaussgrad[2] * (delhvsqsum[i - m1] + delhvsqsum[i + p1] + delhvsqsum[i - p1] + delhvsqsum[i + m1]) +
                     gaussgrad[3] * (delhvsqsum[i - v2] + delhvsqsum[i - 2] + delhvsqsum[i + 2] + delhvsqsum[i + v2]) +
                     gaussgrad[4] * (delhvsqsum[i - v2 - 1] + delhvsqsum[i - v2 + 1] + delhvsqsum[i - TS - 2] + delhvsqsum[i - TS + 2] +
                                     delhvsqsum[i + TS - 2] + delhvsqsum[i + TS + 2] + delhvsqsum[i + v2 - 1] + delhvsqsum[i + v2 + 1]) +
                     gaussgrad[5] * (delhvsqsum[i - m2] + delhvsqsum[i + p2] + delhvsqsum[i - p2] + delhvsqsum[i + m2]));
            }
        for (; cc < cc1 - 6; cc += 2, indx += 2) {
            const int i = indx;
            nyqutest[i >> 1] =
                (gaussodd[0] * cddiffsq[i] +
                 gaussodd[1] * (cddiffsq[i - m1] + cddiffsq[i + p1] + cddiffsq[i - p1] + cddiffsq[i + m1]) +
                 gaussodd[2] * (cddiffsq[i - v2] + cddiffsq[i - 2] + cddiffsq[i + 2] + cddiffsq[i + v2]) +
                 gaussodd[3] * (cddiffsq[i - m2] + cddiffsq[i + p2] + cddiffsq[i - p2] + cddiffsq[i + m2])) -
                (gaussgrad[0] * delhvsqsum[i] +
                 gaussgrad[1] * (delhvsqsum[i - v1] + delhvsqsum[i + 1] + delhvsqsum[i - 1] + delhvsqsum[i + v1]) +
                 gaussgrad[2] * (delhvsqsum[i - m1] + delhvsqsum[i + p1] + delhvsqsum[i - p1] + delhvsqsum[i + m1]) +
                 gaussgrad[3] * (delhvsqsum[i - v2] + delhvsqsum[i - 2] + delhvsqsum[i + 2] + delhvsqsum[i + v2]) +
                 gaussgrad[4] * (delhvsqsum[i - v2 - 1] + delhvsqsum[i - v2 + 1] + delhvsqsum[i - TS - 2] + delhvsqsum[i - TS + 2] +
                                 delhvsqsum[i + TS - 2] + delhvsqsum[i + TS + 2] + delhvsqsum[i + v2 - 1] + delhvsqsum[i + v2 + 1]) +
                 gaussgrad[5] * (delhvsqsum[i - m2] + delhvsqsum[i + p2] + delhvsqsum[i - p2] + delhvsqsum[i + m2]));
        }
    }

    /* ---- nyquist flags and their bounding box (L847-876) */
    int nystartrow = 0, nyendrow = 0, nystartcol = TS + 1, nyendcol = 0;
    for (int rr = 6; rr < rr1 - 6; rr++)
        for (int cc = 6 + (fc_(F, rr, 2) & 1), indx = rr * TS + cc; cc < cc1 - 6; cc += 2, indx += 2)
            if (nyqutest[indx >> 1] > 0.f) {
                nyquist[indx >> 1] = 1;
                nystartrow = nystartrow ? nystartrow : rr;
                nyendrow = rr;
                nystartcol = nystartcol > cc ? cc : nystartcol;
                nyendcol = nyendcol < cc ? cc : nyendcol;
            }
    if (stop_after == 7) return;
    const int doNyquist = nystartrow != nyendrow && nystartcol != nyendcol;
    if (doNyquist) {
        nyendrow++;
        nyendcol++;
        nystartcol -= (nystartcol & 1);
        nystartrow = nystartrow > 8 ? nystartrow : 8;
        nyendrow = nyendrow < rr1 - 8 ? nyendrow : rr1 - 8;
        nystartcol = nystartcol > 8 ? nystartcol : 8;
        nyendcol = nyendcol < cc1 - 8 ? nyendcol : cc1 - 8;
        memset(&nyquist2[4 * TSH], 0, (size_t)(TS - 8) * TSH);     /* L877 */
        /* majority vote over the 8 quincunx neighbours, 16 half-sites (32 columns) per vector (L885-900) */
        for (int rr = nystartrow; rr < nyendrow; rr++)
            for (int indx = rr * TS; indx < rr * TS + cc1; indx += 32)
                for (int b = 0; b < 16; ++b) {
                    /* _mm_adds_epi8 tree (signed, saturating) -- only matters when stray bytes sit in nyquist */
#define NQ(o) ((int)(signed char)nyquist[(((indx) + (o)) >> 1) + b])
#define ADDS(x, y) sat8_((x) + (y))
                    int t1 = ADDS(NQ(-v2), NQ(-m1));
                    int t2 = ADDS(NQ(p1), NQ(-2));
                    const int t3 = ADDS(NQ(2), NQ(-p1));
                    const int t4 = ADDS(NQ(m1), NQ(v2));
                    t1 = ADDS(t1, t3);
                    t2 = ADDS(t2, t4);
                    t1 = ADDS(t1, t2);
#undef NQ
#undef ADDS
                    unsigned char val = nyquist[(indx >> 1) + b];
                    val = t1 > 4 ? 1 : val;
                    val = t1 < 4 ? 0 : val;
                    nyquist2[(indx >> 1) + b] = val;
                }
        /* area interpolation in Nyquist regions (L918-953) */
        for (int rr = nystartrow; rr < nyendrow; rr++)
            for (int indx = rr * TS + nystartcol + (fc_(F, rr, 2) & 1); indx < rr * TS + nyendcol; indx += 2)
                if (nyquist2[indx >> 1]) {
                    float sumcfa = 0.f, sumh = 0.f, sumv = 0.f, sumsqh = 0.f, sumsqv = 0.f, areawt = 0.f;
                    for (int i = -6; i < 7; i += 2) {
                        int indx1 = indx + (i * TS) - 6;
                        for (int j = -6; j < 7; j += 2, indx1 += 2)
                            if (nyquist2[indx1 >> 1]) {
                                const float cfatemp = cfa[indx1];
                                sumcfa += cfatemp;
                                sumh += (cfa[indx1 - 1] + cfa[indx1 + 1]);
                                sumv += (cfa[indx1 - v1] + cfa[indx1 + v1]);
                                sumsqh += sq(cfatemp - cfa[indx1 - 1]) + sq(cfatemp - cfa[indx1 + 1]);
                                sumsqv += sq(cfatemp - cfa[indx1 - v1]) + sq(cfatemp - cfa[indx1 + v1]);
                                areawt += 1;
                            }
                    }
                    sumh = sumcfa - 0.5f * sumh;         /* xdiv2f */
                    sumv = sumcfa - 0.5f * sumv;
                    areawt = 0.5f * areawt;
                    const float hcdvar = epssq + fabsf(areawt * sumsqh - sumh * sumh);
                    const float vcdvar = epssq + fabsf(areawt * sumsqv - sumv * sumv);
                    hvwt[indx >> 1] = hcdvar / (vcdvar + hcdvar);
                }
    }

    if (stop_after == 8) return;
    /* ---- populate G at R/B sites; in-place hvwt recurrence down the rows (L958-974) */
    float* Dgrb0 = S->Dgrb0; float* Dgrb1 = S->Dgrb1; float* Dgrb2 = S->Dgrb2;
    for (int rr = 8; rr < rr1 - 8; rr++)
        for (int indx = rr * TS + 8 + (fc_(F, rr, 2) & 1); indx < rr * TS + cc1 - 8; indx += 2) {
            const float hvwtalt = 0.25f * (hvwt[(indx - m1) >> 1] + hvwt[(indx + p1) >> 1] + hvwt[(indx - p1) >> 1] + hvwt[(indx + m1) >> 1]);  /* xdivf(.,2) */
            hvwt[indx >> 1] = fabsf(0.5f - hvwt[indx >> 1]) < fabsf(0.5f - hvwtalt) ? hvwtalt : hvwt[indx >> 1];
            Dgrb0[indx >> 1] = hvwt[indx >> 1] * vcd[indx] + (1.f - hvwt[indx >> 1]) * hcd[indx];       /* intp, rt_math.h L110 */
            rgbgreen[indx] = cfa[indx] + Dgrb0[indx >> 1];
            Dgrb2[2 * (indx >> 1)] = nyquist2[indx >> 1] ? sq(rgbgreen[indx] - 0.5f * (rgbgreen[indx - 1] + rgbgreen[indx + 1])) : 0.f;
            Dgrb2[2 * (indx >> 1) + 1] = nyquist2[indx >> 1] ? sq(rgbgreen[indx] - 0.5f * (rgbgreen[indx - v1] + rgbgreen[indx + v1])) : 0.f;
        }

    if (stop_after == 10) return;
    /* ---- refine Nyquist areas using G curvatures (L980-999) */
    if (doNyquist)
        for (int rr = nystartrow; rr < nyendrow; rr++)
            for (int indx = rr * TS + nystartcol + (fc_(F, rr, 2) & 1); indx < rr * TS + nyendcol; indx += 2)
                if (nyquist2[indx >> 1]) {
#define D2H(k) Dgrb2[2 * ((k) >> 1)]
#define D2V(k) Dgrb2[2 * ((k) >> 1) + 1]
                    const float gvarh = epssq + (gquinc[0] * D2H(indx) +
                                                 gquinc[1] * (D2H(indx - m1) + D2H(indx + p1) + D2H(indx - p1) + D2H(indx + m1)) +
                                                 gquinc[2] * (D2H(indx - v2) + D2H(indx - 2) + D2H(indx + 2) + D2H(indx + v2)) +
                                                 gquinc[3] * (D2H(indx - m2) + D2H(indx + p2) + D2H(indx - p2) + D2H(indx + m2)));
                    const float gvarv = epssq + (gquinc[0] * D2V(indx) +
                                                 gquinc[1] * (D2V(indx - m1) + D2V(indx + p1) + D2V(indx - p1) + D2V(indx + m1)) +
                                                 gquinc[2] * (D2V(indx - v2) + D2V(indx - 2) + D2V(indx + 2) + D2V(indx + v2)) +
                                                 gquinc[3] * (D2V(indx - m2) + D2V(indx + p2) + D2V(indx - p2) + D2V(indx + m2)));
                    Dgrb0[indx >> 1] = (hcd[indx] * gvarv + vcd[indx] * gvarh) / (gvarv + gvarh);
                    rgbgreen[indx] = cfa[indx] + Dgrb0[indx >> 1];
                }

    if (stop_after == 11) return;
    /* ---- diagonal gradients and colour-difference squares (L1004-1026) */
    float *delp = S->delp, *delm = S->delm, *Dgrbsq1m = S->Dgrbsq1m, *Dgrbsq1p = S->Dgrbsq1p;
    for (int rr = 6; rr < rr1 - 6; rr++) {
        const int odd = (fc_(F, rr, 2) & 1);       /* 0: R/B at even columns */
        for (int cc = 6, indx = rr * TS + cc; cc < cc1 - 6; cc += 8, indx += 8)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l;
                const int g = odd ? i : i + 1;      /* the G member of the column pair */
                const int x = odd ? i + 1 : i;      /* the R/B member */
                const float t = cfa[g];
                const float sp = sq(t - cfa[g - p1]) + sq(t - cfa[g + p1]);
                delp[i >> 1] = fabsf(cfa[x + p1] - cfa[x - p1]);
                delm[i >> 1] = fabsf(cfa[x + m1] - cfa[x - m1]);
                const float sm = sq(t - cfa[g - m1]) + sq(t - cfa[g + m1]);
                Dgrbsq1m[i >> 1] = sm;
                Dgrbsq1p[i >> 1] = sp;
            }
    }

    if (stop_after == 12) return;
    /* ---- diagonal interpolation correction: rbm, rbp, pmwt (L1057-1121) */
    float *rbm = S->rbm, *rbp = S->rbp, *pmwt = S->pmwt, *rbint = S->rbint;
    for (int rr = 8; rr < rr1 - 8; rr++)
        for (int indx = rr * TS + 8 + (fc_(F, rr, 2) & 1), indx1 = indx >> 1; indx < rr * TS + cc1 - 8; indx += 8, indx1 += 4)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l, k = indx1 + l;
                const float cfav = cfa[i];
                float t1 = cfa[i + m1], t2 = cfa[i + m2];
                float rbse = (t1 + t1) / (eps + cfav + t2);
                rbse = fabsf(1.f - rbse) < arthresh ? cfav * rbse : t1 + 0.5f * (cfav - t2);
                t1 = cfa[i - m1]; t2 = cfa[i - m2];
                float rbnw = (t1 + t1) / (eps + cfav + t2);
                rbnw = fabsf(1.f - rbnw) < arthresh ? cfav * rbnw : t1 + 0.5f * (cfav - t2);
                t1 = eps + delm[k];
                const float wtse = t1 + delm[(i + m1) >> 1] + delm[(i + m2) >> 1];
                const float wtnw = t1 + delm[(i - m1) >> 1] + delm[(i - m2) >> 1];
                const float rbmv = (wtse * rbnw + wtnw * rbse) / (wtse + wtnw);
                t1 = vmedian_(rbmv, cfa[i - m1], cfa[i + m1]);
                float wt = ((cfav - rbmv) + (cfav - rbmv)) / (eps + rbmv + cfav);
                t2 = vintpf_(wt, rbmv, t1);
                t2 = (rbmv + rbmv < cfav) ? t1 : t2;
                t2 = (rbmv < cfav) ? t2 : rbmv;
                rbm[k] = (t2 > clip_pt) ? vmedian_(t2, cfa[i - m1], cfa[i + m1]) : t2;

                t1 = cfa[i + p1]; t2 = cfa[i + p2];
                float rbne = (t1 + t1) / (eps + cfav + t2);
                rbne = fabsf(1.f - rbne) < arthresh ? cfav * rbne : t1 + 0.5f * (cfav - t2);
                t1 = cfa[i - p1]; t2 = cfa[i - p2];
                float rbsw = (t1 + t1) / (eps + cfav + t2);
                rbsw = fabsf(1.f - rbsw) < arthresh ? cfav * rbsw : t1 + 0.5f * (cfav - t2);
                t1 = eps + delp[k];
                const float wtne = t1 + delp[(i + p1) >> 1] + delp[(i + p2) >> 1];
                const float wtsw = t1 + delp[(i - p1) >> 1] + delp[(i - p2) >> 1];
                const float rbpv = (wtne * rbsw + wtsw * rbne) / (wtne + wtsw);
                t1 = vmedian_(rbpv, cfa[i - p1], cfa[i + p1]);
                wt = ((cfav - rbpv) + (cfav - rbpv)) / (eps + rbpv + cfav);
                t2 = vintpf_(wt, rbpv, t1);
                t2 = (rbpv + rbpv < cfav) ? t1 : t2;
                t2 = (rbpv < cfav) ? t2 : rbpv;
                rbp[k] = (t2 > clip_pt) ? vmedian_(t2, cfa[i - p1], cfa[i + p1]) : t2;

                const float rbvarm = epssq + (gausseven[0] * (Dgrbsq1m[(i - v1) >> 1] + Dgrbsq1m[(i - 1) >> 1] + Dgrbsq1m[(i + 1) >> 1] + Dgrbsq1m[(i + v1) >> 1]) +
                                              gausseven[1] * (Dgrbsq1m[(i - v2 - 1) >> 1] + Dgrbsq1m[(i - v2 + 1) >> 1] + Dgrbsq1m[(i - 2 - v1) >> 1] + Dgrbsq1m[(i + 2 - v1) >> 1] +
                                                              Dgrbsq1m[(i - 2 + v1) >> 1] + Dgrbsq1m[(i + 2 + v1) >> 1] + Dgrbsq1m[(i + v2 - 1) >> 1] + Dgrbsq1m[(i + v2 + 1) >> 1]));
                pmwt[k] = rbvarm / ((epssq + (gausseven[0] * (Dgrbsq1p[(i - v1) >> 1] + Dgrbsq1p[(i - 1) >> 1] + Dgrbsq1p[(i + 1) >> 1] + Dgrbsq1p[(i + v1) >> 1]) +
                                              gausseven[1] * (Dgrbsq1p[(i - v2 - 1) >> 1] + Dgrbsq1p[(i - v2 + 1) >> 1] + Dgrbsq1p[(i - 2 - v1) >> 1] + Dgrbsq1p[(i + 2 - v1) >> 1] +
                                                              Dgrbsq1p[(i - 2 + v1) >> 1] + Dgrbsq1p[(i + 2 + v1) >> 1] + Dgrbsq1p[(i + v2 - 1) >> 1] + Dgrbsq1p[(i + v2 + 1) >> 1]))) + rbvarm);
            }

    if (stop_after == 13) return;
    /* ---- in-place pmwt recurrence down the rows + rbint (L1213-1223) */
    for (int rr = 10; rr < rr1 - 10; rr++)
        for (int indx = rr * TS + 10 + (fc_(F, rr, 2) & 1), indx1 = indx >> 1; indx < rr * TS + cc1 - 10; indx += 8, indx1 += 4)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l, k = indx1 + l;
                const float pmwtalt = 0.25f * (pmwt[(i - m1) >> 1] + pmwt[(i + p1) >> 1] + pmwt[(i - p1) >> 1] + pmwt[(i + m1) >> 1]);
                float t = pmwt[k];
                t = fabsf(0.5f - t) < fabsf(0.5f - pmwtalt) ? pmwtalt : t;
                pmwt[k] = t;
                rbint[k] = 0.5f * (cfa[i] + vintpf_(t, rbp[k], rbm[k]));
            }

    if (stop_after == 15) return;
    /* ---- G via R+B where the diagonal weights discriminate better (L1241-1294) */
    for (int rr = 12; rr < rr1 - 12; rr++)
        for (int indx = rr * TS + 12 + (fc_(F, rr, 2) & 1), indx1 = indx >> 1; indx < rr * TS + cc1 - 12; indx += 8, indx1 += 4)
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l, k = indx1 + l;
                if (!(fabsf(0.5f - pmwt[k]) >= fabsf(0.5f - hvwt[k]))) continue;      /* copymask lane */
                const float rbintv = rbint[k];
                const float cru = (cfa[i - v1] + cfa[i - v1]) / (eps + rbintv + rbint[k - v1]);
                float gu = rbintv * cru;
                const float gu2 = cfa[i - v1] + 0.5f * (rbintv - rbint[k - v1]);
                gu = fabsf(1.f - cru) < arthresh ? gu : gu2;
                const float crd = (cfa[i + v1] + cfa[i + v1]) / (eps + rbintv + rbint[k + v1]);
                float gd = rbintv * crd;
                const float gd2 = cfa[i + v1] + 0.5f * (rbintv - rbint[k + v1]);
                gd = fabsf(1.f - crd) < arthresh ? gd : gd2;
                float Gintv = (dirwts0[i - v1] * gd + dirwts0[i + v1] * gu) / (dirwts0[i + v1] + dirwts0[i - v1]);
                float Gint1 = vmedian_(Gintv, cfa[i - v1], cfa[i + v1]);
                const float vwt = ((rbintv - Gintv) + (rbintv - Gintv)) / (eps + Gintv + rbintv);
                const float Gint2 = vintpf_(vwt, Gintv, Gint1);
                Gint1 = ((Gintv + Gintv) < rbintv) ? Gint1 : Gint2;
                Gintv = (Gintv < rbintv) ? Gint1 : Gintv;
                Gintv = (Gintv > clip_pt) ? vmedian_(Gintv, cfa[i - v1], cfa[i + v1]) : Gintv;

                const float crl = (cfa[i - 1] + cfa[i - 1]) / (eps + rbintv + rbint[k - 1]);
                float gl = rbintv * crl;
                const float gl2 = cfa[i - 1] + 0.5f * (rbintv - rbint[k - 1]);
                gl = fabsf(1.f - crl) < arthresh ? gl : gl2;
                const float crr = (cfa[i + 1] + cfa[i + 1]) / (eps + rbintv + rbint[k + 1]);
                float gr = rbintv * crr;
                const float gr2 = cfa[i + 1] + 0.5f * (rbintv - rbint[k + 1]);
                gr = fabsf(1.f - crr) < arthresh ? gr : gr2;
                float Ginth = (dirwts1[i - 1] * gr + dirwts1[i + 1] * gl) / (dirwts1[i - 1] + dirwts1[i + 1]);
                float Gint1h = vmedian_(Ginth, cfa[i - 1], cfa[i + 1]);
                const float hwt = ((rbintv - Ginth) + (rbintv - Ginth)) / (eps + Ginth + rbintv);
                const float Gint2h = vintpf_(hwt, Ginth, Gint1h);
                Gint1h = ((Ginth + Ginth) < rbintv) ? Gint1h : Gint2h;
                Ginth = (Ginth < rbintv) ? Gint1h : Ginth;
                Ginth = (Ginth > clip_pt) ? vmedian_(Ginth, cfa[i - 1], cfa[i + 1]) : Ginth;

                const float greenv = vintpf_(hvwt[k], Gintv, Ginth);
                rgbgreen[i] = greenv;
                Dgrb0[k] = greenv - cfa[i];
            }

    if (stop_after == 16) return;
    /* ---- split G-B from G-R at B sites (L1382-1386) */
    for (int rr = 13 - J->ey; rr < rr1 - 12; rr += 2)
        for (int indx1 = (rr * TS + 13 - J->ex) >> 1; indx1 < (rr * TS + cc1 - 12) >> 1; indx1++) {
            Dgrb1[indx1] = Dgrb0[indx1];
            Dgrb0[indx1] = 0;
        }

    if (stop_after == 17) return;
    /* ---- chrominance interpolation to the opposite R/B sites (L1394-1408) */
    for (int rr = 14; rr < rr1 - 14; rr++)
        for (int cc = 14 + (fc_(F, rr, 2) & 1), indx = rr * TS + cc, c = 1 - (int)fc_(F, rr, cc) / 2; cc < cc1 - 14; cc += 8, indx += 8) {
            float* D = c ? Dgrb1 : Dgrb0;
            float res[4];
            for (int l = 0; l < 4; ++l) {
                const int i = indx + 2 * l;
#define DG(o) D[(i + (o)) >> 1]
                const float tempv = eps + fabsf(DG(-m1) - DG(m1));
                const float temp2v = eps + fabsf(DG(p1) - DG(-p1));
                const float wtnw = 1.f / (tempv + fabsf(DG(-m1) - DG(-m3)) + fabsf(DG(m1) - DG(-m3)));
                const float wtne = 1.f / (temp2v + fabsf(DG(p1) - DG(p3)) + fabsf(DG(-p1) - DG(p3)));
                const float wtsw = 1.f / (temp2v + fabsf(DG(-p1) - DG(m3)) + fabsf(DG(p1) - DG(-p3)));
                const float wtse = 1.f / (tempv + fabsf(DG(m1) - DG(-p3)) + fabsf(DG(-m1) - DG(m3)));
                res[l] = (wtnw * (1.325f * DG(-m1) - 0.175f * DG(-m3) - 0.075f * (DG(-m1 - 2) + DG(-m1 - v2))) +
                          wtne * (1.325f * DG(p1) - 0.175f * DG(p3) - 0.075f * (DG(p1 + 2) + DG(p1 + v2))) +
                          wtsw * (1.325f * DG(-p1) - 0.175f * DG(-p3) - 0.075f * (DG(-p1 - 2) + DG(-p1 - v2))) +
                          wtse * (1.325f * DG(m1) - 0.175f * DG(m3) - 0.075f * (DG(m1 + 2) + DG(m1 + v2)))) / (wtnw + wtne + wtsw + wtse);
#undef DG
            }
            for (int l = 0; l < 4; ++l) D[(indx + 2 * l) >> 1] = res[l];
        }

    if (stop_after == 18) return;
    /* ---- write R,B (L1441-1548) and G (L1551-1565).  The vector body and the scalar tails compute the
     *      same expressions, so one scalar loop restates both. */
    for (int rr = 16; rr < rr1 - 16; rr++) {
        const int row = rr + top;
        const int gsite_odd = (fc_(F, rr, 2) & 1);      /* 1: G at even columns (FC(rr,2) is G) */
        for (int cc = 16; cc < cc1 - 16; ++cc) {
            const int indx = rr * TS + cc, col = left + cc;
            const int is_g = ((cc & 1) == 0) ? gsite_odd : !gsite_odd;
            float r, b;
            if (is_g) {
                const float temp = 1.f / (hvwt[(indx - v1) >> 1] + 2.f - hvwt[(indx + 1) >> 1] - hvwt[(indx - 1) >> 1] + hvwt[(indx + v1) >> 1]);
                r = rgbgreen[indx] - ((hvwt[(indx - v1) >> 1]) * Dgrb0[(indx - v1) >> 1] + (1.f - hvwt[(indx + 1) >> 1]) * Dgrb0[(indx + 1) >> 1] + (1.f - hvwt[(indx - 1) >> 1]) * Dgrb0[(indx - 1) >> 1] + (hvwt[(indx + v1) >> 1]) * Dgrb0[(indx + v1) >> 1]) * temp;
                b = rgbgreen[indx] - ((hvwt[(indx - v1) >> 1]) * Dgrb1[(indx - v1) >> 1] + (1.f - hvwt[(indx + 1) >> 1]) * Dgrb1[(indx + 1) >> 1] + (1.f - hvwt[(indx - 1) >> 1]) * Dgrb1[(indx - 1) >> 1] + (hvwt[(indx + v1) >> 1]) * Dgrb1[(indx + v1) >> 1]) * temp;
            } else {
                r = rgbgreen[indx] - Dgrb0[indx >> 1];
                b = rgbgreen[indx] - Dgrb1[indx >> 1];
            }
            J->R[(long)row * J->os + col] = stdmaxf_(0.f, 65535.f * r);
            J->B[(long)row * J->os + col] = stdmaxf_(0.f, 65535.f * b);
            J->G[(long)row * J->os + col] = stdmaxf_(0.f, 65535.f * rgbgreen[indx]);
        }
    }
    (void)vcdalt; (void)hcdalt; (void)smedian_;
}

void artoracle_border_interpolate2(int W, int H, unsigned filters, int bord, const float* raw, long rs,
                                   float* R, float* G, float* B, long os);

static void amz_job_init(amz_job* J, int W, int H, unsigned filters, const float* raw, long rs,
                         float* R, float* G, float* B, long os, float initialGain)
{
    J->W = W; J->H = H; J->filters = filters; J->raw = raw; J->rs = rs; J->R = R; J->G = G; J->B = B; J->os = os;
    J->clip_pt = (float)(1.0 / (double)initialGain);       /* L53-54: const float = 1.0 / initialGain (double) */
    J->clip_pt8 = (float)(0.8 / (double)initialGain);
    /* (ey,ex): offset of the R site in the 2x2 cell (L70-86) */
    if (fc_(filters, 0, 0) == 1) {
        if (fc_(filters, 0, 1) == 0) { J->ey = 0; J->ex = 1; } else { J->ey = 1; J->ex = 0; }
    } else {
        if (fc_(filters, 0, 0) == 0) { J->ey = 0; J->ex = 0; } else { J->ey = 1; J->ex = 1; }
    }
}

/* debug: the scratch block of tile `tile` (row-major over the tile grid) after stage `stop_after` */
size_t artoracle_amaze_slab_bytes(void) { return SCRATCH_BYTES; }
int artoracle_amaze_slab(int W, int H, unsigned filters, const float* raw, long rs, float initialGain,
                         int stop_after, int tile, void* out)
{
    amz_job J;
    amz_job_init(&J, W, H, filters, raw, rs, NULL, NULL, NULL, 0, initialGain);
    const int ntx = (W + 16 + (TS - 32) - 1) / (TS - 32);
    char* buf = (char*)malloc(SCRATCH_BYTES + 63);
    if (!buf) return 1;
    amz_t S;
    amz_bind(&S, (char*)(((uintptr_t)buf + 63) / 64 * 64));
    amaze_tile(&J, &S, -16 + (tile / ntx) * (TS - 32), -16 + (tile % ntx) * (TS - 32), stop_after);
    memcpy(out, S.base, SCRATCH_BYTES);
    free(buf);
    return 0;
}

/* strides in floats */
int artoracle_amaze(int W, int H, unsigned filters, const float* raw, long rs,
                    float* R, float* G, float* B, long os, float initialGain, int border)
{
    amz_job J;
    amz_job_init(&J, W, H, filters, raw, rs, R, G, B, os, initialGain);
    const int nty = (H + 16 + (TS - 32) - 1) / (TS - 32), ntx = (W + 16 + (TS - 32) - 1) / (TS - 32);   /* top=-16+128i < H */
    int fail = 0;
#pragma omp parallel
    {
        char* buf = (char*)malloc(SCRATCH_BYTES + 63);
        if (!buf) {
#pragma omp atomic write
            fail = 1;
        } else {
            amz_t S;
            amz_bind(&S, (char*)(((uintptr_t)buf + 63) / 64 * 64));
#pragma omp for collapse(2) schedule(dynamic)
            for (int ty = 0; ty < nty; ++ty)
                for (int tx = 0; tx < ntx; ++tx) {
                    const int top = -16 + ty * (TS - 32), left = -16 + tx * (TS - 32);
                    if (top < H && left < W) amaze_tile(&J, &S, top, left, 0);
                }
            free(buf);
        }
    }
    if (fail) return 1;
    if (border < 4) artoracle_border_interpolate2(W, H, filters, 3, raw, rs, R, G, B, os);   /* L1587-1589 */
    return 0;
}
