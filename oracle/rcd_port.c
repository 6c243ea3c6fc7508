/*
 * oracle/rcd_port.c -- CPU restatement of the reference's RCD demosaic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under art_b200/ links or calls this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do, and
 * only as the checker.
 *
 * Restates RawImageSource::rcd_demosaic (reference rtengine/rcd_demosaic.cc
 * L51-347) and RawImageSource::border_interpolate2 (rtengine/demosaic_algos.cc
 * L200-353) in plain C.  Pinned against oracle/_ref/libartref_det.so (the
 * reference's own function bodies compiled in place) by tests/test_oracle_rcd.py:
 * bit-exact on every output sample.
 *
 * Tile semantics that matter for exactness (see DESIGN.md "RCD"):
 *   - the reference walks 194x194 tiles at stride 176 (L82-87, L110-124) and
 *     writes tile-local rows/cols [9, T-9) (L305-316);
 *   - VH_Dir exists only on tile-local [4,T-4)x[4,C-4) (L149-166); the G
 *     interpolation on the first/last computed row/col reads it one step
 *     outside that range (L201), where the (zeroed) scratch holds 0.  That 0
 *     reaches the first and last output row/col of every tile, so the output
 *     depends on the tile grid.  We restate the deterministic variant: scratch
 *     is zero at the start of every tile (the stock code leaves stale values
 *     from the thread's previous tile on partial edge tiles).
 * Floating-point expression association follows the reference source exactly;
 * compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TS 194   /* tileSize  L84 */
#define TB 9     /* tileBorder == rcdBorder L82-83 */
#define TN 176   /* tileSizeN L85 */

static inline float sqrf_(float x) { return x * x; }
static inline float maxf_(float a, float b) { return a < b ? b : a; }   /* std::max */
static inline float lim01_(float a) { float m = a < 1.f ? a : 1.f; return 0.f < m ? m : 0.f; } /* rt_math.h L91-94: max(0, min(a,1)) */
static inline float intp_(float a, float b, float c) { return a * b + (1.f - a) * c; } /* rt_math.h L110-118 */

static inline unsigned fc_(unsigned filters, int row, int col)
{   /* rawimage.h L186-189 */
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}

/* colour-difference high-pass along a line with step s (L140, L151, L215-216) */
static inline float hpf_(const float* p, int s)
{
    return sqrf_((p[-3 * s] - p[-s] - p[s] + p[3 * s]) - 3.f * (p[-2 * s] + p[2 * s]) + 6.f * p[0]);
}

typedef struct {
    float *cfa, *vhpf, *hhpf, *vh, *lpf, *ppf, *qpf, *pq, *ch[3];
} rcd_scratch;

static void rcd_tile(int W, int H, unsigned filters, const float* raw, long rs,
                     float* R, float* G, float* B, long os, int r0, int c0, rcd_scratch* s)
{
    const float eps = 1e-5f, epssq = 1e-10f, scale = 65536.f;   /* L90-92 */
    const int rend = (r0 + TS < H) ? r0 + TS : H, cend = (c0 + TS < W) ? c0 + TS : W;
    const int T = rend - r0, C = cend - c0;
    if (r0 + TB == rend - TB || c0 + TB == cend - TB) return;   /* L114-121 */
    const int w1 = TS;
    const size_t n = (size_t)TS * TS;
    float *cfa = s->cfa, *vh = s->vh, *lpf = s->lpf, *pq = s->pq;
    float *rgb0 = s->ch[0], *rgb1 = s->ch[1], *rgb2 = s->ch[2];
    float* rgb[3] = {rgb0, rgb1, rgb2};
    memset(cfa, 0, n * sizeof(float)); memset(vh, 0, n * sizeof(float));
    memset(lpf, 0, n * sizeof(float)); memset(pq, 0, n * sizeof(float));
    memset(s->vhpf, 0, n * sizeof(float)); memset(s->hhpf, 0, n * sizeof(float));
    memset(s->ppf, 0, n * sizeof(float)); memset(s->qpf, 0, n * sizeof(float));
    for (int k = 0; k < 3; ++k) memset(rgb[k], 0, n * sizeof(float));

    /* fill: L126-132 */
    for (int r = 0; r < T; ++r) {
        const int k0 = fc_(filters, r0 + r, c0), k1 = fc_(filters, r0 + r, c0 + 1);
        for (int c = 0; c < C; ++c) {
            const float v = lim01_(raw[(long)(r0 + r) * rs + c0 + c] / scale);
            cfa[r * w1 + c] = v; rgb[k0][r * w1 + c] = v; rgb[k1][r * w1 + c] = v;
        }
    }
    /* step 1: squared HPFs (L138-155) then VH_Dir (L156-162) */
    for (int r = 3; r < T - 3; ++r)
        for (int c = 4; c < C - 4; ++c) s->vhpf[r * w1 + c] = hpf_(cfa + r * w1 + c, w1);
    for (int r = 4; r < T - 4; ++r)
        for (int c = 3; c < C - 3; ++c) s->hhpf[r * w1 + c] = hpf_(cfa + r * w1 + c, 1);
    for (int r = 4; r < T - 4; ++r)
        for (int c = 4; c < C - 4; ++c) {
            const int i = r * w1 + c;
            const float vs = maxf_(epssq, s->vhpf[i - w1] + s->vhpf[i] + s->vhpf[i + w1]);
            const float hs = maxf_(epssq, s->hhpf[i - 1] + s->hhpf[i] + s->hhpf[i + 1]);
            vh[i] = vs / (vs + hs);
        }
    /* step 2: low-pass at R/B sites (L169-175) */
    for (int r = 2; r < T - 2; ++r)
        for (int c = 2 + (fc_(filters, r0 + r, c0) & 1); c < C - 2; c += 2) {
            const int i = r * w1 + c;
            lpf[i] = cfa[i] + 0.5f * (cfa[i - w1] + cfa[i + w1] + cfa[i - 1] + cfa[i + 1])
                     + 0.25f * (cfa[i - w1 - 1] + cfa[i - w1 + 1] + cfa[i + w1 - 1] + cfa[i + w1 + 1]);
        }
    /* step 3: G at R/B (L178-206) */
    for (int r = 4; r < T - 4; ++r)
        for (int c = 4 + (fc_(filters, r0 + r, c0) & 1); c < C - 4; c += 2) {
            const int i = r * w1 + c;
            const float x = cfa[i];
            const float ng = eps + (fabsf(cfa[i - w1] - cfa[i + w1]) + fabsf(x - cfa[i - 2 * w1])) + (fabsf(cfa[i - w1] - cfa[i - 3 * w1]) + fabsf(cfa[i - 2 * w1] - cfa[i - 4 * w1]));
            const float sg = eps + (fabsf(cfa[i - w1] - cfa[i + w1]) + fabsf(x - cfa[i + 2 * w1])) + (fabsf(cfa[i + w1] - cfa[i + 3 * w1]) + fabsf(cfa[i + 2 * w1] - cfa[i + 4 * w1]));
            const float wg = eps + (fabsf(cfa[i - 1] - cfa[i + 1]) + fabsf(x - cfa[i - 2])) + (fabsf(cfa[i - 1] - cfa[i - 3]) + fabsf(cfa[i - 2] - cfa[i - 4]));
            const float eg = eps + (fabsf(cfa[i - 1] - cfa[i + 1]) + fabsf(x - cfa[i + 2])) + (fabsf(cfa[i + 1] - cfa[i + 3]) + fabsf(cfa[i + 2] - cfa[i + 4]));
            const float l = lpf[i];
            const float ne = cfa[i - w1] * (l + l) / (eps + l + lpf[i - 2 * w1]);
            const float se = cfa[i + w1] * (l + l) / (eps + l + lpf[i + 2 * w1]);
            const float we = cfa[i - 1] * (l + l) / (eps + l + lpf[i - 2]);
            const float ee = cfa[i + 1] * (l + l) / (eps + l + lpf[i + 2]);
            const float ve = (sg * ne + ng * se) / (ng + sg);
            const float he = (wg * ee + eg * we) / (eg + wg);
            const float cv = vh[i];
            const float nv = 0.25f * ((vh[i - w1 - 1] + vh[i - w1 + 1]) + (vh[i + w1 - 1] + vh[i + w1 + 1]));
            const float d = fabsf(0.5f - cv) < fabsf(0.5f - nv) ? nv : cv;
            rgb1[i] = intp_(d, he, ve);
        }
    /* step 4.0: diagonal HPFs on the (odd,odd)-from-3 lattice (L213-218) */
    for (int r = 3; r < T - 3; ++r)
        for (int c = 3; c < C - 3; c += 2) {
            const int i = r * w1 + c;
            s->ppf[i] = hpf_(cfa + i, w1 + 1);
            s->qpf[i] = hpf_(cfa + i, w1 - 1);
        }
    /* The reference stores P/Q_CDiff_Hpf at index indx/2, so a sample written for
     * column c is later found by any reader whose (indx/2) matches: column c and,
     * when c is odd, c-1 ... we keep full-resolution planes and resolve that
     * aliasing here: reader (r,c) with index k=(r*w1+c)/2 reads the sample whose
     * writer had the same k on the same row.  Writers sit on odd c = 3,5,..;
     * rows have even length (w1=194) so k never crosses rows. */
#define PQ_AT(plane, two_k) ((plane)[(two_k) + 1])   /* writer with half-index k sits at full index 2k+1 */
    /* step 4.1: PQ_Dir at R/B (L221-227) */
    for (int r = 4; r < T - 4; ++r)
        for (int c = 4 + (fc_(filters, r0 + r, c0) & 1); c < C - 4; c += 2) {
            const int i = r * w1 + c;
            /* indx3=(i-w1-1)/2, indx2=i/2, indx4=(i+w1-1)/2; P uses indx3, indx2, indx4+1; Q uses indx3+1, indx2, indx4 */
            const int k3 = ((i - w1 - 1) / 2) * 2, k2 = (i / 2) * 2, k4 = ((i + w1 - 1) / 2) * 2;
            const float ps = maxf_(epssq, PQ_AT(s->ppf, k3) + PQ_AT(s->ppf, k2) + PQ_AT(s->ppf, k4 + 2));
            const float qs = maxf_(epssq, PQ_AT(s->qpf, k3 + 2) + PQ_AT(s->qpf, k2) + PQ_AT(s->qpf, k4));
            pq[i] = ps / (ps + qs);
        }
    /* step 4.2: R at B, B at R (L230-258).  PQ_Dir shares its buffer with lpf in the
     * reference (L103): where PQ_Dir was not computed a reader finds the lpf sample
     * with the same half-index.  Those reads only feed tile rows/cols < 9 from the
     * edge, which are never written out; we read pq (0 there) instead. */
    for (int r = 4; r < T - 4; ++r)
        for (int c = 4 + (fc_(filters, r0 + r, c0) & 1); c < C - 4; c += 2) {
            const int i = r * w1 + c;
            const int k = 2 - (int)fc_(filters, r0 + r, c0 + c);
            float* ck = rgb[k];
            const float cv = pq[i];
            const float nv = 0.25f * (pq[i - w1 - 1] + pq[i - w1 + 1] + pq[i + w1 - 1] + pq[i + w1 + 1]);
            const float d = (fabsf(0.5f - cv) < fabsf(0.5f - nv)) ? nv : cv;
            const float nwg = eps + fabsf(ck[i - w1 - 1] - ck[i + w1 + 1]) + fabsf(ck[i - w1 - 1] - ck[i - 3 * w1 - 3]) + fabsf(rgb1[i] - rgb1[i - 2 * w1 - 2]);
            const float neg = eps + fabsf(ck[i - w1 + 1] - ck[i + w1 - 1]) + fabsf(ck[i - w1 + 1] - ck[i - 3 * w1 + 3]) + fabsf(rgb1[i] - rgb1[i - 2 * w1 + 2]);
            const float swg = eps + fabsf(ck[i - w1 + 1] - ck[i + w1 - 1]) + fabsf(ck[i + w1 - 1] - ck[i + 3 * w1 - 3]) + fabsf(rgb1[i] - rgb1[i + 2 * w1 - 2]);
            const float seg = eps + fabsf(ck[i - w1 - 1] - ck[i + w1 + 1]) + fabsf(ck[i + w1 + 1] - ck[i + 3 * w1 + 3]) + fabsf(rgb1[i] - rgb1[i + 2 * w1 + 2]);
            const float nwe = ck[i - w1 - 1] - rgb1[i - w1 - 1];
            const float nee = ck[i - w1 + 1] - rgb1[i - w1 + 1];
            const float swe = ck[i + w1 - 1] - rgb1[i + w1 - 1];
            const float see = ck[i + w1 + 1] - rgb1[i + w1 + 1];
            const float pe = (nwg * see + seg * nwe) / (nwg + seg);
            const float qe = (neg * swe + swg * nee) / (neg + swg);
            ck[i] = rgb1[i] + intp_(d, qe, pe);
        }
    /* step 4.3: R and B at G (L261-302) */
    for (int r = 4; r < T - 4; ++r)
        for (int c = 4 + (fc_(filters, r0 + r, c0 + 1) & 1); c < C - 4; c += 2) {
            const int i = r * w1 + c;
            const float cv = vh[i];
            const float nv = 0.25f * ((vh[i - w1 - 1] + vh[i - w1 + 1]) + (vh[i + w1 - 1] + vh[i + w1 + 1]));
            const float d = (fabsf(0.5f - cv) < fabsf(0.5f - nv)) ? nv : cv;
            const float g = rgb1[i];
            const float n1 = eps + fabsf(g - rgb1[i - 2 * w1]);
            const float s1 = eps + fabsf(g - rgb1[i + 2 * w1]);
            const float w1_ = eps + fabsf(g - rgb1[i - 2]);
            const float e1 = eps + fabsf(g - rgb1[i + 2]);
            const float gn = rgb1[i - w1], gs = rgb1[i + w1], gw = rgb1[i - 1], ge = rgb1[i + 1];
            for (int k = 0; k <= 2; k += 2) {
                const float* ck = rgb[k];
                const float sn = fabsf(ck[i - w1] - ck[i + w1]);
                const float ew = fabsf(ck[i - 1] - ck[i + 1]);
                const float ng = n1 + sn + fabsf(ck[i - w1] - ck[i - 3 * w1]);
                const float sg = s1 + sn + fabsf(ck[i + w1] - ck[i + 3 * w1]);
                const float wg = w1_ + ew + fabsf(ck[i - 1] - ck[i - 3]);
                const float eg = e1 + ew + fabsf(ck[i + 1] - ck[i + 3]);
                const float ne = ck[i - w1] - gn, se = ck[i + w1] - gs, we = ck[i - 1] - gw, ee = ck[i + 1] - ge;
                const float ve = (ng * se + sg * ne) / (ng + sg);
                const float he = (eg * we + wg * ee) / (eg + wg);
                rgb[k][i] = g + intp_(d, he, ve);
            }
        }
    /* write-out: L305-316 (both the edge and the interior margin are 9) */
    for (int r = TB; r < T - TB; ++r)
        for (int c = TB; c < C - TB; ++c) {
            const int i = r * w1 + c;
            const long o = (long)(r0 + r) * os + c0 + c;
            R[o] = maxf_(0.f, rgb0[i] * scale);
            G[o] = maxf_(0.f, rgb1[i] * scale);
            B[o] = maxf_(0.f, rgb2[i] * scale);
        }
}

/* demosaic_algos.cc L200-353: 3x3 same-colour mean on the outer `bord` ring.
 * Accumulation order is row-major over the 3x3 window (i1 outer, j1 inner). */
void artoracle_border_interpolate2(int W, int H, unsigned filters, int bord, const float* raw, long rs,
                                   float* R, float* G, float* B, long os)
{
    for (int i = 0; i < H; ++i) {
        const int full_row = (i < bord) || (i >= H - bord);
        for (int j = 0; j < W; ++j) {
            if (!full_row && j >= bord && j < W - bord) { j = W - bord - 1; continue; }
            float sum[6] = {0, 0, 0, 0, 0, 0};
            for (int i1 = i - 1; i1 < i + 2; ++i1)
                for (int j1 = j - 1; j1 < j + 2; ++j1)
                    if (i1 > -1 && i1 < H && j1 > -1 && j1 < W) {
                        const int k = fc_(filters, i1, j1);
                        sum[k] += raw[(long)i1 * rs + j1];
                        sum[k + 3] += 1.f;
                    }
            const int k = fc_(filters, i, j);
            const long o = (long)i * os + j;
            const float x = raw[(long)i * rs + j];
            if (k == 1) { R[o] = sum[0] / sum[3]; G[o] = x; B[o] = sum[2] / sum[5]; }
            else {
                G[o] = sum[1] / sum[4];
                if (k == 0) { R[o] = x; B[o] = sum[2] / sum[5]; }
                else { R[o] = sum[0] / sum[3]; B[o] = x; }
            }
        }
    }
}

/* strides in floats */
int artoracle_rcd(int W, int H, unsigned filters, const float* raw, long rs,
                  float* R, float* G, float* B, long os)
{
    const int nth = H / TN + ((H % TN) ? 1 : 0), ntw = W / TN + ((W % TN) ? 1 : 0);   /* L86-87 */
    int fail = 0;
#pragma omp parallel
    {
        rcd_scratch s;
        const size_t n = (size_t)TS * TS;
        float* pool = (float*)calloc(11 * n, sizeof(float));
        if (!pool) {
#pragma omp atomic write
            fail = 1;
        } else {
            float* p = pool;
            s.cfa = p; s.vhpf = p + n; s.hhpf = p + 2 * n; s.vh = p + 3 * n; s.lpf = p + 4 * n;
            s.ppf = p + 5 * n; s.qpf = p + 6 * n; s.pq = p + 7 * n;
            s.ch[0] = p + 8 * n; s.ch[1] = p + 9 * n; s.ch[2] = p + 10 * n;
#pragma omp for collapse(2) schedule(dynamic, 2)
            for (int tr = 0; tr < nth; ++tr)
                for (int tc = 0; tc < ntw; ++tc)
                    rcd_tile(W, H, filters, raw, rs, R, G, B, os, tr * TN, tc * TN, &s);
            free(pool);
        }
    }
    if (fail) return 1;
    artoracle_border_interpolate2(W, H, filters, TB, raw, rs, R, G, B, os);   /* L342 */
    return 0;
}
