/*
 * oracle/greeneq_port.c -- CPU restatement of the reference's Bayer green equilibration (preprocess).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates RawImageSource::green_equilibrate_global (reference rtengine/green_equil_RT.cc L37-89: the two green phases are scaled to
 * their common mean, sums in double) and RawImageSource::green_equilibrate (L92-250: where the two green populations around a green
 * site differ by more than the local texture explains, the site is pulled half way to a gradient-weighted diagonal interpolation).
 * The SSE2 groups (8 columns from cc = 5 - (FC(rr, 2) & 1) while cc < width - 12) and the scalar tail differ in the association of
 * the two four-term sums d1, d2 and in the condition's product order; both are kept per column.
 * Pinned bit-exact against the reference's own functions compiled in place (oracle/_ref) in tests/test_oracle_greeneq.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline unsigned fc_(unsigned filters, int row, int col) { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }

int artoracle_green_equilibrate_global(float* raw, int W, int H, unsigned filters, int border)
{
    int ng1 = 0, ng2 = 0;
    double avgg1 = 0., avgg2 = 0.;
    for (int i = border; i < H - border; i++) {
        double avgg = 0.;
        for (int j = border + ((fc_(filters, i, border) & 1) ^ 1); j < W - border; j += 2) avgg += raw[(size_t)i * W + j];
        const int ng = (W - 2 * border + (fc_(filters, i, border) & 1)) / 2;
        if (i & 1) { avgg2 += avgg; ng2 += ng; } else { avgg1 += avgg; ng1 += ng; }
    }
    if (ng1 == 0 || avgg1 == 0.0) { ng1 = 1; avgg1 = 1.0; }
    if (ng2 == 0 || avgg2 == 0.0) { ng2 = 1; avgg2 = 1.0; }
    const double corrg1 = (avgg1 / ng1 + avgg2 / ng2) / 2.0 / (avgg1 / ng1);
    const double corrg2 = (avgg1 / ng1 + avgg2 / ng2) / 2.0 / (avgg2 / ng2);
    for (int i = border; i < H - border; i++) {
        const double corrg = (i & 1) ? corrg2 : corrg1;
        for (int j = border + ((fc_(filters, i, border) & 1) ^ 1); j < W - border; j += 2) raw[(size_t)i * W + j] *= corrg;
    }
    return 0;
}

/* thresh: GreenEqulibrateThreshold's constant (0.01 * greenthresh); thresh_map (optional, W x H) stands for a derived threshold class */
int artoracle_green_equilibrate(float* raw, int W, int H, unsigned filters, float thresh, const float* thresh_map)
{
    const int height = H, width = W;
    const int cw = width / 2 + (width & 1);
    float* cfa = (float*)calloc((size_t)cw * height, sizeof(float));
    if (!cfa) return 1;
    for (int i = 0; i < height; ++i)
        for (int j = (fc_(filters, i, 0) & 1) ^ 1; j < width; j += 2) cfa[(size_t)i * cw + (j >> 1)] = raw[(size_t)i * W + j];
#define C(r, x) cfa[(size_t)(r) * cw + (x)]
#define TH(r, c) (thresh_map ? thresh_map[(size_t)(r) * W + (c)] : thresh)
    const float eps = 1.f;
#pragma omp parallel for
    for (int rr = 4; rr < height - 4; rr++) {
        const int c0 = 5 - (fc_(filters, rr, 2) & 1);
        /* columns taken by the 8-wide SSE2 loop: c0, c0 + 8, ... while cc < width - 12; each iteration covers cc, cc+2, cc+4, cc+6 */
        int vec_end = c0;
        while (vec_end < width - 12) vec_end += 8;
        for (int cc = c0; cc < width - 6; cc += 2) {
            const int vec = cc < vec_end;
            const float o1_1 = C(rr - 1, (cc - 1) >> 1), o1_2 = C(rr - 1, (cc + 1) >> 1), o1_3 = C(rr + 1, (cc - 1) >> 1), o1_4 = C(rr + 1, (cc + 1) >> 1);
            const float o2_1 = C(rr - 2, cc >> 1), o2_2 = C(rr + 2, cc >> 1), o2_3 = C(rr, (cc >> 1) - 1), o2_4 = C(rr, (cc >> 1) + 1);
            float d1, d2;
            if (vec) { d1 = ((o1_1 + o1_2) + o1_3) + o1_4; d2 = ((o2_1 + o2_2) + o2_3) + o2_4; }
            else { d1 = (o1_1 + o1_2) + (o1_3 + o1_4); d2 = (o2_1 + o2_2) + (o2_3 + o2_4); }
            const float c1 = (fabsf(o1_1 - o1_2) + fabsf(o1_1 - o1_3) + fabsf(o1_1 - o1_4) + fabsf(o1_2 - o1_3) + fabsf(o1_3 - o1_4) + fabsf(o1_2 - o1_4));
            const float c2 = (fabsf(o2_1 - o2_2) + fabsf(o2_1 - o2_3) + fabsf(o2_1 - o2_4) + fabsf(o2_2 - o2_3) + fabsf(o2_3 - o2_4) + fabsf(o2_2 - o2_4));
            const float tf = TH(rr, cc);
            int hit;
            if (vec) hit = (c1 + c2) < (6.f * tf) * fabsf(d1 - d2);
            else hit = (c1 + c2) < 6 * tf * fabsf(d1 - d2);
            if (!hit) continue;
            const float gin = C(rr, cc >> 1);
            const float gmp2p2 = gin - C(rr + 2, (cc >> 1) + 1), gmm2m2 = gin - C(rr - 2, (cc >> 1) - 1);
            const float gmm2p2 = gin - C(rr - 2, (cc >> 1) + 1), gmp2m2 = gin - C(rr + 2, (cc >> 1) - 1);
            const float gse = o1_4 + 0.5f * gmp2p2, gnw = o1_1 + 0.5f * gmm2m2, gne = o1_2 + 0.5f * gmm2p2, gsw = o1_3 + 0.5f * gmp2m2;
            const float t1 = C(rr + 3, (cc + 3) >> 1) - o1_4, t2 = C(rr - 3, (cc - 3) >> 1) - o1_1, t3 = C(rr - 3, (cc + 3) >> 1) - o1_2, t4 = C(rr + 3, (cc - 3) >> 1) - o1_3;
            const float wtse = 1.f / (eps + gmp2p2 * gmp2p2 + t1 * t1);
            const float wtnw = 1.f / (eps + gmm2m2 * gmm2m2 + t2 * t2);
            const float wtne = 1.f / (eps + gmm2p2 * gmm2p2 + t3 * t3);
            const float wtsw = 1.f / (eps + gmp2m2 * gmp2m2 + t4 * t4);
            const float ginterp = (gse * wtse + gnw * wtnw + gne * wtne + gsw * wtsw) / (wtse + wtnw + wtne + wtsw);
            if (ginterp - gin < tf * (ginterp + gin)) raw[(size_t)rr * W + cc] = 0.5f * (ginterp + gin);
        }
    }
#undef C
#undef TH
    free(cfa);
    return 0;
}
