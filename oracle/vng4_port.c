/*
 * oracle/vng4_port.c -- CPU restatement of the reference's VNG4 Bayer demosaic (the flat-region demosaicer of the dual methods).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates RawImageSource::vng4_demosaic and vng4interpolate_row_redblue (reference rtengine/vng4_demosaic_RT.cc L32-397), four-colour
 * variable-number-of-gradients after dcraw:
 *   1. bilinear fill of the three missing channels of every interior pixel of a 4-channel image (the two greens are colours 1 and 3
 *      of `prefilters`), weights 1 / 2 by distance, sums in raster order of the 3x3 neighbourhood (L112-222);
 *   2. for rows / columns 2 .. n-3: eight directional gradients from the 64-term table (terms that do not join two samples of one
 *      colour, or that lie on the Bayer diagonal, drop out per phase; L224-281), threshold min + max / 2, and the green value from
 *      the directions under it (L300-349) -- only GREEN is kept;
 *   3. red / blue for rows / columns 3 .. n-4 from colour differences against that green (L32-55), clamped at 0;
 *   4. border_interpolate2(W, H, 3) (amaze_port.c).
 * Every pixel of step 2 reads only step-1 values, so the reference's row chunking does not show in the result.
 * Pinned bit-exact against the reference's own function compiled in place (oracle/_ref) in tests/test_oracle_vng4.py.
 * Compile with -ffp-contract=off.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void artoracle_border_interpolate2(int W, int H, unsigned filters, int bord, const float* raw, long rs, float* R, float* G, float* B, long os);

static const signed char TERMS[64 * 6] = {
    -2, -2, +0, -1, 0, 0x01, -2, -2, +0, +0, 1, 0x01, -2, -1, -1, +0, 0, 0x01,
    -2, -1, +0, -1, 0, 0x02, -2, -1, +0, +0, 0, 0x03, -2, -1, +0, +1, 1, 0x01,
    -2, +0, +0, -1, 0, 0x06, -2, +0, +0, +0, 1, 0x02, -2, +0, +0, +1, 0, 0x03,
    -2, +1, -1, +0, 0, 0x04, -2, +1, +0, -1, 1, 0x04, -2, +1, +0, +0, 0, 0x06,
    -2, +1, +0, +1, 0, 0x02, -2, +2, +0, +0, 1, 0x04, -2, +2, +0, +1, 0, 0x04,
    -1, -2, -1, +0, 0, (signed char)0x80, -1, -2, +0, -1, 0, 0x01, -1, -2, +1, -1, 0, 0x01,
    -1, -2, +1, +0, 1, 0x01, -1, -1, -1, +1, 0, (signed char)0x88, -1, -1, +1, -2, 0, 0x40,
    -1, -1, +1, -1, 0, 0x22, -1, -1, +1, +0, 0, 0x33, -1, -1, +1, +1, 1, 0x11,
    -1, +0, -1, +2, 0, 0x08, -1, +0, +0, -1, 0, 0x44, -1, +0, +0, +1, 0, 0x11,
    -1, +0, +1, -2, 1, 0x40, -1, +0, +1, -1, 0, 0x66, -1, +0, +1, +0, 1, 0x22,
    -1, +0, +1, +1, 0, 0x33, -1, +0, +1, +2, 1, 0x10, -1, +1, +1, -1, 1, 0x44,
    -1, +1, +1, +0, 0, 0x66, -1, +1, +1, +1, 0, 0x22, -1, +1, +1, +2, 0, 0x10,
    -1, +2, +0, +1, 0, 0x04, -1, +2, +1, +0, 1, 0x04, -1, +2, +1, +1, 0, 0x04,
    +0, -2, +0, +0, 1, (signed char)0x80, +0, -1, +0, +1, 1, (signed char)0x88, +0, -1, +1, -2, 0, 0x40,
    +0, -1, +1, +0, 0, 0x11, +0, -1, +2, -2, 0, 0x40, +0, -1, +2, -1, 0, 0x20,
    +0, -1, +2, +0, 0, 0x30, +0, -1, +2, +1, 1, 0x10, +0, +0, +0, +2, 1, 0x08,
    +0, +0, +2, -2, 1, 0x40, +0, +0, +2, -1, 0, 0x60, +0, +0, +2, +0, 1, 0x20,
    +0, +0, +2, +1, 0, 0x30, +0, +0, +2, +2, 1, 0x10, +0, +1, +1, +0, 0, 0x44,
    +0, +1, +1, +2, 0, 0x10, +0, +1, +2, -1, 1, 0x40, +0, +1, +2, +0, 0, 0x60,
    +0, +1, +2, +1, 0, 0x20, +0, +1, +2, +2, 0, 0x10, +1, -2, +1, +0, 0, (signed char)0x80,
    +1, -1, +1, +1, 0, (signed char)0x88, +1, +0, +1, +2, 0, 0x08, +1, +0, +2, -1, 0, 0x40,
    +1, +0, +2, +1, 0, 0x10
};
static const signed char CHOOD[16] = {-1, -1, -1, 0, -1, +1, 0, +1, +1, +1, +1, 0, +1, -1, 0, -1};

static inline unsigned fc4(unsigned f, int row, int col) { return (f >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
static inline float max0(float v) { return 0.f < v ? v : 0.f; }      /* std::max(0.f, v) */

typedef struct { int o1, o2; float w; int g1, g2; } vterm;           /* sample offsets (in floats of the 4-channel image), weight, gradient slots */

int artoracle_vng4(int W, int H, unsigned prefilters, const float* raw, float* red, float* green, float* blue)
{
    if (W < 8 || H < 8) return 1;
    const unsigned filters = prefilters & ~((prefilters & 0x55555555u) << 1);      /* the two greens collapsed: ISGREEN / ISBLUE / border */
    const int width = W, height = H;
    float (*image)[4] = (float (*)[4])calloc((size_t)height * width, sizeof *image);
    if (!image) return 1;
    for (int i = 0; i < H; i++) for (int j = 0; j < W; j++) image[(size_t)i * W + j][fc4(prefilters, i, j)] = raw[(size_t)i * W + j];
    /* 1. bilinear fill, L112-222 */
    for (int row = 1; row < height - 1; row++)
        for (int col = 1; col < width - 1; col++) {
            float* pix = image[(size_t)row * width + col];
            float sum[4] = {0.f, 0.f, 0.f, 0.f}, wsum[4] = {0.f, 0.f, 0.f, 0.f};
            for (int y = -1; y <= 1; y++)
                for (int x = -1; x <= 1; x++) {
                    const int shift = (y == 0) + (x == 0);
                    if (shift == 2) continue;
                    const int color = fc4(prefilters, row + y, col + x);
                    sum[color] += pix[(width * y + x) * 4 + color] * (float)(1 << shift);
                    wsum[color] += (float)(1 << shift);
                }
            for (unsigned c = 0; c < 4; c++) if (c != fc4(prefilters, row, col)) pix[c] = sum[c] * (1.f / wsum[c]);
        }
    /* 2. the gradient program of each of the 8 x 2 phases, L224-281 */
    vterm prog[8][2][64];
    int nterm[8][2], hood_o[8][2][8], hood_g[8][2][8];
    for (int row = 0; row < 8; row++)
        for (int col = 0; col < 2; col++) {
            const signed char* cp = TERMS;
            int n = 0;
            for (int t = 0; t < 64; t++) {
                const int y1 = *cp++, x1 = *cp++, y2 = *cp++, x2 = *cp++, weight = *cp++, grads = (unsigned char)*cp++;
                const unsigned color = fc4(prefilters, row + y1, col + x1);
                if (fc4(prefilters, row + y2, col + x2) != color) continue;
                const int diag = (fc4(prefilters, row, col + 1) == color && fc4(prefilters, row + 1, col) == color) ? 2 : 1;
                if (abs(y1 - y2) == diag && abs(x1 - x2) == diag) continue;
                vterm* T = &prog[row][col][n++];
                T->o1 = (y1 * width + x1) * 4 + color;
                T->o2 = (y2 * width + x2) * 4 + color;
                T->w = (float)(1 << weight);
                T->g1 = T->g2 = -1;
                for (int g = 0; g < 8; g++) if (grads & (1 << g)) { if (T->g1 < 0) T->g1 = g; else if (T->g2 < 0) T->g2 = g; }
            }
            nterm[row][col] = n;
            cp = CHOOD;
            for (int g = 0; g < 8; g++) {
                const int y = *cp++, x = *cp++;
                hood_o[row][col][g] = (y * width + x) * 4;
                const unsigned color = fc4(prefilters, row, col);
                hood_g[row][col][g] = (fc4(prefilters, row + y, col + x) != color && fc4(prefilters, row + y * 2, col + x * 2) == color) ? (y * width + x) * 8 + (int)color : 0;
            }
        }
    for (int row = 2; row < height - 2; row++)
        for (int col = 2; col < width - 2; col++) {
            const float* pix = image[(size_t)row * width + col];
            int color = fc4(prefilters, row, col);
            const int pr = row & 7, pc = col & 1;
            float gval[8] = {0.f};
            for (int t = 0; t < nterm[pr][pc]; ++t) {
                const vterm* T = &prog[pr][pc][t];
                const float diff = fabsf(pix[T->o1] - pix[T->o2]) * T->w;
                gval[T->g1] += diff;
                if (T->g2 >= 0) gval[T->g2] += diff;
            }
            /* rtengine::min / max of eight: min(min(a, b), min(rest...)) trees (rt_math.h L60-82); exact for any association */
            float mn = gval[0], mx = gval[0];
            for (int g = 1; g < 8; g++) { if (gval[g] < mn) mn = gval[g]; if (mx < gval[g]) mx = gval[g]; }
            const float thold = mn + mx * 0.5f;
            float sum0 = 0.f, sum1 = 0.f;
            const float greenval = pix[color];
            int num = 0;
            if (color & 1) {
                color ^= 2;
                for (int g = 0; g < 8; g++)
                    if (gval[g] <= thold) {
                        if (hood_g[pr][pc][g]) sum0 += greenval + pix[hood_g[pr][pc][g]];
                        sum1 += pix[hood_o[pr][pc][g] + color];
                        num++;
                    }
                sum0 *= 0.5f;
            } else {
                for (int g = 0; g < 8; g++)
                    if (gval[g] <= thold) {
                        if (hood_g[pr][pc][g]) sum0 += greenval + pix[hood_g[pr][pc][g]];
                        sum1 += pix[hood_o[pr][pc][g] + 1] + pix[hood_o[pr][pc][g] + 3];
                        num++;
                    }
            }
            green[(size_t)row * W + col] = max0(greenval + (sum1 - sum0) / (2 * num));
        }
    /* 3. red / blue from colour differences, L32-55 */
#define RAW(i, j) raw[(size_t)(i) * W + (j)]
    for (int i = 3; i < H - 3; ++i) {
        float *ar = red + (size_t)i * W, *ab = blue + (size_t)i * W;
        const float *pg = green + (size_t)(i - 1) * W, *cg = green + (size_t)i * W, *ng = green + (size_t)(i + 1) * W;
        if (fc4(filters, i, 0) == 2 || fc4(filters, i, 1) == 2) { float* t = ar; ar = ab; ab = t; }
        for (int j = 3; j < W - 3; ++j) {
            if (fc4(filters, i, j) != 1) {
                ar[j] = RAW(i, j);
                float rb = (RAW(i - 1, j - 1) - pg[j - 1] + RAW(i + 1, j - 1) - ng[j - 1]);
                rb += (RAW(i - 1, j + 1) - pg[j + 1] + RAW(i + 1, j + 1) - ng[j + 1]);
                ab[j] = max0(cg[j] + rb * 0.25f);
            } else {
                ar[j] = max0(cg[j] + (RAW(i, j - 1) - cg[j - 1] + RAW(i, j + 1) - cg[j + 1]) / 2);
                ab[j] = max0(cg[j] + (RAW(i - 1, j) - pg[j] + RAW(i + 1, j) - ng[j]) / 2);
            }
        }
    }
#undef RAW
    free(image);
    /* 4. L382 */
    artoracle_border_interpolate2(W, H, filters, 3, raw, W, red, green, blue, W);
    return 0;
}
