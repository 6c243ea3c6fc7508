/*
 * oracle/xtrans_port.c -- plain-C restatement of the reference's X-Trans demosaic (Markesteijn, 1 or 3 passes).
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Follows reference rtengine/xtrans_demosaic.cc: constants L30-40, RawImageSource::cielab L42-116, xtransborder_interpolate
 * L122-173, xtrans_interpolate L181-969.
 *
 * The reference walks every tile row with running column offsets / colour toggles; here every step is restated per pixel
 * (which sites a step visits, with which hexagon, colour and direction plane) so that the steps are visibly data-parallel
 * -- the CUDA kernels are organised the same way.  What has to be kept literally is the per-thread tile buffer: its
 * sub-buffers alias each other (homo and the green min/max table over lab, homosum over drv, homosummax over the last homo
 * map) and the 5x5 homogeneity sums near the image border read homo rows / columns that the homogeneity step never wrote
 * (L829-836: rows MIN(top, 8) - 2 ..., rows up to mrow + 5 - 7), i.e. bytes of Lab / YPbPr floats and min/max floats.
 * Those sums can exceed 255, where the 16-wide SSE2 path saturates (_mm_adds_epu8) and the scalar tail wraps (uint8 store).
 * The canonical reference clears the buffer at the start of every tile (oracle/_ref "det" build): the stock build reads the
 * previous tile's bytes there and is OpenMP-schedule dependent.  Pinned bit-exact against the det build
 * (tests/test_oracle_xtrans.py).  Compile with -ffp-contract=off -fno-strict-aliasing.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TS 114
#define TSH (TS / 2)

typedef struct { float min, max; } minmax_t;

typedef struct {
    int xtrans[6][6];
    float xyz_cam[3][3];
    short allhex[2][3][3][8];
    int sgrow, sgcol;
    int RightShift[3];
    int passes, ndir, useCieLab;
    int W, H;
    const float* raw;
    float *red, *green, *blue;
    const float* cbrt_lut;
} xt_t;

#define FCOL(x, row, col) ((x)->xtrans[(row) % 6][(col) % 6])
#define ISGREEN(x, row, col) ((x)->xtrans[(row) % 3][(col) % 3] & 1)
#define SQRF(a) ((a) * (a))
static inline float limf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }   /* rt_math.h LIM = max(lo, min(v, hi)) */

static const float xyz_rgb[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
static const float d65_white[3] = {0.950456, 1, 1.088754};

static float* g_cbrt = NULL;
static const float* cbrt_table(void)
{   /* cielab L44-63 */
    if (!g_cbrt) {
        float* t = (float*)malloc(sizeof(float) * 0x14000);
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0;
        for (int i = 0; i < 0x14000; i++) {
            const double r = i / 65535.0;
            t[i] = (float)(r > eps ? cbrt(r) : (kappa * r + 16.0) / 116.0);
        }
        g_cbrt = t;
    }
    return g_cbrt;
}
static inline float lut_i(const float* t, int idx) { return t[idx < 0 ? 0 : (idx > 0x14000 - 1 ? 0x14000 - 1 : idx)]; }

static void setup(xt_t* x, const int* xtrans36, const float* rgb_cam12)
{
    static const short orth[12] = {1, 0, 0, 1, -1, 0, 0, -1, 1, 0, 0, 1};
    static const short patt[2][16] = {{0, 1, 0, -1, 2, 0, -1, 0, 1, 1, 1, -1, 0, 0, 0, 0}, {0, 1, 0, -2, 1, 0, -2, 0, 1, 1, -2, -2, 1, -1, -1, 1}};
    memcpy(x->xtrans, xtrans36, sizeof x->xtrans);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += xyz_rgb[i][k] * rgb_cam12[k * 4 + j] / d65_white[i];
            x->xyz_cam[i][j] = s;
        }
    x->sgrow = x->sgcol = 0;
    memset(x->allhex, 0, sizeof x->allhex);
    for (int row = 0; row < 3; row++)
        for (int col = 0; col < 3; col++) {
            const int gint = ISGREEN(x, row, col);
            for (int ng = 0, d = 0; d < 10; d += 2) {
                if (ISGREEN(x, row + orth[d] + 6, col + orth[d + 2] + 6)) ng = 0; else ng++;
                if (ng == 4) { x->sgrow = row; x->sgcol = col; }
                if (ng == gint + 1)
                    for (int c = 0; c < 8; c++) {
                        const int v = orth[d] * patt[gint][c * 2] + orth[d + 1] * patt[gint][c * 2 + 1];
                        const int h = orth[d + 2] * patt[gint][c * 2] + orth[d + 3] * patt[gint][c * 2 + 1];
                        x->allhex[0][row][col][c ^ (gint * 2 & d)] = (short)(h + v * x->W);
                        x->allhex[1][row][col][c ^ (gint * 2 & d)] = (short)(h + v * TS);
                    }
            }
        }
    for (int row = 0; row < 3; row++) {
        int greencount = 0;
        for (int col = 0; col < 3; col++) greencount += ISGREEN(x, row, col);
        x->RightShift[row] = (greencount == 2);
    }
    x->ndir = 4 << (x->passes > 1);
    x->cbrt_lut = cbrt_table();
}

/* the hexagon offsets over the image (W-strided): allhex[0] holds h + v * width as a short, which overflows for wide images
 * exactly as in the reference (short allhex[2][3][3][8], L212/L241) */
#define RAW(x, row, col) ((x)->raw[(size_t)(row) * (x)->W + (col)])

static void cielab_tile(const xt_t* x, const float* rgb /* &rgb[d][4][4][0] */, float* l, float* a, float* b, int height)
{   /* cielab(rgb, l, a, b, width = ts, height, labWidth = ts - 8, xyz_cam), L65-116 */
    const int width = TS, labWidth = TS - 8;
    const float(*m)[3] = x->xyz_cam;
    for (int i = 0; i < height; i++) {
        int j = 0;
        for (; j < labWidth - 3; j += 4)
            for (int k = j; k < j + 4; k++) {
                const float* p = rgb + (size_t)(i * width + k) * 3;
                const float X = p[0] * m[0][0] + p[1] * m[0][1] + p[2] * m[0][2];
                const float Y = p[0] * m[1][0] + p[1] * m[1][1] + p[2] * m[1][2];
                const float Z = p[0] * m[2][0] + p[1] * m[2][1] + p[2] * m[2][2];
                const float fx = lut_i(x->cbrt_lut, (int)lrintf(X)), fy = lut_i(x->cbrt_lut, (int)lrintf(Y)), fz = lut_i(x->cbrt_lut, (int)lrintf(Z));
                l[i * labWidth + k] = 116.f * fy - 16.f;
                a[i * labWidth + k] = 500.f * (fx - fy);
                b[i * labWidth + k] = 200.f * (fy - fz);
            }
        for (; j < labWidth; j++) {
            const float* p = rgb + (size_t)(i * width + j) * 3;
            float xyz[3] = {0.5f, 0.5f, 0.5f};
            for (int c = 0; c < 3; c++) {
                const float val = p[c];
                xyz[0] += m[0][c] * val; xyz[1] += m[1][c] * val; xyz[2] += m[2][c] * val;
            }
            xyz[0] = lut_i(x->cbrt_lut, (int)xyz[0]); xyz[1] = lut_i(x->cbrt_lut, (int)xyz[1]); xyz[2] = lut_i(x->cbrt_lut, (int)xyz[2]);
            l[i * labWidth + j] = 116 * xyz[1] - 16;
            a[i * labWidth + j] = 500 * (xyz[0] - xyz[1]);
            b[i * labWidth + j] = 200 * (xyz[1] - xyz[2]);
        }
    }
}

static void minmax6(const xt_t* x, int row, int col, float* mn, float* mx)
{
    const short* hex = x->allhex[0][row % 3][col % 3];
    const float* pix = &RAW(x, row, col);
    float minval = FLT_MAX, maxval = 0.f;
    for (int c = 0; c < 6; c++) {
        const float val = pix[hex[c]];
        minval = minval < val ? minval : val;
        maxval = maxval > val ? maxval : val;
    }
    *mn = minval; *mx = maxval;
}

static void tile(const xt_t* x, float* buffer, int top, int left)
{
    const int ndir = x->ndir, height = x->H, width = x->W, passes = x->passes;
    const int sgrow = x->sgrow, sgcol = x->sgcol;
    static const short dir[4] = {1, TS, TS + 1, TS - 1};
    memset(buffer, 0, (TS * TS * (ndir * 4 + 3) + 128) * sizeof(float));          /* the canonical (det) reference */
    float(*rgb)[TS][TS][3] = (float(*)[TS][TS][3])buffer;
    float(*lab)[TS - 8][TS - 8] = (float(*)[TS - 8][TS - 8])(buffer + TS * TS * (ndir * 3));
    float(*drv)[TS - 10][TS - 10] = (float(*)[TS - 10][TS - 10])(buffer + TS * TS * (ndir * 3 + 3));
    uint8_t(*homo)[TS][TS] = (uint8_t(*)[TS][TS])lab;
    minmax_t(*gmm)[TSH] = (minmax_t(*)[TSH])lab;
    uint8_t(*homosum)[TS][TS] = (uint8_t(*)[TS][TS])drv;
    uint8_t(*homosummax)[TS] = (uint8_t(*)[TS])homo[ndir - 1];

    int mrow = (top + TS < height - 3) ? top + TS : height - 3;
    int mcol = (left + TS < width - 3) ? left + TS : width - 3;

    /* green min / max, L319-408.  Non-green sites come alone (rows with two greens per three columns) or in horizontal
     * pairs; a pair shares the hexagon of its left pixel when that pixel is inside the tile. */
    for (int row = top; row < mrow; row++)
        for (int col = left; col < mcol; col++) {
            if (ISGREEN(x, row, col)) continue;
            int src = col;
            if (!x->RightShift[row % 3]) {
                const int second = !ISGREEN(x, row, col + 5) /* col - 1 mod 3 */;
                if (second && col - 1 >= left) src = col - 1;
            }
            float mn, mx;
            minmax6(x, row, src, &mn, &mx);
            gmm[row - top][(col - left) >> 1].min = mn;
            gmm[row - top][(col - left) >> 1].max = mx;
        }

    /* rgb[0..3] <- the mosaic, L410-419 */
    for (int row = top; row < mrow; row++)
        for (int col = left; col < mcol; col++)
            rgb[0][row - top][col - left][FCOL(x, row, col)] = RAW(x, row, col);
    for (int c = 0; c < 3; c++) memcpy(rgb[c + 1], rgb[0], sizeof *rgb);

    /* green along the four directions, L421-475 */
    for (int row = top; row < mrow; row++)
        for (int col = left; col < mcol; col++) {
            if (ISGREEN(x, row, col)) continue;
            const short* hex = x->allhex[0][row % 3][col % 3];
            const float* pix = &RAW(x, row, col);
            float color[4];
            color[0] = 0.6796875f * (pix[hex[1]] + pix[hex[0]]) - 0.1796875f * (pix[2 * hex[1]] + pix[2 * hex[0]]);
            color[1] = 0.87109375f * pix[hex[3]] + pix[hex[2]] * 0.12890625f + 0.359375f * (pix[0] - pix[-hex[2]]);
            for (int c = 0; c < 2; c++)
                color[2 + c] = 0.640625f * pix[hex[4 + c]] + 0.359375f * pix[-2 * hex[4 + c]] +
                               0.12890625f * (2.f * pix[0] - pix[3 * hex[4 + c]] - pix[-3 * hex[4 + c]]);
            const int flip = x->RightShift[row % 3] ? 0 : 1;
            const minmax_t mm = gmm[row - top][(col - left) >> 1];
            for (int c = 0; c < 4; c++) rgb[c ^ flip][row - top][col - left][1] = limf(color[c], mm.min, mm.max);
        }

    for (int pass = 0; pass < passes; pass++) {
        if (pass == 1) memcpy(rgb += 4, buffer, 4 * sizeof *rgb);

        /* recalculate green from interpolated values of closer pixels, L483-522 */
        if (pass)
            for (int row = top + 2; row < mrow - 2; row++)
                for (int col = left + 2; col < mcol - 2; col++) {
                    if (ISGREEN(x, row, col)) continue;
                    const int f = FCOL(x, row, col);
                    const short* hex = x->allhex[1][row % 3][col % 3];
                    const int flip = x->RightShift[row % 3] ? 0 : 1;
                    const minmax_t mm = gmm[row - top][(col - left) >> 1];
                    for (int d = 3; d < 6; d++) {
                        float(*rix)[3] = &rgb[(d - 2) ^ flip][row - top][col - left];
                        const float val = 0.33333333f * (rix[-2 * hex[d]][1] + 2 * (rix[hex[d]][1] - rix[hex[d]][f]) - rix[-2 * hex[d]][f]) + rix[0][f];
                        rix[0][1] = limf(val, mm.min, mm.max);
                    }
                }

        /* red and blue for solitary green pixels, L524-561 */
        {
            const int sgstartcol = (left - sgcol + 4) / 3 * 3 + sgcol;
            for (int row = (top - sgrow + 4) / 3 * 3 + sgrow; row < mrow - 2; row += 3)
                for (int col = sgstartcol; col < mcol - 2; col += 3) {
                    int h = FCOL(x, row, col + 1);
                    float(*rix)[3] = &rgb[0][row - top][col - left];
                    float diff[6] = {0.f};
                    float color[3][6];
                    for (int i = 1, d = 0; d < 6; d++, i ^= TS ^ 1, h ^= 2) {
                        for (int c = 0; c < 2; c++, h ^= 2) {
                            const int o = i << c;
                            const float g = rix[0][1] + rix[0][1] - rix[o][1] - rix[-o][1];
                            color[h][d] = g + rix[o][h] + rix[-o][h];
                            if (d > 1) diff[d] += SQRF(rix[o][1] - rix[-o][1] - rix[o][h] + rix[-o][h]) + SQRF(g);
                        }
                        if (d > 2 && (d & 1))
                            if (diff[d - 1] < diff[d])
                                for (int c = 0; c < 2; c++) color[c * 2][d] = color[c * 2][d - 1];
                        if ((d & 1) || d < 2) {
                            for (int c = 0; c < 2; c++) rix[0][c * 2] = 0.5f * color[c * 2][d];
                            rix += TS * TS;
                        }
                    }
                }
        }

        /* red for blue pixels and vice versa, L563-603 */
        for (int row = top + 3; row < mrow - 3; row++) {
            const int c = ((row - sgrow) % 3) ? TS : 1;
            const int h = 3 * (c ^ TS ^ 1);
            for (int col = left + 3; col < mcol - 3; col++) {
                if (ISGREEN(x, row, col)) continue;
                const int f = 2 - FCOL(x, row, col);
                float(*rix)[3] = &rgb[0][row - top][col - left];
                for (int d = 0; d < 4; d++, rix += TS * TS) {
                    const int i = d > 1 || ((d ^ c) & 1) ||
                                  ((fabsf(rix[0][1] - rix[c][1]) + fabsf(rix[0][1] - rix[-c][1])) < 2.f * (fabsf(rix[0][1] - rix[h][1]) + fabsf(rix[0][1] - rix[-h][1]))) ? c : h;
                    rix[0][f] = rix[0][1] + 0.5f * (rix[i][f] + rix[-i][f] - rix[i][1] - rix[-i][1]);
                }
            }
        }

        /* red and blue for the 2x2 blocks of green, L605-650 */
        for (int row = top + 2; row < mrow - 2; row++) {
            if (!((row - sgrow) % 3)) continue;
            for (int col = left + 2; col < mcol - 2; col++) {
                if (!((col - sgcol) % 3)) continue;
                float(*rix)[3] = &rgb[0][row - top][col - left];
                const short* hex = x->allhex[1][row % 3][col % 3];
                for (int d = 0; d < ndir; d += 2, rix += TS * TS) {
                    if (hex[d] + hex[d + 1]) {
                        const float g = 3 * rix[0][1] - 2 * rix[hex[d]][1] - rix[hex[d + 1]][1];
                        for (int c = 0; c < 4; c += 2) rix[0][c] = (g + 2 * rix[hex[d]][c] + rix[hex[d + 1]][c]) * 0.33333333f;
                    } else {
                        const float g = 2 * rix[0][1] - rix[hex[d]][1] - rix[hex[d + 1]][1];
                        for (int c = 0; c < 4; c += 2) rix[0][c] = (g + rix[hex[d]][c] + rix[hex[d + 1]][c]) * 0.5f;
                    }
                }
            }
        }
    }

    rgb = (float(*)[TS][TS][3])buffer;
    mrow -= top;
    mcol -= left;

    /* derivatives of every direction plane in a perceptual space, L657-741 */
    for (int d = 0; d < ndir; d++) {
        if (x->useCieLab)
            cielab_tile(x, &rgb[d][4][4][0], &lab[0][0][0], &lab[1][0][0], &lab[2][0][0], mrow - 8);
        else
            for (int row = 4; row < mrow - 4; row++)
                for (int col = 4; col < mcol - 4; col++) {
                    const float* p = rgb[d][row][col];
                    const float y = 0.2627f * p[0] + 0.6780f * p[1] + 0.0593f * p[2];
                    lab[0][row - 4][col - 4] = y;
                    lab[1][row - 4][col - 4] = (p[2] - y) * 0.56433f;
                    lab[2][row - 4][col - 4] = (p[0] - y) * 0.67815f;
                }
        int f = dir[d & 3];
        f = f == 1 ? 1 : f - 8;
        for (int row = 5; row < mrow - 5; row++)
            for (int col = 5; col < mcol - 5; col++) {
                const float* l = &lab[0][row - 4][col - 4];
                const float* a = &lab[1][row - 4][col - 4];
                const float* b = &lab[2][row - 4][col - 4];
                if (x->useCieLab) {
                    const float g = 2 * l[0] - l[f] - l[-f];
                    drv[d][row - 5][col - 5] = SQRF(g) + SQRF((2 * a[0] - a[f] - a[-f] + g * 2.1551724f)) + SQRF((2 * b[0] - b[f] - b[-f] - g * 0.86206896f));
                } else {
                    drv[d][row - 5][col - 5] = SQRF(2 * l[0] - l[f] - l[-f]) + SQRF(2 * a[0] - a[f] - a[-f]) + SQRF(2 * b[0] - b[f] - b[-f]);
                }
            }
    }

    /* homogeneity maps, L743-811 (the SSE2 and scalar forms agree) */
    for (int row = 6; row < mrow - 6; row++)
        for (int col = 6; col < mcol - 6; col++) {
            float tr = drv[0][row - 5][col - 5] < drv[1][row - 5][col - 5] ? drv[0][row - 5][col - 5] : drv[1][row - 5][col - 5];
            for (int d = 2; d < ndir; d++) tr = (drv[d][row - 5][col - 5] < tr ? drv[d][row - 5][col - 5] : tr);
            tr *= 8;
            for (int d = 0; d < ndir; d++) {
                uint8_t temp = 0;
                for (int v = -1; v <= 1; v++)
                    for (int h = -1; h <= 1; h++) temp += (drv[d][row + v - 5][col + h - 5] <= tr ? 1 : 0);
                homo[d][row][col] = temp;
            }
        }

    if (height - top < TS + 4) mrow = height - top + 2;
    if (width - left < TS + 4) mcol = width - left + 2;

    /* 5x5 sums of the homogeneity maps, L822-868: 16 columns at a time with saturating adds, except on the last row where the
     * columns from startcol + 16 * ceil((mcol - 23 - startcol) / 16) on take the scalar path, whose uint8 store wraps.
     * The reads reach homo rows / columns the step above never wrote (see the file header). */
    const int startrow = top < 8 ? top : 8, startcol = left < 8 ? left : 8;
    for (int d = 0; d < ndir; d++)
        for (int row = startrow; row < mrow - 8; row++) {
            const int endcol = row < mrow - 9 ? mcol - 8 : mcol - 23;
            int vec_end = startcol;
            while (vec_end < endcol) vec_end += 16;
            for (int col = startcol; col < mcol - 8; col++) {
                int sum = 0;
                const uint8_t* base = &homo[d][0][0];
                for (int v = -2; v <= 2; v++)
                    for (int h = -2; h <= 2; h++) sum += base[(row + v) * TS + col + h];
                homosum[d][row][col] = (uint8_t)(col < vec_end ? (sum > 255 ? 255 : sum) : sum);
            }
        }

    /* per-pixel maximum minus an eighth, L870-911 (SSE2 and scalar forms agree) */
    for (int row = startrow; row < mrow - 8; row++)
        for (int col = startcol; col < mcol - 8; col++) {
            uint8_t maxval = homosum[0][row][col];
            for (int d = 1; d < ndir; d++) maxval = maxval < homosum[d][row][col] ? homosum[d][row][col] : maxval;
            maxval -= maxval >> 3;
            homosummax[row][col] = maxval;
        }

    /* average the most homogeneous directions, L914-949 */
    for (int row = startrow; row < mrow - 8; row++)
        for (int col = startcol; col < mcol - 8; col++) {
            uint8_t hm[8] = {0};
            for (int d = 0; d < 4; d++) hm[d] = homosum[d][row][col];
            for (int d = 4; d < ndir; d++) {
                hm[d] = homosum[d][row][col];
                if (hm[d - 4] < hm[d]) hm[d - 4] = 0;
                else if (hm[d - 4] > hm[d]) hm[d] = 0;
            }
            float avg[4] = {0.f};
            const uint8_t maxval = homosummax[row][col];
            for (int d = 0; d < ndir; d++)
                if (hm[d] >= maxval) {
                    for (int c = 0; c < 3; c++) avg[c] += rgb[d][row][col][c];
                    avg[3]++;
                }
            const size_t o = (size_t)(row + top) * width + col + left;
            const float r = avg[0] / avg[3], g = avg[1] / avg[3], b = avg[2] / avg[3];
            x->red[o] = 0.f < r ? r : 0.f;
            x->green[o] = 0.f < g ? g : 0.f;
            x->blue[o] = 0.f < b ? b : 0.f;
        }
}

int artoracle_xtrans_border(int W, int H, const int* xtrans36, int border, const float* raw, float* red, float* green, float* blue)
{   /* xtransborder_interpolate, L122-173 */
    static const float weight[3][3] = {{0.25f, 0.5f, 0.25f}, {0.5f, 0.f, 0.5f}, {0.25f, 0.5f, 0.25f}};
    const int (*xtrans)[6] = (const int (*)[6])xtrans36;
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            if (col == border && row >= border && row < H - border) col = W - border;
            float sum[6] = {0.f};
            for (int y = row - 1 > 0 ? row - 1 : 0, v = row == 0 ? 0 : -1; y <= (row + 1 < H - 1 ? row + 1 : H - 1); y++, v++)
                for (int xx = col - 1 > 0 ? col - 1 : 0, h = col == 0 ? 0 : -1; xx <= (col + 1 < W - 1 ? col + 1 : W - 1); xx++, h++) {
                    const int f = xtrans[y % 6][xx % 6];
                    sum[f] += raw[(size_t)y * W + xx] * weight[v + 1][h + 1];
                    sum[f + 3] += weight[v + 1][h + 1];
                }
            const size_t o = (size_t)row * W + col;
            switch (xtrans[row % 6][col % 6]) {
            case 0: red[o] = raw[o]; green[o] = sum[1] / sum[4]; blue[o] = sum[2] / sum[5]; break;
            case 1:
                if (sum[3] == 0.f) red[o] = green[o] = blue[o] = raw[o];
                else { red[o] = sum[0] / sum[3]; green[o] = raw[o]; blue[o] = sum[2] / sum[5]; }
                break;
            case 2: red[o] = sum[0] / sum[3]; green[o] = sum[1] / sum[4]; blue[o] = raw[o];
            }
        }
    return 0;
}

/* xtrans36: the 6x6 colour matrix (0 R, 1 G, 2 B); rgb_cam12: RawImage::getRgbCam's 3x4; planes contiguous W x H */
int artoracle_xtrans(int W, int H, const int* xtrans36, const float* rgb_cam12, int passes, int useCieLab, const float* raw,
                     float* red, float* green, float* blue)
{
    if (passes != 1 && passes != 3) return 1;
    xt_t x;
    memset(&x, 0, sizeof x);
    x.W = W; x.H = H; x.raw = raw; x.red = red; x.green = green; x.blue = blue; x.passes = passes; x.useCieLab = useCieLab;
    setup(&x, xtrans36, rgb_cam12);
    float* buffer = (float*)malloc((TS * TS * (x.ndir * 4 + 3) + 128) * sizeof(float));
    if (!buffer) return 1;
    for (int top = 3; top < H - 19; top += TS - 16)
        for (int left = 3; left < W - 19; left += TS - 16) tile(&x, buffer, top, left);
    free(buffer);
    return artoracle_xtrans_border(W, H, xtrans36, passes > 1 ? 8 : 11, raw, red, green, blue);
}
