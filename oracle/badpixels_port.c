/*
 * oracle/badpixels_port.c -- CPU restatement of the reference's hot / dead pixel filter.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates RawImageSource::findHotDeadPixels (reference rtengine/badpixels.cc L477-627, with sum5x5 L36-58) and
 * RawImageSource::interpolateBadPixelsBayer (L66-180).
 *   find:  cfablur = raw - median of the 3x3 same-colour neighbourhood (Bayer: samples at distance 2; X-Trans: the first 9 / 7 / 5 / 3 / 1
 *          same-colour samples of the 5x5 window in raster order, L496-518) for rows 2 .. H-3 and columns 2 .. W-3, zero outside (the
 *          reference's five-row ring is cleared at the start, rows >= H-2 are written as zero and the outer columns never written: its row
 *          chunking per thread does not show in the result as long as every thread's static chunk has at least two rows, H - 4 >= 2 x threads).  A pixel is bad when |cfablur| > varthresh * (sum of |cfablur| over its 5x5
 *          neighbourhood - |cfablur|); the SSE2 sum adds the five ring rows in RING ORDER (slot = row % 5): ((s0 + s1) + (s2 + s3)) + s4 per
 *          column, four columns through vhadd ((l0 + l2) + (l1 + l3)), the fifth column after them.
 *   interpolate: gradient-weighted mean of the pairs of good same-colour neighbours, or their plain mean when no pair is good; only good
 *          pixels are read and only bad ones written, so the reference's in-place parallel loop is race-free.
 * Pinned bit-exact against the reference's own functions compiled in place (oracle/_ref) in tests/test_oracle_badpixels.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline unsigned fc_bp(unsigned filters, int row, int col) { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }

static float median_n(float* v, int n)
{   /* exact median of an odd count (the reference's min / max networks, median.h, return the same element for NaN-free input) */
    for (int i = 1; i < n; ++i) { const float x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
    return v[n / 2];
}

/* map: H x W bytes, bad pixels are OR-ed in (PixelsMap::set); xtrans36 == NULL: Bayer.  Returns the number of pixels marked by this call. */
int artoracle_find_hot_dead(const float* raw, int W, int H, const int* xtrans36, float thresh, int hot, int dead, unsigned char* map)
{
    if (W < 5 || H < 5) return 0;
    const int is_xtrans = xtrans36 != NULL;
    const float varthresh = (20.f * (thresh / 100.f) + 1.f) / 24.f * (is_xtrans ? 0.25f : 1.f);
    float* blur = (float*)calloc((size_t)W * H, sizeof(float));
    if (!blur) return -1;
#define RAW(i, j) raw[(size_t)(i) * W + (j)]
#pragma omp parallel for
    for (int i = 2; i < H - 2; ++i)
        for (int j = 2; j < W - 2; ++j) {
            float m[9], med;
            if (!is_xtrans) {
                int n = 0;
                for (int dy = -2; dy <= 2; dy += 2) for (int dx = -2; dx <= 2; dx += 2) m[n++] = RAW(i + dy, j + dx);
                med = median_n(m, 9);
            } else {
                const int c = xtrans36[(i % 6) * 6 + (j % 6)];
                int n = 0;
                for (int y = i - 2; y < i + 3; ++y)
                    for (int x = j - 2; x < j + 3; ++x)
                        if (xtrans36[(y % 6) * 6 + (x % 6)] == c) { if (n < 9) m[n] = RAW(y, x); ++n; }
                med = n >= 9 ? median_n(m, 9) : n >= 7 ? median_n(m, 7) : n >= 5 ? median_n(m, 5) : n >= 3 ? median_n(m, 3) : m[0];
            }
            blur[(size_t)i * W + j] = RAW(i, j) - med;
        }
#undef RAW
    int counter = 0;
    for (int rr = 2; rr < H - 2; ++rr)
        for (int cc = 2; cc < W - 2; ++cc) {
            float pixdev = blur[(size_t)rr * W + cc];
            if (!dead && pixdev <= 0.f) continue;
            if (!hot && pixdev >= 0.f) continue;
            pixdev = fabsf(pixdev);
            const float* in[5];
            for (int r = rr - 2; r <= rr + 2; ++r) in[r % 5] = blur + (size_t)r * W;
            float col[5];
            for (int k = 0; k < 5; ++k) {
                const int x = cc - 2 + k;
                col[k] = ((fabsf(in[0][x]) + fabsf(in[1][x])) + (fabsf(in[2][x]) + fabsf(in[3][x]))) + fabsf(in[4][x]);
            }
            float hfnbrave = -pixdev;
            hfnbrave += (col[0] + col[2]) + (col[1] + col[3]);
            hfnbrave += col[4];
            if (pixdev > varthresh * hfnbrave) { map[(size_t)rr * W + cc] = 1; ++counter; }
        }
    free(blur);
    return counter;
}

int artoracle_interpolate_bad_bayer(float* raw, int W, int H, unsigned filters, const unsigned char* map)
{
    const float eps = 1.f;
    int counter = 0;
#define RAW(i, j) raw[(size_t)(i) * W + (j)]
#define BAD(x, y) (map[(size_t)(y) * W + (x)] != 0)
    for (int row = 2; row < H - 2; ++row)
        for (int col = 2; col < W - 2; ++col) {
            if (!BAD(col, row)) continue;
            float wtdsum = 0.f, norm = 0.f;
            if (fc_bp(filters, row & 1, col & 1) == 1) {
                for (int dx = -1; dx <= 1; dx += 2) {
                    if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                    const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                    wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                    norm += dirwt;
                }
            } else {
                for (int dx = -2; dx <= 2; dx += 4) {
                    if (BAD(col + dx, row - 2) || BAD(col - dx, row + 2)) continue;
                    const float dirwt = 0.35355339f / (fabsf(RAW(row - 2, col + dx) - RAW(row + 2, col - dx)) + eps);
                    wtdsum += dirwt * (RAW(row - 2, col + dx) + RAW(row + 2, col - dx));
                    norm += dirwt;
                }
            }
            if (!(BAD(col - 2, row) || BAD(col + 2, row))) {
                const float dirwt = 0.5f / (fabsf(RAW(row, col - 2) - RAW(row, col + 2)) + eps);
                wtdsum += dirwt * (RAW(row, col - 2) + RAW(row, col + 2));
                norm += dirwt;
            }
            if (!(BAD(col, row - 2) || BAD(col, row + 2))) {
                const float dirwt = 0.5f / (fabsf(RAW(row - 2, col) - RAW(row + 2, col)) + eps);
                wtdsum += dirwt * (RAW(row - 2, col) + RAW(row + 2, col));
                norm += dirwt;
            }
            if (norm > 0.f) {
                RAW(row, col) = wtdsum / (2.f * norm);
                counter++;
            } else {
                int tot = 0;
                float sum = 0.f;
                for (int dy = -2; dy <= 2; dy += 2)
                    for (int dx = -2; dx <= 2; dx += 2) {
                        if (BAD(col + dx, row + dy)) continue;
                        sum += RAW(row + dy, col + dx);
                        tot++;
                    }
                if (tot > 0) { RAW(row, col) = sum / tot; counter++; }
            }
        }
#undef RAW
#undef BAD
    return counter;
}

/* RawImageSource::interpolateBadPixelsXtrans (badpixels.cc L288-475) in raster order -- the stock loop is an OpenMP parallel for whose "virtual
 * pixel" (and distance-2 partner) read neighbours that are not checked against the map and may already have been rewritten in place, so only its
 * one-thread schedule is a function of the input; that schedule is the oracle (pinned against the reference run on one thread).  Also kept: the
 * knight-move scan's inner loop visits d2 = -1 only (`d2 < 1`, L414). */
int artoracle_interpolate_bad_xtrans(float* raw, int W, int H, const int* xt, const unsigned char* map)
{
    const float eps = 1.f;
    int counter = 0;
#define RAW(i, j) raw[(size_t)(i) * W + (j)]
#define BAD(x, y) (map[(size_t)(y) * W + (x)] != 0)
#define XFC(r, c) xt[((r) % 6) * 6 + ((c) % 6)]
    for (int row = 2; row < H - 2; ++row)
        for (int col = 2; col < W - 2; ++col) {
            if (!BAD(col, row)) continue;
            float wtdsum = 0.f, norm = 0.f;
            const int pixelColor = XFC(row, col);
            if (pixelColor == 1) {
                if (XFC(row, col - 1) == XFC(row, col + 1)) {       /* solitary green */
                    for (int dx = -1; dx <= 1; dx += 2) {
                        if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                        const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                        wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                        norm += dirwt;
                    }
                    for (int dx = -1; dx <= 1; dx += 2) {
                        if (BAD(col + dx, row - 2) || BAD(col - dx, row + 2)) continue;
                        const float dirwt = 0.44721359f / (fabsf(RAW(row - 2, col + dx) - RAW(row + 2, col - dx)) + eps);
                        wtdsum += dirwt * (RAW(row - 2, col + dx) + RAW(row + 2, col - dx));
                        norm += dirwt;
                    }
                    for (int dx = -2; dx <= 2; dx += 4) {
                        if (BAD(col + dx, row - 1) || BAD(col - dx, row + 1)) continue;
                        const float dirwt = 0.44721359f / (fabsf(RAW(row - 1, col + dx) - RAW(row + 1, col - dx)) + eps);
                        wtdsum += dirwt * (RAW(row - 1, col + dx) + RAW(row + 1, col - dx));
                        norm += dirwt;
                    }
                } else {                                            /* member of a 2x2 green square */
                    const int offset1 = XFC(row - 1, col - 1) == XFC(row + 1, col + 1) ? 1 : -1;
                    if (!(BAD(col - offset1, row - 1) || BAD(col + offset1, row + 1))) {
                        const float dirwt = 0.70710678f / (fabsf(RAW(row - 1, col - offset1) - RAW(row + 1, col + offset1)) + eps);
                        wtdsum += dirwt * (RAW(row - 1, col - offset1) + RAW(row + 1, col + offset1));
                        norm += dirwt;
                    }
                    int offsety = XFC(row - 1, col) != 1 ? 1 : -1;
                    int offsetx = offset1 * offsety;
                    if (!(BAD(col + offsetx, row) || BAD(col, row + offsety))) {
                        const float dirwt = 1.f / (fabsf(RAW(row, col + offsetx) - RAW(row + offsety, col)) + eps);
                        wtdsum += dirwt * (RAW(row, col + offsetx) + RAW(row + offsety, col));
                        norm += dirwt;
                    }
                    const int offsety2 = -offsety, offsetx2 = -offsetx;
                    offsetx *= 2; offsety *= 2;
                    if (!(BAD(col + offsetx, row + offsety2) || BAD(col + offsetx2, row + offsety))) {
                        const float dirwt = 0.44721359f / (fabsf(RAW(row + offsety2, col + offsetx) - RAW(row + offsety, col + offsetx2)) + eps);
                        wtdsum += dirwt * (RAW(row + offsety2, col + offsetx) + RAW(row + offsety, col + offsetx2));
                        norm += dirwt;
                    }
                }
            } else {                                                /* red / blue */
                for (int d1 = -2, offsety = 3; d1 <= 2; d1 += 4, offsety -= 6)
                    for (int d2 = -1, offsetx = 3; d2 < 1; d2 += 2, offsetx -= 6)
                        if (XFC(row + d1, col + d2) == pixelColor && !(BAD(col + d2, row + d1) || BAD(col + d2 + offsetx, row + d1 + offsety))) {
                            const float dirwt = 0.44721359f / (fabsf(RAW(row + d1, col + d2) - RAW(row + d1 + offsety, col + d2 + offsetx)) + eps);
                            wtdsum += dirwt * (RAW(row + d1, col + d2) + RAW(row + d1 + offsety, col + d2 + offsetx));
                            norm += dirwt;
                        }
                int found = 0, dx, dy;
                for (dx = -2, dy = 0; dx <= 2; dx += 4)
                    if (XFC(row, col + dx) == pixelColor) { found = 1; break; }
                if (!found)
                    for (dx = 0, dy = -2; dy <= 2; dy += 4)
                        if (XFC(row + dy, col) == pixelColor) { found = 1; break; }
                /* no same-colour pixel at distance 2 (not an X-Trans layout): the reference's loops leave dx = 0, dy = 6 and read out of bounds; refuse */
                if (!found) return -1;
                float virtualPixel;
                if (dy == 0) virtualPixel = 0.5f * (RAW(row - 1, col - dx) + RAW(row + 1, col - dx));
                else virtualPixel = 0.5f * (RAW(row - dy, col - 1) + RAW(row - dy, col + 1));
                const float dirwt = 0.5f / (fabsf(virtualPixel - RAW(row + dy, col + dx)) + eps);
                wtdsum += dirwt * (virtualPixel + RAW(row + dy, col + dx));
                norm += dirwt;
            }
            if (norm > 0.f) { RAW(row, col) = wtdsum / (2.f * norm); counter++; }
        }
#undef XFC
#undef RAW
#undef BAD
    return counter;
}
