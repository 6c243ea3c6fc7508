/*
 * oracle/pointwise_port.c -- CPU restatement of the per-pixel stages.  TEST INFRASTRUCTURE ONLY.
 *
 * artoracle_scale_convert: RawImageSource::getImage's gain/clip at skip == 1 (reference
 * rtengine/rawimagesource.cc L943-1025: `rtot *= rm; if (doClip) rtot = CLIP(rtot)`, CLIP = rt_math.h
 * L97-101) followed by the matrix branch of RawImageSource::colorSpaceConversion_ (L3197-3211:
 * float = double coefficient * float sample summed left to right in double).
 * Pinned against oracle/_ref (the reference's own CLIP and its own matrix loop) in
 * tests/test_oracle_pointwise.py: bit-exact.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

static inline float clip65535_(float a)
{   /* LIM(a, 0.f, 65535.f) = max(0, min(a, 65535)), rt_math.h L84-101 */
    const float m = a < 65535.f ? a : 65535.f;
    return 0.f < m ? m : 0.f;
}

int artoracle_scale_convert(int W, int H, float* r, float* g, float* b, long stride,
                            const float* mul, int doClip, const double* mat)
{
#pragma omp parallel for
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const size_t k = (size_t)i * stride + j;
            float x = r[k] * mul[0], y = g[k] * mul[1], z = b[k] * mul[2];
            if (doClip) { x = clip65535_(x); y = clip65535_(y); z = clip65535_(z); }
            if (mat) {
                const float nx = (float)(mat[0] * x + mat[1] * y + mat[2] * z);
                const float ny = (float)(mat[3] * x + mat[4] * y + mat[5] * z);
                const float nz = (float)(mat[6] * x + mat[7] * y + mat[8] * z);
                x = nx; y = ny; z = nz;
            }
            r[k] = x; g[k] = y; b[k] = z;
        }
    return 0;
}

/* RawImageSource::scaleColors, Bayer branch (rawimagesource.cc L2731-2772, dynamicRowNoiseFilter off):
 * val = max(0.f, raw - cblacksom[c4]) * scale_mul[c4]; chmax[c] = max(chmax[c], val). */
static inline unsigned fc_pw_(unsigned filters, int row, int col)
{
    return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
}
int artoracle_scale_colors_bayer(int W, int H, unsigned filters, float* raw, long stride,
                                 const float* cblacksom, const float* scale_mul, float* chmax)
{
    float mx[3] = {0.f, 0.f, 0.f};
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const int c = fc_pw_(filters, row, col);
            const int c4 = (c == 1 && !(row & 1)) ? 3 : c;
            const float d = raw[(size_t)row * stride + col] - cblacksom[c4];
            const float val = (0.f < d ? d : 0.f) * scale_mul[c4];
            raw[(size_t)row * stride + col] = val;
            mx[c] = mx[c] < val ? val : mx[c];
        }
    chmax[0] = mx[0]; chmax[1] = mx[1]; chmax[2] = mx[2];
    return 0;
}

/* RawImageSource::scaleColors, X-Trans branch (rawimagesource.cc L2795-2826): c = XTRANSFC(row, col) = xtrans[row % 6][col % 6] */
int artoracle_scale_colors_xtrans(int W, int H, const int* xtrans36, float* raw, long stride,
                                  const float* cblacksom, const float* scale_mul, float* chmax)
{
    float mx[3] = {0.f, 0.f, 0.f};
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const int c = xtrans36[(row % 6) * 6 + (col % 6)];
            const float d = raw[(size_t)row * stride + col] - cblacksom[c];
            const float val = (0.f < d ? d : 0.f) * scale_mul[c];
            raw[(size_t)row * stride + col] = val;
            mx[c] = mx[c] < val ? val : mx[c];
        }
    chmax[0] = mx[0]; chmax[1] = mx[1]; chmax[2] = mx[2];
    return 0;
}

/* RawImageSource::HLRecovery_blend (rawimagesource.cc L3613-3747; hlRecovery L3752-3755 calls it with maxval 65535), the "Blend" highlight
 * reconstruction getImage applies per line (L1014): clipped pixels get their chroma (in an opponent space) scaled to the unclipped
 * estimate's, faded in above half the lowest clip point; the tail mixes float and double arithmetic (double literals). */
static inline float hl_min(float a, float b) { return b < a ? b : a; }      /* rtengine::min */
int artoracle_hl_blend(float* rin, float* gin, float* bin, int width, float maxval, const float* hlmax)
{
    static const float trans[3][3] = {{1, 1, 1}, {1.7320508, -1.7320508, 0}, {-1, -1, 2}};
    static const float itrans[3][3] = {{1, 0.8660254, -0.5}, {1, -0.8660254, -0.5}, {1, 0, 1}};
    const float minpt = hl_min(hl_min(hlmax[0], hlmax[1]), hlmax[2]);
    const float maxave = (hlmax[0] + hlmax[1] + hlmax[2]) / 3;
    const float clipthresh = 0.95, fixthresh = 0.5;
    float clip[3];
    for (int c = 0; c < 3; c++) clip[c] = hl_min(maxave, hlmax[c]);
    const float clippt = clipthresh * maxval;
    const float fixpt = fixthresh * minpt;
    for (int col = 0; col < width; col++) {
        float rgb[3], cam[2][3], lab[2][3], sum[2], chratio, lratio = 0;
        float L, C, H;
        rgb[0] = rin[col]; rgb[1] = gin[col]; rgb[2] = bin[col];
        int c;
        for (c = 0; c < 3; c++) if (rgb[c] > clippt) break;
        if (c == 3) continue;
        for (c = 0; c < 3; c++) { lratio += hl_min(rgb[c], clip[c]); cam[0][c] = rgb[c]; cam[1][c] = hl_min(cam[0][c], maxval); }
        for (int i = 0; i < 2; i++) {
            for (c = 0; c < 3; c++) {
                lab[i][c] = 0;
                for (int j = 0; j < 3; j++) lab[i][c] += trans[c][j] * cam[i][j];
            }
            sum[i] = 0;
            for (c = 1; c < 3; c++) sum[i] += lab[i][c] * lab[i][c];
        }
        chratio = sqrtf(sum[1] / sum[0]);
        for (c = 1; c < 3; c++) lab[0][c] *= chratio;
        for (c = 0; c < 3; c++) {
            cam[0][c] = 0;
            for (int j = 0; j < 3; j++) cam[0][c] += itrans[c][j] * lab[0][j];
        }
        for (c = 0; c < 3; c++) rgb[c] = cam[0][c] / 3;
        if (rin[col] > fixpt) { const float t = (hl_min(clip[0], rin[col]) - fixpt) / (clip[0] - fixpt), f = t * t; rin[col] = hl_min(maxave, f * rgb[0] + (1 - f) * rin[col]); }
        if (gin[col] > fixpt) { const float t = (hl_min(clip[1], gin[col]) - fixpt) / (clip[1] - fixpt), f = t * t; gin[col] = hl_min(maxave, f * rgb[1] + (1 - f) * gin[col]); }
        if (bin[col] > fixpt) { const float t = (hl_min(clip[2], bin[col]) - fixpt) / (clip[2] - fixpt), f = t * t; bin[col] = hl_min(maxave, f * rgb[2] + (1 - f) * bin[col]); }
        const float tot = (rin[col] + gin[col] + bin[col]);
        lratio /= tot;
        L = tot / 3 / lratio;
        C = lratio * 1.732050808 * (rin[col] - gin[col]);
        H = lratio * (2 * bin[col] - rin[col] - gin[col]);
        rin[col] = L - H / 6.0 + C / 3.464101615;
        gin[col] = L - H / 6.0 - C / 3.464101615;
        bin[col] = L + H / 3.0;
    }
    return 0;
}

/* RawImageSource::getImage at skip == 1 for a standard CCD (rawimagesource.cc L943-1025, L1079-1086): per source line the gains / clip, the optional
 * "Blend" highlight reconstruction (hlRecovery, L1014), the coarse rotation of transLineStandard -> rotateLine (L57-87) into the output image, and after
 * the loop the horizontal / vertical mirror (PlanarRGBData::hflip / vflip, iimage.h L868-915).  tran = TR_R90 1 | TR_R180 2 | TR_R270 3 | TR_VFLIP 4 | TR_HFLIP 8.
 * Output: W x H, or H wide and W high for the quarter turns.  Pinned against the reference's own rotateLine / CLIP / HLRecovery_blend in
 * tests/test_oracle_getimage.py. */
int artoracle_getimage(int W, int H, const float* r, const float* g, const float* b, long stride, const float* mul, int doClip,
                       int doHr, const float* hlmax, int tran, float* outr, float* outg, float* outb, long ostride)
{
    const int rot = tran & 3, swap = rot == 1 || rot == 3;
    const int ow = swap ? H : W, oh = swap ? W : H;
    float* lr = (float*)malloc(sizeof(float) * (size_t)W * 3);
    if (!lr) return 1;
    float *lg = lr + W, *lb = lg + W;
    for (int i = 0; i < H; ++i) {
        for (int j = 0; j < W; ++j) {
            const size_t k = (size_t)i * stride + j;
            float x = r[k] * mul[0], y = g[k] * mul[1], z = b[k] * mul[2];
            if (doClip) { x = clip65535_(x); y = clip65535_(y); z = clip65535_(z); }
            lr[j] = x; lg[j] = y; lb[j] = z;
        }
        if (doHr) artoracle_hl_blend(lr, lg, lb, W, 65535.0f, hlmax);
        for (int j = 0; j < W; ++j) {
            int row, col;
            switch (rot) {
            case 2: row = H - 1 - i; col = W - 1 - j; break;      /* channel(h - 1 - i, w - 1 - j) = line[j] */
            case 1: row = j; col = H - 1 - i; break;              /* channel(j, h - 1 - i) */
            case 3: row = W - 1 - j; col = i; break;              /* channel(w - 1 - j, i) */
            default: row = i; col = j;
            }
            if (tran & 8) col = ow - 1 - col;                     /* hflip, then vflip: both are involutions on disjoint axes */
            if (tran & 4) row = oh - 1 - row;
            const size_t o = (size_t)row * ostride + col;
            outr[o] = lr[j]; outg[o] = lg[j]; outb[o] = lb[j];
        }
    }
    free(lr);
    return 0;
}

/* RawImageSource::transformRect for a standard CCD (rawimagesource.cc L664-751; no D1X, no Fuji SuperCCD): the source rectangle of a PreviewProps
 * window (x, y, w, h in the coordinates of the TRANSFORMED, border-cropped image; skip) under the coarse transform, and the size of the image
 * getImage renders from it.  out4 = sx1, sy1, width, height. */
int artoracle_transform_rect(int W, int H, int border, int x, int y, int w, int h, int skip, int tran, int* out4)
{
    int pp_x = x + border, pp_y = y + border, pp_width = w, pp_height = h;
    const int rot = tran & 3;
    int sw = W, sh = H;
    if (rot == 1 || rot == 3) { sw = H; sh = W; }
    if (pp_width > sw - 2 * border) pp_width = sw - 2 * border;
    if (pp_height > sh - 2 * border) pp_height = sh - 2 * border;
    int ppx = pp_x, ppy = pp_y;
    if (tran & 8) { ppx = sw - pp_x - pp_width; if (ppx < 0) ppx = 0; }
    if (tran & 4) { ppy = sh - pp_y - pp_height; if (ppy < 0) ppy = 0; }
#define MIN_(a, b) ((a) < (b) ? (a) : (b))
#define MAX_(a, b) ((a) > (b) ? (a) : (b))
    int sx1 = ppx, sy1 = ppy, sx2 = MIN_(ppx + pp_width, W - 1), sy2 = MIN_(ppy + pp_height, H - 1);
    if (rot == 2) {
        sx1 = MAX_(W - ppx - pp_width, 0); sy1 = MAX_(H - ppy - pp_height, 0);
        sx2 = MIN_(sx1 + pp_width, W - 1); sy2 = MIN_(sy1 + pp_height, H - 1);
    } else if (rot == 1) {
        sx1 = ppy; sy1 = MAX_(H - ppx - pp_width, 0);
        sx2 = MIN_(sx1 + pp_height, W - 1); sy2 = MIN_(sy1 + pp_width, H - 1);
    } else if (rot == 3) {
        sx1 = MAX_(W - ppy - pp_height, 0); sy1 = ppx;
        sx2 = MIN_(sx1 + pp_height, W - 1); sy2 = MIN_(sy1 + pp_width, H - 1);
    }
#undef MIN_
#undef MAX_
    out4[0] = sx1; out4[1] = sy1;
    out4[2] = (sx2 + 1 - sx1) / skip + ((sx2 + 1 - sx1) % skip > 0);
    out4[3] = (sy2 + 1 - sy1) / skip + ((sy2 + 1 - sy1) % skip > 0);
    return 0;
}

/* RawImageSource::getImage's line loop at any skip (rawimagesource.cc L943-1025): output pixel (ix, j) is the skip x skip box sum (rows outer,
 * columns inner, from 0) of the demosaiced planes at (min(sy1 + skip ix, H - skip), min(sx1 + skip j, W - skip)) times the gain (the caller's
 * rm / gm / bm already hold the 1 / skip^2 of L928-931), clipped, optionally highlight-reconstructed, then turned / mirrored as at skip 1. */
int artoracle_getimage_pp(int W, int H, const float* r, const float* g, const float* b, long stride, const float* mul, int doClip,
                          int doHr, const float* hlmax, int tran, int sx1, int sy1, int imwidth, int imheight, int skip,
                          float* outr, float* outg, float* outb, long ostride)
{
    const int rot = tran & 3, swap = rot == 1 || rot == 3;
    const int ow = swap ? imheight : imwidth, oh = swap ? imwidth : imheight;
    float* lr = (float*)malloc(sizeof(float) * (size_t)imwidth * 3);
    if (!lr) return 1;
    float *lg = lr + imwidth, *lb = lg + imwidth;
    for (int ix = 0; ix < imheight; ++ix) {
        int i = sy1 + skip * ix;
        if (i > H - skip) i = H - skip;
        for (int j = 0; j < imwidth; ++j) {
            int jx = sx1 + skip * j;
            if (jx > W - skip) jx = W - skip;
            float rtot = 0.f, gtot = 0.f, btot = 0.f;
            for (int m = 0; m < skip; m++)
                for (int n = 0; n < skip; n++) {
                    const size_t k = (size_t)(i + m) * stride + (jx + n);
                    rtot += r[k]; gtot += g[k]; btot += b[k];
                }
            rtot *= mul[0]; gtot *= mul[1]; btot *= mul[2];
            if (doClip) { rtot = clip65535_(rtot); gtot = clip65535_(gtot); btot = clip65535_(btot); }
            lr[j] = rtot; lg[j] = gtot; lb[j] = btot;
        }
        if (doHr) artoracle_hl_blend(lr, lg, lb, imwidth, 65535.0f, hlmax);
        for (int j = 0; j < imwidth; ++j) {
            int row, col;
            switch (rot) {
            case 2: row = imheight - 1 - ix; col = imwidth - 1 - j; break;
            case 1: row = j; col = imheight - 1 - ix; break;
            case 3: row = imwidth - 1 - j; col = ix; break;
            default: row = ix; col = j;
            }
            if (tran & 8) col = ow - 1 - col;
            if (tran & 4) row = oh - 1 - row;
            const size_t o = (size_t)row * ostride + col;
            outr[o] = lr[j]; outg[o] = lg[j]; outb[o] = lb[j];
        }
    }
    free(lr);
    return 0;
}
