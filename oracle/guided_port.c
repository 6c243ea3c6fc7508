/*
 * oracle/guided_port.c -- CPU restatement of the float box blur and the fast guided filter.
 * TEST INFRASTRUCTURE ONLY.
 *
 * artoracle_boxblur: rtengine::boxblur(float** src, float** dst, int radius, int W, int H, bool) (reference
 *   rtengine/boxblur.h L318-556), SSE2 build: running mean along rows (dividing by the window length,
 *   L350-381) then along columns (multiplying by 1/len in the steady state, L383-443 and L500-553 -- the
 *   8-column vector groups and the leftover columns perform the same float operations).
 * artoracle_guided_filter: rtengine::guidedFilter (rtengine/guidedfilter.cc L80-241): He/Sun fast guided filter
 *   with the subsampling rule of calculate_subsampling (L58-75), bilinear resampling of rescale.h L27-72.
 * Pinned bit-exact against the reference's own boxblur.h / guidedfilter.cc compiled in place (oracle/_ref) in
 * tests/test_oracle_guided.py.  Compile with -ffp-contract=off.
 */
#include <stdlib.h>
#include <string.h>

/* one line of the running mean; x may alias y (ring of radius+1 old samples kept in `ring`) */
static void box_line(const float* x, long xs, float* y, long ys, int n, int radius, float* ring, int use_rlen)
{
    float len = radius + 1;
    float t = x[0];
    ring[0] = t;
    for (int j = 1; j <= radius; j++) t += x[j * xs];
    t /= len;
    /* keep the samples we are about to overwrite when running in place */
    for (int c = 1; c <= radius; c++) ring[c] = x[c * xs];
    y[0] = t;
    for (int c = 1; c <= radius; c++) {
        t = (t * len + x[(long)(c + radius) * xs]) / (len + 1);
        y[c * ys] = t;
        ++len;
    }
    const float rlen = 1.f / len;
    int pos = 0;
    for (int c = radius + 1; c < n - radius; c++) {
        const float old = ring[pos];
        ring[pos] = x[c * xs];
        if (use_rlen) t = t + (x[(long)(c + radius) * xs] - old) * rlen;
        else t = t + (x[(long)(c + radius) * xs] - old) / len;
        y[c * ys] = t;
        ++pos;
        pos = pos <= radius ? pos : 0;
    }
    for (int c = n - radius; c < n; c++) {
        t = (t * len - ring[pos]) / (len - 1);
        y[c * ys] = t;
        --len;
        ++pos;
        pos = pos <= radius ? pos : 0;
    }
}

/* strides in floats; src == dst allowed */
int artoracle_boxblur(const float* src, long ss, float* dst, long ds, int W, int H, int radius)
{
    if (radius == 0) {
        if (src != dst) for (int i = 0; i < H; i++) memcpy(dst + (long)i * ds, src + (long)i * ss, (size_t)W * sizeof(float));
        return 0;
    }
    if (2 * radius + 1 > W || 2 * radius + 1 > H) return 1;
    float* ring = (float*)malloc(sizeof(float) * (size_t)(radius + 1));
    if (!ring) return 1;
    for (int r = 0; r < H; r++) {
        /* the horizontal pass reads src[col + radius] ahead of the write position and the ring for the samples
         * behind it, so it is in-place safe exactly like the reference's lineBuffer */
        const float* x = src + (long)r * ss;
        float* y = dst + (long)r * ds;
        if (x == y) {
            box_line(y, 1, y, 1, W, radius, ring, 0);
        } else {
            box_line(x, 1, y, 1, W, radius, ring, 0);
        }
    }
    for (int c = 0; c < W; c++) box_line(dst + c, ds, dst + c, ds, H, radius, ring, 1);
    free(ring);
    return 0;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static float bilinear(const float* s, int W, int H, float x, float y)
{   /* getBilinearValue, rescale.h L27-51 */
    const int xi = imin((int)x, W - 1), yi = imin((int)y, H - 1);
    const float xf = x - xi, yf = y - yi;
    const int xi1 = imin(xi + 1, W - 1), yi1 = imin(yi + 1, H - 1);
    const float bl = s[(long)yi * W + xi], br = s[(long)yi * W + xi1], tl = s[(long)yi1 * W + xi], tr = s[(long)yi1 * W + xi1];
    const float b = xf * br + (1.f - xf) * bl;
    const float t = xf * tr + (1.f - xf) * tl;
    return yf * t + (1.f - yf) * b;
}

static void resample(const float* s, long ss, int Ws, int Hs, float* d, int Wd, int Hd, float* packed)
{   /* f_subsample, guidedfilter.cc L144-159: copy when the sizes agree, else rescaleBilinear (rescale.h L54-72) */
    if (Ws == Wd && Hs == Hd) {
        for (int y = 0; y < Hs; y++) memcpy(d + (long)y * Wd, s + (long)y * ss, (size_t)Ws * sizeof(float));
        return;
    }
    for (int y = 0; y < Hs; y++) memcpy(packed + (long)y * Ws, s + (long)y * ss, (size_t)Ws * sizeof(float));
    const float col_scale = (float)Ws / (float)Wd, row_scale = (float)Hs / (float)Hd;
    for (int y = 0; y < Hd; y++) {
        const float ymrs = y * row_scale;
        for (int x = 0; x < Wd; x++) d[(long)y * Wd + x] = bilinear(packed, Ws, Hs, x * col_scale, ymrs);
    }
}

int artoracle_guided_subsampling(int w, int h, int r)
{   /* calculate_subsampling, guidedfilter.cc L58-75 */
    if (r == 1) return 1;
    if (imax(w, h) <= 600) return 1;
    for (int s = 5; s > 0; --s) if (r % s == 0) return s;
    return imax(2, imin(r / 2, 4));
}

/* guide, src, dst: W x H planes with a common stride (floats); dst may alias src or guide */
int artoracle_guided_filter(const float* guide, const float* src, float* dst, long stride, int W, int H,
                            int r, float epsilon, int subsampling)
{
    if (subsampling <= 0) subsampling = artoracle_guided_subsampling(W, H, r);
    const int w = W / subsampling, h = H / subsampling;
    const size_t n = (size_t)w * h;
    float* pool = (float*)malloc(sizeof(float) * (4 * n + (size_t)W * H));
    if (!pool) return 1;
    float *I1 = pool, *p1 = pool + n, *meanI = pool + 2 * n, *meanp = pool + 3 * n, *packed = pool + 4 * n;
    resample(guide, stride, W, H, I1, w, h, packed);
    resample(src, stride, W, H, p1, w, h, packed);
    const float r1 = (float)r / subsampling;
    /* f_mean: rad = LIM(rad, 0, (min(w,h) - 1) / 2 - 1), L165-169 (int rad <- float r1) */
    int rad = (int)r1;
    { const int hi = (imin(w, h) - 1) / 2 - 1; rad = imax(0, imin(rad, hi)); }
    int rc = 0;
    rc |= artoracle_boxblur(I1, w, meanI, w, w, h, rad);
    rc |= artoracle_boxblur(p1, w, meanp, w, w, h, rad);
    for (size_t k = 0; k < n; k++) p1[k] = I1[k] * p1[k];                 /* corrIp = I1 * p1 (MUL) */
    rc |= artoracle_boxblur(p1, w, p1, w, w, h, rad);
    for (size_t k = 0; k < n; k++) I1[k] = I1[k] * I1[k];                 /* corrI = I1 * I1 */
    rc |= artoracle_boxblur(I1, w, I1, w, w, h, rad);
    for (size_t k = 0; k < n; k++) I1[k] = I1[k] - (meanI[k] * meanI[k]);   /* varI = corrI - meanI*meanI (SUBMUL) */
    for (size_t k = 0; k < n; k++) p1[k] = p1[k] - (meanI[k] * meanp[k]);   /* covIp = corrIp - meanI*meanp */
    for (size_t k = 0; k < n; k++) I1[k] = p1[k] / (I1[k] + epsilon);       /* a = covIp / (varI + eps) (DIVEPSILON) */
    for (size_t k = 0; k < n; k++) p1[k] = meanp[k] - (I1[k] * meanI[k]);   /* b = meanp - a*meanI */
    rc |= artoracle_boxblur(I1, w, I1, w, w, h, rad);                       /* meana */
    rc |= artoracle_boxblur(p1, w, p1, w, w, h, rad);                       /* meanb */
    const float col_scale = (float)w / (float)W, row_scale = (float)h / (float)H;
    for (int y = 0; y < H; y++) {
        const float ymrs = y * row_scale;
        for (int x = 0; x < W; x++)
            dst[(long)y * stride + x] = bilinear(I1, w, h, x * col_scale, ymrs) * guide[(long)y * stride + x] + bilinear(p1, w, h, x * col_scale, ymrs);
    }
    free(pool);
    return rc;
}
