/*
 * oracle/gauss_port.c -- CPU restatement of the reference's gaussianBlur, GAUSS_STANDARD, no box buffer.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates gaussianBlurImpl's dispatch (reference rtengine/gauss.cc L1387-1567) and the kernels it reaches on
 * an x86-64 (SSE2) build for gausstype == GAUSS_STANDARD and buffer == nullptr:
 *   sigma < 0.25            copy (L1438-1444)
 *   sigma < 0.6, src != dst gauss3x3 (L129-174), coefficients computed in double, applied in float
 *   sigma < 0.6, src == dst gaussHorizontal3 (L446-464) + gaussVertical3 (L467-526)
 *   sigma < 25              gaussHorizontalSse (L554-665) + gaussVerticalSse (L716-856): Young-van Vliet IIR with
 *                           Triggs-Sdika boundaries; rows in the 4-row vector groups (columns in the 8-column
 *                           groups) use float arithmetic with float coefficients, the leftover H%4 rows (W%8
 *                           columns) use the scalar loop, which computes in double with double coefficients and
 *                           rounds to float on every store to the float scratch
 *   sigma >= 25             gaussHorizontal (L669-713) + gaussVertical (L1148-1225): everything in double
 * artoracle_gauss_iir: the recursive branch of GAUSS_MULT / GAUSS_DIV (L1490-1511: gaussHorizontalSse + gaussVerticalSsemult L860-999 /
 * gaussVerticalSsediv L1002-1144), the forms deconvsharpening reaches above sigma 1.15.
 * Pinned bit-exact against the reference's own gauss.cc compiled in place (oracle/_ref) in tests/test_oracle_gauss.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void yvv_factors(double sigma, double* b1, double* b2, double* b3, double* B, double M[3][3])
{   /* calculateYvVFactors<double>, L94-126 */
    double q;
    if (sigma < 2.5) q = 3.97156 - 4.14554 * sqrt(1.0 - 0.26891 * sigma);
    else q = 0.98711 * sigma - 0.96330;
    double b0 = 1.57825 + 2.44413 * q + 1.4281 * q * q + 0.422205 * q * q * q;
    *b1 = 2.44413 * q + 2.85619 * q * q + 1.26661 * q * q * q;
    *b2 = -1.4281 * q * q - 1.26661 * q * q * q;
    *b3 = 0.422205 * q * q * q;
    *B = 1.0 - (*b1 + *b2 + *b3) / b0;
    *b1 /= b0; *b2 /= b0; *b3 /= b0;
    const double c1 = *b1, c2 = *b2, c3 = *b3;
    M[0][0] = -c3 * c1 + 1.0 - c3 * c3 - c2;
    M[0][1] = (c3 + c1) * (c2 + c3 * c1);
    M[0][2] = c3 * (c1 + c3 * c2);
    M[1][0] = c1 + c3 * c2;
    M[1][1] = -(c2 - 1.0) * (c2 + c3 * c1);
    M[1][2] = -(c3 * c1 + c3 * c3 + c2 - 1.0) * c3;
    M[2][0] = c3 * c1 + c2 + c1 * c1 - c2 * c2;
    M[2][1] = c1 * c2 + c3 * c2 * c2 - c1 * c3 * c3 - c3 * c3 * c3 - c3 * c2 + c3;
    M[2][2] = c3 * (c1 + c3 * c2);
}

/* One line (row or column) of the SSE float form: x[k*xs] -> y[k*ys], n samples.  tmp: n floats. */
static void yvv_line_float(const float* x, long xs, float* y, long ys, int n, float* tmp,
                           float B, float b1, float b2, float b3, const float Mf[3][3], int vertical)
{
    /* causal pass (L586-604 / L755-790) */
    float T = x[0];
    float Tm3 = T * (B + b1 + b2 + b3);
    tmp[0] = Tm3;
    float Tm2;
    if (vertical) Tm2 = x[xs] * B + Tm3 * b1 + T * (b2 + b3);        /* L766: LVFU(src[1]) * Bv + Rv * b1v + Tv * (b2v + b3v) */
    else Tm2 = x[xs] * B + Tm3 * b1 + T * (b2 + b3);                   /* L590 */
    tmp[1] = Tm2;
    float R = x[2 * xs] * B + Tm2 * b1 + Tm3 * b2 + T * b3;
    tmp[2] = R;
    for (int j = 3; j < n; j++) {
        T = R;
        R = x[j * xs] * B + T * b1 + Tm2 * b2 + Tm3 * b3;
        tmp[j] = R;
        Tm3 = Tm2;
        Tm2 = T;
    }
    /* boundary (L606-624 / L792-815) */
    T = x[(long)(n - 1) * xs];
    const float t2Wp1 = T + Mf[2][0] * (R - T) + Mf[2][1] * (Tm2 - T) + Mf[2][2] * (Tm3 - T);
    const float t2W = T + Mf[1][0] * (R - T) + Mf[1][1] * (Tm2 - T) + Mf[1][2] * (Tm3 - T);
    R = T + Mf[0][0] * (R - T) + Mf[0][1] * (Tm2 - T) + Mf[0][2] * (Tm3 - T);
    float* out = vertical ? NULL : tmp;     /* horizontal keeps writing tmp and copies out at the end; vertical writes dst */
    if (out) out[n - 1] = R; else y[(long)(n - 1) * ys] = R;
    Tm2 = B * Tm2 + b1 * R + b2 * t2W + b3 * t2Wp1;
    if (out) out[n - 2] = Tm2; else y[(long)(n - 2) * ys] = Tm2;
    Tm3 = B * Tm3 + b1 * Tm2 + b2 * R + b3 * t2W;
    if (out) out[n - 3] = Tm3; else y[(long)(n - 3) * ys] = Tm3;
    T = R; R = Tm3; Tm3 = T;
    /* anticausal pass (L626-632 / L824-834) */
    for (int j = n - 4; j >= 0; j--) {
        T = R;
        R = tmp[j] * B + T * b1 + Tm2 * b2 + Tm3 * b3;
        if (out) out[j] = R; else y[(long)j * ys] = R;
        Tm3 = Tm2;
        Tm2 = T;
    }
    if (out) for (int j = 0; j < n; j++) y[(long)j * ys] = tmp[j];
}

/* One line of the scalar remainder loops (L646-664 / L843-855): double arithmetic, float scratch */
static void yvv_line_mixed(const float* x, long xs, float* y, long ys, int n, float* tmp,
                           double B, double b1, double b2, double b3, double M[3][3])
{
    tmp[0] = x[0] * (B + b1 + b2 + b3);
    tmp[1] = B * x[xs] + b1 * tmp[0] + x[0] * (b2 + b3);
    tmp[2] = B * x[2 * xs] + b1 * tmp[1] + b2 * tmp[0] + b3 * x[0];
    for (int j = 3; j < n; j++) tmp[j] = B * x[j * xs] + b1 * tmp[j - 1] + b2 * tmp[j - 2] + b3 * tmp[j - 3];
    const float xl = x[(long)(n - 1) * xs];
    const float t2Wm1 = xl + M[0][0] * (tmp[n - 1] - xl) + M[0][1] * (tmp[n - 2] - xl) + M[0][2] * (tmp[n - 3] - xl);
    const float t2W = xl + M[1][0] * (tmp[n - 1] - xl) + M[1][1] * (tmp[n - 2] - xl) + M[1][2] * (tmp[n - 3] - xl);
    const float t2Wp1 = xl + M[2][0] * (tmp[n - 1] - xl) + M[2][1] * (tmp[n - 2] - xl) + M[2][2] * (tmp[n - 3] - xl);
    tmp[n - 1] = t2Wm1;
    tmp[n - 2] = B * tmp[n - 2] + b1 * tmp[n - 1] + b2 * t2W + b3 * t2Wp1;
    tmp[n - 3] = B * tmp[n - 3] + b1 * tmp[n - 2] + b2 * tmp[n - 1] + b3 * t2W;
    for (int j = n - 4; j >= 0; j--) tmp[j] = B * tmp[j] + b1 * tmp[j + 1] + b2 * tmp[j + 2] + b3 * tmp[j + 3];
    for (int j = 0; j < n; j++) y[(long)j * ys] = tmp[j];
}

/* One line of the all-double form (gaussHorizontal L685-711, gaussVertical L1168-1223) */
static void yvv_line_double(const float* x, long xs, float* y, long ys, int n, double* t,
                            double B, double b1, double b2, double b3, double M[3][3])
{
    t[0] = B * x[0] + b1 * x[0] + b2 * x[0] + b3 * x[0];
    t[1] = B * x[xs] + b1 * t[0] + b2 * x[0] + b3 * x[0];
    t[2] = B * x[2 * xs] + b1 * t[1] + b2 * t[0] + b3 * x[0];
    for (int j = 3; j < n; j++) t[j] = B * x[j * xs] + b1 * t[j - 1] + b2 * t[j - 2] + b3 * t[j - 3];
    const float xl = x[(long)(n - 1) * xs];
    const double tm1 = xl + M[0][0] * (t[n - 1] - xl) + M[0][1] * (t[n - 2] - xl) + M[0][2] * (t[n - 3] - xl);
    const double tw = xl + M[1][0] * (t[n - 1] - xl) + M[1][1] * (t[n - 2] - xl) + M[1][2] * (t[n - 3] - xl);
    const double tp1 = xl + M[2][0] * (t[n - 1] - xl) + M[2][1] * (t[n - 2] - xl) + M[2][2] * (t[n - 3] - xl);
    t[n - 1] = tm1;
    t[n - 2] = B * t[n - 2] + b1 * t[n - 1] + b2 * tw + b3 * tp1;
    t[n - 3] = B * t[n - 3] + b1 * t[n - 2] + b2 * t[n - 1] + b3 * tw;
    for (int j = n - 4; j >= 0; j--) t[j] = B * t[j] + b1 * t[j + 1] + b2 * t[j + 2] + b3 * t[j + 3];
    for (int j = 0; j < n; j++) y[(long)j * ys] = (float)t[j];
}

static void gauss3x3(const float* src, long ss, float* dst, long ds, int W, int H, float c0, float c1, float c2, float b0, float b1)
{   /* L129-174 */
#define S(i, j) src[(long)(i) * ss + (j)]
#define D(i, j) dst[(long)(i) * ds + (j)]
    D(0, 0) = S(0, 0);
    for (int j = 1; j < W - 1; j++) D(0, j) = b1 * (S(0, j - 1) + S(0, j + 1)) + b0 * S(0, j);
    D(0, W - 1) = S(0, W - 1);
    for (int i = 1; i < H - 1; i++) {
        D(i, 0) = b1 * (S(i - 1, 0) + S(i + 1, 0)) + b0 * S(i, 0);
        for (int j = 1; j < W - 1; j++)
            D(i, j) = c2 * (S(i - 1, j - 1) + S(i - 1, j + 1) + S(i + 1, j - 1) + S(i + 1, j + 1)) + c1 * (S(i - 1, j) + S(i, j - 1) + S(i, j + 1) + S(i + 1, j)) + c0 * S(i, j);
        D(i, W - 1) = b1 * (S(i - 1, W - 1) + S(i + 1, W - 1)) + b0 * S(i, W - 1);
    }
    D(H - 1, 0) = S(H - 1, 0);
    for (int j = 1; j < W - 1; j++) D(H - 1, j) = b1 * (S(H - 1, j - 1) + S(H - 1, j + 1)) + b0 * S(H - 1, j);
    D(H - 1, W - 1) = S(H - 1, W - 1);
#undef S
#undef D
}

/* src == dst allowed (pass the same pointer).  Strides in floats. */
int artoracle_gauss(const float* src, long ss, float* dst, long ds, int W, int H, double sigma)
{
    if (W < 4 || H < 4) return 1;
    const int inplace = (src == dst);
    if (sigma < 0.25) {
        if (!inplace) for (int i = 0; i < H; i++) memcpy(dst + (long)i * ds, src + (long)i * ss, (size_t)W * sizeof(float));
        return 0;
    }
    if (sigma < 0.6) {
        if (!inplace) {   /* L1446-1478 */
            double c0 = 1.0, c1 = exp(-0.5 * ((1.0 / sigma) * (1.0 / sigma))), c2 = exp(-((1.0 / sigma) * (1.0 / sigma)));
            const double sum = c0 + 4.0 * (c1 + c2);
            c0 /= sum; c1 /= sum; c2 /= sum;
            double b1 = exp(-1.0 / (2.0 * sigma * sigma));
            const double bsum = 2.0 * b1 + 1.0;
            b1 /= bsum;
            const double b0 = 1.0 / bsum;
            gauss3x3(src, ss, dst, ds, W, H, (float)c0, (float)c1, (float)c2, (float)b0, (float)b1);
        } else {          /* L1479-1487: separable 3-tap, in place */
            double c1d = exp(-1.0 / (2.0 * sigma * sigma));
            const double csum = 2.0 * c1d + 1.0;
            c1d /= csum;
            const float c1 = (float)c1d, c0 = (float)(1.0 / csum);
            float* t = (float*)malloc(sizeof(float) * (size_t)(W > H ? W : H));
            if (!t) return 1;
            for (int i = 0; i < H; i++) {         /* gaussHorizontal3 L446-464 */
                float* r = dst + (long)i * ds;
                for (int j = 1; j < W - 1; j++) t[j] = (float)(c1 * (r[j - 1] + r[j + 1]) + c0 * r[j]);
                memcpy(r + 1, t + 1, (size_t)(W - 2) * sizeof(float));
            }
            for (int i = 0; i < W; i++) {         /* gaussVertical3 L467-526 (vector and scalar forms agree) */
                for (int j = 1; j < H - 1; j++) t[j] = c1 * (dst[(long)(j + 1) * ds + i] + dst[(long)(j - 1) * ds + i]) + dst[(long)j * ds + i] * c0;
                for (int j = 1; j < H - 1; j++) dst[(long)j * ds + i] = t[j];
            }
            free(t);
        }
        return 0;
    }
    double b1, b2, b3, B, M[3][3];
    yvv_factors(sigma, &b1, &b2, &b3, &B, M);      /* NOTE: the float paths call it with (double)(float)sigma */
    const int n = W > H ? W : H;
    if (sigma < 25.0) {
        const float sigf = (float)sigma;             /* gaussHorizontalSse(..., const float sigma) */
        yvv_factors(sigf, &b1, &b2, &b3, &B, M);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {           /* L559-563 */
                M[i][j] *= (1.0 + b2 + (b1 - b3) * b3);
                M[i][j] /= (1.0 + b1 - b2 + b3) * (1.0 - b1 - b2 - b3);
            }
        float Mf[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Mf[i][j] = (float)M[i][j];
        const float Bf = (float)B, b1f = (float)b1, b2f = (float)b2, b3f = (float)b3;
        int fail = 0;
#pragma omp parallel
        {
            float* tmp = (float*)malloc(sizeof(float) * (size_t)n);
            if (!tmp) {
#pragma omp atomic write
                fail = 1;
            } else {
#pragma omp for
                for (int i = 0; i < H; i++) {
                    const float* x = (inplace ? dst + (long)i * ds : src + (long)i * ss);
                    if (i < H - (H % 4)) yvv_line_float(x, 1, dst + (long)i * ds, 1, W, tmp, Bf, b1f, b2f, b3f, Mf, 0);
                    else yvv_line_mixed(x, 1, dst + (long)i * ds, 1, W, tmp, B, b1, b2, b3, M);
                }
#pragma omp for
                for (int i = 0; i < W; i++) {
                    if (i < W - (W % 8)) yvv_line_float(dst + i, ds, dst + i, ds, H, tmp, Bf, b1f, b2f, b3f, Mf, 1);
                    else yvv_line_mixed(dst + i, ds, dst + i, ds, H, tmp, B, b1, b2, b3, M);
                }
                free(tmp);
            }
        }
        return fail;
    }
    /* sigma >= 25: double (L669-713, L1148-1225) */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i][j] /= (1.0 + b1 - b2 + b3) * (1.0 + b2 + (b1 - b3) * b3);
    int fail = 0;
#pragma omp parallel
    {
        double* t = (double*)malloc(sizeof(double) * (size_t)n);
        if (!t) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for
            for (int i = 0; i < H; i++)
                yvv_line_double(inplace ? dst + (long)i * ds : src + (long)i * ss, 1, dst + (long)i * ds, 1, W, t, B, b1, b2, b3, M);
#pragma omp for
            for (int i = 0; i < W; i++) yvv_line_double(dst + i, ds, dst + i, ds, H, t, B, b1, b2, b3, M);
            free(t);
        }
    }
    return fail;
}

/* gaussianBlur(src, dst, W, H, sigma, nullptr, type, divb) in the recursive branch, 0.6 <= sigma < 25, src != dst (strides in floats):
 *   type 1  GAUSS_MULT  gaussHorizontalSse(src, src) IN PLACE (L1496), then dst *= vertical(src)             (L860-999)
 *   type 2  GAUSS_DIV   gaussHorizontalSse(src, dst), then dst = divb / (v > 0 ? v : 1) with v = vertical(dst) (L1002-1144):
 *                       clamped at 0 by vmaxf for rows < H-3 of the 8-column groups (the three boundary rows L1078-1091 are
 *                       stored unclamped) and by rtengine::max on every row of the W%8 remainder columns (L1139-1141) */
int artoracle_gauss_iir(float* src, long ss, float* dst, long ds, const float* divb, long vs, int W, int H, double sigma, int type)
{
    if (W < 4 || H < 4 || !(sigma >= 0.6) || sigma >= 25.0 || (type != 1 && type != 2) || (type == 2 && !divb)) return 1;
    double b1, b2, b3, B, M[3][3];
    const float sigf = (float)sigma;
    yvv_factors(sigf, &b1, &b2, &b3, &B, M);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            M[i][j] *= (1.0 + b2 + (b1 - b3) * b3);
            M[i][j] /= (1.0 + b1 - b2 + b3) * (1.0 - b1 - b2 - b3);
        }
    float Mf[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Mf[i][j] = (float)M[i][j];
    const float Bf = (float)B, b1f = (float)b1, b2f = (float)b2, b3f = (float)b3;
    const int n = W > H ? W : H;
    float* hp = type == 1 ? src : dst;              /* plane holding the horizontal result */
    const long hs = type == 1 ? ss : ds;
    int fail = 0;
#pragma omp parallel
    {
        float* tmp = (float*)malloc(sizeof(float) * (size_t)n * 2);
        if (!tmp) {
#pragma omp atomic write
            fail = 1;
        } else {
            float* col = tmp + n;
#pragma omp for
            for (int i = 0; i < H; i++) {
                if (i < H - (H % 4)) yvv_line_float(src + (long)i * ss, 1, hp + (long)i * hs, 1, W, tmp, Bf, b1f, b2f, b3f, Mf, 0);
                else yvv_line_mixed(src + (long)i * ss, 1, hp + (long)i * hs, 1, W, tmp, B, b1, b2, b3, M);
            }
#pragma omp for
            for (int i = 0; i < W; i++) {
                const int vec = i < W - (W % 8);
                if (vec) yvv_line_float(hp + i, hs, col, 1, H, tmp, Bf, b1f, b2f, b3f, Mf, 1);
                else yvv_line_mixed(hp + i, hs, col, 1, H, tmp, B, b1, b2, b3, M);
                for (int j = 0; j < H; j++) {
                    float* o = dst + (long)j * ds + i;
                    if (type == 1) *o = *o * col[j];
                    else {
                        float q = divb[(long)j * vs + i] / (col[j] > 0.f ? col[j] : 1.f);
                        if (!vec) q = q < 0.f ? 0.f : q;                    /* rtengine::max(q, 0.f) */
                        else if (j < H - 3) q = q > 0.f ? q : 0.f;          /* vmaxf(q, ZEROV) */
                        *o = q;
                    }
                }
            }
            free(tmp);
        }
    }
    return fail;
}
