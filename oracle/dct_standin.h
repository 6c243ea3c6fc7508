/*
 * oracle/dct_standin.h -- stand-in for the four FFTW entry points rtengine/FTblockDN.cc calls (L1921-1931, L1604,
 * L1614, L2661-2664).  TEST INFRASTRUCTURE ONLY.
 *
 * fftw3f is an external dependency of the reference (CMakeLists.txt L447, version unpinned) and is absent from this
 * image, so the reference's block DCT cannot be run here: **parity is unpinned at this boundary**.  This header
 * restates the published definitions (FFTW manual, "1d Real-even DFTs", unnormalised):
 *   REDFT10:  Y_k = 2 sum_j X_j cos(pi (j + 1/2) k / n)
 *   REDFT01:  Y_k = X_0 + 2 sum_{j>=1} X_j cos(pi j (k + 1/2) / n)
 * applied separably to each n0 x n1 block, accumulated in double and rounded once to float.  FFTW's float
 * codelets differ from this by a few ulp of the block's largest coefficient; tests on anything downstream of the DCT
 * therefore carry a tolerance (1e-4 relative, BASELINE.json north_star) instead of bit equality.
 * Used by oracle/_ref's shim (as <fftw3.h>) and by oracle/denoise_port.c.
 */
#ifndef ART_ORACLE_DCT_STANDIN_H
#define ART_ORACLE_DCT_STANDIN_H
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef enum { FFTW_REDFT00 = 3, FFTW_REDFT01 = 4, FFTW_REDFT10 = 5 } fftw_r2r_kind;
#define FFTW_MEASURE 0u
#define FFTW_ESTIMATE (1u << 6)
#define FFTW_DESTROY_INPUT (1u << 0)

typedef struct artdct_plan_s {
    int n0, n1, howmany, dist;
    fftw_r2r_kind k0, k1;
    double* c0;   /* [n0][n0] transform matrix along dimension 0: out[k] = sum_j c0[k][j] in[j] */
    double* c1;
} * fftwf_plan;

static inline void* fftwf_malloc(size_t n) { void* p = NULL; return posix_memalign(&p, 64, n) ? NULL : p; }
static inline void fftwf_free(void* p) { free(p); }

static inline double* artdct_matrix(int n, fftw_r2r_kind kind)
{
    double* c = (double*)malloc(sizeof(double) * (size_t)n * n);
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) {
            if (kind == FFTW_REDFT10) c[k * n + j] = 2.0 * cos(pi * (j + 0.5) * k / n);
            else /* REDFT01 */        c[k * n + j] = j == 0 ? 1.0 : 2.0 * cos(pi * j * (k + 0.5) / n);
        }
    return c;
}

static inline fftwf_plan fftwf_plan_many_r2r(int rank, const int* n, int howmany, float* in, const int* inembed, int istride, int idist,
                                             float* out, const int* onembed, int ostride, int odist, const fftw_r2r_kind* kind, unsigned flags)
{
    (void)in; (void)inembed; (void)out; (void)onembed; (void)flags;
    if (rank != 2 || istride != 1 || ostride != 1 || idist != odist) return NULL;
    if ((kind[0] != FFTW_REDFT10 && kind[0] != FFTW_REDFT01) || (kind[1] != FFTW_REDFT10 && kind[1] != FFTW_REDFT01)) return NULL;
    fftwf_plan p = (fftwf_plan)malloc(sizeof(*p));
    p->n0 = n[0]; p->n1 = n[1]; p->howmany = howmany; p->dist = idist; p->k0 = kind[0]; p->k1 = kind[1];
    p->c0 = artdct_matrix(n[0], kind[0]);
    p->c1 = artdct_matrix(n[1], kind[1]);
    return p;
}

static inline void fftwf_destroy_plan(fftwf_plan p) { if (p) { if (p->k0 != FFTW_REDFT00) { free(p->c0); free(p->c1); } free(p); } }

/* one n0 x n1 block, in -> out (may alias) */
static inline void artdct_block(const struct artdct_plan_s* p, const float* in, float* out)
{
    const int n0 = p->n0, n1 = p->n1;
    double* t = (double*)malloc(sizeof(double) * (size_t)n0 * n1);
    double* u = (double*)malloc(sizeof(double) * (size_t)n0 * n1);
    for (int i = 0; i < n0; ++i)            /* along dimension 1 (rows) */
        for (int k = 0; k < n1; ++k) {
            double s = 0.0;
            const double* c = p->c1 + (size_t)k * n1;
            for (int j = 0; j < n1; ++j) s += c[j] * (double)in[i * n1 + j];
            t[i * n1 + k] = s;
        }
    for (int k = 0; k < n0; ++k) {          /* along dimension 0 (columns) */
        const double* c = p->c0 + (size_t)k * n0;
        for (int x = 0; x < n1; ++x) u[k * n1 + x] = 0.0;
        for (int j = 0; j < n0; ++j) {
            const double cj = c[j];
            for (int x = 0; x < n1; ++x) u[k * n1 + x] += cj * t[j * n1 + x];
        }
    }
    for (int i = 0; i < n0 * n1; ++i) out[i] = (float)u[i];
    free(t); free(u);
}

static inline void fftwf_execute_r2r(const fftwf_plan p, float* in, float* out)
{
    for (int b = 0; b < p->howmany; ++b) artdct_block(p, in + (size_t)b * p->dist, out + (size_t)b * p->dist);
}

/* ---------------------------------------------------------------------------------------------------------------
 * REDFT00 (DCT-I), 2-D, out of place or in place: the two calls in rtengine/tmo_fattal02.cc L768-772 and L784-788
 * (fftwf_plan_r2r_2d(height, width, in, out, FFTW_REDFT00, FFTW_REDFT00, FFTW_ESTIMATE) + fftwf_execute).
 * Published definition (FFTW manual, unnormalised):  Y_k = X_0 + (-1)^k X_{n-1} + 2 sum_{j=1}^{n-2} X_j cos(pi j k/(n-1)).
 * Evaluated in double through the length-2(n-1) DFT of the even extension (mixed-radix recursion below; any prime
 * factor falls back to the O(p^2) definition), applied separably, rounded ONCE to float at the end.
 * tests/test_oracle_fattal.py checks this routine against the literal cosine sum.  Parity unpinned at this boundary
 * for the same reason as above (fftw3f absent).
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct { double re, im; } artdct_cplx;

static inline void artdct_fft_rec(int n, const artdct_cplx* in, int stride, artdct_cplx* out, const artdct_cplx* tw, int twstep)
{   /* out[k] = sum_j in[j*stride] * w^(j k),  w = exp(-2 pi i / n) = tw[twstep] */
    if (n == 1) { out[0] = in[0]; return; }
    int p = 2;
    while (p * p <= n && n % p) ++p;
    if (n % p) p = n;
    const int m = n / p;
    for (int q = 0; q < p; ++q) artdct_fft_rec(m, in + (size_t)q * stride, stride * p, out + (size_t)q * m, tw, twstep * p);
    artdct_cplx tstack[16];
    artdct_cplx* t = p <= 16 ? tstack : (artdct_cplx*)malloc(sizeof(artdct_cplx) * p);
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q) t[q] = out[(size_t)q * m + k];
        for (int j = 0; j < p; ++j) {
            const int kk = k + j * m;
            double sr = 0.0, si = 0.0;
            for (int q = 0; q < p; ++q) {
                const artdct_cplx w = tw[(size_t)(((long long)q * kk) % n) * twstep];
                sr += t[q].re * w.re - t[q].im * w.im;
                si += t[q].re * w.im + t[q].im * w.re;
            }
            out[kk].re = sr; out[kk].im = si;
        }
    }
    if (t != tstack) free(t);
}

/* 1-D DCT-I of n samples x[0..n-1] (stride xs) into y (stride ys), double */
static inline void artdct_redft00_1d(int n, const double* x, int xs, double* y, int ys, const artdct_cplx* tw, artdct_cplx* a, artdct_cplx* b)
{
    const int N = n - 1, M = 2 * N;
    for (int j = 0; j <= N; ++j) { a[j].re = x[(size_t)j * xs]; a[j].im = 0.0; }
    for (int j = N + 1; j < M; ++j) { a[j].re = x[(size_t)(M - j) * xs]; a[j].im = 0.0; }
    artdct_fft_rec(M, a, 1, b, tw, 1);
    for (int k = 0; k <= N; ++k) y[(size_t)k * ys] = b[k].re;
}

static inline artdct_cplx* artdct_twiddles(int M)
{
    artdct_cplx* tw = (artdct_cplx*)malloc(sizeof(artdct_cplx) * M);
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < M; ++k) { tw[k].re = cos(2.0 * pi * k / M); tw[k].im = -sin(2.0 * pi * k / M); }
    return tw;
}

/* 2-D DCT-I of an n0 x n1 (rows x cols) float array, in -> out (may alias) */
static inline void artdct_redft00_2d(int n0, int n1, const float* in, float* out)
{
    double* t = (double*)malloc(sizeof(double) * (size_t)n0 * n1);
    for (size_t i = 0; i < (size_t)n0 * n1; ++i) t[i] = (double)in[i];
    artdct_cplx* tw1 = artdct_twiddles(2 * (n1 - 1));
    artdct_cplx* tw0 = artdct_twiddles(2 * (n0 - 1));
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        const int mx = 2 * ((n0 > n1 ? n0 : n1) - 1);
        artdct_cplx* a = (artdct_cplx*)malloc(sizeof(artdct_cplx) * mx);
        artdct_cplx* b = (artdct_cplx*)malloc(sizeof(artdct_cplx) * mx);
        double* line = (double*)malloc(sizeof(double) * (n0 > n1 ? n0 : n1));
#ifdef _OPENMP
#pragma omp for
#endif
        for (int i = 0; i < n0; ++i) {
            artdct_redft00_1d(n1, t + (size_t)i * n1, 1, line, 1, tw1, a, b);
            memcpy(t + (size_t)i * n1, line, sizeof(double) * n1);
        }
#ifdef _OPENMP
#pragma omp for
#endif
        for (int x = 0; x < n1; ++x) {
            artdct_redft00_1d(n0, t + x, n1, line, 1, tw0, a, b);
            for (int k = 0; k < n0; ++k) t[(size_t)k * n1 + x] = line[k];
        }
        free(a); free(b); free(line);
    }
    for (size_t i = 0; i < (size_t)n0 * n1; ++i) out[i] = (float)t[i];
    free(t); free(tw0); free(tw1);
}

/* the literal cosine sum, for checking the routine above at small sizes */
static inline void artdct_redft00_1d_naive(int n, const double* x, double* y)
{
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < n; ++k) {
        double s = x[0] + ((k & 1) ? -x[n - 1] : x[n - 1]);
        for (int j = 1; j < n - 1; ++j) s += 2.0 * x[j] * cos(pi * (double)j * (double)k / (double)(n - 1));
        y[k] = s;
    }
}

static inline fftwf_plan fftwf_plan_r2r_2d(int n0, int n1, float* in, float* out, fftw_r2r_kind k0, fftw_r2r_kind k1, unsigned flags)
{
    (void)flags;
    if (k0 != FFTW_REDFT00 || k1 != FFTW_REDFT00) return NULL;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->n0 = n0; p->n1 = n1; p->k0 = k0; p->k1 = k1;
    p->c0 = (double*)in; p->c1 = (double*)out;      /* REDFT00 plans carry their arrays (fftwf_execute takes none) */
    return p;
}
static inline void fftwf_execute(const fftwf_plan p) { artdct_redft00_2d(p->n0, p->n1, (const float*)p->c0, (float*)p->c1); }
#endif
