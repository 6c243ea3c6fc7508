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

static inline void fftwf_destroy_plan(fftwf_plan p) { if (p) { free(p->c0); free(p->c1); free(p); } }

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
#endif
