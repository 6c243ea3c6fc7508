// Wavelet decomposition / reconstruction for sm_100a.
//
// Replaces rtengine::wavelet_decomposition (reference rtengine/cplx_wavelet_dec.h L97-270) and the filters of
// rtengine::wavelet_level (rtengine/cplx_wavelet_level.h L205-763) for Daub4Len == 6, skipcrop == 1, as
// FTblockDN uses them (subsampling == 1: level 0 decimated Daub4, levels >= 1 undecimated Haar with tap
// spacing 1,2,4,...).  The reference filters a row at a time through two line buffers (vertical then
// horizontal); here every level is ONE kernel, one thread per output coefficient, that evaluates the
// vertical stage for exactly the columns its horizontal stage needs -- the same float operations in the same
// order (bit-exact), no intermediate plane: per level the HBM traffic is one read of the input and one write
// of the four subbands.  Compiled with -fmad=false.
#include "ctx.h"

namespace {

constexpr int TAPS = 6, OFFS = 2;
__constant__ float c_anal[2][TAPS] = {
    {0.f, 0.f, 0.34150635f, 0.59150635f, 0.15849365f, -0.091506351f},           // cplx_wavelet_filter_coeffs.h Daub4_anal
    {-0.091506351f, -0.15849365f, 0.59150635f, -0.34150635f, 0.f, 0.f}};

struct LvArgs {
    const float* src; size_t sp;      // input plane (pitch in floats)
    float *lo, *b1, *b2, *b3;         // outputs, dense w2 x h2
    int w, h, w2, h2, skip;
};

// decimated analysis: AnalysisFilterSubsampVertical (L331-398) + AnalysisFilterSubsampHorizontal (L301-329)
__global__ void __launch_bounds__(256) k_wav_an_sub(LvArgs a)
{
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (x2 >= a.w2) return;
    for (int y2 = blockIdx.y; y2 < a.h2; y2 += gridDim.y) {
        const int row = 2 * y2, i = 2 * x2;
        const float* rp[TAPS];
        #pragma unroll
        for (int j = 0; j < TAPS; ++j) rp[j] = a.src + (size_t)max(0, min(row + a.skip * (OFFS - j), a.h - 1)) * a.sp;
        float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
        #pragma unroll
        for (int jh = 0; jh < TAPS; ++jh) {
            const int c = max(0, min(i + a.skip * (OFFS - jh), a.w - 1));
            float tlo = 0.f, thi = 0.f;                 // the line-buffer samples tmpLo[c], tmpHi[c]
            #pragma unroll
            for (int jv = 0; jv < TAPS; ++jv) {
                const float s = rp[jv][c];
                tlo += c_anal[0][jv] * s;
                thi += c_anal[1][jv] * s;
            }
            ll += c_anal[0][jh] * tlo;
            lh += c_anal[1][jh] * tlo;
            hl += c_anal[0][jh] * thi;
            hh += c_anal[1][jh] * thi;
        }
        const size_t o = (size_t)y2 * a.w2 + x2;
        a.lo[o] = ll; a.b1[o] = lh; a.b2[o] = hl; a.b3[o] = hh;
    }
}

// undecimated Haar analysis: AnalysisFilterHaarVertical (L223-239) + AnalysisFilterHaarHorizontal (L205-221)
__global__ void __launch_bounds__(256) k_wav_an_haar(LvArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.w) return;
    const int xn = x < a.w - a.skip ? x + a.skip : x - a.skip;
    for (int y = blockIdx.y; y < a.h; y += gridDim.y) {
        const int yn = y < a.h - a.skip ? y + a.skip : y - a.skip;
        const float* r0 = a.src + (size_t)y * a.sp;
        const float* r1 = a.src + (size_t)yn * a.sp;
        const float tlo0 = 0.25f * (r0[x] + r1[x]), thi0 = 0.25f * (r0[x] - r1[x]);
        const float tlo1 = 0.25f * (r0[xn] + r1[xn]), thi1 = 0.25f * (r0[xn] - r1[xn]);
        const size_t o = (size_t)y * a.w + x;
        a.lo[o] = tlo0 + tlo1; a.b1[o] = tlo0 - tlo1;
        a.b2[o] = thi0 + thi1; a.b3[o] = thi0 - thi1;
    }
}

struct SyArgs {
    const float *lo, *b1, *b2, *b3;   // dense sw x sh
    float* dst; size_t dp;            // output (pitch in floats)
    int sw, sh, dw, dh, skip;
    float blend;
};

// Haar synthesis: SynthesisFilterHaarHorizontal (L244-264) on (b2,b3) and (lo,b1), SynthesisFilterHaarVertical (L266-298)
__device__ __forceinline__ float haar_h(const float* lo, const float* hi, size_t r, int c, int skip)
{
    return c < skip ? lo[r + c] + hi[r + c] : 0.5f * (lo[r + c] + hi[r + c] + lo[r + c - skip] - hi[r + c - skip]);
}
__global__ void __launch_bounds__(256) k_wav_sy_haar(SyArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.sw) return;
    for (int y = blockIdx.y; y < a.sh; y += gridDim.y) {
        const size_t r = (size_t)y * a.sw;
        const float tl = haar_h(a.lo, a.b1, r, x, a.skip), th = haar_h(a.b2, a.b3, r, x, a.skip);
        float v;
        if (y < a.skip) v = tl + th;
        else {
            const size_t rm = (size_t)(y - a.skip) * a.sw;
            const float tlm = haar_h(a.lo, a.b1, rm, x, a.skip), thm = haar_h(a.b2, a.b3, rm, x, a.skip);
            v = 0.5f * (tl + th + tlm - thm);
        }
        a.dst[(size_t)y * a.dp + x] = v;
    }
}

// decimated synthesis: SynthesisFilterSubsampHorizontal (L447-515) + SynthesisFilterSubsampVertical (L518-597);
// synthesis taps are the analysis taps reversed (cplx_wavelet_dec.h L112-113)
__device__ __forceinline__ float sub_h(const float* lo, const float* hi, size_t r, int i, int sw, int skip)
{
    const int shift = skip * (TAPS - OFFS - 1);
    const int i_src = (i + shift) / 2, begin = (i + shift) % 2;
    float tot = 0.f;
    for (int j = begin, l = 0; j < TAPS; j += 2, l += skip) {
        const int arg = max(0, min(i_src - l, sw - 1));
        tot += ((c_anal[0][TAPS - 1 - j] * lo[r + arg] + c_anal[1][TAPS - 1 - j] * hi[r + arg]));
    }
    return tot;
}
__global__ void __launch_bounds__(256) k_wav_sy_sub(SyArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.dw) return;
    const int shift = a.skip * (TAPS - OFFS - 1);
    const float srcFactor = 1.f - a.blend;
    for (int y = blockIdx.y; y < a.dh; y += gridDim.y) {
        const int i_src = (y + shift) / 2, begin = (y + shift) % 2;
        float tot = 0.f;
        for (int j = begin, l = 0; j < TAPS; j += 2, l += a.skip) {
            const size_t r = (size_t)max(0, min(i_src - l, a.sh - 1)) * a.sw;
            const float tl = sub_h(a.lo, a.b1, r, x, a.sw, a.skip), th = sub_h(a.b2, a.b3, r, x, a.sw, a.skip);
            tot += ((c_anal[0][TAPS - 1 - j] * tl + c_anal[1][TAPS - 1 - j] * th));
        }
        float* d = a.dst + (size_t)y * a.dp + x;
        *d = *d * srcFactor + a.blend * 4.f * tot;      // L546: blends into the destination
    }
}


// The same synthesis for tap spacing 1 (level 0, the only decimated level FTblockDN uses), one thread per 2x2 output quad:
// the two output columns (rows) of a quad draw on source columns (rows) m-1 .. m+2, so the quad shares its 4 x 4 x 4 loads and
// its horizontal stage; every output is still the reference's expression, term for term.
__global__ void __launch_bounds__(256) k_wav_sy_sub_quad(SyArgs a)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * m >= a.dw) return;
    const float srcFactor = 1.f - a.blend;
    int col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) col[k] = max(0, min(m - 1 + k, a.sw - 1));
    const bool has_x1 = 2 * m + 1 < a.dw;
    for (int n = blockIdx.y; 2 * n < a.dh; n += gridDim.y) {
        float tl[4][2], th[4][2];            // horizontal stage of source rows n-1 .. n+2 for the even / odd output column
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t r = (size_t)max(0, min(n - 1 + k, a.sh - 1)) * a.sw;
            float lo[4], b1[4], b2[4], b3[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { lo[q] = a.lo[r + col[q]]; b1[q] = a.b1[r + col[q]]; b2[q] = a.b2[r + col[q]]; b3[q] = a.b3[r + col[q]]; }
            // x = 2m: i_src = m + 1, taps j = 1, 3, 5 on columns m+1, m, m-1;  x = 2m + 1: i_src = m + 2, taps j = 0, 2, 4 on m+2, m+1, m
            float t = 0.f;
            t += ((c_anal[0][4] * lo[2] + c_anal[1][4] * b1[2])); t += ((c_anal[0][2] * lo[1] + c_anal[1][2] * b1[1])); t += ((c_anal[0][0] * lo[0] + c_anal[1][0] * b1[0]));
            tl[k][0] = t;
            t = 0.f;
            t += ((c_anal[0][5] * lo[3] + c_anal[1][5] * b1[3])); t += ((c_anal[0][3] * lo[2] + c_anal[1][3] * b1[2])); t += ((c_anal[0][1] * lo[1] + c_anal[1][1] * b1[1]));
            tl[k][1] = t;
            t = 0.f;
            t += ((c_anal[0][4] * b2[2] + c_anal[1][4] * b3[2])); t += ((c_anal[0][2] * b2[1] + c_anal[1][2] * b3[1])); t += ((c_anal[0][0] * b2[0] + c_anal[1][0] * b3[0]));
            th[k][0] = t;
            t = 0.f;
            t += ((c_anal[0][5] * b2[3] + c_anal[1][5] * b3[3])); t += ((c_anal[0][3] * b2[2] + c_anal[1][3] * b3[2])); t += ((c_anal[0][1] * b2[1] + c_anal[1][1] * b3[1]));
            th[k][1] = t;
        }
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const int y = 2 * n + py;
            if (y >= a.dh) break;
            float* d = a.dst + (size_t)y * a.dp + 2 * m;
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                if (px && !has_x1) break;
                float tot = 0.f;
                if (py == 0) {          // y = 2n: i_src = n + 1, taps j = 1, 3, 5 on rows n+1, n, n-1
                    tot += ((c_anal[0][4] * tl[2][px] + c_anal[1][4] * th[2][px]));
                    tot += ((c_anal[0][2] * tl[1][px] + c_anal[1][2] * th[1][px]));
                    tot += ((c_anal[0][0] * tl[0][px] + c_anal[1][0] * th[0][px]));
                } else {                // y = 2n + 1: i_src = n + 2, taps j = 0, 2, 4 on rows n+2, n+1, n
                    tot += ((c_anal[0][5] * tl[3][px] + c_anal[1][5] * th[3][px]));
                    tot += ((c_anal[0][3] * tl[2][px] + c_anal[1][3] * th[2][px]));
                    tot += ((c_anal[0][1] * tl[1][px] + c_anal[1][1] * th[1][px]));
                }
                d[px] = d[px] * srcFactor + a.blend * 4.f * tot;
            }
        }
    }
}

}  // namespace

struct WLevel { int w, h, w2, h2, skip, sub; float* band[4]; };
struct art_hp_wavelet {
    art_hp_ctx* ctx;
    int nlev, W, H, subsamp;
    WLevel lev[10];
    float* block;         // one device allocation: all subbands + two lowpass ping-pong buffers
    float* buf[2];
    float* coeff0;        // lowpass of the last level (one of buf[])
    int consumed;
};

extern "C" {

int art_hp_wavelet_decompose_dev(art_hp_ctx* ctx, const float* d_src, size_t pitch, int W, int H, int maxlvl, int subsampling,
                                 art_hp_wavelet** out)
{
    if (!ctx || !out) return ART_HP_ERR_INVALID;
    *out = nullptr;
    if (!d_src || W < 8 || H < 8 || pitch < (size_t)W) return ctx->fail(ART_HP_ERR_INVALID, "bad geometry %dx%d", W, H);
    if (maxlvl < 1 || maxlvl > 10) return ctx->fail(ART_HP_ERR_INVALID, "maxlvl %d out of [1,10]", maxlvl);
    if (!(subsampling & 1)) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "level 0 must be decimated (the reference's buffers assume it, cplx_wavelet_dec.h L159-175)");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    art_hp_wavelet* wv = new art_hp_wavelet();
    wv->ctx = ctx; wv->nlev = maxlvl; wv->W = W; wv->H = H; wv->subsamp = subsampling; wv->consumed = 0;
    size_t total = 0;
    int w = W, h = H;
    for (int l = 0; l < maxlvl; ++l) {
        WLevel& L = wv->lev[l];
        L.w = w; L.h = h; L.sub = (subsampling >> l) & 1;
        L.skip = 1;                                   // cplx_wavelet_level.h L82-104, skipcrop == 1
        for (int n = 0; n < l; ++n) L.skip *= 2 - ((subsampling >> n) & 1);
        L.w2 = L.sub ? (w + 1) / 2 : w;
        L.h2 = L.sub ? (h + 1) / 2 : h;
        if (!L.sub && (w < 2 * L.skip || h < 2 * L.skip)) {
            delete wv;
            return ctx->fail(ART_HP_ERR_INVALID, "level %d: %dx%d is smaller than twice the tap spacing %d", l, w, h, L.skip);
        }
        total += 3 * round_up((size_t)L.w2 * L.h2, 64);
        w = L.w2; h = L.h2;
    }
    const size_t nb = round_up((size_t)wv->lev[0].w2 * wv->lev[0].h2, 64);
    total += 2 * nb;
    { void* blk = nullptr; const int prc = art_pool_alloc(ctx, total * sizeof(float), &blk); if (prc) { delete wv; return prc; } wv->block = (float*)blk; }
    float* p = wv->block;
    for (int l = 0; l < maxlvl; ++l) {
        WLevel& L = wv->lev[l];
        const size_t n = round_up((size_t)L.w2 * L.h2, 64);
        L.band[0] = nullptr;
        for (int j = 1; j < 4; ++j) { L.band[j] = p; p += n; }
    }
    wv->buf[0] = p; wv->buf[1] = p + nb;
    cudaStream_t st = ctx->stream;
    int bi = 0;
    for (int l = 0; l < maxlvl; ++l) {
        WLevel& L = wv->lev[l];
        LvArgs a;
        if (l == 0) { a.src = d_src; a.sp = pitch; }
        else { bi ^= 1; a.src = wv->buf[bi]; a.sp = L.w; }
        a.lo = wv->buf[bi ^ 1]; a.b1 = L.band[1]; a.b2 = L.band[2]; a.b3 = L.band[3];
        a.w = L.w; a.h = L.h; a.w2 = L.w2; a.h2 = L.h2; a.skip = L.skip;
        const dim3 grid((L.w2 + 255) / 256, std::min(L.h2, 148 * 8));
        art_prof_begin(ctx, L.sub ? "k_wav_an_sub" : "k_wav_an_haar");
        if (L.sub) k_wav_an_sub<<<grid, 256, 0, st>>>(a); else k_wav_an_haar<<<grid, 256, 0, st>>>(a);
        art_prof_end(ctx);
        ctx->launches++;
    }
    wv->coeff0 = wv->buf[bi ^ 1];
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { art_pool_free(ctx, wv->block); delete wv; return ctx->fail(ART_HP_ERR_CUDA, "wavelet kernels: %s", cudaGetErrorString(e)); }
    *out = wv;
    return ART_HP_OK;
}

int art_hp_wavelet_maxlevel(const art_hp_wavelet* w) { return w ? w->nlev : 0; }

int art_hp_wavelet_level_dims(const art_hp_wavelet* w, int level, int* width, int* height, int* stride)
{
    if (!w || level < 0 || level >= w->nlev) return ART_HP_ERR_INVALID;
    if (width) *width = w->lev[level].w2;
    if (height) *height = w->lev[level].h2;
    if (stride) *stride = w->lev[level].skip;
    return ART_HP_OK;
}

float* art_hp_wavelet_band_dev(const art_hp_wavelet* w, int level, int dir)
{
    if (!w || dir < 0 || dir > 3) return nullptr;
    if (dir == 0) return w->coeff0;
    if (level < 0 || level >= w->nlev) return nullptr;
    return w->lev[level].band[dir];
}

int art_hp_wavelet_reconstruct_dev(art_hp_wavelet* w, float* d_dst, size_t pitch, float blend)
{
    if (!w || !d_dst) return ART_HP_ERR_INVALID;
    art_hp_ctx* ctx = w->ctx;
    if (w->consumed) return ctx->fail(ART_HP_ERR_INVALID, "the decomposition was already reconstructed (the reference consumes it too)");
    if (pitch < (size_t)w->W) return ctx->fail(ART_HP_ERR_INVALID, "pitch smaller than width");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    float* cur = w->coeff0;
    float* other = (cur == w->buf[0]) ? w->buf[1] : w->buf[0];
    for (int l = w->nlev - 1; l >= 0; --l) {
        WLevel& L = w->lev[l];
        SyArgs a;
        a.lo = cur; a.b1 = L.band[1]; a.b2 = L.band[2]; a.b3 = L.band[3];
        a.sw = L.w2; a.sh = L.h2; a.dw = L.w; a.dh = L.h; a.skip = L.skip;
        if (l == 0) { a.dst = d_dst; a.dp = pitch; a.blend = blend; }
        else { a.dst = other; a.dp = L.w; a.blend = 1.f; }
        const dim3 grid((L.w + 255) / 256, std::min(L.h, 148 * 8));
        art_prof_begin(ctx, L.sub ? "k_wav_sy_sub" : "k_wav_sy_haar");
        if (L.sub) {
            if (l != 0) {
                // a decimated level above 0 reconstructs over the lowpass buffer itself (dec.h L221: dst == coeff0):
                // start from a copy so that dst * (1 - blend) reads the same values
                ART_CUDA(ctx, cudaMemcpyAsync(other, cur, (size_t)L.w2 * L.h2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
            if (a.skip == 1) {
                const dim3 qgrid(((L.w + 1) / 2 + 255) / 256, std::min((L.h + 1) / 2, 148 * 8));
                k_wav_sy_sub_quad<<<qgrid, 256, 0, st>>>(a);
            } else k_wav_sy_sub<<<grid, 256, 0, st>>>(a);
        } else k_wav_sy_haar<<<grid, 256, 0, st>>>(a);
        art_prof_end(ctx);
        ctx->launches++;
        if (l != 0) { float* t = cur; cur = other; other = t; }
    }
    w->consumed = 1;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

static int band_geometry(const art_hp_wavelet* w, int level, int dir, float** p, size_t* n)
{
    if (!w || dir < 0 || dir > 3 || level < 0 || level >= w->nlev) return ART_HP_ERR_INVALID;
    const WLevel& L = w->lev[dir == 0 ? w->nlev - 1 : level];
    *p = dir == 0 ? w->coeff0 : w->lev[level].band[dir];
    *n = (size_t)L.w2 * L.h2;
    return ART_HP_OK;
}

int art_hp_wavelet_get_band(const art_hp_wavelet* w, int level, int dir, float* host)
{
    float* p; size_t n;
    if (!host || band_geometry(w, level, dir, &p, &n)) return ART_HP_ERR_INVALID;
    art_hp_ctx* ctx = w->ctx;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    ART_CUDA(ctx, cudaMemcpyAsync(host, p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

int art_hp_wavelet_set_band(art_hp_wavelet* w, int level, int dir, const float* host)
{
    float* p; size_t n;
    if (!host || band_geometry(w, level, dir, &p, &n)) return ART_HP_ERR_INVALID;
    art_hp_ctx* ctx = w->ctx;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    ART_CUDA(ctx, cudaMemcpyAsync(p, host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ART_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ART_HP_OK;
}

void art_hp_wavelet_destroy(art_hp_wavelet* w)
{
    if (!w) return;
    art_pool_free(w->ctx, w->block);      // stream-ordered reuse: no synchronisation, no cudaFree
    delete w;
}

}  // extern "C"
