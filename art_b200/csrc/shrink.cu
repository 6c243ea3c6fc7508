// FTblockDN wavelet shrinkage for sm_100a.
//
// Replaces (reference rtengine/FTblockDN.cc) MadRgb L569-603, ShrinkAllL L638-726, ShrinkAllAB L729-839 and the
// drivers WaveletDenoiseAllL L1111-1167 (edge == 0) / WaveletDenoiseAllAB L1170-1221 on device-resident
// wavelet decompositions (wavelet.cu).  Everything stays in HBM: the MAD values live in a small device table
// that the shrink kernels read, so a whole denoise pass needs no host synchronisation.
//   * MAD: int32 histogram of |trunc(coeff)| (exact by construction: integer counts), hot low bins privatised in
//     shared memory, then a one-block prefix scan finds the median bin and interpolates like the reference.
//   * shrink factor / apply: element-wise; coefficients the reference handles in its 4-wide SSE loops use the
//     vector xexpf (rtengine/sleefsseavx.h L1326-1345) and the vector expression association, the n % 4 tail
//     uses the scalar xexpf (rtengine/sleef.h L1247-1266) and the scalar association -- bit-exact.
//   * the local averaging is the flat boxblur (rtengine/boxblur.h L558-742) with its three column classes.
// Compiled with -fmad=false.
#include "ctx.h"
#include "sleef_dev.cuh"

struct WLevel { int w, h, w2, h2, skip, sub; float* band[4]; };
struct art_hp_wavelet {
    art_hp_ctx* ctx;
    int nlev, W, H, subsamp;
    WLevel lev[10];
    float* block;
    float* buf[2];
    float* coeff0;
    int consumed;
};

namespace {
using sleef::xexpf_scalar;
using sleef::xexpf_vector;

// ------------------------------------------------------------------ MAD (MadRgb, L569-603)
constexpr int NB = 65536, HOT = 4096;
__global__ void __launch_bounds__(512) k_mad_hist(const float* __restrict__ data, int n, int* __restrict__ histo)
{
    __shared__ int hot[HOT];
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) hot[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) {
        int v = abs(__float2int_rz(data[i]));      // abs(static_cast<int>(x))
        v = v < 65535 ? v : 65535;
        if (v < HOT) atomicAdd(&hot[v], 1); else atomicAdd(&histo[v], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) if (hot[i]) atomicAdd(&histo[i], hot[i]);
}

// one block: median bin by prefix scan, then the reference's interpolation; writes SQR(mad) * premul to out[0]
__global__ void __launch_bounds__(1024) k_mad_median(const int* __restrict__ histo, int n, float* out, int square)
{
    __shared__ int part[1024];
    const int t = threadIdx.x;
    int s = 0;
    for (int k = 0; k < NB / 1024; ++k) s += histo[t * (NB / 1024) + k];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        float r = 0.f;
        if (n > 1) {
            const int half = n / 2;
            int count = 0, seg = 0;
            while (seg < 1024 && count + part[seg] < half) { count += part[seg]; ++seg; }
            int median = seg * (NB / 1024);
            // while (count < datalen / 2) { count += histo[median]; ++median; }
            while (count < half) { count += histo[median]; ++median; }
            const int count_ = count - histo[median - 1];
            r = (float)((double)((median - 1) + (half - count_) / ((float)(count - count_))) / 0.6745);
        }
        out[0] = square ? r * r : r;
    }
}

// ------------------------------------------------------------------ shrink factors
struct ShArgs {
    float* c; const float* cL; const float* nv; float* sf; const float* sfd;
    const float* mad; int n; float lvlmul; float noisevar_ab; int useCCurve;
    const float* madab;
    int nv_uniform; float nv_value;      // the noise-variance map holds one value everywhere: it is passed instead of read
};

__global__ void __launch_bounds__(256) k_sf_L(ShArgs a)
{   // L669-684
    const float eps = 0.01f;
    const float levelFactor = a.mad[0] * 5.f / a.lvlmul;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float x = a.c[i];
        const float nvi = a.nv_uniform ? a.nv_value : a.nv[i];
        const float mag = x * x;
        float r;
        if ((i & ~3) < a.n - 3) {      // handled by a full 4-wide vector in the reference (for (i = 0; i < n - 3; i += 4))
            const float mad = nvi * levelFactor;
            r = mag / (mag + mad * xexpf_vector(-mag / (9.0f * mad)) + eps);
        } else {
            r = mag / (mag + levelFactor * nvi * xexpf_scalar(-mag / (9 * levelFactor * nvi)) + eps);
        }
        a.sf[i] = r;
    }
}

__global__ void __launch_bounds__(256) k_sf_AB(ShArgs a)
{   // L762-786
    const float mad_L = a.mad[0];
    float madab = a.madab[0];
    madab = a.useCCurve ? madab : madab * a.noisevar_ab;
    const float rmadLm9 = 1.f / (mad_L * 9.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float xl = a.cL[i], xab = a.c[i];
        const float nvi = a.nv_uniform ? a.nv_value : a.nv[i];
        float r;
        if ((i & ~3) < a.n - 3) {
            const float mad_ab = nvi * madab;
            const float mag_ab = xab * xab;
            const float mag_L = (xl * xl) * rmadLm9;
            r = 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
        } else {
            const float mag_L = xl * xl, mag_ab = xab * xab;
            r = (1.f - xexpf_scalar(-(mag_ab / (nvi * madab)) - (mag_L / (9.f * mad_L))));
        }
        a.sf[i] = r;
    }
}

// WaveletDenoiseAll_BiShrinkAB's "simple" shrinkage of the levels below the coarsest one (L1046-1087): in place, no local averaging,
// the factor squared, mad_abr = (useNoiseCCurve ? noisevar_ab : SQR(noisevar_ab)) * madab
__global__ void __launch_bounds__(256) k_sf_AB_simple(ShArgs a)
{
    const float mad_L = a.mad[0];
    const float madab = a.madab[0];
    const float mad_abr = a.useCCurve ? a.noisevar_ab * madab : (a.noisevar_ab * a.noisevar_ab) * madab;
    const float rmadLm9 = 1.f / (mad_L * 9.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float xl = a.cL[i], xab = a.c[i];
        const float nvi = a.nv_uniform ? a.nv_value : a.nv[i];
        if ((i & ~3) < a.n - 3) {
            const float mad_ab = nvi * mad_abr;
            const float mag_ab = xab * xab;
            const float mag_L = (xl * xl) * rmadLm9;
            const float f = 1.f - xexpf_vector(-(mag_ab / mad_ab) - (mag_L));
            a.c[i] = xab * (f * f);
        } else {
            const float mag_L = xl * xl, mag_ab = xab * xab;
            const float f = 1.f - xexpf_scalar(-(mag_ab / (nvi * mad_abr)) - (mag_L / (9.f * mad_L)));
            a.c[i] = xab * (f * f);
        }
    }
}

__global__ void __launch_bounds__(256) k_sf_apply(ShArgs a)
{   // L692-709 / L791-813
    const float eps = 0.01f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float s = a.sf[i], d = a.sfd[i], x = a.c[i];
        a.c[i] = ((i & ~3) < a.n - 3) ? x * (d * d + s * s) / (d + s + eps) : x * ((d * d + s * s) / (d + s + eps));
    }
}

// ------------------------------------------------------------------ flat boxblur (boxblur.h L558-742), radx == rady >= 1
struct FbArgs { const float* x; float* y; int W, H, rad; };

// Both passes are running sums -- t += (x[n + rad] - x[n - rad - 1]) * (1 / len) -- whose fp32 association is part of the
// result, so a chain (a row for the horizontal pass, a column for the vertical one) is inherently serial.  The work is
// split so that only the additions are serial: a CTA owns FB_CH chains and walks them in tiles of FB_T steps; all
// 256 threads load the tile with coalesced reads and form the per-step differences in parallel, FB_CH lanes run the
// additions out of shared memory, all threads store the tile.  The ramps at both ends of a chain (rad + 1 and rad steps)
// are done by the chain lanes straight from global memory.
// horizontal pass: 8 rows x 256 steps per tile (342 CTAs per 45 MP subband); vertical pass: 32 columns x 64 steps (128-byte segments)
constexpr int FB_NT = 256, FB_CH_H = 8, FB_T_H = 256, FB_CH_V = 32, FB_T_V = 64;

template <bool VERT>
__global__ void __launch_bounds__(FB_NT) k_fbox(FbArgs a)
{
    constexpr int FB_CH = VERT ? FB_CH_V : FB_CH_H, FB_T = VERT ? FB_T_V : FB_T_H, FB_PER = FB_CH * FB_T / FB_NT;   // horizontal: boxblur.h L571-602; vertical: L614-710 (columns < W - W % 4 follow the 4-/8-wide code, the rest the scalar tail)
    __shared__ float tile[FB_T][FB_CH + 1];
    const int W = a.W, H = a.H, rad = a.rad;
    const int nchains = VERT ? W : H, nsteps = VERT ? H : W;
    const int chain0 = blockIdx.x * FB_CH;
    const int tid = threadIdx.x;
    auto at = [&](int ch, int st) -> size_t { return VERT ? (size_t)st * W + ch : (size_t)ch * W + st; };
    const bool chain_thread = tid < FB_CH && chain0 + tid < nchains;
    const int mych = chain0 + tid;
    const bool scalar_class = VERT && mych >= W - (W % 4);
    float t = 0.f;
    const float full = (float)(2 * rad + 1);
    if (chain_thread) {
        if (!VERT) {
            int len = rad + 1;
            t = a.x[at(mych, 0)];
            for (int j = 1; j <= rad; j++) t += a.x[at(mych, j)];
            t = t / len;
            a.y[at(mych, 0)] = t;
            for (int st = 1; st <= rad; st++) {
                t = (t * len + a.x[at(mych, st + rad)]) / (len + 1);
                a.y[at(mych, st)] = t;
                len++;
            }
        } else if (!scalar_class) {
            float len = (float)(rad + 1);
            t = a.x[at(mych, 0)];
            for (int i = 1; i <= rad; i++) t = t + a.x[at(mych, i)];
            t = t / len;
            a.y[at(mych, 0)] = t;
            for (int st = 1; st <= rad; st++) {
                const float lp1 = len + 1.f;
                t = (t * len + a.x[at(mych, st + rad)]) / lp1;
                a.y[at(mych, st)] = t;
                len = lp1;
            }
        } else {
            int len = rad + 1;
            t = a.x[at(mych, 0)] / len;
            for (int i = 1; i <= rad; i++) t += a.x[at(mych, i)] / len;
            a.y[at(mych, 0)] = t;
            for (int st = 1; st <= rad; st++) {
                t = (t * len + a.x[at(mych, st + rad)]) / (len + 1);
                a.y[at(mych, st)] = t;
                len++;
            }
        }
    }
    const float rlen = 1.f / full;
    const int first = rad + 1, last = nsteps - rad;          // main region [first, last)
    for (int s0 = first; s0 < last; s0 += FB_T) {
        const int nst = min(FB_T, last - s0);
        {   // all loads of the tile first (FB_PER independent pairs per thread in flight), then the differences
            float va[FB_PER], vb[FB_PER];
#pragma unroll
            for (int u = 0; u < FB_PER; ++u) {
                const int i = tid + u * FB_NT;
                const int cl = VERT ? i % FB_CH : i / FB_T, sl = VERT ? i / FB_CH : i % FB_T;
                const int ch = chain0 + cl, st = s0 + sl;
                const bool ok = ch < nchains && sl < nst;
                va[u] = ok ? a.x[at(ch, st + rad)] : 0.f;
                vb[u] = ok ? a.x[at(ch, st - rad - 1)] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < FB_PER; ++u) {
                const int i = tid + u * FB_NT;
                const int cl = VERT ? i % FB_CH : i / FB_T, sl = VERT ? i / FB_CH : i % FB_T;
                const float diff = va[u] - vb[u];
                tile[sl][cl] = (VERT && chain0 + cl >= W - (W % 4)) ? diff / (float)(2 * rad + 1) : diff * rlen;
            }
        }
        __syncthreads();
        if (chain_thread) {
#pragma unroll 8
            for (int sl = 0; sl < nst; ++sl) {
                t = t + tile[sl][tid];
                tile[sl][tid] = t;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < FB_PER; ++u) {
            const int i = tid + u * FB_NT;
            const int cl = VERT ? i % FB_CH : i / FB_T, sl = VERT ? i / FB_CH : i % FB_T;
            const int ch = chain0 + cl;
            if (ch < nchains && sl < nst) a.y[at(ch, s0 + sl)] = tile[sl][cl];
        }
        __syncthreads();
    }
    if (chain_thread) {
        if (VERT && !scalar_class) {
            float len = full;
            for (int st = max(last, first); st < nsteps; st++) {
                const float lm1 = len - 1.f;
                t = (t * len - a.x[at(mych, st - rad - 1)]) / lm1;
                a.y[at(mych, st)] = t;
                len = lm1;
            }
        } else {
            int len = 2 * rad + 1;
            for (int st = max(last, first); st < nsteps; st++) {
                t = (t * len - a.x[at(mych, st - rad - 1)]) / (len - 1);
                a.y[at(mych, st)] = t;
                len--;
            }
        }
    }
}

int mad_of(art_hp_ctx* ctx, const float* band, int n, int* d_histo, float* d_out, int square)
{
    cudaStream_t st = ctx->stream;
    ART_CUDA(ctx, cudaMemsetAsync(d_histo, 0, NB * sizeof(int), st));
    art_prof_begin(ctx, "k_mad_hist");
    k_mad_hist<<<148 * 4, 512, 0, st>>>(band, n, d_histo);
    art_prof_end(ctx);
    k_mad_median<<<1, 1024, 0, st>>>(d_histo, n, d_out, square);
    ctx->launches += 2;
    return ART_HP_OK;
}

int blur_radius(int level, double scale) { const int r = (int)((level + 2) / scale); return r > 1 ? r : 1; }

// scratch, one set per lane: [histogram 65536 ints][madab 64 floats][sf n][sfd n][tmp n]
struct Scratch { int* histo; float* madab; float *sf, *sfd, *tmp; };
constexpr int NL = art_hp_ctx::NLANES;
int scratch_for(art_hp_ctx* ctx, size_t n, Scratch s[NL])
{
    const size_t np = round_up(n, 64);
    const size_t one = round_up(NB * sizeof(int) + 256 + 3 * np * sizeof(float), 256);
    int rc = art_reserve(ctx, ctx->d_scratch, NL * one);
    if (rc) return rc;
    for (int i = 0; i < NL; ++i) {
        char* p = (char*)ctx->d_scratch.p + i * one;
        s[i].histo = (int*)p; p += NB * sizeof(int);
        s[i].madab = (float*)p; p += 256;
        s[i].sf = (float*)p; s[i].sfd = s[i].sf + np; s[i].tmp = s[i].sfd + np;
    }
    return ART_HP_OK;
}

// Fork the context's stream into NL lanes (one per subband of a channel) and join them again.  Between the two calls
// `ctx->stream` is pointed at a lane with lane_of(): every helper queues on ctx->stream, so nothing else changes.
struct Lanes {
    art_hp_ctx* ctx; cudaStream_t main; bool serial;
    int begin(art_hp_ctx* c)
    {
        ctx = c; main = c->stream;
        serial = c->profiling;       // per-kernel timing (art_hp_profile_*) wants every launch alone on the stream its events are on
        if (serial) return ART_HP_OK;
        if (!c->lane[0]) {
            for (int i = 0; i < NL; ++i) {
                ART_CUDA(c, cudaStreamCreateWithFlags(&c->lane[i], cudaStreamNonBlocking));
                ART_CUDA(c, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
            }
            ART_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        }
        ART_CUDA(c, cudaEventRecord(c->ev_fork, main));
        for (int i = 0; i < NL; ++i) ART_CUDA(c, cudaStreamWaitEvent(c->lane[i], c->ev_fork, 0));
        return ART_HP_OK;
    }
    void lane_of(int i) { if (!serial) ctx->stream = ctx->lane[i]; }
    int end()
    {
        ctx->stream = main;
        if (serial) return ART_HP_OK;
        for (int i = 0; i < NL; ++i) {
            ART_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], ctx->lane[i]));
            ART_CUDA(ctx, cudaStreamWaitEvent(main, ctx->ev_join[i], 0));
        }
        return ART_HP_OK;
    }
};

int shrink_band(art_hp_ctx* ctx, const Scratch& s, ShArgs a, int W, int H, int rad, bool ab)
{
    cudaStream_t st = ctx->stream;
    const int grid = std::min((a.n + 255) / 256, 148 * 16);
    a.sf = s.sf; a.sfd = s.sfd;
    art_prof_begin(ctx, ab ? "k_sf_AB" : "k_sf_L");
    if (ab) k_sf_AB<<<grid, 256, 0, st>>>(a); else k_sf_L<<<grid, 256, 0, st>>>(a);
    art_prof_end(ctx);
    if (2 * rad + 1 > W || 2 * rad + 1 > H) return ctx->fail(ART_HP_ERR_INVALID, "blur radius %d does not fit a %dx%d subband", rad, W, H);
    art_prof_begin(ctx, "k_fbox_h");
    k_fbox<false><<<(H + FB_CH_H - 1) / FB_CH_H, FB_NT, 0, st>>>(FbArgs{s.sf, s.tmp, W, H, rad});
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_fbox_v");
    k_fbox<true><<<(W + FB_CH_V - 1) / FB_CH_V, FB_NT, 0, st>>>(FbArgs{s.tmp, s.sfd, W, H, rad});
    art_prof_end(ctx);
    art_prof_begin(ctx, "k_sf_apply");
    k_sf_apply<<<grid, 256, 0, st>>>(a);
    art_prof_end(ctx);
    ctx->launches += 4;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

}  // namespace

extern "C" {

// madL[lvl][dir-1] = SQR(MadRgb(level_coeffs(lvl)[dir])) for every level (FTblockDN.cc L2311-2320) into a device table
int art_hp_wavelet_mad_dev(art_hp_ctx* ctx, const art_hp_wavelet* w, float* d_madL)
{
    if (!ctx || !w || !d_madL) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch s[NL];
    int rc = scratch_for(ctx, 64, s);
    if (rc) return rc;
    Lanes lanes;
    if ((rc = lanes.begin(ctx))) return rc;
    for (int l = 0; l < w->nlev && !rc; ++l)
        for (int d = 1; d < 4 && !rc; ++d) {
            const int ln = (3 * l + d - 1) % NL;
            lanes.lane_of(ln);
            rc = mad_of(ctx, w->lev[l].band[d], w->lev[l].w2 * w->lev[l].h2, s[ln].histo, d_madL + 3 * l + (d - 1), 1);
        }
    const int rc2 = lanes.end();
    if (rc || rc2) return rc ? rc : rc2;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

}  // extern "C"

// internal forms: `uniform` != nullptr says the noise-variance map is that one value everywhere (the map pointer is then unused)
int art_wavelet_denoise_L(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* uniform, const float* d_madL, double scale);
int art_wavelet_denoise_AB(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom, const float* uniform,
                           const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale, int bishrink = 0);

extern "C" {

int art_hp_wavelet_denoise_L_dev(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* d_madL, double scale)
{
    if (!d_noisevarlum) return ART_HP_ERR_INVALID;
    return art_wavelet_denoise_L(ctx, wL, d_noisevarlum, nullptr, d_madL, scale);
}

}  // extern "C"

int art_wavelet_denoise_L(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* uniform, const float* d_madL, double scale)
{
    if (!ctx || !wL || (!d_noisevarlum && !uniform) || !d_madL || !(scale > 0)) return ART_HP_ERR_INVALID;
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    const int maxlvl = std::min(wL->nlev, 5);      // L1115
    Scratch s[NL];
    int rc = scratch_for(ctx, (size_t)wL->lev[0].w2 * wL->lev[0].h2, s);
    if (rc) return rc;
    Lanes lanes;
    if ((rc = lanes.begin(ctx))) return rc;
    for (int l = 0; l < maxlvl && !rc; ++l)
        for (int d = 1; d < 4 && !rc; ++d) {
            const WLevel& L = wL->lev[l];
            ShArgs a{};
            a.c = L.band[d]; a.nv = d_noisevarlum; a.mad = d_madL + 3 * l + (d - 1); a.n = L.w2 * L.h2; a.lvlmul = (float)(l + 1);
            a.nv_uniform = uniform != nullptr; a.nv_value = uniform ? *uniform : 0.f;
            const int ln = (3 * l + d - 1) % NL;
            lanes.lane_of(ln);
            rc = shrink_band(ctx, s[ln], a, L.w2, L.h2, blur_radius(l, scale), false);
        }
    const int rc2 = lanes.end();
    return rc ? rc : rc2;
}

extern "C" int art_hp_wavelet_denoise_AB_dev(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom,
                                             const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale)
{
    if (!d_noisevarchrom) return ART_HP_ERR_INVALID;
    return art_wavelet_denoise_AB(ctx, wL, wab, d_noisevarchrom, nullptr, d_madL, noisevar_ab, useNoiseCCurve, autoch, scale, 0);
}

// bishrink != 0: WaveletDenoiseAll_BiShrinkAB (L976-1108) instead of WaveletDenoiseAllAB: the coarsest level goes through ShrinkAllAB
// as usual (its MAD is the same whether computed up front or now), every finer level through the simple shrinkage
int art_wavelet_denoise_AB(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom, const float* uniform,
                           const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale, int bishrink)
{
    if (!ctx || !wL || !wab || (!d_noisevarchrom && !uniform) || !d_madL || !(scale > 0)) return ART_HP_ERR_INVALID;
    if (wL->nlev != wab->nlev || wL->W != wab->W || wL->H != wab->H) return ctx->fail(ART_HP_ERR_INVALID, "L and ab decompositions differ in shape");
    ART_CUDA(ctx, cudaSetDevice(ctx->device));
    if (autoch && noisevar_ab <= 0.001f) noisevar_ab = 0.02f;     // L737-739
    Scratch s[NL];
    int rc = scratch_for(ctx, (size_t)wab->lev[0].w2 * wab->lev[0].h2, s);
    if (rc) return rc;
    Lanes lanes;
    if ((rc = lanes.begin(ctx))) return rc;
    for (int l = 0; l < wL->nlev && !rc; ++l)
        for (int d = 1; d < 4 && !rc; ++d) {
            const WLevel& L = wab->lev[l];
            const int n = L.w2 * L.h2;
            if (!(noisevar_ab > 0.001f)) continue;                   // L761 (MadRgb is computed but unused)
            const int ln = (3 * l + d - 1) % NL;
            lanes.lane_of(ln);
            const Scratch& sc = s[ln];
            if ((rc = mad_of(ctx, L.band[d], n, sc.histo, sc.madab, 1))) break;
            ShArgs a{};
            a.c = L.band[d]; a.cL = wL->lev[l].band[d]; a.nv = d_noisevarchrom; a.mad = d_madL + 3 * l + (d - 1); a.n = n;
            a.noisevar_ab = noisevar_ab; a.useCCurve = useNoiseCCurve; a.madab = sc.madab;
            a.nv_uniform = uniform != nullptr; a.nv_value = uniform ? *uniform : 0.f;
            if (bishrink && l != wL->nlev - 1) {
                art_prof_begin(ctx, "k_sf_AB_simple");
                k_sf_AB_simple<<<std::min((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(a);
                art_prof_end(ctx);
                ctx->launches++;
                continue;
            }
            rc = shrink_band(ctx, sc, a, L.w2, L.h2, blur_radius(l, scale), true);
        }
    const int rc2 = lanes.end();
    return rc ? rc : rc2;
}
