// denoise::RGB_denoise for sm_100a: the path ART's driver takes (ipdenoise.cc L1165: kall = 0, isRAW = true).
//
// Replaces (reference) rtengine/FTblockDN.cc RGB_denoise L1638-2689 (colorSpace RGB and LAB, standard and aggressive quality, chrominance
// method MANUAL, Tile_calc L442-478 = one tile), Noise_residualAB L607-635, detail_recovery L1479-1635 with
// RGBtile_denoise L494-525 / RGBoutput_tile_row L531-558 / boxabsblur (boxblur.h L745-888), Color::gammaf2lut
// (color.cc L1128-1170), gammaf / rgb2yuv / yuv2rgb (color.h L782-796, L1202-1205), rgbxyz + XYZ2Lab (color.cc L833,
// L1247-1274, L1382-1399) for the chroma noise curve.  Everything stays in HBM between stages:
//   k_dn_gamma_lut x2, [k_dn_ccalc], k_dn_split (gamma LUT + YUV + noise-variance maps), wavelet.cu / shrink.cu for
//   the three decompositions, k_dn_blocks5 (dn_blocks.cu: pairs of 64x64 DCT blocks on tcgen05 -- gather + window, DCT-II,
//   |.| box blur, shrink, DCT-III -> block store), k_dn_gather (ordered overlap-add + normalise), k_dn_merge (chroma boost, YUV->RGB,
//   inverse gamma).
// Bit-exact with the reference except the two block DCTs, which the reference delegates to FFTW (fp32 codelets,
// absent here) and which run here as 3xTF32 tcgen05 matrix products (fp32 accuracy) against cosine tables rounded from double.
// The reference's detail_recovery overlap-adds block rows from several OpenMP threads without synchronisation; the
// one-thread order (vblk, then hblk ascending) is the one reproduced.
// Compiled with -fmad=false.
#include <cmath>
#include "ctx.h"
#include "dn_blocks.h"
#include "sleef_dev.cuh"

struct art_hp_wavelet;

namespace {

constexpr int TS = 64, OFFSET = 25, BLKRAD = 1;

__device__ __forceinline__ float lut_clip_below(const float* __restrict__ data, int size, float index)
{   // LUTf(size, LUT_CLIP_BELOW)::operator[](float), LUT.h L437-459
    int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) idx = size - 2;
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
__device__ __forceinline__ float lut_clip_both(const float* __restrict__ data, int size, float index)
{
    const int idx = (int)index;
    if (index < 0.f || !(index == index)) return data[0];
    else if (index > (float)(size - 2)) return data[size - 1];
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}

// Color::gammaf2lut, SSE2 build (color.cc L1128-1163)
__global__ void __launch_bounds__(256) k_dn_gamma_lut(float* lut, float gamma, float start, float slope, float divisor, float factor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536) return;
    const float gammav = 1.f / gamma;
    const float slopev = (slope / divisor) * factor;
    const float divisorv = sleef::xlogf_scalar(divisor);
    const float comparev = start * divisor;
    const int border = (int)(start * divisor);
    const int border1 = border - (border & 3);
    const int border2 = border1 + 4;
    const float iv = (float)i;
    float r;
    if (i < border1) r = iv * slopev;
    else if (i < border2) r = iv <= comparev ? iv * slopev : sleef::xexpf_vector((sleef::xlogf_vector(iv) - divisorv) * gammav) * factor;
    else r = sleef::xexpf_nocheck((sleef::xlogf_nocheck(iv) - divisorv) * gammav) * factor;
    lut[i] = r;
}

struct Gam { const float* lut; float gam, thresh, slope, top; };     // top: 65535 for the forward curve, 65536 for the inverse (L1813, L1822)
__device__ __forceinline__ float gammaf_(float x, float gamma, float start, float slope)
{
    return x <= start ? x * slope : sleef::xexpf_scalar(sleef::xlogf_scalar(x) / gamma);
}
__device__ __forceinline__ float apply_gam(const Gam& g, float outer_gam, float v)
{   // apply_gamma / apply_igamma, L1809-1826
    if (outer_gam > 1.f && v > 0.f) v = v < g.top ? lut_clip_below(g.lut, 65536, v) : (gammaf_(v / 65535.f, g.gam, g.thresh, g.slope) * 65535.f);
    return v;
}

// chroma noise-curve map on the half-resolution calclum image, L1722-1764
struct CcArgs { const float *r, *g, *b; size_t cp; int w2, h2; float wp[9]; const float* cachef; const float* curve; float* ccalc; };
__device__ __forceinline__ float computeXYZ2Lab(const float* cachef, float f)
{   // color.cc L1247-1259
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * (((24389.0 / 27.0) * f / (double)65535.f + 16.0) / 116.0));
    else if (f > 65535.f) return 327.68f * sleef::xcbrtf_scalar(f / 65535.f);
    return lut_clip_below(cachef, 65536, f);
}
__device__ __forceinline__ float computeXYZ2LabY(const float* cachefy, float f)
{   // color.cc L1262-1274
    if (f != f) return f;
    if (f < 0.f) return (float)(327.68 * ((24389.0 / 27.0) * f / (double)65535.f));
    else if (f > 65535.f) return 327.68f * (116.f * sleef::xcbrtf_scalar(f / 65535.f) - 16.f);
    return lut_clip_below(cachefy, 65536, f);
}
__device__ __forceinline__ float lut_noclip(const float* __restrict__ data, int size, float index)
{   // LUTf(size, 0)::operator[](float), LUT.h L437-459: extrapolates at both ends
    int idx = (int)index;
    if (index < 0.f || !(index == index)) idx = 0;
    else if (index > (float)(size - 2)) idx = size - 2;
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
// DenoiseParams::ColorSpace::LAB: Color::denoiseIGammaTab / denoiseGammaTab, cachef / cachefy, working-space matrix and its inverse
struct LabTabs { int on; const float *dn_igamma, *dn_gamma, *cachef, *cachefy; float wp[9], iwp[9]; };
__device__ __forceinline__ float f2xyz(float f)
{   // color.h L767-770
    const float epsilonExpInv3f = (float)(6.0 / 29.0), kappaInvf = (float)(27.0 / 24389.0);
    return (f > epsilonExpInv3f) ? f * f * f : (116.f * f - 16.f) * kappaInvf;
}

__global__ void __launch_bounds__(256) k_dn_ccalc(CcArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.w2) return;
    const float RL = a.r[(size_t)y * a.cp + x], GL = a.g[(size_t)y * a.cp + x], BL = a.b[(size_t)y * a.cp + x];
    const float XL = ((a.wp[0] * RL + a.wp[1] * GL + a.wp[2] * BL));
    const float YL = ((a.wp[3] * RL + a.wp[4] * GL + a.wp[5] * BL));
    const float ZL = ((a.wp[6] * RL + a.wp[7] * GL + a.wp[8] * BL));
    const float fx = computeXYZ2Lab(a.cachef, XL / 0.9642f), fy = computeXYZ2Lab(a.cachef, YL), fz = computeXYZ2Lab(a.cachef, ZL / 0.8249f);
    const float AA = (500.0f * (fx - fy)), BB = (200.0f * (fy - fz));
    const float cN = sqrtf(AA * AA + BB * BB);
    const float cn100 = 1.f + 1.f * (4.f * lut_clip_both(a.curve, 501, 100.f / 60.f));
    const float c = 1.f + 1.f * (4.f * lut_clip_both(a.curve, 501, cN / 60.f));
    a.ccalc[(size_t)y * a.w2 + x] = cN > 100 ? c * c : cn100 * cn100;
}

// gamma + rgb2yuv + noise-variance maps, L2079-2126
struct SplitArgs {
    const float *r, *g, *b; size_t ip; int W, H;
    float *L, *a, *bb;            // dense W x H
    float *nvl, *nvc; const float* ccalc; int w2;
    Gam gam; float outer_gam, gain, wy0, wy1, wy2, noisevarL, maxNoiseVarab; int useCC;
    LabTabs lab;
};
__global__ void __launch_bounds__(256) k_dn_split(SplitArgs s)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= s.W) return;
    const size_t i = (size_t)y * s.ip + x, o = (size_t)y * s.W + x;
    float X = s.gain * s.r[i], Y = s.gain * s.g[i], Z = s.gain * s.b[i];
    if (s.lab.on) { X = lut_noclip(s.lab.dn_igamma, 65536, X); Y = lut_noclip(s.lab.dn_igamma, 65536, Y); Z = lut_noclip(s.lab.dn_igamma, 65536, Z); }   // L2093-2097
    X = apply_gam(s.gam, s.outer_gam, X); Y = apply_gam(s.gam, s.outer_gam, Y); Z = apply_gam(s.gam, s.outer_gam, Z);
    if (s.lab.on) {     // Color::rgb2lab(X, Y, Z, l, v, u, wpi): rgbxyz (color.cc L833-838) + XYZ2Lab (L1382-1399)
        const float* w = s.lab.wp;
        const float x = (w[0] * X + w[1] * Y + w[2] * Z), y = (w[3] * X + w[4] * Y + w[5] * Z), z = (w[6] * X + w[7] * Y + w[8] * Z);
        const float fx = computeXYZ2Lab(s.lab.cachef, x / 0.9642f), fy = computeXYZ2Lab(s.lab.cachef, y), fz = computeXYZ2Lab(s.lab.cachef, z / 0.8249f);
        s.L[o] = computeXYZ2LabY(s.lab.cachefy, y); s.a[o] = 500.0f * (fx - fy); s.bb[o] = 200.0f * (fy - fz);
    } else {
        const float l = X * s.wy0 + Y * s.wy1 + Z * s.wy2;
        s.L[o] = l; s.a[o] = X - l; s.bb[o] = l - Z;
    }
    if (((x | y) & 1) == 0) {
        const size_t h = (size_t)(y >> 1) * s.w2 + (x >> 1);
        s.nvl[h] = s.noisevarL;
        s.nvc[h] = s.useCC ? s.maxNoiseVarab * s.ccalc[h] : 1.f;
    }
}

// ordered overlap-add (RGBoutput_tile_row L531-558 + totwt L1577) and L += Ldetail / totwt (L1628-1632)
struct GatherArgs { float* L; const float* blocks; const float *tin, *tout; int width, height, nbw, nbh, nbw_out; };
__global__ void __launch_bounds__(256) k_dn_gather(GatherArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.width) return;
    const float DCTnorm = 1.0f / (4 * TS * TS);
    float det = 0.f, tw = 0.f;
    // blocks with top <= y < top + 64, top = (vblk - 1) * 25
    const int v0 = max(0, (y - TS + OFFSET * BLKRAD + OFFSET) / OFFSET), v1 = min(a.nbh - 1, (y + OFFSET * BLKRAD) / OFFSET);
    const int h0 = max(0, (x - TS + OFFSET * BLKRAD + OFFSET) / OFFSET), h1 = (x + OFFSET * BLKRAD) / OFFSET;
    for (int vb = v0; vb <= v1; ++vb) {
        const int i = y - (vb - BLKRAD) * OFFSET;
        if (i < 0 || i >= TS) continue;
        for (int hb = h0; hb <= h1; ++hb) {
            const int j = x - (hb - BLKRAD) * OFFSET;
            if (j < 0 || j >= TS) continue;
            const float ti = a.tin[i * TS + j], to = a.tout[i * TS + j];
            if (hb < a.nbw) tw += ti * to;
            if (hb < a.nbw_out) det += to * a.blocks[((size_t)vb * a.nbw + hb) * (TS * TS) + i * TS + j] * DCTnorm;
        }
    }
    float* p = a.L + (size_t)y * a.width + x;
    *p += det / tw;
}

// chroma boost, yuv2rgb, inverse gamma, L2489-2543
struct MergeArgs {
    const float *L, *a, *bb; float *r, *g, *b; size_t op; int W, H;
    Gam igam; float outer_gam, newGain, w10, w11, w12, realred, realblue, qhighFactor;
    LabTabs lab;
};
__global__ void __launch_bounds__(256) k_dn_merge(MergeArgs m)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= m.W) return;
    const size_t i = (size_t)y * m.W + x, o = (size_t)y * m.op + x;
    float av = m.a[i], bv = m.bb[i];
    const float c_h = sqrtf(av * av + bv * bv);
    if (c_h > 3000.f) {
        av *= 1.f + m.qhighFactor * m.realred / 100.f;
        bv *= 1.f + m.qhighFactor * m.realblue / 100.f;
    }
    const float Yv = m.L[i];
    float X, Y, Z;
    if (m.lab.on) {     // Color::lab2rgb(L, a, b, X, Y, Z, wpi_inverse): Lab2XYZ (color.cc L1203-1214) + xyz2rgb (L880-885)
        const float c1By116 = (float)(1.0 / 116.0), c16By116 = (float)(16.0 / 116.0);
        const float LL = Yv / 327.68f, aa = av / 327.68f, b2 = bv / 327.68f;
        const float fy = (c1By116 * LL) + c16By116;
        const float fx = (0.002f * aa) + fy;
        const float fz = fy - (0.005f * b2);
        const float x = 65535.0f * f2xyz(fx) * 0.9642f;
        const float z = 65535.0f * f2xyz(fz) * 0.8249f;
        const float y = ((double)LL > 8.0) ? 65535.0f * fy * fy * fy : (float)((double)(65535.0f * LL) / (24389.0 / 27.0));
        const float* w = m.lab.iwp;
        X = (w[0] * x + w[1] * y + w[2] * z); Y = (w[3] * x + w[4] * y + w[5] * z); Z = (w[6] * x + w[7] * y + w[8] * z);
    } else {
        Z = Yv - bv;
        X = av + Yv;
        Y = (Yv - X * m.w10 - Z * m.w12) / m.w11;
    }
    X = apply_gam(m.igam, m.outer_gam, X); Y = apply_gam(m.igam, m.outer_gam, Y); Z = apply_gam(m.igam, m.outer_gam, Z);
    if (m.lab.on) { X = lut_noclip(m.lab.dn_gamma, 65536, X); Y = lut_noclip(m.lab.dn_gamma, 65536, Y); Z = lut_noclip(m.lab.dn_gamma, 65536, Z); }   // L2533-2537
    m.r[o] = m.newGain * X; m.g[o] = m.newGain * Y; m.b[o] = m.newGain * Z;
}

inline float sqrf(float x) { return x * x; }

}  // namespace

// shrink.cu, internal forms: `uniform` != nullptr says the noise-variance map is that one value everywhere
int art_wavelet_denoise_L(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* uniform, const float* d_madL, double scale);
int art_wavelet_denoise_LAB(art_hp_ctx* ctx, art_hp_wavelet* wL, art_hp_wavelet* wa, art_hp_wavelet* wb, const float* d_noisevarchrom, const float* uniform_c,
                            float noisevar_a, float noisevar_b, int useNoiseCCurve, const float* d_noisevarlum, const float* uniform_l,
                            const float* d_madL, double scale, int with_L);
int art_wavelet_denoise_AB(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom, const float* uniform,
                           const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale, int bishrink = 0);

extern "C" {
int art_hp_wavelet_decompose_dev(art_hp_ctx* ctx, const float* d_src, size_t pitch, int W, int H, int maxlvl, int subsampling, art_hp_wavelet** out);
int art_hp_wavelet_maxlevel(const art_hp_wavelet* w);
int art_hp_wavelet_level_dims(const art_hp_wavelet* w, int level, int* width, int* height, int* stride);
float* art_hp_wavelet_band_dev(const art_hp_wavelet* w, int level, int dir);
int art_hp_wavelet_reconstruct_dev(art_hp_wavelet* w, float* d_dst, size_t pitch, float blend);
void art_hp_wavelet_destroy(art_hp_wavelet* w);
int art_hp_wavelet_mad_dev(art_hp_ctx* ctx, const art_hp_wavelet* w, float* d_madL);
int art_hp_wavelet_denoise_L_dev(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* d_madL, double scale);
int art_hp_wavelet_denoise_AB_dev(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom,
                                  const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale);
}

// tilemask_in / tilemask_out (L1833-1849) and the two DCT matrices, built on the host in double like the reference
static void build_tables(float* tin, float* tout, float* dctf, float* dctb)
{
    const float epsilon = 0.001f / (TS * TS);
    const int border = std::max(2, TS / 16);
    const double RT_PI = 3.14159265358979323846;
    for (int i = 0; i < TS; ++i) {
        const float i1 = (float)std::abs((i > TS / 2 ? i - TS + 1 : i));
        const float vmask = (i1 < border ? (float)(std::sin((RT_PI * i1) / (2 * border)) * std::sin((RT_PI * i1) / (2 * border))) : 1.0f);
        const float vmask2 = (i1 < 2 * border ? (float)(std::sin((RT_PI * i1) / (2 * border)) * std::sin((RT_PI * i1) / (2 * border))) : 1.0f);
        for (int j = 0; j < TS; ++j) {
            const float j1 = (float)std::abs((j > TS / 2 ? j - TS + 1 : j));
            const double sj = std::sin((RT_PI * j1) / (2 * border));
            tin[i * TS + j] = (float)((vmask * (j1 < border ? sj * sj : (double)1.0f)) + epsilon);
            tout[i * TS + j] = (float)((vmask2 * (j1 < 2 * border ? sj * sj : (double)1.0f)) + epsilon);
        }
    }
    for (int k = 0; k < TS; ++k)
        for (int j = 0; j < TS; ++j) {
            dctf[k * TS + j] = (float)(2.0 * std::cos(RT_PI * (j + 0.5) * k / TS));                       // REDFT10
            dctb[k * TS + j] = j == 0 ? 1.0f : (float)(2.0 * std::cos(RT_PI * j * (k + 0.5) / TS));       // REDFT01
        }
}

static float host_compute_detail(float d)
{
    const float a = static_cast<float>(((100. - d) * (100. - d)) + 50. * (100. - d)) * TS * 0.5f;
    return a * a;
}

// noise_residual: MAD^2 of the 3 * nlev bands into d_out[0 .. 3 nlev)
static int residual_mads(art_hp_ctx* ctx, const art_hp_wavelet* w, float* d_out) { return art_hp_wavelet_mad_dev(ctx, w, d_out); }

int art_rgb_denoise_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_denoise_params* P,
                        const double* wprof, const float* cl_r, const float* cl_g, const float* cl_b, size_t cp, float* nresi_highresi)
{
    cudaStream_t st = ctx->stream;
    const double scale = P->scale;
    const bool have_curve = P->noiseCCurve != nullptr;
    if (ctx->band.active && nresi_highresi) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "residual statistics are whole-frame: not available on a row band");
    if (P->luminance == 0 && P->chrominance == 0 && !have_curve) return ART_HP_OK;          // L1654-1666
    const bool useCC = have_curve && P->noiseCCurveSum > 5.f;
    if (useCC && !(cl_r && cl_g && cl_b)) return ctx->fail(ART_HP_ERR_INVALID, "the chroma noise curve needs the half-resolution calclum planes");
    const float noiseluma = static_cast<float>(P->luminance);
    const float noisevarL = static_cast<float>(((noiseluma / 125.0) * (1.0 + noiseluma / 25.0)) * ((noiseluma / 125.0) * (1.0 + noiseluma / 25.0)));
    const bool denoiseLuminance = (noisevarL > 0.00001f);
    float wp[9];
    for (int i = 0; i < 9; ++i) wp[i] = static_cast<float>(wprof[i]);
    if (!(P->luminance != 0 || P->chrominance != 0)) return ART_HP_OK;                      // curve set but both sliders 0: nothing visible happens

    const size_t n = (size_t)W * H;
    const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
    const size_t nh = (size_t)w2 * h2;
    const int nbw = (int)std::ceil(((float)W) / OFFSET) + 2 * BLKRAD, nbh = (int)std::ceil(((float)H) / OFFSET) + 2 * BLKRAD;
    const size_t nblk = denoiseLuminance ? (size_t)nbw * nbh * TS * TS : 0;
    const bool use_mask = denoiseLuminance && P->luminanceDetailThreshold > 0;
    // work buffer: L, a, b, Lin (n each), [mask n], nvl, nvc, ccalc (nh each), LUTs, tables, small
    const size_t small_floats = 2 * 65536 + 65536 + 512 + 4 * TS * TS + 64 + 64;
    size_t floats = round_up(n, 64) * (4 + (use_mask ? 1 : 0)) + round_up(nh, 64) * 3 + small_floats + nblk + 2 * round_up((size_t)(W / 4 + 1) * (H / 4 + 1), 64);
    int rc = art_reserve(ctx, ctx->d_dn, floats * sizeof(float));
    if (rc) return rc;
    float* p = (float*)ctx->d_dn.p;
    auto take = [&p](size_t k) { float* q = p; p += round_up(k, 64); return q; };
    float *Lp = take(n), *ap = take(n), *bp = take(n), *Lin = take(n), *mask = use_mask ? take(n) : nullptr;
    float *nvl = take(nh), *nvc = take(nh), *ccalc = take(nh);
    float *gamcurve = take(65536), *igamcurve = take(65536), *cachef = take(65536), *curve = take(512);
    float *tin = take(TS * TS), *tout = take(TS * TS), *dctf = take(TS * TS), *dctb = take(TS * TS);
    float *madL = take(64), *resid = take(64);
    float* quarter = take(2 * (size_t)(W / 4 + 1) * (H / 4 + 1));
    float* blocks = take(nblk);

    // gamma curves, L1779-1806
    const float gam = static_cast<float>(P->gamma);
    const float gamthresh = 0.001f;
    const float gamslope = std::exp(std::log(static_cast<double>(gamthresh)) / gam) / gamthresh;
    const float igam = 1.f / gam;
    const float igamthresh = gamthresh * gamslope;
    const float igamslope = 1.f / gamslope;
    art_prof_begin(ctx, "k_dn_tables");
    k_dn_gamma_lut<<<256, 256, 0, st>>>(gamcurve, gam, gamthresh, gamslope, 65535.f, 65535.f);
    k_dn_gamma_lut<<<256, 256, 0, st>>>(igamcurve, igam, igamthresh, igamslope, 65535.f, 65535.f);
    art_prof_end(ctx);
    ctx->launches += 2;
    const float gain = std::pow(2.0f, float(0.0));
    const float params_Ldetail = std::min(float(P->luminanceDetail), 99.9f);

    if (denoiseLuminance) {
        // the window / DCT tables are constants: built and uploaded once per context
        constexpr size_t NT0 = 4 * TS * TS;                             // tilemask_in, tilemask_out, REDFT10 and REDFT01 matrices, dense
        constexpr size_t NT = NT0 + DN_SPLIT_WORDS;                     // + the pre-split tcgen05 operand images of the two DCT matrices
        int trc = art_reserve(ctx, ctx->d_dn_tables, NT * sizeof(float));
        if (trc) return trc;
        float* tb = (float*)ctx->d_dn_tables.p;
        if (!ctx->dn_tables_ready) {
            std::vector<float> host(NT, 0.f);
            build_tables(host.data(), host.data() + TS * TS, host.data() + 2 * TS * TS, host.data() + 3 * TS * TS);
            art_dn_blocks_split_tables(host.data() + 2 * TS * TS, host.data() + 3 * TS * TS, reinterpret_cast<unsigned*>(host.data() + NT0));
            ART_CUDA(ctx, cudaMemcpyAsync(tb, host.data(), sizeof(float) * NT, cudaMemcpyHostToDevice, st));
            ART_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->dn_tables_ready = true;
        }
        tin = tb; tout = tb + TS * TS; dctf = tb + 2 * TS * TS; dctb = tb + 3 * TS * TS;
    }
    LabTabs labt{};
    if (P->colorSpace == 1) {      // DenoiseParams::ColorSpace::LAB
        if (!P->wprof_inverse) return ctx->fail(ART_HP_ERR_INVALID, "colorSpace LAB needs wprof_inverse");
        int lrc = art_reserve(ctx, ctx->d_dn_labtabs, 4 * 65536 * sizeof(float));
        if (lrc) return lrc;
        float* lt = (float*)ctx->d_dn_labtabs.p;
        if (!ctx->dn_labtabs_ready) {      // color.cc L188-189, L205-233, L278-292 (host libm, as in the reference)
            std::vector<float> h(4 * 65536);
            const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
            const int epsmaxint = (int)(MAXVALF * eps);
            for (int i = 0; i < 65536; i++) {
                const double x = i / 65535.0;
                h[i] = (float)(65535.0 * (x <= 0.131889 ? x / 10.0 : std::exp(std::log((x + 0.593503) / 1.593503) * 5.5)));           // denoiseIGammaTab
                h[65536 + i] = (float)(65535.0 * (x <= 0.013189 ? x * 10.0 : 1.593503 * std::exp(std::log(x) / 5.5) - 0.593503));   // denoiseGammaTab
                if (i <= epsmaxint) { h[2 * 65536 + i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); h[3 * 65536 + i] = (float)(327.68 * (kappa * i / MAXVALF)); }
                else { h[2 * 65536 + i] = (float)(327.68 * std::cbrt((double)i / MAXVALF)); h[3 * 65536 + i] = (float)(327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0)); }
            }
            ART_CUDA(ctx, cudaMemcpyAsync(lt, h.data(), sizeof(float) * 4 * 65536, cudaMemcpyHostToDevice, st));
            ART_CUDA(ctx, cudaStreamSynchronize(st));
            ctx->dn_labtabs_ready = true;
        }
        labt.on = 1; labt.dn_igamma = lt; labt.dn_gamma = lt + 65536; labt.cachef = lt + 2 * 65536; labt.cachefy = lt + 3 * 65536;
        for (int i = 0; i < 9; ++i) { labt.wp[i] = wp[i]; labt.iwp[i] = static_cast<float>(P->wprof_inverse[i]); }
    }
    std::vector<float> hcache;
    if (useCC) {       // L1706-1770
        hcache.resize(65536);
        const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
        const int epsmaxint = (int)(MAXVALF * eps);
        int i = 0;
        for (; i <= epsmaxint; i++) hcache[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0));      // color.cc L205-216
        for (; i < 65536; i++) hcache[i] = (float)(327.68 * std::cbrt((double)i / MAXVALF));
        ART_CUDA(ctx, cudaMemcpyAsync(cachef, hcache.data(), sizeof(float) * 65536, cudaMemcpyHostToDevice, st));
        ART_CUDA(ctx, cudaMemcpyAsync(curve, P->noiseCCurve, sizeof(float) * 501, cudaMemcpyHostToDevice, st));
        CcArgs c{};
        c.r = cl_r; c.g = cl_g; c.b = cl_b; c.cp = cp; c.w2 = w2; c.h2 = h2; c.cachef = cachef; c.curve = curve; c.ccalc = ccalc;
        for (int k = 0; k < 9; ++k) c.wp[k] = wp[k];
        art_prof_begin(ctx, "k_dn_ccalc");
        k_dn_ccalc<<<dim3((w2 + 255) / 256, h2), 256, 0, st>>>(c);
        art_prof_end(ctx);
        ctx->launches += 1;
    }
    if (useCC) ART_CUDA(ctx, cudaStreamSynchronize(st));        // hcache (a local) must outlive its copy

    // chroma sliders, L2026-2068
    float interm_med = static_cast<float>(P->chrominance) / 10.0;
    float intermred = P->chrominanceRedGreen > 0. ? (P->chrominanceRedGreen / 10.) : static_cast<float>(P->chrominanceRedGreen) / 7.0;
    float intermblue = P->chrominanceBlueYellow > 0. ? (P->chrominanceBlueYellow / 10.) : static_cast<float>(P->chrominanceBlueYellow) / 7.0;
    float realred = interm_med + intermred;
    if (realred <= 0.f) realred = 0.001f;
    float realblue = interm_med + intermblue;
    if (realblue <= 0.f) realblue = 0.001f;
    const float noisevarab_r = sqrf(realred), noisevarab_b = sqrf(realblue);
    const float maxNoiseVarab = std::max(noisevarab_b, noisevarab_r);

    SplitArgs s{};
    s.r = r; s.g = g; s.b = b; s.ip = ip; s.W = W; s.H = H; s.L = Lp; s.a = ap; s.bb = bp; s.nvl = nvl; s.nvc = nvc; s.ccalc = ccalc; s.w2 = w2;
    s.gam = Gam{gamcurve, gam, gamthresh, gamslope, 65535.f}; s.outer_gam = gam; s.gain = gain;
    s.wy0 = wp[3]; s.wy1 = wp[4]; s.wy2 = wp[5]; s.noisevarL = noisevarL; s.maxNoiseVarab = maxNoiseVarab; s.useCC = useCC; s.lab = labt;
    const dim3 gfull((W + 255) / 256, H);
    art_prof_begin(ctx, "k_dn_split");
    k_dn_split<<<gfull, 256, 0, st>>>(s);
    art_prof_end(ctx);
    ctx->launches += 1;
    ART_CUDA(ctx, cudaGetLastError());

    // wavelet levels, L2246-2293
    int levwav = 5;
    const float maxreal = std::max(realred, realblue);
    if (maxreal < 8.f) levwav = 5; else if (maxreal < 10.f) levwav = 6; else if (maxreal < 15.f) levwav = 7; else levwav = 8;
    const bool aggressive = P->aggressive != 0;         // nrQuality == QUALITY_HIGH, L1671
    if (aggressive) levwav += 2;                         // L2260-2262
    if (levwav > 8) levwav = 8;
    levwav = std::max(5, int(levwav - std::ceil(std::log(scale))));
    const int minsizetile = std::min(W, ctx->band.active ? ctx->band.H_full : H);      // a row band follows its frame's wavelet depth
    int maxlev2 = 8;
    if (minsizetile < 256) maxlev2 = 7;
    if (minsizetile < 128) maxlev2 = 6;
    if (minsizetile < 64) maxlev2 = 5;
    if (minsizetile < 32) maxlev2 = 4;
    if (minsizetile < 16) maxlev2 = 3;
    levwav = std::min(maxlev2, levwav);

    art_hp_wavelet* Ldec = nullptr;
    if ((rc = art_hp_wavelet_decompose_dev(ctx, Lp, W, W, H, levwav, 1, &Ldec))) return rc;
    ART_CUDA(ctx, cudaMemsetAsync(madL, 0, 64 * sizeof(float), st));
    if ((rc = art_hp_wavelet_mad_dev(ctx, Ldec, madL))) { art_hp_wavelet_destroy(Ldec); return rc; }
    float* chan[2] = {ap, bp};
    const float nv[2] = {noisevarab_r, noisevarab_b};
    // The standard quality runs the shrinkage of a, b and L as one batch (shrink.cu: art_wavelet_denoise_LAB); ART_HP_SHRINK_MERGE = 0 restores the
    // channel-by-channel order of the reference (same bits either way), 1 merges a and b only.
    static const int merge_mode = [] { const char* e = getenv("ART_HP_SHRINK_MERGE"); return e ? atoi(e) : 2; }();
    const int maxlvl = art_hp_wavelet_maxlevel(Ldec);
    (void)maxlvl;
    if (!aggressive && merge_mode > 0) {
        art_hp_wavelet* dec[2] = {nullptr, nullptr};
        for (int c = 0; c < 2 && !rc; ++c) rc = art_hp_wavelet_decompose_dev(ctx, chan[c], W, W, H, levwav, 1, &dec[c]);
        const float one = 1.f;
        const bool with_L = denoiseLuminance && merge_mode > 1;
        if (!rc) rc = art_wavelet_denoise_LAB(ctx, Ldec, dec[0], dec[1], nvc, useCC ? nullptr : &one, nv[0], nv[1], useCC, nvl, &noisevarL, madL, scale, with_L);
        for (int c = 0; c < 2; ++c) {
            if (!rc && nresi_highresi) rc = residual_mads(ctx, dec[c], resid + 24 * c);
            if (!rc) rc = art_hp_wavelet_reconstruct_dev(dec[c], chan[c], W, 1.f);
            if (dec[c]) art_hp_wavelet_destroy(dec[c]);
        }
        if (!rc && denoiseLuminance) {
            if (!with_L) rc = art_wavelet_denoise_L(ctx, Ldec, nvl, &noisevarL, madL, scale);
            if (!rc && cudaMemcpyAsync(Lin, Lp, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = ctx->fail(ART_HP_ERR_CUDA, "Lin copy failed");
            if (!rc) rc = art_hp_wavelet_reconstruct_dev(Ldec, Lp, W, 1.f);
        }
    } else {
    for (int c = 0; c < 2; ++c) {
        art_hp_wavelet* dec = nullptr;
        if ((rc = art_hp_wavelet_decompose_dev(ctx, chan[c], W, W, H, levwav, 1, &dec))) { art_hp_wavelet_destroy(Ldec); return rc; }
        {   // without the chroma noise curve the chroma variance map is 1 everywhere (L2098-2100): passed as a value, not read
            const float one = 1.f;
            // QUALITY_HIGH: WaveletDenoiseAll_BiShrinkAB and then WaveletDenoiseAllAB as well (L2339-2349)
            if (aggressive) rc = art_wavelet_denoise_AB(ctx, Ldec, dec, nvc, useCC ? nullptr : &one, madL, nv[c], useCC, 0, scale, 1);
            if (!rc) rc = art_wavelet_denoise_AB(ctx, Ldec, dec, nvc, useCC ? nullptr : &one, madL, nv[c], useCC, 0, scale, 0);
        }
        if (!rc && nresi_highresi) rc = residual_mads(ctx, dec, resid + 24 * c);
        if (!rc) rc = art_hp_wavelet_reconstruct_dev(dec, chan[c], W, 1.f);
        art_hp_wavelet_destroy(dec);
        if (rc) { art_hp_wavelet_destroy(Ldec); return rc; }
    }
    if (denoiseLuminance) {
        // QUALITY_HIGH: WaveletDenoiseAll_BiShrinkL (the same computation as WaveletDenoiseAllL) and then WaveletDenoiseAllL (L2412-2421)
        if (aggressive) rc = art_wavelet_denoise_L(ctx, Ldec, nvl, &noisevarL, madL, scale);
        if (!rc) rc = art_wavelet_denoise_L(ctx, Ldec, nvl, &noisevarL, madL, scale);       // the luminance variance map is noisevarL everywhere (L2097)
        if (!rc) {
            if (cudaMemcpyAsync(Lin, Lp, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = ctx->fail(ART_HP_ERR_CUDA, "Lin copy failed");
        }
        if (!rc) rc = art_hp_wavelet_reconstruct_dev(Ldec, Lp, W, 1.f);
    }
    }
    art_hp_wavelet_destroy(Ldec);
    if (rc) return rc;

    if (denoiseLuminance) {
        if (use_mask) {
            const float amount = std::max(0.f, std::min(float(P->luminanceDetailThreshold) / 100.f, 1.f));
            if ((rc = art_detail_mask_dev(ctx, Lp, W, mask, W, W, H, 65535.f, 25.f, 10000.f, amount, 2, 25.f / scale, quarter))) return rc;
        }
        DnBlocksArgs b{};
        b.Lin = Lin; b.L = Lp; b.mask = mask; b.width = W; b.height = H; b.nbw = nbw; b.nbh = nbh; b.tin = tin;
        b.fwd_split = reinterpret_cast<const unsigned*>(tin + 4 * TS * TS); b.bwd_split = b.fwd_split + DN_SPLIT_WORDS / 2;
        b.blocks = blocks; b.detail_hi = host_compute_detail(params_Ldetail); b.detail_lo = host_compute_detail(0.f); b.params_Ldetail = params_Ldetail;
        b.use_mask = use_mask; b.blur_rad = std::max(1, int(3 / scale));
        art_prof_begin(ctx, "k_dn_blocks");
        rc = art_dn_blocks_launch(ctx, b);
        art_prof_end(ctx);
        if (rc) return rc;
        GatherArgs ga{};
        ga.L = Lp; ga.blocks = blocks; ga.tin = tin; ga.tout = tout; ga.width = W; ga.height = H; ga.nbw = nbw; ga.nbh = nbh;
        ga.nbw_out = (int)std::ceil(((float)W) / OFFSET);
        art_prof_begin(ctx, "k_dn_gather");
        k_dn_gather<<<gfull, 256, 0, st>>>(ga);
        art_prof_end(ctx);
        ctx->launches += 2;
        ART_CUDA(ctx, cudaGetLastError());
    }

    MergeArgs m{};
    m.L = Lp; m.a = ap; m.bb = bp; m.r = r; m.g = g; m.b = b; m.op = ip; m.W = W; m.H = H;
    m.igam = Gam{igamcurve, igam, igamthresh, igamslope, 65536.f}; m.outer_gam = gam; m.newGain = 1.f / gain;
    m.lab = labt; m.w10 = wp[3]; m.w11 = wp[4]; m.w12 = wp[5]; m.realred = realred; m.realblue = realblue; m.qhighFactor = aggressive ? 1.f / static_cast<float>(0.9) : 1.0f;
    art_prof_begin(ctx, "k_dn_merge");
    k_dn_merge<<<gfull, 256, 0, st>>>(m);
    art_prof_end(ctx);
    ctx->launches += 1;
    ART_CUDA(ctx, cudaGetLastError());

    if (nresi_highresi) {       // Noise_residualAB + L2398-2404
        float h[48];
        ART_CUDA(ctx, cudaMemcpyAsync(h, resid, sizeof h, cudaMemcpyDeviceToHost, st));
        ART_CUDA(ctx, cudaStreamSynchronize(st));
        float res[2], mx[2];
        for (int c = 0; c < 2; ++c) {
            float rs = 0.f, mr = 0.f;
            for (int lvl = 0; lvl < maxlvl; ++lvl)
                for (int d = 0; d < 3; ++d) { const float madC = h[24 * c + 3 * lvl + d]; rs += madC; if (madC > mr) mr = madC; }
            res[c] = rs; mx[c] = mr;
        }
        float chresid = res[1] + res[0], chmaxresid = mx[1] + mx[0];
        chresid = std::sqrt(chresid / (6 * (levwav)));
        nresi_highresi[1] = chresid + 0.66f * (std::sqrt(chmaxresid) - chresid);
        nresi_highresi[0] = chresid;
    }
    return ART_HP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Automatic chroma estimator: ImProcFunctions::denoiseComputeParams for DenoiseParams::ChrominanceMethod::AUTOMATIC, the reference's
// default (rtengine/ipdenoise.cc L800-1093; procparams.cc L1909).  Nine crops of HALF the frame's width and height (Tile_calc returns
// one tile, so crW = widIm / 2, crH = heiIm / 2: L866-877, FTblockDN.cc L442-480) at 50 px from the corners / centred, each through
// RGB_denoise_info (ipdenoise.cc L227-669): getImage (gain / clip, camera space) -> gamma LUT + rgb2yuv -> two 5-level wavelet
// decompositions -> WaveletDenoiseAll_info / ShrinkAll_info (FTblockDN.cc L1227-1364): the MAD of every subband and, on level 1, running
// statistics of the crop's half-resolution chroma / hue / luminance maps (provicalc through the camera->working matrix, XYZ2Lab).
// Those statistics are float sums in raster order whose terms depend on the running mean (`dev += SQR(c - chro / nc)`), so the order is
// part of the result: k_auto_stats keeps it -- per crop one CTA streams the maps through shared memory; two lanes carry the running sums
// (one dependent add per element), all threads form the terms, two more lanes add them, pipelined one chunk apart.
// The scalar tail (calcautodn_info L66-206 and the nine-crop combination L964-1075) is ~200 flops and runs on the host after one
// download of 9 x (30 MADs + 10 sums): the number of wavelet levels of RGB_denoise depends on the result, so the host has to see it.
namespace {

struct AutoPrepArgs {
    const float *r, *g, *b; size_t ip;          // demosaiced planes, camera space, addressed at the crop's origin
    int W, H, wid;                               // crop size, half-resolution width
    float mul[3]; int do_clip, do_mat; double mat[9];
    float wp[9]; const float *cachef, *cachefy, *gamcurve;
    float gam, gamthresh, gamslope, gain;
    float *la, *lb;                              // dense W x H
    float *nvc, *nvh, *nvl;                      // dense wid x hei
};

__device__ __forceinline__ float clip65535_dn(float a) { const float m = a < 65535.f ? a : 65535.f; return 0.f < m ? m : 0.f; }

__global__ void __launch_bounds__(256) k_auto_prep(AutoPrepArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t i = (size_t)y * a.ip + x;
        float R = a.r[i] * a.mul[0], G = a.g[i] * a.mul[1], B = a.b[i] * a.mul[2];          // getImage, rawimagesource.cc L943-1025
        if (a.do_clip) { R = clip65535_dn(R); G = clip65535_dn(G); B = clip65535_dn(B); }
        if (!((x | y) & 1)) {      // provicalc: the even pixels through convertColorSpace's matrix, then Lab (ipdenoise.cc L271-287, L390-420)
            float RL = R, GL = G, BL = B;
            if (a.do_mat) {
                const double dr = R, dg = G, db = B;
                RL = (float)(a.mat[0] * dr + a.mat[1] * dg + a.mat[2] * db);
                GL = (float)(a.mat[3] * dr + a.mat[4] * dg + a.mat[5] * db);
                BL = (float)(a.mat[6] * dr + a.mat[7] * dg + a.mat[8] * db);
            }
            const float XL = ((a.wp[0] * RL + a.wp[1] * GL + a.wp[2] * BL));
            const float YL = ((a.wp[3] * RL + a.wp[4] * GL + a.wp[5] * BL));
            const float ZL = ((a.wp[6] * RL + a.wp[7] * GL + a.wp[8] * BL));
            const float fx = computeXYZ2Lab(a.cachef, XL / 0.9642f), fy = computeXYZ2Lab(a.cachef, YL), fz = computeXYZ2Lab(a.cachef, ZL / 0.8249f);
            const float aN = (500.0f * (fx - fy)), bN = (200.0f * (fy - fz));
            float cN = sqrtf(aN * aN + bN * bN);
            const size_t o = (size_t)(y >> 1) * a.wid + (x >> 1);
            a.nvh[o] = sleef::xatan2f(bN, aN);
            if (cN < 100.f) cN = 100.f;
            a.nvc[o] = cN;
            float Llum = computeXYZ2LabY(a.cachefy, YL);
            Llum = Llum < 2.f ? 2.f : Llum;
            Llum = Llum > 32768.f ? 32768.f : Llum;
            a.nvl[o] = Llum;
        }
        float X = a.gain * R, Y = a.gain * G, Z = a.gain * B;                                  // ipdenoise.cc L430-447
        X = X < 65535.f ? lut_noclip(a.gamcurve, 65536, X) : (gammaf_(X / 65535.f, a.gam, a.gamthresh, a.gamslope) * 32768.f);
        Y = Y < 65535.f ? lut_noclip(a.gamcurve, 65536, Y) : (gammaf_(Y / 65535.f, a.gam, a.gamthresh, a.gamslope) * 32768.f);
        Z = Z < 65535.f ? lut_noclip(a.gamcurve, 65536, Z) : (gammaf_(Z / 65535.f, a.gam, a.gamthresh, a.gamslope) * 32768.f);
        const float l = X * a.wp[3] + Y * a.wp[4] + Z * a.wp[5];
        const size_t o = (size_t)y * a.W + x;
        a.la[o] = X - l;
        a.lb[o] = l - Z;
    }
}

// ShrinkAll_info's level-1 statistics (FTblockDN.cc L1262-1316) over rows [0, Hab) x columns [0, Wab) of maps with row stride `wid`.
// out: chro, dev, lume, devL, red_yel, skin_c, then nc, nry, nsk as floats (counts are exact below 2^24; nc is passed back as an int too)
struct AutoStatsArgs { const float *nvc, *nvh, *nvl; int wid, Wab, Hab; float* out; int* out_n; };
constexpr int AS_CHUNK = 1024, AS_THREADS = 256;      // 10 arrays x 1024 floats = 40 KB of static shared memory

__global__ void __launch_bounds__(AS_THREADS) k_auto_stats(const AutoStatsArgs* __restrict__ crops)
{
    const AutoStatsArgs a = crops[blockIdx.x];
    __shared__ float sc[2][AS_CHUNK], sh[2][AS_CHUNK], sl[2][AS_CHUNK];       // chroma, hue, luminance of a chunk (double buffered)
    __shared__ float tdev[2][AS_CHUNK], tdevL[2][AS_CHUNK];                   // the terms SQR(x - running mean)
    const long long total = (long long)a.Wab * a.Hab;
    const int nchunks = (int)((total + AS_CHUNK - 1) / AS_CHUNK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float chro = 0.f, lume = 0.f, dev = 0.f, devL = 0.f, red_yel = 0.f, skin_c = 0.f;     // live in lane 0 of warps 0..3
    int nry = 0, nsk = 0;
    auto load = [&](int c, int buf) {
        const long long base = (long long)c * AS_CHUNK;
        for (int k = tid; k < AS_CHUNK; k += AS_THREADS) {
            const long long e = base + k;
            if (e < total) {
                const int i = (int)(e / a.Wab), j = (int)(e - (long long)i * a.Wab);
                const size_t o = (size_t)i * a.wid + j;
                sc[buf][k] = a.nvc[o]; sh[buf][k] = a.nvh[o]; sl[buf][k] = a.nvl[o];
            }
        }
    };
    if (nchunks > 0) load(0, 0);
    __syncthreads();
    // iteration c: lanes run the serial sums of chunk c while the other threads fetch chunk c + 1; then everybody forms chunk c's terms;
    // the term sums of chunk c run in iteration c + 1 beside the running sums of chunk c + 1
    for (int c = 0; c <= nchunks; ++c) {
        const int buf = c & 1;
        const int cnt = c < nchunks ? (int)min((long long)AS_CHUNK, total - (long long)c * AS_CHUNK) : 0;
        const int pcnt = c > 0 ? (int)min((long long)AS_CHUNK, total - (long long)(c - 1) * AS_CHUNK) : 0;
        if (lane == 0) {
            if (warp == 0) {          // chro += c: the running sum after each element goes where the term will be formed (the mean's numerator)
#pragma unroll 8
                for (int k = 0; k < cnt; ++k) { chro += sc[buf][k]; tdev[buf][k] = chro; }
            } else if (warp == 1) {
#pragma unroll 8
                for (int k = 0; k < cnt; ++k) { lume += sl[buf][k]; tdevL[buf][k] = lume; }
            } else if (warp == 2) {
                for (int k = 0; k < cnt; ++k) {
                    const float cv = sc[buf][k], hv = sh[buf][k];
                    if (hv > -0.8f && hv < 2.0f && cv > 10000.f) { red_yel += cv; ++nry; }
                    if (hv > 0.f && hv < 1.6f && cv < 10000.f) { skin_c += cv; ++nsk; }
                }
            } else if (warp == 3) {
#pragma unroll 8
                for (int k = 0; k < pcnt; ++k) dev += tdev[buf ^ 1][k];
            } else if (warp == 4) {
#pragma unroll 8
                for (int k = 0; k < pcnt; ++k) devL += tdevL[buf ^ 1][k];
            }
        }
        if (warp >= 5 && c + 1 < nchunks) {
            // warps 5..7 fetch the next chunk into the other buffer's inputs: its terms (tdev[buf ^ 1]) are still being summed by warps
            // 3 / 4, but its inputs sc / sh / sl were consumed in the previous iteration
            const long long base = (long long)(c + 1) * AS_CHUNK;
            for (int k = tid - 160; k < AS_CHUNK; k += AS_THREADS - 160) {
                const long long e = base + k;
                if (e < total) {
                    const int i = (int)(e / a.Wab), j = (int)(e - (long long)i * a.Wab);
                    const size_t o = (size_t)i * a.wid + j;
                    sc[buf ^ 1][k] = a.nvc[o]; sh[buf ^ 1][k] = a.nvh[o]; sl[buf ^ 1][k] = a.nvl[o];
                }
            }
        }
        __syncthreads();
        // terms of chunk c: SQR(x - sum / n) with n = elements so far (float(sum) / int, L1268 / L1288: the int is converted to float)
        const long long base = (long long)c * AS_CHUNK;
        for (int k = tid; k < cnt; k += AS_THREADS) {
            const float n = (float)(int)(base + k + 1);
            const float d1 = sc[buf][k] - (tdev[buf][k] / n), d2 = sl[buf][k] - (tdevL[buf][k] / n);
            tdev[buf][k] = d1 * d1; tdevL[buf][k] = d2 * d2;
        }
        __syncthreads();
    }
    if (tid == 0) { a.out[0] = chro; a.out_n[0] = (int)total; }
    if (tid == 32) a.out[2] = lume;
    if (tid == 64) { a.out[4] = red_yel; a.out[5] = skin_c; a.out_n[1] = nry; a.out_n[2] = nsk; }
    if (tid == 96) a.out[1] = dev;
    if (tid == 128) a.out[3] = devL;
}

// calcautodn_info, ipdenoise.cc L66-206 (mode 1, lissage 0, levaut 0 as denoiseComputeParams calls it; the other branches are kept)
void calcautodn_info_host(bool aggressive, float& chaut, float& delta, int Nb, int levaut, float maxmax, float lumema, float chromina, int mode, int lissage,
                          float redyel, float skinc, float nsknc)
{
    float reducdelta = 1.f;
    if (aggressive) reducdelta = static_cast<float>(0.9);
    chaut = (chaut * Nb - maxmax) / (Nb - 1);
    if ((redyel > 5000.f || skinc > 1000.f) && nsknc < 0.4f && chromina > 3000.f) chaut *= 0.45f;
    else if ((redyel > 12000.f || skinc > 1200.f) && nsknc < 0.3f && chromina > 3000.f) chaut *= 0.3f;
    if (mode == 0 || mode == 2) {
        if (chromina > 10000.f) chaut *= 0.7f; else if (chromina > 6000.f) chaut *= 0.9f; else if (chromina < 3000.f) chaut *= 1.2f; else if (chromina < 2000.f) chaut *= 1.5f;
        if (lumema < 2500.f) chaut *= 1.3f; else if (lumema < 5000.f) chaut *= 1.2f; else if (lumema > 20000.f) chaut *= 0.9f;
    } else if (mode == 1) {
        if (chromina > 10000.f) chaut *= 0.8f; else if (chromina > 6000.f) chaut *= 0.9f; else if (chromina < 3000.f) chaut *= 1.5f; else if (chromina < 2000.f) chaut *= 2.2f;
        if (lumema < 2500.f) chaut *= 1.2f; else if (lumema < 5000.f) chaut *= 1.1f; else if (lumema > 20000.f) chaut *= 0.9f;
    }
    if (levaut == 0 && chaut > 300.f) chaut = 0.714286f * chaut + 85.71428f;
    delta = maxmax - chaut;
    delta *= reducdelta;
    if (lissage == 1 || lissage == 2) {
        if (chaut < 200.f && delta < 200.f) delta *= 0.95f; else if (chaut < 200.f && delta < 400.f) delta *= 0.5f; else if (chaut < 200.f && delta >= 400.f) delta = 200.f;
        else if (chaut < 400.f && delta < 400.f) delta *= 0.4f; else if (chaut < 400.f && delta >= 400.f) delta = 120.f;
        else if (chaut < 550.f) delta *= 0.15f; else if (chaut < 650.f) delta *= 0.1f; else delta *= 0.07f;
        if (mode == 0 || mode == 2) { if (chromina < 6000.f) delta *= 1.4f; if (lumema < 5000.f) delta *= 1.4f; }
        else if (mode == 1) { if (chromina < 6000.f) delta *= 1.2f; if (lumema < 5000.f) delta *= 1.2f; }
    }
    if (lissage == 0) {
        if (chaut < 200.f && delta < 200.f) delta *= 0.95f; else if (chaut < 200.f && delta < 400.f) delta *= 0.7f; else if (chaut < 200.f && delta >= 400.f) delta = 280.f;
        else if (chaut < 400.f && delta < 400.f) delta *= 0.6f; else if (chaut < 400.f && delta >= 400.f) delta = 200.f;
        else if (chaut < 550.f) delta *= 0.3f; else if (chaut < 650.f) delta *= 0.2f; else delta *= 0.15f;
        if (mode == 0 || mode == 2) { if (chromina < 6000.f) delta *= 1.4f; if (lumema < 5000.f) delta *= 1.4f; }
        else if (mode == 1) { if (chromina < 6000.f) delta *= 1.2f; if (lumema < 5000.f) delta *= 1.2f; }
    }
}

}  // namespace

int art_denoise_auto_chroma_dev(art_hp_ctx* ctx, const float* r, const float* g, const float* b, size_t ip, int widIm, int heiIm,
                                const float mul[3], int doClip, const double* cam2work, const double* wprof, double gamma, int aggressive,
                                float out3[3], float* stats_out)
{
    cudaStream_t st = ctx->stream;
    const int crW = widIm / 2, crH = heiIm / 2;                    // L866-877 with Tile_calc's single tile
    const int levwav = 5;                                           // max(2, 5 - ceil(log(scale))) at RGB_denoise_info's scale 1 (L935, L449)
    if (crW < 128 || crH < 128)
        return ctx->fail(ART_HP_ERR_UNSUPPORTED, "automatic chroma needs a frame of at least 256 x 256 after the border crop (got %dx%d)", widIm, heiIm);
    const int wid = (crW + 1) / 2, hei = (crH + 1) / 2;
    const size_t n = (size_t)crW * crH, nh = (size_t)wid * hei;
    // scratch: la, lb (one crop at a time), nine sets of half-resolution maps, LUTs, results
    const size_t floats = 2 * round_up(n, 64) + 27 * round_up(nh, 64) + 3 * 65536 + 9 * 64 + 9 * 16 + 1024;
    void* blk = nullptr;
    int rc = art_pool_alloc(ctx, floats * sizeof(float), &blk);
    if (rc) return rc;
    float* p = (float*)blk;
    auto take = [&p](size_t k) { float* q = p; p += round_up(k, 64); return q; };
    float *la = take(n), *lb = take(n);
    float* maps = take(27 * round_up(nh, 64));
    float *cachef = take(65536), *cachefy = take(65536), *gamcurve = take(65536);
    float* mads = take(9 * 64);                                     // per crop: a[8][3] at 0, b[8][3] at 24 (art_hp_wavelet_mad_dev's layout)
    float* sums = take(9 * 16);                                     // per crop: 6 floats + 3 ints
    AutoStatsArgs* d_crops = reinterpret_cast<AutoStatsArgs*>(take(512));
    auto done = [&](int code) { art_pool_free(ctx, blk); return code; };

    {   // Color::cachef / cachefy, color.cc L205-233 (host libm cbrt, as in the reference)
        std::vector<float>& h = ctx->h_auto_tabs;
        if (h.empty()) {
            h.resize(2 * 65536);
            const double eps = 216.0 / 24389.0, kappa = 24389.0 / 27.0, MAXVALF = 65535.f;
            const int epsmaxint = (int)(MAXVALF * eps);
            int i = 0;
            for (; i <= epsmaxint; i++) { h[i] = (float)(327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0)); h[65536 + i] = (float)(327.68 * (kappa * i / MAXVALF)); }
            for (; i < 65536; i++) { h[i] = (float)(327.68 * std::cbrt((double)i / MAXVALF)); h[65536 + i] = (float)(327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0)); }
        }
        if (cudaMemcpyAsync(cachef, h.data(), sizeof(float) * 2 * 65536, cudaMemcpyHostToDevice, st) != cudaSuccess) return done(ctx->fail(ART_HP_ERR_CUDA, "table upload failed"));
    }
    // RGB_denoise_infoGamCurve, L209-225 (isRAW)
    const float gam = static_cast<float>(gamma);
    const float gamthresh = 0.001f;
    const float gamslope = std::exp(std::log(static_cast<double>(gamthresh)) / gam) / gamthresh;
    k_dn_gamma_lut<<<256, 256, 0, st>>>(gamcurve, gam, gamthresh, gamslope, 65535.f, 32768.f);
    ctx->launches++;
    const float gain = std::pow(2.0f, float(std::log(5.f) / std::log(2.f)));      // expcomp = log(5) / log(2), L939 -> L293
    if (cudaMemsetAsync(mads, 0, 9 * 64 * sizeof(float), st) != cudaSuccess) return done(ctx->fail(ART_HP_ERR_CUDA, "memset failed"));

    const int begW = 50, begH = 50;
    const int coordW[3] = {begW, widIm / 2 - crW / 2, widIm - crW - begW};
    const int coordH[3] = {begH, heiIm / 2 - crH / 2, heiIm - crH - begH};
    AutoStatsArgs hc[9];
    if (!ctx->ev_auto) { if (cudaEventCreateWithFlags(&ctx->ev_auto, cudaEventDisableTiming) != cudaSuccess) return done(ctx->fail(ART_HP_ERR_CUDA, "event")); }
    for (int wcr = 0; wcr < 3 && !rc; ++wcr)
        for (int hcr = 0; hcr < 3 && !rc; ++hcr) {
            const int k = hcr * 3 + wcr;
            const size_t org = (size_t)coordH[hcr] * ip + coordW[wcr];
            AutoPrepArgs a{};
            a.r = r + org; a.g = g + org; a.b = b + org; a.ip = ip; a.W = crW; a.H = crH; a.wid = wid;
            for (int i = 0; i < 3; ++i) a.mul[i] = mul[i];
            a.do_clip = doClip; a.do_mat = cam2work != nullptr;
            for (int i = 0; i < 9; ++i) { a.mat[i] = cam2work ? cam2work[i] : 0.0; a.wp[i] = static_cast<float>(wprof[i]); }
            a.cachef = cachef; a.cachefy = cachefy; a.gamcurve = gamcurve; a.gam = gam; a.gamthresh = gamthresh; a.gamslope = gamslope; a.gain = gain;
            a.la = la; a.lb = lb;
            a.nvc = maps + (size_t)(3 * k) * round_up(nh, 64); a.nvh = a.nvc + round_up(nh, 64); a.nvl = a.nvh + round_up(nh, 64);
            art_prof_begin(ctx, "k_auto_prep");
            k_auto_prep<<<dim3((crW + 255) / 256, std::min(crH, 148 * 4)), 256, 0, st>>>(a);
            art_prof_end(ctx);
            ctx->launches++;
            art_hp_wavelet *adec = nullptr, *bdec = nullptr;
            if ((rc = art_hp_wavelet_decompose_dev(ctx, la, crW, crW, crH, levwav, 1, &adec))) break;
            if (art_hp_wavelet_maxlevel(adec) < levwav) { art_hp_wavelet_destroy(adec); rc = ctx->fail(ART_HP_ERR_UNSUPPORTED, "crop %dx%d is too small for %d wavelet levels", crW, crH, levwav); break; }
            int Wab = 0, Hab = 0, stride = 0;
            art_hp_wavelet_level_dims(adec, 1, &Wab, &Hab, &stride);
            hc[k] = AutoStatsArgs{a.nvc, a.nvh, a.nvl, wid, Wab, Hab, sums + 16 * k, reinterpret_cast<int*>(sums + 16 * k + 8)};
            rc = art_hp_wavelet_mad_dev(ctx, adec, mads + 64 * k);
            art_hp_wavelet_destroy(adec);
            if (rc) break;
            if ((rc = art_hp_wavelet_decompose_dev(ctx, lb, crW, crW, crH, levwav, 1, &bdec))) break;
            rc = art_hp_wavelet_mad_dev(ctx, bdec, mads + 64 * k + 24);
            art_hp_wavelet_destroy(bdec);
        }
    if (rc) return done(rc);
    // the nine serial statistics run side by side, one CTA each
    if (cudaMemcpyAsync(d_crops, hc, sizeof hc, cudaMemcpyHostToDevice, st) != cudaSuccess) return done(ctx->fail(ART_HP_ERR_CUDA, "upload failed"));
    art_prof_begin(ctx, "k_auto_stats");
    k_auto_stats<<<9, AS_THREADS, 0, st>>>(d_crops);
    art_prof_end(ctx);
    ctx->launches++;
    float hm[9 * 64], hs[9 * 16];
    if (cudaMemcpyAsync(hm, mads, sizeof hm, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaMemcpyAsync(hs, sums, sizeof hs, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return done(ctx->fail(ART_HP_ERR_CUDA, "automatic chroma: download failed: %s", cudaGetErrorString(cudaGetLastError())));
    art_pool_free(ctx, blk);

    // ShrinkAll_info's scalar part per crop, then calcautodn_info and the nine-crop combination
    const float reduc = aggressive ? static_cast<float>(0.9) : 1.f;
    float ch_M[9], max_r[9], max_b[9], min_r[9], min_b[9], lumL[9], chromC[9], ry[9], sk[9], pcsk[9], delta[9];
    int Nb[9];
    for (int k = 0; k < 9; ++k) {
        const float* ma = hm + 64 * k; const float* mb = ma + 24;
        const float* s = hs + 16 * k; const int* sn = reinterpret_cast<const int*>(s + 8);
        float chromina = 0.f, sigma = 0.f, lumema = 0.f, sigma_L = 0.f, redyel = 0.f, skinc = 0.f, nsknc = 0.f;
        const int nc = sn[0], nry = sn[1], nsk = sn[2], nL = nc;
        if (nc > 0) { chromina = s[0] / nc; sigma = std::sqrt(s[1] / nc); nsknc = (float)nsk / (float)nc; } else nsknc = (float)nsk;
        if (nL > 0) { lumema = s[2] / nL; sigma_L = std::sqrt(s[3] / nL); }
        if (nry > 0) redyel = s[4] / nry;
        if (nsk > 0) skinc = s[5] / nsk;
        float chau = 0.f, chred = 0.f, chblue = 0.f, maxchred = 0.f, maxchblue = 0.f, minchred = 100000000.f, minchblue = 100000000.f;
        float chaut = 0.f, redaut = 0.f, blueaut = 0.f, maxredaut = 0.f, maxblueaut = 0.f, minredaut = 0.f, minblueaut = 0.f;
        int nb = 0;
        for (int lvl = 0; lvl < levwav; ++lvl)
            for (int dir = 0; dir < 3; ++dir) {
                const float mada = ma[3 * lvl + dir], madb = mb[3 * lvl + dir];
                chred += mada;
                if (mada > maxchred) maxchred = mada;
                if (mada < minchred) minchred = mada;
                maxredaut = std::sqrt(reduc * maxchred);
                minredaut = std::sqrt(reduc * minchred);
                chblue += madb;
                if (madb > maxchblue) maxchblue = madb;
                if (madb < minchblue) minchblue = madb;
                maxblueaut = std::sqrt(reduc * maxchblue);
                minblueaut = std::sqrt(reduc * minchblue);
                chau += (mada + madb);
                ++nb;
                chaut = std::sqrt(reduc * chau / (nb + nb));
                redaut = std::sqrt(reduc * chred / nb);
                blueaut = std::sqrt(reduc * chblue / nb);
            }
        Nb[k] = nb; ch_M[k] = 1.0f * chaut; max_r[k] = 1.0f * maxredaut; max_b[k] = 1.0f * maxblueaut; min_r[k] = 1.0f * minredaut; min_b[k] = 1.0f * minblueaut;
        lumL[k] = lumema; chromC[k] = chromina; ry[k] = redyel; sk[k] = skinc; pcsk[k] = nsknc;
        if (stats_out) {
            float* o = stats_out + 15 * k;
            o[0] = chaut; o[1] = (float)nb; o[2] = redaut; o[3] = blueaut; o[4] = maxredaut; o[5] = maxblueaut; o[6] = minredaut; o[7] = minblueaut;
            o[8] = chromina; o[9] = sigma; o[10] = lumema; o[11] = sigma_L; o[12] = redyel; o[13] = skinc; o[14] = nsknc;
        }
    }
    const float autoNR = 10, autoNRmax = 40, lowdenoise = 1.f, adjustr = 1.f, multip = 1.f;      // isRAW
    const int levaut = 0, mode = 1, lissage = 0;
    float Max_R[9] = {0}, Max_B[9] = {0}, Min_R[9], Min_B[9];
    for (int k = 0; k < 9; ++k) {
        const float maxmax = max_r[k] < max_b[k] ? max_b[k] : max_r[k];
        calcautodn_info_host(aggressive != 0, ch_M[k], delta[k], Nb[k], levaut, maxmax, lumL[k], chromC[k], mode, lissage, ry[k], sk[k], pcsk[k]);
    }
    for (int k = 0; k < 9; ++k) {
        if (max_r[k] > max_b[k]) {
            Max_R[k] = (delta[k]) / ((autoNRmax * multip * adjustr * lowdenoise) / 2.f);
            Min_B[k] = -(ch_M[k] - min_b[k]) / (autoNRmax * multip * adjustr * lowdenoise);
            Max_B[k] = 0.f; Min_R[k] = 0.f;
        } else {
            Max_B[k] = (delta[k]) / ((autoNRmax * multip * adjustr * lowdenoise) / 2.f);
            Min_R[k] = -(ch_M[k] - min_r[k]) / (autoNRmax * multip * adjustr * lowdenoise);
            Min_B[k] = 0.f; Max_R[k] = 0.f;
        }
    }
    float chM = 0.f, MaxR = 0.f, MaxB = 0.f, MinR = 100000000000.f, MinB = 100000000000.f, maxr, maxb;
    float MaxRMoy = 0.f, MaxBMoy = 0.f, MinRMoy = 0.f, MinBMoy = 0.f;
    for (int k = 0; k < 9; ++k) {
        chM += ch_M[k]; MaxBMoy += Max_B[k]; MaxRMoy += Max_R[k]; MinRMoy += Min_R[k]; MinBMoy += Min_B[k];
        if (Max_R[k] > MaxR) MaxR = Max_R[k];
        if (Max_B[k] > MaxB) MaxB = Max_B[k];
        if (Min_R[k] < MinR) MinR = Min_R[k];
        if (Min_B[k] < MinB) MinB = Min_B[k];
    }
    chM /= 9; MaxBMoy /= 9; MaxRMoy /= 9; MinBMoy /= 9; MinRMoy /= 9;
    if (MaxR > MaxB) { maxr = MaxRMoy + (MaxR - MaxRMoy) * 0.66f; maxb = MinBMoy + (MinB - MinBMoy) * 0.66f; }
    else { maxb = MaxBMoy + (MaxB - MaxBMoy) * 0.66f; maxr = MinRMoy + (MinR - MinRMoy) * 0.66f; }
    out3[0] = chM / (autoNR * multip * adjustr);
    out3[1] = maxr;
    out3[2] = maxb;
    return ART_HP_OK;
}
