// The whole-frame path of simpleprocess.cc's normal pipeline, device resident end to end:
//   RawImageSource::demosaic            (rtengine/simpleprocess.cc L215-222)
//   RawImageSource::getImage gains + convertColorSpace matrix branch   (L255-259 region; rawimagesource.cc L943-1025, L3184-3213)
//   ImProcFunctions::denoise            (rtengine/ipdenoise.cc L1096-1189: half-resolution calclum L1119-1131,
//                                        adjust_params L35-63, RGB_denoise, optional NL-means on Y)
//   ImProcFunctions::process STAGE_0 -> dynamicRangeCompression        (rtengine/improcfun.cc L580-583)
// One H2D of the CFA plane, one D2H of the three result planes; everything between stays in HBM on one stream.
#include "ctx.h"

#include <cmath>

namespace {

// calclum of ipdenoise.cc L1119-1131: every second sample of every second row, then the camera->working matrix
// (convertColorSpace -> colorSpaceConversion_ matrix branch: double coefficients, float samples, rounded once)
__global__ void k_calclum(const float* __restrict__ r, const float* __restrict__ g, const float* __restrict__ b, size_t ip, int W, int H,
                          float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb, size_t cp, int do_mat,
                          double m0, double m1, double m2, double m3, double m4, double m5, double m6, double m7, double m8)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
    if (x >= w2 || y >= h2) return;
    const size_t i = (size_t)(2 * y) * ip + 2 * x;
    float vr = r[i], vg = g[i], vb = b[i];
    if (do_mat) {
        const double dr = vr, dg = vg, db = vb;
        vr = (float)(m0 * dr + m1 * dg + m2 * db);
        vg = (float)(m3 * dr + m4 * dg + m5 * db);
        vb = (float)(m6 * dr + m7 * dg + m8 * db);
    }
    const size_t o = (size_t)y * cp + x;
    cr[o] = vr; cg[o] = vg; cb[o] = vb;
}

// Imagefloat::setMode(YUV) / setMode(RGB) around NLMeans (ipdenoise.cc L1173-1177): Color::rgb2yuv / yuv2rgb with the
// working-space matrix as float (imagefloat.cc rgb_to_yuv / yuv_to_rgb); Y lives in the g plane, u in b, v in r
__global__ void k_rgb2yuv(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H, float w0, float w1, float w2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * ip + x;
    const float R = r[i], G = g[i], B = b[i];
    const float Y = R * w0 + G * w1 + B * w2;
    g[i] = Y; b[i] = Y - B; r[i] = R - Y;
}
__global__ void k_yuv2rgb(float* __restrict__ r, float* __restrict__ g, float* __restrict__ b, size_t ip, int W, int H, float w0, float w1, float w2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t i = (size_t)y * ip + x;
    const float Y = g[i], u = b[i], v = r[i];
    const float B = Y - u, R = v + Y;
    const float G = (Y - R * w0 - B * w2) / w1;
    r[i] = R; g[i] = G; b[i] = B;
}

double adj_c(double x, double f)
{   // the lambda of adjust_params, ipdenoise.cc L41-47: SGN, LIM01, intp
    const int s = (x > 0) - (x < 0);
    double y = std::fabs(x) / 100.0;
    y = y < 0.0 ? 0.0 : y > 1.0 ? 1.0 : y;
    return s * (y * (y * f) + (1.0 - y) * y) * 100.0;
}

}  // namespace

void art_adjust_denoise_params(art_hp_denoise_params* p)
{   // adjust_params, ipdenoise.cc L35-63
    const double scale = p->scale;
    if (scale <= 1.0) return;
    const double scale_factor = 1.0 / scale;
    const double noise_factor_c = std::pow(scale_factor, 0.46);
    const double noise_factor_l = std::pow(scale_factor, 0.62) * scale_factor;
    p->luminance = adj_c(p->luminance, noise_factor_l);
    p->luminanceDetail *= (1.0 + std::pow(1.0 - scale_factor, 2.2));
    p->chrominance = adj_c(p->chrominance, noise_factor_c);
    p->chrominanceRedGreen = adj_c(p->chrominanceRedGreen, noise_factor_c);
    p->chrominanceBlueYellow = adj_c(p->chrominanceBlueYellow, noise_factor_c);
}

int art_calclum_dev(art_hp_ctx* ctx, const float* r, const float* g, const float* b, size_t ip, int W, int H,
                    float* cr, float* cg, float* cb, size_t cp, const double* mat)
{
    const dim3 blk(32, 8), grid(((W + 1) / 2 + 31) / 32, ((H + 1) / 2 + 7) / 8);
    static const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double* m = mat ? mat : ident;
    art_prof_begin(ctx, "k_calclum");
    k_calclum<<<grid, blk, 0, ctx->stream>>>(r, g, b, ip, W, H, cr, cg, cb, cp, mat != nullptr, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

// ImProcFunctions::denoise on device planes (ipdenoise.cc L1096-1189).  ecomp = params->exposure.enabled ? expcomp : 0: when
// positive the stage runs between expcomp(+ecomp) and expcomp(-ecomp) (L1155-1163, L1181-1184; ExposureParams() has black 0).
// guidedChromaRadius / nlStrength are 0 unless smoothingEnabled (L1170-1178).
int art_denoise_stage_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_denoise_params* dn,
                          int nlStrength, int nlDetail, int guidedChromaRadius, double ecomp, const double* cam2work, const double* wprof)
{
    art_hp_denoise_params P = *dn;
    art_adjust_denoise_params(&P);
    auto bracket = [&](double ev) -> int {       // expcomp, ipexposure.cc L29-73: exp_scale = pow(2.f, expcomp) evaluated in double, stored to float
        art_hp_chain_params e{};
        e.exposure_enabled = 1; e.exp_scale = (float)std::pow(2.0, ev); e.black = 0.f;
        return art_chain_dev(ctx, W, H, r, g, b, ip, &e);
    };
    const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
    const size_t cp = round_up((size_t)w2, 32);
    float* cl[3] = {nullptr, nullptr, nullptr};
    int rc;
    if (P.noiseCCurve) {
        if ((rc = art_reserve(ctx, ctx->d_small2, 3 * cp * (size_t)h2 * sizeof(float)))) return rc;
        for (int c = 0; c < 3; ++c) cl[c] = (float*)ctx->d_small2.p + (size_t)c * cp * h2;
        if ((rc = art_calclum_dev(ctx, r, g, b, ip, W, H, cl[0], cl[1], cl[2], cp, cam2work))) return rc;
    }
    if (ecomp > 0 && (rc = bracket(ecomp))) return rc;
    if ((rc = art_rgb_denoise_dev(ctx, r, g, b, ip, W, H, &P, wprof, cl[0], cl[1], cl[2], cp, nullptr))) return rc;
    if (guidedChromaRadius && (rc = art_guided_smoothing_dev(ctx, r, g, b, ip, W, H, wprof, guidedChromaRadius, dn->scale > 0 ? dn->scale : 1.0))) return rc;
    if (nlStrength) {
        const dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8);
        const float w0 = (float)wprof[3], w1 = (float)wprof[4], w2f = (float)wprof[5];
        art_prof_begin(ctx, "k_rgb2yuv");
        k_rgb2yuv<<<grid, blk, 0, ctx->stream>>>(r, g, b, ip, W, H, w0, w1, w2f);
        art_prof_end(ctx);
        if ((rc = art_nlmeans_dev(ctx, g, ip, W, H, 65535.f, nlStrength, nlDetail, (float)P.scale))) return rc;
        art_prof_begin(ctx, "k_yuv2rgb");
        k_yuv2rgb<<<grid, blk, 0, ctx->stream>>>(r, g, b, ip, W, H, w0, w1, w2f);
        art_prof_end(ctx);
        ctx->launches += 2;
    }
    if (ecomp > 0 && (rc = bracket(-ecomp))) return rc;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}

int art_develop_geometry2(const art_hp_develop_params* p, int W, int H, art_dev_geo* g)
{   // RawImageSource::border: 4 for Bayer sensors (the caller's value), 7 for X-Trans (rawimagesource.cc L1345-1352, computeFullSize L1163-1175)
    const bool xt = p->method == ART_HP_XTRANS_3PASS || p->method == ART_HP_XTRANS_1PASS;
    const int bd = p->full_frame ? 0 : (xt ? 7 : (p->border > 0 ? p->border : 0));
    const int rot = p->tran & 3;
    const bool turned = rot == 1 || rot == 3;
    // the transformed, border-cropped full image (getFullSize)
    const int fw = (turned ? H : W) - 2 * bd, fh = (turned ? W : H) - 2 * bd;
    g->bd = bd; g->window = p->pp_skip > 0;
    const int skip = g->window ? p->pp_skip : 1;
    const int px = g->window ? p->pp_x : 0, py = g->window ? p->pp_y : 0;
    const int pw = g->window ? p->pp_width : fw, ph = g->window ? p->pp_height : fh;
    g->skip = skip;
    if (px < 0 || py < 0 || pw < 1 || ph < 1 || px + pw > fw || py + ph > fh) { g->Wo = g->Ho = g->iw = g->ih = 0; g->sx1 = g->sy1 = 0; return ART_HP_ERR_INVALID; }
    // RawImageSource::getSize(pp, w, h), L1199-1203: the image getImage fills
    g->Wo = pw / skip + (pw % skip > 0);
    g->Ho = ph / skip + (ph % skip > 0);
    // RawImageSource::transformRect for a standard CCD, L664-751 (the window lies inside the image, so no clamp of its size fires)
    const int sw = turned ? H : W, sh = turned ? W : H;
    int ppx = px + bd, ppy = py + bd;
    if (p->tran & 8) ppx = std::max(sw - (px + bd) - pw, 0);
    if (p->tran & 4) ppy = std::max(sh - (py + bd) - ph, 0);
    int sx1 = ppx, sy1 = ppy;
    if (rot == 2) { sx1 = std::max(W - ppx - pw, 0); sy1 = std::max(H - ppy - ph, 0); }
    else if (rot == 1) { sx1 = ppy; sy1 = std::max(H - ppx - pw, 0); }
    else if (rot == 3) { sx1 = std::max(W - ppy - ph, 0); sy1 = ppx; }
    g->sx1 = sx1; g->sy1 = sy1;
    // imwidth / imheight: transformRect's size clamped to the image's (L855-861) = the image's, in the source orientation
    g->iw = turned ? g->Ho : g->Wo;
    g->ih = turned ? g->Wo : g->Ho;
    return ART_HP_OK;
}

void art_develop_geometry(const art_hp_develop_params* p, int W, int H, int* b, int* Wo, int* Ho)
{
    art_dev_geo g;
    art_develop_geometry2(p, W, H, &g);
    *b = g.bd; *Wo = g.Wo; *Ho = g.Ho;
}

int art_develop_dev(art_hp_ctx* ctx, const art_hp_develop_params* p, int Wr, int Hr, const float* raw, size_t rp,
                    float* r, float* g, float* b, size_t op)
{
    int rc;
    art_dev_geo geo;
    if (art_develop_geometry2(p, Wr, Hr, &geo)) return ctx->fail(ART_HP_ERR_INVALID, "the PreviewProps window (%d, %d, %d x %d, skip %d) leaves the developed frame", p->pp_x, p->pp_y, p->pp_width, p->pp_height, p->pp_skip);
    const int bd = geo.bd, W = geo.Wo, H = geo.Ho;
    const bool oop = bd || p->tran || p->hr_blend || geo.window;        // getImage's stage runs out of place: crop, coarse transform, highlight reconstruction, preview window
    // then the demosaicer writes context-owned planes and getImage's stage crops / turns them into the caller's planes
    float* dm[3] = {r, g, b};
    size_t dmp = op;
    if (oop) {
        dmp = round_up((size_t)Wr, 32);
        for (int c = 0; c < 3; ++c) {
            if ((rc = art_reserve(ctx, ctx->d_dm[c], dmp * (size_t)Hr * sizeof(float)))) return rc;
            dm[c] = (float*)ctx->d_dm[c].p;
        }
    }
    if (p->method == ART_HP_XTRANS_3PASS || p->method == ART_HP_XTRANS_1PASS)
        rc = art_xtrans_dev(ctx, p->method == ART_HP_XTRANS_3PASS ? 3 : 1, p->method == ART_HP_XTRANS_3PASS, Wr, Hr, p->xtrans, p->rgb_cam, raw, rp, dm[0], dm[1], dm[2], dmp);
    else if (p->method == ART_HP_BAYER_AMAZE) rc = art_amaze_dev(ctx, Wr, Hr, p->filters, raw, rp, dm[0], dm[1], dm[2], dmp, p->initialGain, p->border, 0, Hr);
    else rc = art_rcd_dev(ctx, Wr, Hr, p->filters, raw, rp, dm[0], dm[1], dm[2], dmp, 0, Hr);      // ends with its own border_interpolate2(9)
    if (rc) return rc;
    // simpleprocess.cc L254-256: denoiseComputeParams measures the demosaiced frame (camera space, through getImage per crop) before
    // getImage converts it; the estimate replaces the chroma sliders of this frame's denoise parameters
    art_hp_denoise_params dn_resolved;
    const art_hp_denoise_params* dn = p->denoise;
    if (dn && dn->chrominanceMethod == 1) {
        float est[3];
        const size_t off0 = (size_t)bd * dmp + bd;
        if (p->tran) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "the automatic chroma estimator with a coarse transform: its crops are cut from the turned frame (pass the estimate, chrominanceMethod 0)");
        if (geo.window) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "the automatic chroma estimator on a preview window: it measures crops of the whole frame (run it once on the frame and pass the estimate, chrominanceMethod 0)");
        if ((rc = art_denoise_auto_chroma_dev(ctx, dm[0] + off0, dm[1] + off0, dm[2] + off0, dmp, W, H, p->mul, p->doClip, p->cam2work, p->wprof,
                                              dn->gamma, dn->aggressive, est, nullptr))) return rc;
        dn_resolved = *dn;
        const double f = dn->chrominanceAutoFactor > 0 ? dn->chrominanceAutoFactor : 1.0;
        dn_resolved.chrominance = est[0] * f;                  // float * double, as `store.chrominance * dnparams.chrominanceAutoFactor` (L1064-1066)
        dn_resolved.chrominanceRedGreen = est[1] * f;
        dn_resolved.chrominanceBlueYellow = est[2] * f;
        dn_resolved.chrominanceMethod = 0;
        dn = &dn_resolved;
    }
    if (oop) {
        // the source lines are geo.iw x geo.ih from (sx1, sy1); W x H is the turned image.  At skip 1 the planes are passed at that origin,
        // at skip > 1 un-offset (the box sum clamps its own origin against the frame, L945 / L949)
        const size_t off = geo.skip > 1 ? 0 : (size_t)geo.sy1 * dmp + geo.sx1;
        rc = art_scale_convert_crop_dev(ctx, geo.iw, geo.ih, dm[0] + off, dm[1] + off, dm[2] + off, dmp, r, g, b, op, p->mul, p->doClip, p->cam2work,
                                        p->tran, p->hr_blend, p->hlmax, geo.skip, geo.sx1, geo.sy1, Wr, Hr);
    } else rc = art_scale_convert_dev(ctx, W, H, r, g, b, op, p->mul, p->doClip, p->cam2work);
    if (rc) return rc;
    if (dn) {
        if ((rc = art_denoise_stage_dev(ctx, r, g, b, op, W, H, dn, p->nlStrength, p->nlDetail, p->guidedChromaRadius, p->denoise_expcomp, p->cam2work, p->wprof))) return rc;
    }
    if (p->fattal_enabled) {
        if ((rc = art_fattal_dev(ctx, r, g, b, op, W, H, p->fattal_threshold, p->fattal_amount, p->fattal_satcontrol, p->wprof))) return rc;
    }
    // ipf.process(STAGE_1 .. STAGE_3), improcfun.cc L584-625: exposure | sharpening | saturationVibrance, toneCurve, rgbCurves, labAdjustments
    const bool sharpen = p->sharpen && p->sharpen->amount >= 1;
    if (p->chain && sharpen && p->chain->exposure_enabled) {
        art_hp_chain_params e{};
        e.exposure_enabled = 1; e.exp_scale = p->chain->exp_scale; e.black = p->chain->black;
        if ((rc = art_chain_dev(ctx, W, H, r, g, b, op, &e))) return rc;
    }
    if (sharpen) {
        if (!p->wprof) return ctx->fail(ART_HP_ERR_INVALID, "sharpening needs wprof");
        if ((rc = art_usm_dev(ctx, r, g, b, op, W, H, p->sharpen, p->wprof))) return rc;
    }
    if (p->chain) {
        art_hp_chain_params c = *p->chain;
        if (sharpen) c.exposure_enabled = 0;
        if ((rc = art_chain_dev(ctx, W, H, r, g, b, op, &c))) return rc;
    }
    return ART_HP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// One frame across GPUs: a rank's row band (include/art_hotpath.h, ABI version 3)
// ---------------------------------------------------------------------------------------------------------------------------
int art_band_plan(const art_hp_develop_params* p, int Wr, int Hr, int own_begin, int own_end, int halo, art_hp_band_plan* out)
{
    int bd, W, H, P = 0, O = 0;
    art_develop_geometry(p, Wr, Hr, &bd, &W, &H);
    if (art_hp_band_align(p->method, &P, &O)) return ART_HP_ERR_UNSUPPORTED;
    if (own_begin < 0 || own_end > H || own_begin >= own_end || (own_begin & 1) || ((own_end & 1) && own_end != H) || halo < 150) return ART_HP_ERR_INVALID;
    art_hp_band_plan b{};
    b.own_begin = own_begin; b.own_end = own_end;
    b.band_begin = own_begin - halo > 0 ? (own_begin - halo) / 50 * 50 : 0;       // 2: decimated wavelet level; 25: the DCT block grid
    b.band_end = std::min(H, own_end + halo);
    // demosaiced rows [band_begin + bd, band_end + bd) of the raw frame, widened to the method's tile grid
    const int lo = b.band_begin + bd, hi = b.band_end + bd;
    b.dm_begin = lo <= O + P ? 0 : O + (lo - O) / P * P;
    b.dm_end = O + (hi - O + P - 1) / P * P;
    if (b.dm_end >= Hr || b.dm_end <= O) b.dm_end = Hr;
    const int dh = art_hp_band_halo(p->method);
    b.raw_begin = std::max(0, b.dm_begin - dh);
    b.raw_end = std::min(Hr, b.dm_end + dh);
    if (b.dm_begin == 0) b.raw_end = std::max(b.raw_end, std::min(Hr, 33));         // the mirrored rows at the frame's top / bottom (dist.row_bands)
    if (b.dm_end == Hr) b.raw_begin = std::min(b.raw_begin, std::max(0, Hr - 17));
    *out = b;
    return ART_HP_OK;
}

int art_develop_band_dev(art_hp_ctx* ctx, const art_hp_develop_params* p, int Wr, int Hr, const float* raw, size_t rp,
                         float* r, float* g, float* b, size_t op, const art_hp_band_plan* plan)
{
    int rc, bd, W, H;
    art_develop_geometry(p, Wr, Hr, &bd, &W, &H);
    if (p->method != ART_HP_BAYER_AMAZE && p->method != ART_HP_BAYER_RCD) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: Bayer methods only");
    if (p->tran) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: the coarse transform is not split (rows of the turned frame are columns of the raw frame)");
    if (p->pp_skip > 0) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: a preview window is one GPU's work");
    if (p->fattal_enabled) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: Fattal's Poisson solve is a transform of the whole frame");
    if (p->denoise && p->denoise->chrominanceMethod == 1) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: the automatic chroma estimator measures crops of the whole frame");
    if (p->nlStrength) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: NL-means is not split");
    if (p->denoise && p->denoise->aggressive) return ctx->fail(ART_HP_ERR_UNSUPPORTED, "row bands: the aggressive (twice shrunk) mode is not split");
    {
        const art_hp_band_plan& q = *plan;
        int P = 0, O = 0;
        art_hp_band_align(p->method, &P, &O);
        auto aligned = [&](int row, int edge) { return row == edge || (row > O && row < Hr && (row - O) % P == 0); };
        const bool ok = q.own_begin >= 0 && q.own_begin < q.own_end && q.own_end <= H && !(q.own_begin & 1) && (!(q.own_end & 1) || q.own_end == H) &&
                        q.band_begin >= 0 && q.band_begin % 50 == 0 && q.band_begin <= q.own_begin && q.band_end >= q.own_end && q.band_end <= H &&
                        (q.band_begin == 0 || q.own_begin - q.band_begin >= 150) && (q.band_end == H || q.band_end - q.own_end >= 150) &&
                        aligned(q.dm_begin, 0) && aligned(q.dm_end, Hr) && q.dm_begin <= q.band_begin + bd && q.dm_end >= q.band_end + bd;
        if (!ok) return ctx->fail(ART_HP_ERR_INVALID, "inconsistent band plan: own [%d,%d) band [%d,%d) demosaic [%d,%d) of %d rows", q.own_begin, q.own_end,
                                  q.band_begin, q.band_end, q.dm_begin, q.dm_end, H);
    }
    // demosaic on the frame's tile grid into context-owned full-frame planes (only the band's rows are touched)
    const size_t dmp = round_up((size_t)Wr, 32);
    float* dm[3];
    for (int c = 0; c < 3; ++c) {
        if ((rc = art_reserve(ctx, ctx->d_dm[c], dmp * (size_t)Hr * sizeof(float)))) return rc;
        dm[c] = (float*)ctx->d_dm[c].p;
    }
    if (p->method == ART_HP_BAYER_AMAZE) rc = art_amaze_dev(ctx, Wr, Hr, p->filters, raw, rp, dm[0], dm[1], dm[2], dmp, p->initialGain, p->border, plan->dm_begin, plan->dm_end);
    else rc = art_rcd_dev(ctx, Wr, Hr, p->filters, raw, rp, dm[0], dm[1], dm[2], dmp, plan->dm_begin, plan->dm_end);
    if (rc) return rc;
    // from here on the band is a frame of Hb rows whose first row is row band_begin of the developed frame
    const int y0 = plan->band_begin, Hb = plan->band_end - plan->band_begin;
    const size_t off = (size_t)(y0 + bd) * dmp + bd, oo = (size_t)y0 * op;
    float *br = r + oo, *bg = g + oo, *bb = b + oo;
    if ((rc = art_scale_convert_crop_dev(ctx, W, Hb, dm[0] + off, dm[1] + off, dm[2] + off, dmp, br, bg, bb, op, p->mul, p->doClip, p->cam2work,
                                         0, p->hr_blend, p->hlmax))) return rc;
    ctx->band.active = true;
    ctx->band.own0 = plan->own_begin - y0; ctx->band.own1 = plan->own_end - y0; ctx->band.H_full = H;
    rc = ART_HP_OK;
    if (p->denoise)
        rc = art_denoise_stage_dev(ctx, br, bg, bb, op, W, Hb, p->denoise, 0, 0, p->guidedChromaRadius, p->denoise_expcomp, p->cam2work, p->wprof);
    ctx->band.active = false;
    if (rc) return rc;
    const bool sharpen = p->sharpen && p->sharpen->amount >= 1;
    if (p->chain && sharpen && p->chain->exposure_enabled) {
        art_hp_chain_params e{};
        e.exposure_enabled = 1; e.exp_scale = p->chain->exp_scale; e.black = p->chain->black;
        if ((rc = art_chain_dev(ctx, W, Hb, br, bg, bb, op, &e))) return rc;
    }
    if (sharpen) {
        if (!p->wprof) return ctx->fail(ART_HP_ERR_INVALID, "sharpening needs wprof");
        art_hp_sharpen_params sp = *p->sharpen;
        if ((rc = art_usm_dev(ctx, br, bg, bb, op, W, Hb, &sp, p->wprof))) return rc;
    }
    if (p->chain) {
        art_hp_chain_params c = *p->chain;
        if (sharpen) c.exposure_enabled = 0;
        if ((rc = art_chain_dev(ctx, W, Hb, br, bg, bb, op, &c))) return rc;
    }
    return ART_HP_OK;
}
