/*
 * art_hotpath.h -- C-ABI of the B200-native raw-development hot path.
 *
 * This is the drop-in boundary for the reference's rtengine hot path
 * (artpixls/ART).  The reference has no plugin/FFI layer for this path: the
 * call sites are C++ member functions reached from rtengine/simpleprocess.cc
 * (ImageProcessor, L112-490).  Each entry point below names the reference
 * member it replaces (file:line under /root/reference) and keeps that member's
 * argument meaning: caller-owned buffers, `float* const*` row-pointer tables
 * exactly as rtengine::array2D<float> (rtengine/array2D.h L74-296) and
 * PlanarPtr hold them, in-place mutation where the reference mutates in place.
 * INTEGRATION.md shows the <=10-line patch per call site.
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns an
 *     art_hp_status (0 = OK) and art_hp_last_error() gives the text;
 *   - a context owns one GPU (one process per GPU, one context per process
 *     is the intended deployment), its stream, device scratch and pinned
 *     staging; a context is thread-compatible, not thread-safe (the
 *     reference serialises the same way: one ImProcFunctions per image);
 *   - `*_dev` variants take device pointers + a pitch in floats and do no
 *     host<->device copies: they are what a resident pipeline chains;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     returns ART_HP_ERR_NO_DEVICE.
 */
#ifndef ART_HOTPATH_H
#define ART_HOTPATH_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ART_HP_ABI_VERSION 5

typedef enum art_hp_status {
    ART_HP_OK = 0,
    ART_HP_ERR_INVALID = 1,      /* bad argument (null pointer, non-RGB CFA, size out of range) */
    ART_HP_ERR_NO_DEVICE = 2,    /* no usable CUDA device */
    ART_HP_ERR_CUDA = 3,         /* a CUDA runtime call failed; see art_hp_last_error */
    ART_HP_ERR_NOMEM = 4,
    ART_HP_ERR_UNSUPPORTED = 5
} art_hp_status;

/* RAWParams::BayerSensor::Method subset (reference rtengine/procparams.h; dispatch in
 * rtengine/rawimagesource.cc L1872-1924) */
typedef enum art_hp_bayer_method {
    ART_HP_BAYER_AMAZE = 0,      /* RawImageSource::amaze_demosaic_RT, rtengine/amaze_demosaic_RT.cc L41-1595 */
    ART_HP_BAYER_RCD = 1,        /* RawImageSource::rcd_demosaic,      rtengine/rcd_demosaic.cc L51-347   */
    /* art_hp_develop only: RawImageSource::xtrans_interpolate(3, true) / (1, false), rtengine/rawimagesource.cc L1920-1924 */
    ART_HP_XTRANS_3PASS = 2,
    ART_HP_XTRANS_1PASS = 3
} art_hp_bayer_method;

typedef struct art_hp_ctx art_hp_ctx;

/* ---- context ---------------------------------------------------------- */
int  art_hp_abi_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int  art_hp_device_count(void);
/* create a context on CUDA device `device_id` */
int  art_hp_create(art_hp_ctx** out, int device_id);
void art_hp_destroy(art_hp_ctx* ctx);
const char* art_hp_last_error(const art_hp_ctx* ctx);
/* run on a caller-owned cudaStream_t (NULL restores the context's own stream) */
int  art_hp_set_stream(art_hp_ctx* ctx, void* cuda_stream);
void* art_hp_get_stream(art_hp_ctx* ctx);
/* block until everything queued on the context's stream is done */
int  art_hp_sync(art_hp_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
unsigned long long art_hp_launch_count(const art_hp_ctx* ctx);

/* Optional per-kernel timing (the reference's BENCHFUN/StopWatch, rtengine/StopWatch.h L33-75, is the
 * analogue): when enabled every kernel launch is bracketed by CUDA events on the context's stream.
 * art_hp_profile_collect() synchronises, folds the spans into per-kernel totals and returns how many
 * distinct kernels were seen; art_hp_profile_entry() reads one (name is owned by the context). */
int  art_hp_profile_enable(art_hp_ctx* ctx, int on);
int  art_hp_profile_collect(art_hp_ctx* ctx);
int  art_hp_profile_entry(art_hp_ctx* ctx, int index, const char** name, double* total_ms, int* calls);

/* pinned host memory for callers that want zero-staging transfers
 * (what an AlignedBuffer, rtengine/alignedbuffer.h, would be backed by) */
void* art_hp_host_alloc(size_t bytes);
void  art_hp_host_free(void* p);

/* ---- demosaic --------------------------------------------------------- */
/*
 * Replaces RawImageSource::amaze_demosaic_RT(0,0,W,H,rawData,red,green,blue)
 * (rtengine/amaze_demosaic_RT.cc L41) and RawImageSource::rcd_demosaic()
 * (rtengine/rcd_demosaic.cc L51), as dispatched by RawImageSource::demosaic
 * (rtengine/rawimagesource.cc L1872-1924).
 *   filters      dcraw CFA descriptor, e.g. 0x94949494 for RGGB
 *                (RawImage::FC, rtengine/rawimage.h L186-189); must describe an RGB 2x2 pattern
 *   rawData      H row pointers, W floats each, values in the scaleColors
 *                domain 0..65535 (rtengine/rawimagesource.cc L2677-2859)
 *   red/green/blue  H row pointers each, W floats per row, fully overwritten
 *   initialGain  RawImageSource::initialGain (AMaZE clip points, amaze_demosaic_RT.cc L53-54)
 *   border       RawImageSource::border (AMaZE: border<4 => 3-px border_interpolate2, L1587-1589)
 * Host pointers; the call returns when the outputs are complete.
 */
int art_hp_demosaic_bayer(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                          const float* const* rawData,
                          float* const* red, float* const* green, float* const* blue,
                          double initialGain, int border);

/* Device-resident form: planes already in HBM, pitch in floats, asynchronous on the
 * context's stream. */
int art_hp_demosaic_bayer_dev(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                              const float* d_raw, size_t raw_pitch,
                              float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                              double initialGain, int border);

/*
 * art_hp_demosaic_xtrans   replaces RawImageSource::xtrans_interpolate(passes, useCieLab) including its closing
 *                      xtransborder_interpolate(passes > 1 ? 8 : 11) (rtengine/xtrans_demosaic.cc L181-969, L122-173, cielab
 *                      L42-116; dispatch rtengine/rawimagesource.cc L1916-1924: "1-pass" = (1, false), "3-pass (best)" =
 *                      (3, true)).  xtrans = RawImage::getXtransMatrix (0 R, 1 G, 2 B), rgb_cam = RawImage::getRgbCam.
 *                      passes must be 1 or 3; 23 <= W <= 16380 (the reference keeps its hexagon offsets in shorts), H >= 23.
 *                      Bit-identical to the reference with its per-thread tile buffer cleared per tile (the stock
 *                      reference reads the previous tile's bytes near the image border and is schedule dependent there).
 */
int art_hp_demosaic_xtrans(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                           float* const* rawData, float* const* red, float* const* green, float* const* blue);
int art_hp_demosaic_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                               const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch);

/* Row-band form for sharding one frame across GPUs: computes only the output
 * rows [row_begin, row_end) of the full W x H frame.  Bands must be cut on the method's reference tile
 * grid so that results are identical to the full-frame call: art_hp_band_align(method) gives the period P
 * and offset O; row_begin must be 0 or O + k*P, row_end must be H or O + k*P (AMaZE: P=128, O=0 --
 * amaze_demosaic_RT.cc L182; RCD: P=176, O=9 -- rcd_demosaic.cc L82-87, L305-316).
 * d_raw/d_red/... are the addresses of ROW 0 of the frame; only rows
 * [row_begin - art_hp_band_halo(method), row_end + halo) of d_raw (clipped to the frame, plus the
 * mirror rows [0,33) / [H-17,H) when the band touches the top / bottom edge) are read and only rows
 * [row_begin,row_end) of the outputs are written, so a rank may back the rest with nothing. */
int art_hp_band_align(int method, int* period, int* offset);
int art_hp_band_halo(int method);
int art_hp_demosaic_bayer_rows_dev(art_hp_ctx* ctx, int method, int W, int H, unsigned filters,
                                   const float* d_raw, size_t raw_pitch,
                                   float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                   double initialGain, int border, int row_begin, int row_end);

/* Replaces RawImageSource::border_interpolate2(W,H,lborders,rawData,red,green,blue)
 * (rtengine/demosaic_algos.cc L200-353). Device-resident. */
int art_hp_border_interpolate2_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, int lborders,
                                   const float* d_raw, size_t raw_pitch,
                                   float* d_red, float* d_green, float* d_blue, size_t out_pitch);

/* ---- raw scaling ---------------------------------------------------------- */
/*
 * Replaces the Bayer branch of RawImageSource::scaleColors (rtengine/rawimagesource.cc L2731-2772):
 * rawData = max(0, rawData - cblacksom[c4]) * scale_mul[c4] in place (c4: 0=R, 1=G on odd rows, 2=B,
 * 3=G on even rows) and chmax[c] = per-colour maximum of the result.  cblacksom / scale_mul are the four
 * numbers the reference computes on the host (L2711-2719); dynamicRowNoiseFilter is not supported.
 */
int art_hp_scale_colors_bayer(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData,
                              const float cblacksom[4], const float scale_mul[4], float chmax[3]);
int art_hp_scale_colors_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch,
                                  const float cblacksom[4], const float scale_mul[4], float chmax[3]);
/* The X-Trans branch (L2795-2826): c = XTRANSFC(row, col) = xtrans[row % 6][col % 6], cblacksom / scale_mul per colour. */
int art_hp_scale_colors_xtrans(art_hp_ctx* ctx, int W, int H, const int xtrans[36], float* const* rawData,
                               const float cblacksom[3], const float scale_mul[3], float chmax[3]);
int art_hp_scale_colors_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int xtrans[36], float* d_raw, size_t pitch,
                                   const float cblacksom[3], const float scale_mul[3], float chmax[3]);

/* ---- Gaussian blur --------------------------------------------------------- */
/*
 * Replaces gaussianBlur(src, dst, W, H, sigma, buffer = nullptr, GAUSS_STANDARD) (rtengine/gauss.h L25,
 * rtengine/gauss.cc L1387-1574): every sigma branch (copy < 0.25, 3x3 / separable 3-tap < 0.6, Young-van
 * Vliet recursive filter in float < 25, in double above).  src == dst (the same row table / the same device
 * pointer) selects the reference's in-place variants.  gausstype: 0 = GAUSS_STANDARD; GAUSS_MULT / GAUSS_DIV
 * and the box-blur `buffer` variant are not on the hot path and return ART_HP_ERR_UNSUPPORTED.
 * W, H >= 4.
 */
int art_hp_gauss(art_hp_ctx* ctx, float* const* src, float* const* dst, int W, int H, double sigma, int gausstype);
int art_hp_gauss_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch,
                     int W, int H, double sigma, int gausstype);

/* ---- wavelet decomposition ------------------------------------------------- */
/*
 * Replaces rtengine::wavelet_decomposition (rtengine/cplx_wavelet_dec.h L37-95, ctor L97-198, reconstruct
 * L201-270) as FTblockDN uses it (rtengine/FTblockDN.cc L2296-2438): Daub4Len == 6, skipcrop == 1, `subsampling`
 * bit l set = level l decimated (bit 0 must be set).  Device-resident: the object owns its subbands in HBM;
 * callers (the shrink stages) read and write them through art_hp_wavelet_band_dev.
 *   art_hp_wavelet_decompose_dev   = the constructor (asynchronous on the context's stream)
 *   art_hp_wavelet_maxlevel        = maxlevel()
 *   art_hp_wavelet_level_dims      = level_W(), level_H(), level_stride()
 *   art_hp_wavelet_band_dev        = level_coeffs(level)[dir] for dir 1..3 (dense, width = level_W); dir 0 = coeff0,
 *                                    the lowpass of the LAST level
 *   art_hp_wavelet_reconstruct_dev = reconstruct(dst, blend); like the reference it consumes the decomposition and,
 *                                    for blend != 1, blends into the existing contents of dst
 */
typedef struct art_hp_wavelet art_hp_wavelet;
int    art_hp_wavelet_decompose_dev(art_hp_ctx* ctx, const float* d_src, size_t pitch, int W, int H, int maxlvl,
                                    int subsampling, art_hp_wavelet** out);
int    art_hp_wavelet_maxlevel(const art_hp_wavelet* w);
int    art_hp_wavelet_level_dims(const art_hp_wavelet* w, int level, int* width, int* height, int* stride);
float* art_hp_wavelet_band_dev(const art_hp_wavelet* w, int level, int dir);
int    art_hp_wavelet_reconstruct_dev(art_hp_wavelet* w, float* d_dst, size_t pitch, float blend);
/* host access to one subband (dense level_W x level_H floats), for callers that keep a stage on the CPU */
int    art_hp_wavelet_get_band(const art_hp_wavelet* w, int level, int dir, float* host);
int    art_hp_wavelet_set_band(art_hp_wavelet* w, int level, int dir, const float* host);
void   art_hp_wavelet_destroy(art_hp_wavelet* w);

/* ---- wavelet shrinkage (FTblockDN) ------------------------------------------ */
/*
 * art_hp_wavelet_mad_dev        madL[lvl][dir-1] = SQR(MadRgb(level_coeffs(lvl)[dir], W*H)) for all levels
 *                               (rtengine/FTblockDN.cc MadRgb L569-603, table fill L2311-2320) into d_madL, a device
 *                               array of 8*3 floats.
 * art_hp_wavelet_denoise_L_dev  WaveletDenoiseAllL(scale, L, noisevarlum, madL, nullptr, 0) (L1111-1167 -> ShrinkAllL L638-726)
 * art_hp_wavelet_denoise_AB_dev WaveletDenoiseAllAB(scale, L, ab, noisevarchrom, madL, noisevar_ab, useNoiseCCurve, autoch)
 *                               (L1170-1221 -> ShrinkAllAB L729-839)
 * d_noisevarlum / d_noisevarchrom: device arrays of level_W(0)*level_H(0) floats, as in the reference.
 * All asynchronous on the context's stream; nothing is copied to the host.
 */
int art_hp_wavelet_mad_dev(art_hp_ctx* ctx, const art_hp_wavelet* w, float* d_madL);
int art_hp_wavelet_denoise_L_dev(art_hp_ctx* ctx, art_hp_wavelet* wL, const float* d_noisevarlum, const float* d_madL, double scale);
int art_hp_wavelet_denoise_AB_dev(art_hp_ctx* ctx, const art_hp_wavelet* wL, art_hp_wavelet* wab, const float* d_noisevarchrom,
                                  const float* d_madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale);

/* ---- RGB_denoise ------------------------------------------------------------- */
/*
 * art_hp_rgb_denoise   rtengine::denoise::RGB_denoise(im, kall = 0, src = dst = img, calclum, ..., isRAW = true, dnparams,
 *                      expcomp = 0, noiseLCurve (unset), noiseCCurve, nresi, highresi) (rtengine/FTblockDN.cc L1638-2689)
 *                      as ImProcFunctions::denoise calls it (rtengine/ipdenoise.cc L1165), in place on three planes.
 * Parameters mirror procparams::DenoiseParams (rtengine/procparams.h) for the fields RGB_denoise reads.  Supported:
 * colorSpace RGB (0) | LAB (1), aggressive 0 | 1 (QUALITY_STANDARD | QUALITY_HIGH, FTblockDN.cc L1671), chrominanceMethod MANUAL (0); AUTOMATIC (1)
 * needs the camera-space frame and is resolved by art_hp_develop (or by the caller through art_hp_denoise_compute_params below): here it
 * returns ART_HP_ERR_UNSUPPORTED, like anything else.
 * `scale` is ImProcData::scale.  noiseCCurve: the 501-entry LUT NoiseCurve::Set builds (rtengine/ipdenoise.cc L684-705) and
 * its sum, host pointers, or NULL for "curve not set"; when given, the half-resolution calclum image of
 * ipdenoise.cc L1119-1131 (3 planes of ((H+1)/2) x ((W+1)/2)) must be supplied too.
 * wprof: ICCStore::workingSpaceMatrix(params->icm.workingProfile), row-major 3x3 doubles.
 * nresi_highresi: optional host float[2] receiving nresi, highresi (forces a stream synchronisation).
 * The block DCT of detail_recovery (L1479-1635) is FFTW's in the reference; here it is an fp32 matrix product on the
 * GPU -- the one stage whose results are within tolerance instead of bit-identical.
 */
typedef struct art_hp_denoise_params {
    double luminance, luminanceDetail;
    int luminanceDetailThreshold;
    double chrominance, chrominanceRedGreen, chrominanceBlueYellow;
    double gamma;
    double scale;
    int colorSpace, aggressive, chrominanceMethod;
    const float* noiseCCurve;
    float noiseCCurveSum;
    const double* wprof_inverse;    /* ICCStore::workingSpaceInverseMatrix, 9 doubles: needed when colorSpace == 1 (LAB), else may be NULL */
    /* ---- ABI version 2 ---- */
    double chrominanceAutoFactor;   /* chrominanceMethod == 1 (AUTOMATIC, the reference default): the estimate is multiplied by it; 0 reads as 1 */
} art_hp_denoise_params;
int art_hp_rgb_denoise(art_hp_ctx* ctx, float* const* r, float* const* g, float* const* b, int W, int H,
                       const art_hp_denoise_params* params, const double wprof[9],
                       float* const* calclum_r, float* const* calclum_g, float* const* calclum_b, float* nresi_highresi);
int art_hp_rgb_denoise_dev(art_hp_ctx* ctx, float* d_r, float* d_g, float* d_b, size_t pitch, int W, int H,
                           const art_hp_denoise_params* params, const double wprof[9],
                           const float* d_calclum_r, const float* d_calclum_g, const float* d_calclum_b, size_t calclum_pitch,
                           float* nresi_highresi);

/*
 * art_hp_denoise_compute_params   ImProcFunctions::denoiseComputeParams for DenoiseParams::ChrominanceMethod::AUTOMATIC -- the reference's default
 *                      (rtengine/ipdenoise.cc L800-1093; RGB_denoise_info L227-669, calcautodn_info L66-206, WaveletDenoiseAll_info /
 *                      ShrinkAll_info rtengine/FTblockDN.cc L1227-1364): nine crops of half the frame each are measured (wavelet MADs,
 *                      chroma / hue / luminance statistics in the reference's summation order) and combined into
 *                      out3 = DenoiseInfoStore::chrominance, chrominanceRedGreen, chrominanceBlueYellow (the caller multiplies by
 *                      chrominanceAutoFactor as L1064-1066 do).  red / green / blue: the DEMOSAICED planes in camera space, W x H = the size
 *                      getFullSize reports (after the raw border crop); mul / doClip / cam2work as in art_hp_scale_convert (getImage and
 *                      convertColorSpace run inside, per crop, as in the reference); wprof = ICCStore::workingSpaceMatrix; gamma / aggressive =
 *                      DenoiseParams'.  stats (optional): 9 x 15 floats, crop k = hcr * 3 + wcr: chaut, Nb, redaut, blueaut, maxredaut,
 *                      maxblueaut, minredaut, minblueaut, chromina, sigma, lumema, sigma_L, redyel, skinc, nsknc.  W, H >= 256.
 *                      Bit-identical to the reference.  Synchronises the context's stream.
 */
int art_hp_denoise_compute_params(art_hp_ctx* ctx, int W, int H, float* const* red, float* const* green, float* const* blue,
                                  const float mul[3], int doClip, const double cam2work[9], const double wprof[9], double gamma, int aggressive,
                                  float out3[3], float* stats);
int art_hp_denoise_compute_params_dev(art_hp_ctx* ctx, int W, int H, const float* d_red, const float* d_green, const float* d_blue, size_t pitch,
                                      const float mul[3], int doClip, const double cam2work[9], const double wprof[9], double gamma, int aggressive,
                                      float out3[3], float* stats);

/* ---- detail mask / NL-means --------------------------------------------------- */
/*
 * art_hp_detail_mask   rtengine::denoise::detail_mask(src, mask, scaling, threshold, ceiling, factor, blur_type, blur, mt)
 *                      (rtengine/FTblockDN.cc L1408-1476; laplacian L1366-1403; rtengine/rescale.h rescaleBilinear L53-77).
 *                      blur_type follows denoise::BlurType (rtengine/ipdenoise.h L81-85): 0 OFF, 1 BOX, 2 GAUSS.
 * art_hp_nlmeans       rtengine::denoise::NLMeans(img, normcoeff, strength, detail_thresh, scale, mt)
 *                      (rtengine/nlmeans.cc L50-280), in place on one plane.  strength == 0 returns at once like the
 *                      reference.  The reference's tile grid (150/136), shift order, 4-wide vector / scalar LUT lookups
 *                      and flush-to-zero arithmetic are reproduced: results are bit-identical.
 * The host forms take array2D-style row tables (row i at rows[i], W floats); the _dev forms device planes with a
 * pitch in floats, asynchronous on the context's stream.
 */
int art_hp_detail_mask(art_hp_ctx* ctx, float* const* src, float* const* mask, int W, int H,
                       float scaling, float threshold, float ceiling, float factor, int blur_type, float blur);
int art_hp_detail_mask_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_mask, size_t mask_pitch, int W, int H,
                           float scaling, float threshold, float ceiling, float factor, int blur_type, float blur);
int art_hp_nlmeans(art_hp_ctx* ctx, float* const* img, int W, int H, float normcoeff, int strength, int detail_thresh, float scale);
int art_hp_nlmeans_dev(art_hp_ctx* ctx, float* d_img, size_t pitch, int W, int H, float normcoeff, int strength, int detail_thresh, float scale);

/* ---- Fattal tone mapping ----------------------------------------------------- */
/*
 * art_hp_fattal        rtengine::ImProcFunctions::dynamicRangeCompression(Imagefloat*) -> ToneMapFattal02
 *                      (rtengine/tmo_fattal02.cc L1053-1215, L1220-1225; declared rtengine/improcfun.h L147), in place on the
 *                      three planes of the working-space image.  threshold / amount / satcontrol are
 *                      procparams::FattalToneMappingParams (rtengine/procparams.h L835-840); ws is
 *                      ICCStore::workingSpaceMatrix(params->icm.workingProfile), row-major 3x3 doubles.
 *                      alpha <= 0 or beta <= 0 returns at once, like the reference (L1068-1070).
 *                      Bit-identical to the reference except for its two external-library calls: the 2-D REDFT00
 *                      transforms (FFTW there; an fp64 shared-memory FFT here, rounded to float where the reference's plan
 *                      stores floats) and pow() in calculateFiMatrix (glibc powf there; fp64 pow rounded to float here).
 *                      Padded sides (find_fast_dim(W) + 1, L1094-1095) above 14337 return ART_HP_ERR_UNSUPPORTED.
 * art_hp_fattal_fast_dim   find_fast_dim (L1014-1050).
 * art_hp_median_denoise    denoise::Median_Denoise(src, dst, [upperBound,] W, H, medianType, iterations = 1, ...)
 *                      (rtengine/FTblockDN.cc L87-445, rtengine/ipdenoise.h L60-75): type 0..5 = denoise::Median
 *                      {3X3_SOFT, 3X3_STRONG, 5X5_SOFT, 5X5_STRONG, 7X7, 9X9}; use_upper selects the overload that only
 *                      filters samples <= upper_bound.  src == dst allowed for the host form (the call tone mapping makes).
 * art_hp_redft00_2d    the transform at tmo_fattal02.cc L768-772 by itself (fftwf_plan_r2r_2d(n0, n1, in, out, REDFT00,
 *                      REDFT00)), contiguous host arrays; exposed so the transform can be checked against its definition.
 */
int art_hp_fattal(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                  int threshold, int amount, int satcontrol, const double ws[9]);
int art_hp_fattal_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                      int threshold, int amount, int satcontrol, const double ws[9]);
int art_hp_fattal_fast_dim(int dim);
int art_hp_median_denoise(art_hp_ctx* ctx, float* const* src, float* const* dst, int W, int H, int median_type,
                          int use_upper, float upper_bound);
int art_hp_median_denoise_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch, int W, int H,
                              int median_type, int use_upper, float upper_bound);
int art_hp_redft00_2d(art_hp_ctx* ctx, int n0, int n1, const float* in, float* out);

/* ---- box blur / guided filter ----------------------------------------------- */
/*
 * art_hp_boxblur*: replaces rtengine::boxblur(float** src, float** dst, int radius, int W, int H, bool multiThread)
 * (rtengine/boxblur.h L318-556); src == dst selects the in-place use.  Requires 2*radius+1 <= min(W,H).
 * art_hp_guided_filter*: replaces rtengine::guidedFilter(guide, src, dst, r, epsilon, multithread, subsampling)
 * (rtengine/guidedfilter.h L26, rtengine/guidedfilter.cc L80-241); subsampling <= 0 selects the reference's
 * own rule (calculate_subsampling, L58-75).  dst may be the same plane as src (the reference is called that
 * way by guidedFilterLog, L256).
 */
int art_hp_boxblur(art_hp_ctx* ctx, float* const* src, float* const* dst, int radius, int W, int H);
int art_hp_boxblur_dev(art_hp_ctx* ctx, const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch,
                       int radius, int W, int H);
int art_hp_guided_filter(art_hp_ctx* ctx, int W, int H, float* const* guide, float* const* src, float* const* dst,
                         int r, float epsilon, int subsampling);
int art_hp_guided_filter_dev(art_hp_ctx* ctx, int W, int H, const float* d_guide, size_t guide_pitch,
                             const float* d_src, size_t src_pitch, float* d_dst, size_t dst_pitch,
                             int r, float epsilon, int subsampling);

/*
 * art_hp_denoise_guided_smoothing   rtengine::denoise::denoiseGuidedSmoothing(im, rgb) (rtengine/ipsmoothing.cc L875-897; guided_smoothing
 *                      L334-409 with Channel::C, guidedFilterLog rtengine/guidedfilter.cc L243-263), in place on the three planes of the
 *                      working-space image: chroma smoothing guided by the log luminance, the input luminance kept.
 *                      guidedChromaRadius = params->denoise.guidedChromaRadius (0 returns at once, like the reference), scale =
 *                      ImProcData::scale, ws = ICCStore::workingSpaceMatrix.  Bit-identical to the reference's SSE2 build.
 */
int art_hp_denoise_guided_smoothing(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                                    const double ws[9], int guidedChromaRadius, double scale);
int art_hp_denoise_guided_smoothing_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                                        const double ws[9], int guidedChromaRadius, double scale);

/*
 * Bayer green equilibration (preprocess; RawImageSource::preprocess calls them when raw.bayersensor.greenthresh > 0, rtengine/rawimagesource.cc
 * L1720-1745), on the rawData plane in place, before the demosaic:
 *   art_hp_green_equilibrate_global   RawImageSource::green_equilibrate_global (rtengine/green_equil_RT.cc L37-89): both green phases scaled to their
 *                      common mean over the frame less `border`.  The row sums are added in row order (the reference's OpenMP reduction
 *                      moves with the schedule in the last bits; its one-thread order is the one reproduced).
 *   art_hp_green_equilibrate          RawImageSource::green_equilibrate(thresh, rawData) (L92-250): thresh = the constant of
 *                      GreenEqulibrateThreshold (0.01 * greenthresh); thresh_map (optional, W x H floats at map_pitch) stands for a derived
 *                      threshold class (per-pixel values, e.g. the PDAF-lines one).
 * Bit-identical to the reference's SSE2 build.
 */
int art_hp_green_equilibrate_global(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData, int border);
int art_hp_green_equilibrate_global_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch, int border);
int art_hp_green_equilibrate(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData, float thresh, const float* const* thresh_map);
int art_hp_green_equilibrate_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t pitch, float thresh,
                                 const float* d_thresh_map, size_t map_pitch);

/* ---- gain / clip / camera->working colour space ------------------------- */
/*
 * Replaces the per-pixel part of RawImageSource::getImage (rtengine/rawimagesource.cc L943-1025, full
 * resolution i.e. skip == 1, no flips -- art_hp_develop's `tran` / `hr_blend` cover those: `x *= mul[c]; if (doClip) x = CLIP(x)`) followed by the matrix
 * branch of RawImageSource::colorSpaceConversion_ (L3184-3213): mat = workingSpaceInverse * camMatrix, row
 * major double[9], applied as (float)(m0*r + m1*g + m2*b) in double.  mat == NULL skips the matrix.
 * In place on three planes.  The caller computes mul[] (rm,gm,bm, L790-928) and mat exactly as the
 * reference does on the host; they are 3 + 9 numbers.
 */
int art_hp_scale_convert(art_hp_ctx* ctx, int W, int H,
                         float* const* red, float* const* green, float* const* blue,
                         const float mul[3], int doClip, const double mat[9]);
int art_hp_scale_convert_dev(art_hp_ctx* ctx, int W, int H,
                             float* d_red, float* d_green, float* d_blue, size_t pitch,
                             const float mul[3], int doClip, const double mat[9]);

/* ---- per-pixel colour / curve chain ------------------------------------------------ */
/*
 * art_hp_color_chain   the per-pixel stages of ImProcFunctions::process (rtengine/improcfun.cc L567-641) fused into one pass,
 *                      in the reference's order: STAGE_1 exposure = expcomp (rtengine/ipexposure.cc L29-73: the caller passes
 *                      exp_scale = pow(2, expcomp) and black = params black * 2000, L33-34); STAGE_3 saturationVibrance
 *                      (rtengine/ipsaturation.cc L44-83), toneCurve (rtengine/iptonecurve.cc L553-716 for its per-pixel LUT
 *                      branch: filmlike_clip then StandardToneCurve::Apply [tonecurve_mode 0] or AdobeToneCurve::Apply [1],
 *                      rtengine/curves.h L360-368, L425-472; or, tonecurve_mode 2, the reference's default single NEUTRAL curve:
 *                      NeutralToneCurve::BatchApply, rtengine/curves.cc L891-1037, base curve LINEAR, no filmlike_clip pass before
 *                      it, iptonecurve.cc L581-589; or, after filmlike_clip like STD, tonecurve_mode 3 WeightedStdToneCurve::Apply
 *                      (curves.h L499-562), 4 SatAndValueBlendingToneCurve::Apply (L634-668), 5 LuminanceToneCurve::Apply (L474-496);
 *                      TcMode::PERCEPTUAL returns ART_HP_ERR_UNSUPPORTED; then apply_satcurve, L398-441), rgbCurves (rtengine/iprgbcurves.cc L113-146) and
 *                      labAdjustments (rtengine/iplabadjustments.cc L252-283 between Imagefloat::setMode(LAB) and the
 *                      setMode(RGB) of the next stage, rtengine/imagefloat.cc L841-878, L949-972).
 *                      Curves stay host-built (rtengine/curves.cc) and are passed by pointer as the LUT<float> data they fill:
 *                      65536 floats each, lab_lcurve 32770.  A NULL LUT / a zero `*_enabled` skips that stage like the
 *                      reference's `enabled == false` / identity-curve early-outs.  In place on three planes.
 *                      Bit-identical to the reference's SSE2 build, 4-pixel groups and scalar row tails included.
 */
/* One stage of the Curve::getVal chain the reference composes with DoubleCurve (rtengine/iptonecurve.cc L525-551, L652-658).
 * curves::setLutVal (rtengine/curves.h L224-231) reads the 65536-entry LUT for samples <= 65535 and calls Curve::getVal above
 * it -- and NeutralToneCurve::BatchApply never clips (its filmlike_clip runs with Lmax = 65535 * whitecoeff on [0, 1] data,
 * rtengine/curves.cc L893, L989), so over-range samples take that branch even at white point 1.  The curve objects stay
 * host-built; a stage is what one of them evaluates:
 *   kind 0  identity (DiagonalCurve of kind DCT_Empty)
 *   kind 1  DiagonalCurve::getVal for DCT_CatmullRom (rtengine/diagonalcurves.cc L511-522) -- the kind every tone curve has
 *           after ImProcFunctions::toneCurve's `adjust` (iptonecurve.cc L604-650): nearest point of the polyline poly_x / poly_y
 *           (n doubles each, the curve's own members), the last y above the last x
 *   kind 2  ContrastCurve::getVal (iptonecurve.cc L104-121): lin2log(pow(LIM(x, 0, w) / w, a), b) * w */
typedef struct art_hp_curve_stage {
    int kind;
    const double* poly_x; const double* poly_y; int n;
    double a, b, w;
} art_hp_curve_stage;

typedef struct art_hp_chain_params {
    int   exposure_enabled;  float exp_scale, black;
    int   saturation_enabled, saturation, vibrance;      /* procparams::SaturationParams, integers as in the GUI */
    int   tonecurve_mode;    const float* tonecurve_lut;  /* 0 = STD, 1 = FILMLIKE, 2 = NEUTRAL (the reference default), 3 = WEIGHTEDSTD,
                                                             4 = SATANDVALBLENDING, 5 = LUMINANCE (needs ws); NULL = no tone curve */
    const float *rcurve, *gcurve, *bcurve;                /* rgbCurves LUTs, each may be NULL */
    int   lab_enabled;       const float *lab_lcurve, *lab_acurve, *lab_bcurve;  float lab_chroma;
    const double* ws;        /* ICCStore::workingSpaceMatrix, 9 doubles (saturation luminance, rgb -> Lab) */
    const double* iws;       /* ICCStore::workingSpaceInverseMatrix, 9 doubles (Lab -> rgb) */
    /* ---- ABI version 2 ---- */
    float tonecurve_whitept;                              /* ToneCurve::whitecoeff (params->toneCurve.whitePoint when hasWhitePoint()); 0 reads as 1 */
    const art_hp_curve_stage* tonecurve_stages;           /* the curve above the LUT, applied first to last; NULL / 0 = `curve == nullptr` (LUT only) */
    int   tonecurve_nstages;
    const float *neutral_to_out, *neutral_to_work;        /* NeutralToneCurve::ApplyState::to_out / to_work (curves.cc L866-875), 9 floats each;
                                                             NULL = identity (no matrix for the output profile) */
    const float* satcurve_lut;                            /* apply_satcurve's table (satcurve_lut, iptonecurve.cc L365-374), 65536 floats; NULL = identity
                                                             saturation curve.  White point 1 and identity `saturation2` only */
    /* ---- ABI version 5 ---- */
    const float* softlight_lut;                           /* ImProcFunctions::softLight (rtengine/ipsoftlight.cc L44-81, the stage after labAdjustments): its
                                                             table f[i] = sl(strength / 100, i), 65536 floats, read for samples <= 65535; NULL = disabled */
} art_hp_chain_params;
int art_hp_color_chain(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                       const art_hp_chain_params* params);
int art_hp_color_chain_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                           const art_hp_chain_params* params);
/*
 * art_hp_lab_histogram   the L histogram ImProcFunctions::labAdjustments takes when labCurve.contrast != 0 (rtengine/iplabadjustments.cc L307-334:
 *                        hist16[(int)L]++ over the Lab image) and hands to the host's get_L_curve: the stages of `params` that precede the Lab
 *                        stage run on the fly, the frame goes to Lab exactly as Imagefloat::setMode(LAB) does, the planes are NOT modified and
 *                        the Lab curves of `params` are not read.  The caller builds lab_lcurve from it and calls art_hp_color_chain.  Exact (integer).
 */
int art_hp_lab_histogram(art_hp_ctx* ctx, int W, int H, const float* const* r, const float* const* g, const float* const* b,
                         const art_hp_chain_params* params, unsigned hist16[65536]);
int art_hp_lab_histogram_dev(art_hp_ctx* ctx, int W, int H, const float* d_r, const float* d_g, const float* d_b, size_t pitch,
                             const art_hp_chain_params* params, unsigned hist16[65536]);

/* ---- sharpening ------------------------------------------------------------------- */
/*
 * art_hp_sharpen_usm   ImProcFunctions::sharpening -> doSharpening (rtengine/ipsharpen.cc L711-790; declared
 *                      rtengine/improcfun.h) for SharpeningParams::method == "usm", in place on the three planes of the
 *                      working-space image: get_luminance (rtengine/rt_algo.cc L942-956), buildBlendMask with
 *                      autoContrast = false (L315-496), unsharp_mask (ipsharpen.cc L232-312: apply_gamma 3,
 *                      gaussianBlur(radius / scale), Threshold<int>::multiply, rtengine/procparams.h L445-503), multiply
 *                      (rt_algo.cc L958-975).  Fields are procparams::SharpeningParams (rtengine/procparams.h L669-693,
 *                      defaults rtengine/procparams.cc L1756-1776); threshold = {bottom_left, top_left, bottom_right,
 *                      top_right}; scale = ImProcFunctions::scale (1 for full-resolution output); ws =
 *                      ICCStore::workingSpaceMatrix.  amount < 1 or an image under 8x8 returns untouched, like the
 *                      reference (L716-718).  method 1 selects the "rld" route (the reference's default method).  halocontrol runs sharpenHaloCtrl (L80-141); edgesonly takes the difference image on a bilateral-filtered copy (bilateral<float, float>, rtengine/bilateral2.h L38-547).  Bit-identical to the
 *                      reference's SSE2 build.
 */
typedef struct art_hp_sharpen_params {
    double contrast;            /* 20 */
    double radius;              /* 0.5 */
    int    amount;              /* 200 */
    int    threshold[4];        /* 20, 80, 2000, 1200 */
    int    edgesonly;           /* 0 | 1 */
    int    halocontrol;         /* 0 | 1 */
    int    halocontrol_amount;
    double scale;               /* 1 */
    int    method;              /* 0 = "usm", 1 = "rld" (RL deconvolution: markImpulse + deconvsharpening, ipsharpen.cc L144-230, L747-771) */
    double deconvradius;        /* 0.75; rld is on the hot path for 0.25 <= deconvradius / scale < 25 (3x3 / 5x5 / 7x7 and recursive GAUSS_DIV / GAUSS_MULT) */
    int    deconvamount;        /* 100 */
    double deconvCornerBoost;   /* 0; > 0.01 * scale mixes a second deconvolution (radius + boost) in towards the corners (CornerBoostMask, L313-338) */
    int    deconvCornerLatitude;            /* 25 */
    int    offset_x, offset_y, full_width, full_height;   /* ImProcFunctions' viewport (improcfun.h L227-230); full_* <= 0 means the image itself */
    double edges_radius;        /* 1.9; edgesonly: bilateral sigma = edges_radius / scale */
    int    edges_tolerance;     /* 1800; edgesonly: range sigma of the bilateral filter, >= 1 */
} art_hp_sharpen_params;
int art_hp_sharpen_usm(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b,
                       const art_hp_sharpen_params* params, const double ws[9]);
int art_hp_sharpen_usm_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch,
                           const art_hp_sharpen_params* params, const double ws[9]);

/* ---- output packing --------------------------------------------------------------- */
/*
 * art_hp_scanlines     Imagefloat::getScanline(row, buffer, bps, isFloat) for row = 0 .. H - 1 (rtengine/imagefloat.cc L125-169; CLIP and
 *                      uint16ToUint8Rounded rtengine/rt_math.h L97-100, L144-147; DNG_FloatToHalf rtengine/halffloat.h L9-47): planar float
 *                      RGB in [0, 65535] to interleaved rows, row `y` at out + y * out_stride_bytes.  (bps, isFloat) in {(16, 0), (8, 0),
 *                      (32, 1), (16, 1)}.  Bit-identical to the reference.
 */
int art_hp_scanlines(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, int bps, int isFloat,
                     void* out, size_t out_stride_bytes);
int art_hp_scanlines_dev(art_hp_ctx* ctx, int W, int H, const float* d_r, const float* d_g, const float* d_b, size_t pitch, int bps, int isFloat,
                         void* d_out, size_t out_stride_bytes);

/* ---- hot / dead pixel filter ------------------------------------------------------- */
/*
 * art_hp_find_hot_dead_pixels        RawImageSource::findHotDeadPixels(bpMap, thresh, findHotPixels, findDeadPixels) (rtengine/badpixels.cc
 *                                    L477-627; called by RawImageSource::preprocess, rtengine/rawimagesource.cc L1394-1420, with
 *                                    raw.hotdeadpix_thresh): a sample is bad when it deviates from the median of its same-colour
 *                                    neighbourhood by more than varthresh x the mean deviation around it.  xtrans = NULL: Bayer (any 2x2
 *                                    CFA: only the distance-2 neighbours are used), else the 6x6 X-Trans matrix.  `map` plays PixelsMap: H
 *                                    rows of W bytes, row y at map + y * map_stride, non-zero = bad; bad pixels found are OR-ed in (the caller
 *                                    pre-fills it with the bad pixels it knows from its files, as the reference does).  *count = pixels
 *                                    marked by this call (the return value of the reference function).
 * art_hp_interpolate_bad_pixels_bayer  RawImageSource::interpolateBadPixelsBayer(bitmapBads, rawData) (L66-180), in place; *count = pixels
 *                                    interpolated.
 * Bit-identical to the reference (for findHotDeadPixels: whenever each of its OpenMP threads owns at least two rows; below that the stock
 * function depends on its thread count).  interpolateBadPixelsXtrans (L288-475) reads pixels its own loop may already have rewritten and is
 * not reproduced.
 */
int art_hp_find_hot_dead_pixels(art_hp_ctx* ctx, int W, int H, const int* xtrans, const float* const* rawData, float thresh,
                                int findHotPixels, int findDeadPixels, unsigned char* map, size_t map_stride, int* count);
int art_hp_find_hot_dead_pixels_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans, const float* d_raw, size_t raw_pitch, float thresh,
                                    int findHotPixels, int findDeadPixels, unsigned char* d_map, size_t map_pitch, int* count);
int art_hp_interpolate_bad_pixels_bayer(art_hp_ctx* ctx, int W, int H, unsigned filters, float* const* rawData,
                                        const unsigned char* map, size_t map_stride, int* count);
int art_hp_interpolate_bad_pixels_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* d_raw, size_t raw_pitch,
                                            const unsigned char* d_map, size_t map_pitch, int* count);
/* RawImageSource::interpolateBadPixelsXtrans(bitmapBads) (rtengine/badpixels.cc L288-475), in place on the CFA plane, in the reference's ONE-THREAD
 * (raster) order: the function reads the neighbours of a red / blue site's "virtual pixel" and its distance-2 partner without checking them
 * against the map, so its OpenMP schedule decides whether such a neighbour was already rewritten -- only the serial order is a function of the
 * input.  Bit-identical to the reference run on one thread (the device iterates to the fixed point of that order; one stream synchronisation per
 * pass, two or three passes unless bad pixels form long chains).  `xtrans` must be an X-Trans layout (ART_HP_ERR_INVALID otherwise). */
int art_hp_interpolate_bad_pixels_xtrans(art_hp_ctx* ctx, int W, int H, const int* xtrans, float* const* rawData,
                                         const unsigned char* map, size_t map_stride, int* count);
int art_hp_interpolate_bad_pixels_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans, float* d_raw, size_t raw_pitch,
                                             const unsigned char* d_map, size_t map_pitch, int* count);

/* ---- channel mixer ----------------------------------------------------------------- */
/*
 * art_hp_channel_mixer   the per-pixel loop of ImProcFunctions::channelMixer (rtengine/ipchmixer.cc L152-232), in place on linear RGB planes:
 *                        out = max(M rgb, 0), M row-major RR RG RB / GR GG GB / BR BG BB in float.  ChannelMixerParams::RGB_MATRIX: M[i] =
 *                        float(slider[i]) / 1000.f (L156-164); PRIMARIES_CHROMA: the caller passes get_mixer_matrix's result (L33-149: host
 *                        colour science on the working profile, not part of the hot path).  Bit-identical to the reference, including its two
 *                        clamps (SSE2 groups of four turn NaN into 0, the row tail keeps it).
 */
int art_hp_channel_mixer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const float matrix[9]);
int art_hp_channel_mixer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const float matrix[9]);

/* ---- HSL equalizer ----------------------------------------------------------------- */
/*
 * art_hp_hsl_equalizer   ImProcFunctions::hslEqualizer (rtengine/iphsl.cc L29-221; STAGE_1 of ImProcFunctions::process, improcfun.cc
 *                        L580-584) in place on working-space RGB planes in [0, 65535], including the Imagefloat::setMode(YUV) it starts with
 *                        (rtengine/imagefloat.cc L700-725) and the setMode(RGB) the next stage applies to the YUV image it leaves (L779-803):
 *                        hue / saturation of the chroma plane (Color::yuv2hsl, rtengine/color.cc L6691-6695), then per enabled curve -- S, L, H --
 *                        the hue-indexed FlatCurve value smoothed by guidedFilter(Y, mask, mask, radius, eps) (radius 4 / scale * smooth for S and H,
 *                        25 / scale * smooth for L; smooth = pow(10, LIM01(smoothing / 10)) - 1) and applied to saturation, luminance or hue.
 *                        Curves stay host-built: each is the polyline the reference's FlatCurve constructor produces (poly_x, poly_y, dyByDx of
 *                        rtengine/curves.h; FlatCurve(points, true, CURVES_MIN_POLY_POINTS / scale)), n = 0 for an identity curve (isIdentity(), the
 *                        reference then skips that curve); `coeff` is the function's local FlatCurve of L119-123 built the same way (needed with a
 *                        saturation curve).  Bit-identical to the reference.
 */
typedef struct art_hp_flat_curve {
    int n;                                        /* points of the polyline; 0 = identity */
    const double *poly_x, *poly_y, *dy_by_dx;     /* n doubles each (dy_by_dx: n - 1 used) */
} art_hp_flat_curve;
typedef struct art_hp_hsl_params {
    art_hp_flat_curve hcurve, scurve, lcurve;     /* params->hsl.hCurve / sCurve / lCurve */
    art_hp_flat_curve coeff;
    int smoothing;                                /* params->hsl.smoothing */
    double scale;                                 /* ImProcFunctions::scale */
    const double* ws;                             /* ICCStore::workingSpaceMatrix, 9 doubles */
} art_hp_hsl_params;
int art_hp_hsl_equalizer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_hsl_params* params);
int art_hp_hsl_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_hsl_params* params);

/* ---- tone equalizer ---------------------------------------------------------------- */
/*
 * art_hp_tone_equalizer   ImProcFunctions::toneEqualizer (rtengine/iptoneequalizer.cc L343-371; STAGE_1 of ImProcFunctions::process, improcfun.cc
 *                         L584) with tone_eq() (L68-338), in place on working-space RGB planes in [0, 65535]: the frame is scaled by
 *                         1 / 65535 * 2^-pivot, the luminance LIM(Y, 1e-5, 32) is smoothed -- regularization > 0: guidedFilterLog(10, Y, 5 / scale
 *                         + 0.5, 0.014); regularization > 1: posterised to 1/5 EV and guided by the unposterised Y with radius 350 / scale (and once
 *                         more with (reg - 1) times that radius, reg = 5 - min(regularization, 4)) -- and every pixel is multiplied by the
 *                         correction of its luminance: five sliders spread over twelve 2-EV gaussian bands (bands[0] blacks .. bands[4] whites,
 *                         -100 .. 100).  The colour-map preview (show_colormap, PREVIEW pipeline, lcms2) is not reproduced.  Bit-identical to the reference, SSE2 groups and scalar row tails included.
 */
typedef struct art_hp_toneeq_params {
    int    bands[5];                              /* params->toneEqualizer.bands */
    int    regularization;                        /* params->toneEqualizer.regularization */
    double pivot;                                 /* params->toneEqualizer.pivot */
    double scale;                                 /* ImProcFunctions::scale */
    const double* ws;                             /* ICCStore::workingSpaceMatrix, 9 doubles */
} art_hp_toneeq_params;
int art_hp_tone_equalizer(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_toneeq_params* params);
int art_hp_tone_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_toneeq_params* params);

/* ---- proPhotoBlue ------------------------------------------------------------------ */
/*
 * art_hp_prophoto_blue   proPhotoBlue(img, multiThread) (rtengine/improcfun.cc L312-357), the step ImProcFunctions::process STAGE_1 ends with when
 *                        params->icm.workingProfile == "ProPhoto" (L585-587), in place on working-space RGB planes: a pixel with r == 0 or g == 0 and no
 *                        negative channel loses 1 % of its HSV saturation (Color::rgb2hsv / hsv2rgb, rtengine/color.cc L586-622, L654-694).
 *                        Bit-identical to the reference.
 */
int art_hp_prophoto_blue(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b);
int art_hp_prophoto_blue_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch);

/* ---- black and white --------------------------------------------------------------- */
/*
 * art_hp_black_and_white   the pixel loops of ImProcFunctions::blackAndWhite (rtengine/ipbw.cc L283-312, L343-362; the last step of STAGE_3), in place
 *                          on working-space RGB planes: r = g = b = (bwr r' + bwg g' + bwb b') kcorec with r', g', b' read through the gamma tables when
 *                          given, then -- with ulut / vlut -- the colour cast: Imagefloat::setMode(YUV), u += ulut[Y], v += vlut[Y], and the
 *                          setMode(RGB) the next stage applies.  bwr / bwg / bwb / kcorec are computeBWMixerConstants' results (L50-217) and the tables
 *                          are the LUTf(65536) data the function fills (gamma_r/g/b: L264-272, all three or none; ulut / vlut: L321-341, both or none):
 *                          host code of the reference, passed by pointer.  Bit-identical to the reference, SSE2 groups and scalar row tails included.
 */
typedef struct art_hp_bw_params {
    float bwr, bwg, bwb, kcorec;
    const float *gamma_r, *gamma_g, *gamma_b;     /* 65536 floats each, or all NULL (hasgammabw == false) */
    const float *ulut, *vlut;                     /* 65536 floats each, or both NULL (colorCast.getBottom() == 0) */
    const double* ws;                             /* ICCStore::workingSpaceMatrix, 9 doubles; needed with the colour cast */
} art_hp_bw_params;
int art_hp_black_and_white(art_hp_ctx* ctx, int W, int H, float* const* r, float* const* g, float* const* b, const art_hp_bw_params* params);
int art_hp_black_and_white_dev(art_hp_ctx* ctx, int W, int H, float* d_r, float* d_g, float* d_b, size_t pitch, const art_hp_bw_params* params);

/* ---- dual demosaic ----------------------------------------------------------------- */
/*
 * art_hp_demosaic_vng4        RawImageSource::vng4_demosaic(rawData, red, green, blue) (rtengine/vng4_demosaic_RT.cc L32-397): the four-colour
 *                             VNG demosaic the dual methods use in flat regions.  prefilters = RawImage::prefilters, the CFA with the
 *                             second green of each 2x2 as colour 3 (RGGB: 0xb4b4b4b4).  W, H >= 8.  Bit-identical to the reference run on
 *                             one thread (on several threads the stock function races with its own border pass in rows / columns 3 and
 *                             n - 4 of its row chunks).
 * art_hp_dual_demosaic_bayer  RawImageSource::dual_demosaic_RT(isBayer = true, ...) (rtengine/dual_demosaic_RT.cc L39-152) for
 *                             Method::AMAZEBILINEAR / AMAZEVNG4 / RCDBILINEAR / RCDVNG4: the first demosaicer (method = ART_HP_BAYER_AMAZE |
 *                             ART_HP_BAYER_RCD), Color::RGB2L of its frame (rtengine/color.cc L1343-1380), buildBlendMask(L, blend, W, H,
 *                             contrast / 100, 1, autoContrast) (rtengine/rt_algo.cc L317-494, blur radius 2) and the flat-region demosaicer
 *                             mixed in by intp(blend, first, flat): second = ART_HP_DUAL_BILINEAR (bayer_bilinear_demosaic(blend, ...),
 *                             rtengine/bayer_bilinear_demosaic.cc L33-75) or ART_HP_DUAL_VNG4.  *contrast is in percent, in / out as in the
 *                             reference (L108-112): with autoContrast != 0 (the reference default, procparams.cc L2928) it returns the
 *                             threshold buildBlendMask chose from the flattest tile of the frame (W, H >= 80 then).  contrast == 0 without
 *                             autoContrast is the first demosaicer alone (L43-71).  Bit-identical to the reference.
 * art_hp_dual_demosaic_xtrans the same for X-Trans (isBayer = false): xtrans_interpolate(passes, useCieLab) (FOUR_PASS: 3, true; TWO_PASS:
 *                             1, false) and fast_xtrans_interpolate_blend (rtengine/xtrans_demosaic.cc L1033-1092).
 * The _dev forms take the threshold by value and, optionally, a device float that receives the threshold used (contrast / 100): they do not
 * synchronise, the automatic search runs as a chain of small kernels on the context's stream.
 */
#define ART_HP_DUAL_BILINEAR 0
#define ART_HP_DUAL_VNG4     1
int art_hp_demosaic_vng4(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* const* rawData,
                         float* const* red, float* const* green, float* const* blue);
int art_hp_demosaic_vng4_dev(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* d_raw, size_t raw_pitch,
                             float* d_red, float* d_green, float* d_blue, size_t out_pitch);
int art_hp_dual_demosaic_bayer(art_hp_ctx* ctx, int method, int second, int W, int H, unsigned filters, unsigned prefilters,
                               const float* const* rawData, float* const* red, float* const* green, float* const* blue,
                               double initialGain, int border, double* contrast, int autoContrast);
int art_hp_dual_demosaic_bayer_dev(art_hp_ctx* ctx, int method, int second, int W, int H, unsigned filters, unsigned prefilters,
                                   const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                   double initialGain, int border, double contrast, int autoContrast, float* d_threshold_out);
int art_hp_dual_demosaic_xtrans(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                                const float* const* rawData, float* const* red, float* const* green, float* const* blue,
                                double* contrast, int autoContrast);
int art_hp_dual_demosaic_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int xtrans[36], const float rgb_cam[12],
                                    const float* d_raw, size_t raw_pitch, float* d_red, float* d_green, float* d_blue, size_t out_pitch,
                                    double contrast, int autoContrast, float* d_threshold_out);

/* ---- resize ------------------------------------------------------------------------ */
/*
 * art_hp_resize_lanczos   ImProcFunctions::Lanczos(src, dst, scale) (rtengine/ipresize.cc L38-207; called by ImProcFunctions::resize L365-394
 *                         with the destination size of resizeScale L230-362, int(w * scale + 0.5)): three planes sW x sH -> dW x dH, 3 lobes,
 *                         int(6 / min(scale, 1)) + 1 taps.  The reference runs it between src->setMode(LAB) and dst->setMode(mode): those
 *                         per-pixel conversions are not part of this entry (all three planes are resampled alike, as the reference does with
 *                         its L / a / b slots).  Bit-identical to the reference on the same planes.  ART_HP_ERR_UNSUPPORTED below a scale of
 *                         about 0.001 (the taps of one output pixel no longer fit a shared-memory line).
 */
int art_hp_resize_lanczos(art_hp_ctx* ctx, int sW, int sH, float* const* s0, float* const* s1, float* const* s2,
                          int dW, int dH, float* const* d0, float* const* d1, float* const* d2, float scale);
int art_hp_resize_lanczos_dev(art_hp_ctx* ctx, int sW, int sH, const float* d_s0, const float* d_s1, const float* d_s2, size_t src_pitch,
                              int dW, int dH, float* d_d0, float* d_d1, float* d_d2, size_t dst_pitch, float scale);

/* ---- whole frame ------------------------------------------------------------------ */
/*
 * art_hp_develop       the stages of simpleprocess.cc's normal pipeline that are on the hot path, back to back on the
 *                      device: imgsrc->demosaic (rtengine/simpleprocess.cc L215-222), imgsrc->getImage gains +
 *                      convertColorSpace matrix branch, ipf.denoise (rtengine/ipdenoise.cc L1096-1189: half-resolution
 *                      calclum through the same matrix when a chroma noise curve is given, adjust_params for scale > 1,
 *                      RGB_denoise, then -- nlStrength != 0, i.e. smoothingEnabled with guidedChromaRadius 0 --
 *                      NLMeans(Y, 65535, nlStrength, nlDetail, scale) between Imagefloat::setMode(YUV) and setMode(RGB)),
 *                      ipf.process(STAGE_0) = dynamicRangeCompression (rtengine/improcfun.cc L580-583), and the per-pixel /
 *                      sharpening steps of STAGE_1..3 (see `sharpen` and `chain` below).
 *                      One host->device copy of the CFA plane, one device->host copy of the three planes.
 *                      denoise == NULL and fattal_enabled == 0 skip their stages, like `enabled = false` does.
 *                      Geometry: like the reference, the frame is cropped by RawImageSource::border after the demosaic --
 *                      getImage reads from (border, border) and every later stage sees (W - 2 border) x (H - 2 border)
 *                      (rtengine/rawimagesource.cc transformRect L664-700, computeFullSize L1163-1175; border = `border` for Bayer
 *                      sensors, i.e. 4, and 7 for X-Trans).  red / green / blue are therefore tables of H - 2 border rows of
 *                      W - 2 border floats; art_hp_develop_size gives the numbers.  full_frame = 1 keeps W x H (no crop).
 *                      denoise->chrominanceMethod AUTOMATIC (the reference default) runs denoiseComputeParams on the demosaiced frame first
 *                      (simpleprocess.cc L254-256) -- one stream synchronisation, as its result sets RGB_denoise's wavelet depth.
 *                      ImProcFunctions::denoise is reproduced with its expcomp(+ecomp) / expcomp(-ecomp) bracket
 *                      (ipdenoise.cc L1155-1163, L1181-1184: denoise_expcomp) and denoiseGuidedSmoothing (L1171-1172:
 *                      guidedChromaRadius).
 */
typedef struct art_hp_develop_params {
    int method;                 /* ART_HP_BAYER_AMAZE | ART_HP_BAYER_RCD | ART_HP_XTRANS_3PASS | ART_HP_XTRANS_1PASS */
    unsigned filters;
    double initialGain;
    int border;
    float mul[3];               /* rm, gm, bm of getImage (rawimagesource.cc L790-928) */
    int doClip;
    const double* cam2work;     /* 9 doubles, row major; NULL = no matrix */
    const art_hp_denoise_params* denoise;
    int nlStrength, nlDetail;
    int fattal_enabled, fattal_threshold, fattal_amount, fattal_satcontrol;
    const double* wprof;        /* ICCStore::workingSpaceMatrix(workingProfile), 9 doubles */
    /* ipf.process(STAGE_1..3) (rtengine/improcfun.cc L584-625): exposure, then sharpening ("usm"), then saturationVibrance,
     * toneCurve, rgbCurves, labAdjustments.  Both may be NULL.  Without sharpening the whole chain is one fused pass; with it
     * the exposure stage runs before the sharpening and the rest after, in the reference's order. */
    const art_hp_sharpen_params* sharpen;
    const art_hp_chain_params* chain;
    /* method == ART_HP_XTRANS_*: RawImage::getXtransMatrix (36 ints) and RawImage::getRgbCam (12 floats); `filters`, `initialGain`
     * and `border` are not used */
    const int* xtrans;
    const float* rgb_cam;
    /* ---- ABI version 2 ---- */
    int full_frame;             /* 0 = the reference's geometry (crop by the border, see above); 1 = outputs are W x H */
    int guidedChromaRadius;     /* params->denoise.smoothingEnabled ? guidedChromaRadius : 0 (default 3): denoiseGuidedSmoothing between
                                   RGB_denoise and NLMeans; likewise nlStrength above is 0 unless smoothingEnabled */
    double denoise_expcomp;     /* params->exposure.enabled ? params->exposure.expcomp : 0; > 0 brackets the denoise stage */
    /* ---- ABI version 3: the rest of RawImageSource::getImage (rtengine/rawimagesource.cc L781-1104, a standard CCD at skip == 1) ---- */
    int tran;                   /* coarse transform: TR_R90 1 | TR_R180 2 | TR_R270 3, | TR_VFLIP 4, | TR_HFLIP 8 (iimage.h L34-40): transLineStandard ->
                                   rotateLine per line (L57-95), then hflip / vflip (L1079-1086).  The quarter turns swap the output's width and height
                                   (art_hp_develop_size reports it); every later stage sees the turned frame, as in the reference */
    int hr_blend;               /* 1 = ExposureParams::HR_BLEND: hlRecovery -> HLRecovery_blend on every line after the gains (L1013-1015, L3613-3755);
                                   the reference then leaves doClip off (L882) -- the caller passes doClip as it computed it */
    float hlmax[3];             /* clmax[c] * {rm, gm, bm} (L916-918), used by hr_blend */
    /* ---- ABI version 4: the preview path (rtengine/improccoordinator.cc L377, rtengine/dcrop.cc L204: imgsrc->getImage(wb, tr, image, pp, ...)) ---- */
    int pp_x, pp_y, pp_width, pp_height, pp_skip;
                                /* PreviewProps: a window of the transformed, border-cropped full image and the subsampling step.  pp_skip <= 0: the
                                   whole frame at skip 1 (versions 1-3).  getImage renders ceil(pp_width / pp_skip) x ceil(pp_height / pp_skip) pixels
                                   (getSize, L1199-1203): each the pp_skip x pp_skip box sum of the demosaiced frame from transformRect's origin
                                   (L664-751, L943-968) times mul[] -- the caller folds the 1 / pp_skip^2 of L928-931 into mul[], as it folds the
                                   other factors.  The later stages take their `scale` (= pp_skip) in their own parameter blocks, as the reference's
                                   ImProcFunctions does.  The window must lie inside the full image.  Not with the automatic chroma estimator (it
                                   measures the whole frame: run it once, pass the estimate) and not on a band. */
} art_hp_develop_params;
/* output size of art_hp_develop for a W x H raw frame: *out_w = W - 2 b, *out_h = H - 2 b, b = the border it crops (0 with full_frame); with a
 * PreviewProps window (pp_skip > 0) the size of that window's image; ART_HP_ERR_INVALID when the window leaves the full image */
int art_hp_develop_size(const art_hp_develop_params* params, int W, int H, int* out_w, int* out_h, int* border);
int art_hp_develop(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                   float* const* red, float* const* green, float* const* blue);
int art_hp_develop_dev(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, const float* d_raw, size_t raw_pitch,
                       float* d_red, float* d_green, float* d_blue, size_t out_pitch);
/*
 * Batch-queue form of art_hp_develop, for the loop of rtgui/batchqueue.cc (BatchQueue::startProcessing ->
 * rtengine::startBatchProcessing, one job after the other, L586-676): art_hp_develop_submit queues one frame -- upload, kernels,
 * download on three streams -- and returns; art_hp_develop_wait blocks until the OLDEST queued frame's planes are in host
 * memory.  At most two frames are in flight, so a caller alternates submit(k), wait() [collects k-1] and the copies of one
 * frame overlap the kernels of its neighbours.  Planes must be pinned (art_hp_host_alloc) with a constant row stride and must
 * stay untouched until the frame is collected; otherwise ART_HP_ERR_UNSUPPORTED is returned and art_hp_develop is the call
 * to make.  Results are identical to art_hp_develop.  `params` is consumed before submit returns.
 */
int art_hp_develop_submit(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                          float* const* red, float* const* green, float* const* blue);
/* The same, with the developed frame leaving in the reference's wire format: Imagefloat::getScanline(row, buffer, bps, isFloat)
 * (rtengine/imagefloat.cc L125-169) runs on the device for every row and `out` -- pinned, H_out rows of 3 * W_out samples at
 * out_stride_bytes -- receives interleaved RGB: bps 16 / 8 integer (clamped; truncated / rounded), bps 32 / 16 float (v / 65535; half).
 * 6 (or 3) bytes per pixel cross PCIe instead of 12. */
int art_hp_develop_submit_packed(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, float* const* rawData,
                                 int bps, int isFloat, void* out, size_t out_stride_bytes);
int art_hp_develop_wait(art_hp_ctx* ctx);
int art_hp_develop_pending(const art_hp_ctx* ctx);

/* ---- ABI version 3: ONE frame across the GPUs of a box (row bands; SURVEY.md section 8e, BASELINE.json configs[2]) --------------
 *
 * Every rank (one process and one context per GPU) develops a band of rows of the same frame.  The reference has no
 * counterpart: its frame lives in one address space and OpenMP threads share it.  What a band needs from the rest of the frame:
 *   - demosaic: raw rows within the method's tile halo (art_hp_band_halo) -- the band is demosaiced on the frame's own
 *     tile grid (art_hp_demosaic_bayer_rows_dev), bit-identical to the whole frame;
 *   - getImage gains / matrix, colour chain: nothing (per pixel);
 *   - RGB_denoise: MadRgb of every wavelet subband is a whole-frame statistic -- each rank histograms the coefficient rows
 *     it owns and the int32 histograms are SUMMED OVER THE RANKS (one ncclAllReduce per decomposition, 15 x 65536
 *     counters), so every rank shrinks with the frame's MAD; everything else in the stage (wavelet filters, box blurs,
 *     64 x 64 DCT blocks anchored at the frame origin, unsharp mask) has a bounded reach, and the band carries `halo`
 *     extra rows either side that are computed and thrown away.  The band's first row is a multiple of 50 (2: the
 *     decimated wavelet level; 25: the DCT block grid), so the band's grids coincide with the frame's.
 * The result on the owned rows equals the single-GPU frame up to the rounding of the box blurs' running sums, which restart
 * at the band's first row (~1e-7 relative; tests hold 1e-5) -- within north_star's 1e-4, not bit-identical.
 * Not available on a band (ART_HP_ERR_UNSUPPORTED): Fattal tone mapping (a global 2-D transform), the automatic chroma
 * estimator (crops of the whole frame), NL-means, X-Trans, RCD.
 */
typedef struct art_hp_band_plan {
    int own_begin, own_end;      /* rows of the developed frame (after the border crop) this rank delivers; even or the frame's end */
    int band_begin, band_end;    /* rows it computes: own rows + halo, band_begin a multiple of 50, clipped to the frame */
    int dm_begin, dm_end;        /* rows of the raw frame it demosaics (on the method's tile grid) */
    int raw_begin, raw_end;      /* rows of the raw frame it reads (dm rows + the demosaic halo; at the frame's top / bottom also the
                                    rows the demosaicer mirrors) */
} art_hp_band_plan;
/* rows for a rank that owns [own_begin, own_end) of the developed frame of a W x H raw frame; halo >= 150 (200 is a good value) */
int art_hp_band_plan_rows(const art_hp_develop_params* params, int W, int H, int own_begin, int own_end, int halo, art_hp_band_plan* plan);
/* The collective: sum an int32 device buffer in place over the ranks that share the frame, ordered on `cuda_stream`.  Either
 * art_hp_comm_init (ncclAllReduce on the context's stream; libnccl.so.2 is loaded on first use) or a caller's own function. */
typedef int (*art_hp_allreduce_fn)(void* user, int* d_buf, size_t count, void* cuda_stream);
int art_hp_set_allreduce(art_hp_ctx* ctx, art_hp_allreduce_fn fn, void* user);
int art_hp_comm_unique_id(unsigned char id[128]);                                /* rank 0 makes it, every rank gets a copy */
int art_hp_comm_init(art_hp_ctx* ctx, const unsigned char id[128], int rank, int nranks);
int art_hp_comm_destroy(art_hp_ctx* ctx);
/* d_raw / d_red / ... are the addresses of ROW 0 of the raw frame and of the developed frame (as for
 * art_hp_demosaic_bayer_rows_dev); raw rows [raw_begin, raw_end) must be valid, rows
 * [band_begin, band_end) of the outputs are written and rows [own_begin, own_end) are the result.  Every rank of the
 * communicator must make the call (the all-reduces match up). */
int art_hp_develop_band_dev(art_hp_ctx* ctx, const art_hp_develop_params* params, int W, int H, const float* d_raw, size_t raw_pitch,
                            float* d_red, float* d_green, float* d_blue, size_t out_pitch, const art_hp_band_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* ART_HOTPATH_H */
