// Internal context of libart_hotpath.so -- not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/art_hotpath.h"

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct PoolBlk { void* p; size_t bytes; bool used; };

struct ProfSpan { const char* name; cudaEvent_t e0, e1; };
struct ProfStat { std::string name; double ms = 0.0; int calls = 0; };

struct art_hp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // the stream work is queued on (own or caller's)
    cudaStream_t copy_stream = nullptr;  // host->device copies that overlap compute
    cudaStream_t d2h_stream = nullptr;   // device->host copies that overlap compute (other DMA engine)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long launches = 0;
    // cudaFuncSetAttribute is per device: the opt-ins to > 48 KB of dynamic shared memory are remembered per context, not per process
    enum { ATTR_RCD = 1, ATTR_XTRANS = 2, ATTR_DN_BLOCKS = 4, ATTR_NLM = 8, ATTR_FATTAL = 16, ATTR_DN_TC = 32, ATTR_SHRINK = 64, ATTR_AMAZE = 128 };
    unsigned attrs_set = 0;
    void* d_nlm_dbg = nullptr;
    // optional per-kernel timing (art_hp_profile_*): CUDA events around every launch
    bool profiling = false;
    std::vector<ProfSpan> spans;
    std::vector<ProfStat> stats;
    std::string err;
    // device scratch, grown on demand and kept across calls
    DevBuf d_dm[3];                      // art_hp_develop: the demosaiced W x H planes the cropped getImage stage reads
    DevBuf d_raw, d_out[3], d_scratch, d_small, d_work, d_dn, d_fattal, d_small2;
    // blocks handed to short-lived device objects (wavelet decompositions): reused across calls instead of
    // cudaMalloc/cudaFree per frame (both synchronise the device and serialise with NVML queries)
    std::vector<PoolBlk> pool;
    bool dn_tables_ready = false;        // the constant window / DCT tables of detail_recovery are uploaded once
    DevBuf d_dn_tables;
    DevBuf d_dn_labtabs;                 // colorSpace LAB: denoiseIGammaTab, denoiseGammaTab, cachef, cachefy
    bool dn_labtabs_ready = false;
    // pinned staging (two halves for double buffering)
    // colour chain: device LUT slots, their pinned staging and the event that says the staging may be rewritten
    DevBuf d_chain;
    void* h_chain = nullptr;
    cudaEvent_t ev_chain = nullptr;
    bool chain_cache_ready = false;
    std::vector<float> h_auto_tabs;      // automatic chroma: cachef / cachefy, built once
    cudaEvent_t ev_auto = nullptr;
    DevBuf d_chain_stages;               // Curve::getVal stages above the tone-curve LUT (descriptors + polylines)
    std::vector<char> h_chain_stages;
    DevBuf d_usm_tables;                 // apply_gamma's two 65536-entry LUTs (gamma 1/3 and 3), built once on the device
    bool usm_tables_ready = false;
    DevBuf d_bl_lut;                     // edges-only sharpening: the bilateral filter's range LUT (0x20000 floats), host-built
    int bl_lut_scale = 0, bl_lut_sens = 0;
    DevBuf d_xt_cbrt;                    // cielab's 0x14000-entry cube-root LUT of the X-Trans demosaic
    DevBuf d_dual_tabs;                  // dual demosaic: cachefy, VNG4's gradient programs, the automatic threshold's state
    bool dual_tabs_ready = false, dual_prog_ready = false;
    unsigned dual_prog_filters = 0;
    bool xt_cbrt_ready = false;
    // batch queue (art_hp_develop_submit / _wait): two frames in flight, each with its own raw + output planes
    struct QSlot { DevBuf raw, out[3], packed; cudaEvent_t up = nullptr, done = nullptr, down = nullptr; };
    QSlot q[2];
    unsigned long long q_submitted = 0, q_collected = 0;
    // side streams for work that is independent per wavelet subband (shrink.cu): the flat box blurs are serial recurrences
    // along a line, so one subband cannot fill the GPU -- the subbands of a channel run side by side
    static constexpr int NLANES = 15;    // one per (level, direction) of a 5-level decomposition
    cudaStream_t lane[NLANES] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NLANES] = {};
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_bytes = 0;
    // one frame across GPUs (art_hp_develop_band_dev): the planes the stages see are a row band of the frame; `own0 .. own1` are the
    // band-local rows this rank owns (whole-frame statistics count only those), H_full the frame's height after the border crop
    struct Band { bool active = false; int own0 = 0, own1 = 0, H_full = 0; } band;
    // sum of an int32 device buffer over the ranks sharing the frame, queued on `stream`: ncclAllReduce (art_hp_comm_init) or the caller's hook
    art_hp_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define ART_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return (ctx)->fail(ART_HP_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                               cudaGetErrorString(e_), __FILE__, __LINE__);                   \
    } while (0)

// RAII-free span helpers used by the launch sites
static inline void art_prof_begin(art_hp_ctx* ctx, const char* name)
{
    if (!ctx->profiling) return;
    ProfSpan sp{name, nullptr, nullptr};
    cudaEventCreate(&sp.e0);
    cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, ctx->stream);
    ctx->spans.push_back(sp);
}
static inline void art_prof_end(art_hp_ctx* ctx)
{
    if (!ctx->profiling || ctx->spans.empty()) return;
    cudaEventRecord(ctx->spans.back().e1, ctx->stream);
}

static inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// grow-only device buffer
int art_reserve(art_hp_ctx* ctx, DevBuf& b, size_t bytes);
// stream-ordered block pool: work that used a block was queued on ctx->stream, so the next user (same stream) is ordered after it
int art_pool_alloc(art_hp_ctx* ctx, size_t bytes, void** out);
void art_pool_free(art_hp_ctx* ctx, void* p);

// kernels (device-resident planes, pitch in floats); each returns an art_hp_status
// row_begin/row_end: output rows to produce (tile-grid aligned, see art_hp_demosaic_bayer_rows_dev)
int art_rcd_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                float* R, float* G, float* B, size_t op, int row_begin, int row_end);
int art_border_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, int bord, const float* raw, size_t rp,
                   float* R, float* G, float* B, size_t op, int row_begin, int row_end);
int art_amaze_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                  float* R, float* G, float* B, size_t op, double initialGain, int border, int row_begin, int row_end);
// Same, but in bands of `band_tile_rows` reference tile rows (0 = as many as the scratch budget allows);
// `cb(user, 0, 0, need_end)` is called before a band is queued: the band reads raw rows [0, need_end);
// `cb(user, 1, row0, row1)` after its kernels are queued: the band completes image rows [row0, row1).
// The host entry uses both to overlap host->device and device->host copies with the other bands.
typedef int (*art_band_cb)(void* user, int phase, int row0, int row1);
int art_amaze_dev_banded(art_hp_ctx* ctx, int W, int H, unsigned filters, const float* raw, size_t rp,
                         float* R, float* G, float* B, size_t op, double initialGain, int border,
                         int band_tile_rows, art_band_cb cb, void* user, int row_begin, int row_end);

// xtrans_interpolate(passes, useCieLab) + xtransborder_interpolate (xtrans.cu); xtrans36 = 6x6 colours, rgb_cam12 = 3x4
int art_xtrans_dev(art_hp_ctx* ctx, int passes, int useCieLab, int W, int H, const int* xtrans36, const float* rgb_cam12,
                   const float* raw, size_t rp, float* R, float* G, float* B, size_t op);
// getImage gain/clip + colorSpaceConversion_ matrix branch, in place on device planes
int art_scale_convert_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch,
                          const float mul[3], int doClip, const double* mat);
int art_scale_convert_crop_dev(art_hp_ctx* ctx, int W, int H, const float* sr, const float* sg, const float* sb, size_t sp,
                               float* r, float* g, float* b, size_t pitch, const float mul[3], int doClip, const double* mat,
                               int tran = 0, int hr_blend = 0, const float* hlmax = nullptr,
                               int skip = 1, int sx1 = 0, int sy1 = 0, int maxx = 0, int maxy = 0);      // skip > 1: sr / sg / sb are the un-offset planes
// denoise::denoiseGuidedSmoothing (smoothing.cu), planes in place
int art_guided_smoothing_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const double* ws9, int guidedChromaRadius, double scale);
// scaleColors (Bayer): in place; d_chmax_bits = 3 device ints receiving the float bit patterns of chmax[0..2]
int art_scale_colors_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch,
                         const float black[4], const float mul[4], int* d_chmax_bits);
// vng4_demosaic_RT.cc / dual_demosaic_RT.cc (dual.cu).  second: 0 bilinear (cfa = filters), 1 VNG4 (cfa = prefilters), 2 fast X-Trans (xtrans36)
int art_vng4_dev(art_hp_ctx* ctx, int W, int H, unsigned prefilters, const float* raw, size_t rp, float* R, float* G, float* B, size_t op);
const float* art_dual_threshold_slot(art_hp_ctx* ctx);
int art_dual_blend_dev(art_hp_ctx* ctx, int second, int W, int H, unsigned cfa, const int* xtrans36, const float* raw, size_t rp,
                       float* R, float* G, float* B, size_t op, double contrast, int auto_contrast, float* d_threshold_out);
// ImProcFunctions::channelMixer's loop (pointwise.cu), planes in place; m = RR RG RB / GR GG GB / BR BG BB
int art_prophoto_blue_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip);
int art_bw_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_bw_params* p);
int art_tone_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_toneeq_params* p);
int art_hsl_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_hsl_params* p);
int art_channel_mixer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const float m[9]);
int art_scale_colors_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, float* raw, size_t pitch,
                                const float black[3], const float mul[3], int* d_chmax_bits);
// badpixels.cc (badpixels.cu): findHotDeadPixels (xtrans36 == nullptr: Bayer) into a byte map (bad pixels OR-ed in), interpolateBadPixelsBayer;
// d_count = one device int receiving the number of pixels marked / interpolated
int art_find_hot_dead_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, const float* raw, size_t rp, float thresh, int hot, int dead,
                          unsigned char* map, size_t mp, int* d_count);
int art_interpolate_bad_xtrans_dev(art_hp_ctx* ctx, int W, int H, const int* xtrans36, float* raw, size_t rp, const unsigned char* map, size_t mp, int* d_count);
int art_interpolate_bad_bayer_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t rp, const unsigned char* map, size_t mp, int* d_count);
// ipresize.cc (resize.cu): ImProcFunctions::Lanczos on three planes
int art_lanczos_dev(art_hp_ctx* ctx, const float* s0, const float* s1, const float* s2, size_t sp, int sW, int sH,
                    float* d0, float* d1, float* d2, size_t dp, int dW, int dH, float scale);
// green_equil_RT.cc (greeneq.cu): global and local green equilibration of the Bayer plane, in place
int art_green_equilibrate_global_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch, int border);
int art_green_equilibrate_dev(art_hp_ctx* ctx, int W, int H, unsigned filters, float* raw, size_t pitch, float thresh, const float* thresh_map, size_t map_pitch);
// gaussianBlur, GAUSS_STANDARD (src == dst allowed)
int art_gauss_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, double sigma);
int art_gauss_divmult_dev(art_hp_ctx* ctx, float* src, size_t sp, float* dst, size_t dp, const float* div, size_t vp, int W, int H, double sigma, int type);
// rtengine::boxblur(float**, float**, radius, W, H) and rtengine::guidedFilter (src == dst allowed for boxblur;
// for the guided filter dst may alias src but not guide... both are read before dst is written only in the
// final pass, pixel by pixel, so dst may alias either)
int art_boxblur_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, int radius);
int art_guided_subsampling(int w, int h, int r);
// denoise::detail_mask / denoise::NLMeans (nlmeans.cu).  scratch: 2 * (W/4) * (H/4) floats; blur_type 0 off, 1 box, 2 gauss
int art_detail_mask_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* mask, size_t mp, int W, int H,
                        float scaling, float threshold, float ceiling, float factor, int blur_type, float blur, float* scratch);
// denoise::RGB_denoise (denoise.cu); planes r/g/b in place, calclum planes optional (needed with the chroma noise curve)
int art_rgb_denoise_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_denoise_params* P,
                        const double* wprof, const float* cl_r, const float* cl_g, const float* cl_b, size_t cp, float* nresi_highresi);
int art_nlmeans_dev(art_hp_ctx* ctx, float* img, size_t ip, int W, int H, float normcoeff, int strength, int detail_thresh, float scale);
int art_guided_dev(art_hp_ctx* ctx, const float* guide, size_t gp, const float* src, size_t sp, float* dst, size_t dp,
                   int W, int H, int r, float epsilon, int subsampling);
// ToneMapFattal02 (fattal.cu), planes in place; ws = working-space matrix, row-major 3x3
int art_fattal_dev(art_hp_ctx* ctx, float* R, float* G, float* B, size_t ip, int W, int H, int threshold, int amount, int satcontrol,
                   const double* ws9);
int art_fattal_fast_dim(int dim);
// denoise::Median_Denoise, one iteration, src != dst; type 0..5 = denoise::Median
int art_median_dev(art_hp_ctx* ctx, const float* src, size_t sp, float* dst, size_t dp, int W, int H, int type, int useUpper, float upper);
// 2-D REDFT00 of a contiguous n0 x n1 float array (the transform of tmo_fattal02.cc L768-772 alone)
int art_redft00_2d_dev(art_hp_ctx* ctx, const float* in, float* out, int n0, int n1);
// ImProcFunctions::process per-pixel chain (chain.cu), planes in place
int art_chain_lab_hist_dev(art_hp_ctx* ctx, int W, int H, const float* r, const float* g, const float* b, size_t pitch, const art_hp_chain_params* p, unsigned* d_hist);
int art_chain_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t pitch, const art_hp_chain_params* p);
// doSharpening, "usm" route (usm.cu), planes in place
int art_usm_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_sharpen_params* p, const double* ws9);
// develop.cu: ImProcFunctions::denoise (calclum, adjust_params, RGB_denoise, NL-means on Y) and the whole-frame pipeline
int art_denoise_stage_dev(art_hp_ctx* ctx, float* r, float* g, float* b, size_t ip, int W, int H, const art_hp_denoise_params* dn,
                          int nlStrength, int nlDetail, int guidedChromaRadius, double ecomp, const double* cam2work, const double* wprof);
// ImProcFunctions::denoiseComputeParams for the AUTOMATIC chroma method (denoise.cu): r / g / b = the demosaiced camera-space planes at the frame's
// origin (after the border crop); out3 = store.chrominance, chrominanceRedGreen, chrominanceBlueYellow; stats_out (optional) = 9 x 15 per-crop values.
// Synchronises the stream (the result decides RGB_denoise's wavelet depth).
int art_denoise_auto_chroma_dev(art_hp_ctx* ctx, const float* r, const float* g, const float* b, size_t ip, int widIm, int heiIm,
                                const float mul[3], int doClip, const double* cam2work, const double* wprof, double gamma, int aggressive,
                                float out3[3], float* stats_out);
// Imagefloat::getScanline for every row (pack.cu): planar float -> interleaved 16-bit / 8-bit / float / half rows
int art_scanline_mode(int bps, int is_float);
int art_scanlines_dev(art_hp_ctx* ctx, int W, int H, const float* r, const float* g, const float* b, size_t ip, int bps, int is_float,
                      void* out, size_t stride_bytes);
// output size of art_hp_develop for a W x H raw frame
void art_develop_geometry(const art_hp_develop_params* p, int W, int H, int* b, int* Wo, int* Ho);
// the same with getImage's source rectangle: (sx1, sy1) = transformRect's origin in the raw frame, iw x ih = imwidth x imheight (the source-orientation
// line geometry; Wo x Ho is the turned image), skip.  ART_HP_ERR_INVALID when the PreviewProps window leaves the frame.
struct art_dev_geo { int bd, Wo, Ho, sx1, sy1, iw, ih, skip; bool window; };
int art_develop_geometry2(const art_hp_develop_params* p, int W, int H, art_dev_geo* g);
int art_develop_dev(art_hp_ctx* ctx, const art_hp_develop_params* p, int W, int H, const float* raw, size_t rp,
                    float* r, float* g, float* b, size_t op);
// one frame across GPUs: geometry of a rank's band and the band pipeline (develop.cu); the collective (comm.cu)
int art_band_plan(const art_hp_develop_params* p, int W, int H, int own_begin, int own_end, int halo, art_hp_band_plan* out);
int art_develop_band_dev(art_hp_ctx* ctx, const art_hp_develop_params* p, int W, int H, const float* raw, size_t rp,
                         float* r, float* g, float* b, size_t op, const art_hp_band_plan* plan);
int art_allreduce_i32(art_hp_ctx* ctx, int* d_buf, size_t count);
