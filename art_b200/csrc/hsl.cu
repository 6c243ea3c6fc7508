// ImProcFunctions::hslEqualizer (reference rtengine/iphsl.cc L29-221), STAGE_1 of ImProcFunctions::process (improcfun.cc L580-584), with
// the mode changes around it: Imagefloat::setMode(YUV) on entry (rgb_to_yuv, imagefloat.cc L700-725), normalizeFloatTo1 / To65535
// (L396-438), and the setMode(RGB) of the next stage on the YUV image the function leaves (yuv_to_rgb, L779-803).
//   r plane = v -> hue (Color::yuv2hsl, color.cc L6691-6695: sleef xatan2f), g plane = Y, b plane = u -> saturation (sqrt)
//   S, L, H curves in this order, each: mask = FlatCurve::getVal(hue01(h)) (flatcurves.cc L339-365, fp64 over the host-built polyline),
//   guidedFilter(Y, mask, mask, radius, eps) (guided.cu, automatic subsampling), per-pixel update through tolin()
//   Color::hsl2yuv (sleef xsincosf), scale back, yuv2rgb (color.h L790-796)
// One kernel does "apply the update of the curve before + build the mask of the curve after" so that a frame with all three curves is
// four point-wise passes around three guided filters (the reference: eleven passes).  Bit-identical to the reference.
#include "ctx.h"
#include "sleef_dev.cuh"

#include <cmath>

namespace {

// start: for t in [0, 2), start[(int)(512 t)] = the interval FlatCurve::getVal's binary search returns at the bucket's lower edge; the interval of t
// is reached from there in a step or two instead of ten dependent probes
constexpr int HSL_BUCKETS = 1024;
struct FlatDev { int n; const double *px, *py, *dy; const int* start; };
struct HslArgs {
    float *r, *g, *b; size_t ip;
    float* mask; size_t mp;
    int W, H;
    int apply;          // 0 = entry (RGB -> h, Y, s), 1 = S update, 2 = L update, 3 = H update
    int next;           // 1 / 2 / 3 = build that curve's mask, 4 = exit (h, Y, s -> RGB)
    FlatDev cur, coeff; // the curve whose mask is built; the S update's local `coeff` curve
    float w0, w1, w2;
};

__device__ __forceinline__ double flat_getval(const FlatDev& c, double t)
{   // FCT_MinMaxCPoints, flatcurves.cc L344-365
    if (t < c.px[0]) t += 1.0;
    unsigned k_lo;
    if (t >= 0.0 && t < 2.0) {
        // the search below ends on the largest k <= n - 2 with px[k] <= t (0 when there is none): the same k, walked to from the bucket's start
        k_lo = (unsigned)c.start[(int)(t * 512.0)];
        while (k_lo + 2 < (unsigned)c.n && c.px[k_lo + 1] <= t) ++k_lo;
    } else {
        k_lo = 0;
        unsigned k_hi = (unsigned)c.n - 1;
        while (k_hi > 1 + k_lo) {
            const unsigned k = (k_hi + k_lo) / 2;
            if (c.px[k] > t) k_hi = k; else k_lo = k;
        }
    }
    return c.py[k_lo] + (t - c.px[k_lo]) * c.dy[k_lo];
}
__device__ __forceinline__ float lim01f(float a) { const float m = 1.f < a ? 1.f : a; return 0.f < m ? m : 0.f; }
__device__ __forceinline__ float sgnf(float a) { return (float)((0.f < a) - (a < 0.f)); }
__device__ __forceinline__ float hue01(float h)
{
    const float pi2 = 2.f * (float)3.14159265358979323846;
    const float v = h / pi2;
    if (v < 0.f) return 1.f + v;
    else if (v > 1.f) return v - 1.f;
    return v;
}
__device__ __forceinline__ float tolin(float y, float base)
{
    const float v = (y - 0.5f) * 2.f;
    return sgnf(v) * lim01f(sleef::xlog2lin_scalar(fabsf(v), base));
}

__global__ void __launch_bounds__(256) k_hsl(const HslArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    const float PI_F = (float)3.14159265358979323846;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t i = (size_t)y * a.ip + x, m = (size_t)y * a.mp + x;
        float h, Y, s;
        if (a.apply == 0) {
            const float R = a.r[i], G = a.g[i], B = a.b[i];
            const float down = 1.f / 65535.f;
            const float l = R * a.w0 + G * a.w1 + B * a.w2;
            const float u = (l - B) * down, v = (R - l) * down;
            Y = l * down;
            s = sqrtf(u * u + v * v);
            h = sleef::xatan2f(u, v);
        } else {
            h = a.r[i]; Y = a.g[i]; s = a.b[i];
            const float mk = a.mask[m];
            if (a.apply == 1) {
                const float f = tolin(mk, 2.f);
                const float e = (float)(1.f + (f < 0 ? flat_getval(a.coeff, s) : 1.f - flat_getval(a.coeff, s)));
                s *= 1.f + sgnf(f) * sleef::pow_F_scalar(lim01f(fabsf(f)), e);
            } else if (a.apply == 2) {
                Y *= 1.f + tolin(mk, 10.f);
            } else {
                h += tolin(mk, 32.f) * PI_F;
            }
        }
        if (a.next == 4) {
            float sn, cs;
            sleef::xsincosf(h, sn, cs);
            const float u = s * sn * 65535.f, v = s * cs * 65535.f, Yo = Y * 65535.f;
            const float bb = Yo - u, rr = v + Yo;
            a.r[i] = rr; a.b[i] = bb;
            a.g[i] = (Yo - rr * a.w0 - bb * a.w2) / a.w1;
        } else {
            a.r[i] = h; a.g[i] = Y; a.b[i] = s;
            a.mask[m] = (float)flat_getval(a.cur, hue01(h));
        }
    }
}

}  // namespace

int art_hsl_equalizer_dev(art_hp_ctx* ctx, int W, int H, float* r, float* g, float* b, size_t ip, const art_hp_hsl_params* p)
{
    if (!p->ws) return ctx->fail(ART_HP_ERR_INVALID, "the HSL equalizer needs the working-space matrix");
    if (!(p->scale > 0.0)) return ctx->fail(ART_HP_ERR_INVALID, "scale must be positive");
    const art_hp_flat_curve* cv[4] = {&p->hcurve, &p->scurve, &p->lcurve, &p->coeff};
    size_t doubles = 0;
    for (int c = 0; c < 4; ++c) {
        if (cv[c]->n < 0 || (cv[c]->n > 0 && (cv[c]->n < 2 || !cv[c]->poly_x || !cv[c]->poly_y || !cv[c]->dy_by_dx)))
            return ctx->fail(ART_HP_ERR_INVALID, "flat curve %d: a polyline needs at least two points and its three arrays", c);
        doubles += 3 * (size_t)cv[c]->n;
    }
    if (p->scurve.n > 0 && p->coeff.n == 0) return ctx->fail(ART_HP_ERR_INVALID, "the saturation curve needs the coeff curve (iphsl.cc L119-123)");
    cudaStream_t st = ctx->stream;
    const size_t mp = round_up((size_t)W, 32), n = mp * (size_t)H;
    void* blk = nullptr;
    const size_t curve_bytes = round_up(doubles * sizeof(double), 256);
    int rc = art_pool_alloc(ctx, round_up(n * sizeof(float), 256) + curve_bytes + 4 * HSL_BUCKETS * sizeof(int), &blk);
    if (rc) return rc;
    float* mask = (float*)blk;
    double* dcur = (double*)((char*)blk + round_up(n * sizeof(float), 256));
    int* dstart = (int*)((char*)dcur + curve_bytes);
    FlatDev dev[4];
    {
        std::vector<double> host(doubles);
        std::vector<int> hstart(4 * HSL_BUCKETS, 0);
        size_t off = 0;
        for (int c = 0; c < 4; ++c) {
            const int k = cv[c]->n;
            dev[c].n = k;
            dev[c].px = dcur + off; dev[c].py = dcur + off + k; dev[c].dy = dcur + off + 2 * (size_t)k;
            dev[c].start = dstart + c * HSL_BUCKETS;
            for (int b = 0; k && b < HSL_BUCKETS; ++b) {      // the binary search of flatcurves.cc L351-362 at t = b / 512
                const double t = b / 512.0;
                unsigned lo = 0, hi = (unsigned)k - 1;
                while (hi > 1 + lo) { const unsigned m = (hi + lo) / 2; if (cv[c]->poly_x[m] > t) hi = m; else lo = m; }
                hstart[c * HSL_BUCKETS + b] = (int)lo;
            }
            if (k) {
                memcpy(&host[off], cv[c]->poly_x, k * sizeof(double));
                memcpy(&host[off + k], cv[c]->poly_y, k * sizeof(double));
                memcpy(&host[off + 2 * (size_t)k], cv[c]->dy_by_dx, k * sizeof(double));
            }
            off += 3 * (size_t)k;
        }
        // pageable source: cudaMemcpyAsync returns once the bytes are staged, `host` may go out of scope
        if (doubles) {
            cudaError_t e = cudaMemcpyAsync(dcur, host.data(), doubles * sizeof(double), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dstart, hstart.data(), hstart.size() * sizeof(int), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) { art_pool_free(ctx, blk); return ctx->fail(ART_HP_ERR_CUDA, "curve upload failed: %s", cudaGetErrorString(e)); }
        }
    }
    // iphsl.cc L82: smooth = pow(10.f, LIM01(smoothing / 10.f)) - 1.f (host powf, as in the reference); radii L116, L150, L178
    const float t01 = std::max(0.f, std::min(p->smoothing / 10.f, 1.f));
    const float smooth = std::pow(10.f, t01) - 1.f;
    const int radius_of[4] = {0, (int)(4 / p->scale * smooth + 0.5), (int)(25 / p->scale * smooth + 0.5), (int)(4 / p->scale * smooth + 0.5)};
    const float eps_of[4] = {0.f, 0.001f, 0.0001f, 0.001f};
    const int order[3] = {1, 2, 3};             // S, L, H
    const FlatDev* curve_of[4] = {nullptr, &dev[1], &dev[2], &dev[0]};
    HslArgs a{};
    a.r = r; a.g = g; a.b = b; a.ip = ip; a.mask = mask; a.mp = mp; a.W = W; a.H = H;
    a.w0 = (float)p->ws[3]; a.w1 = (float)p->ws[4]; a.w2 = (float)p->ws[5];
    a.coeff = dev[3];
    const dim3 grid((W + 255) / 256, std::min(H, 148 * 8));
    int apply = 0;
    for (int k = 0; k <= 3 && !rc; ++k) {
        int next = 4;
        if (k < 3) { next = order[k]; if (curve_of[next]->n == 0) continue; }     // isIdentity(): the reference skips the curve
        a.apply = apply; a.next = next;
        if (next < 4) a.cur = *curve_of[next];
        art_prof_begin(ctx, "k_hsl");
        k_hsl<<<grid, 256, 0, st>>>(a);
        art_prof_end(ctx);
        ctx->launches++;
        if (next < 4 && radius_of[next] > 0) rc = art_guided_dev(ctx, g, ip, mask, mp, mask, mp, W, H, radius_of[next], eps_of[next], 0);
        apply = next;
    }
    art_pool_free(ctx, blk);
    if (rc) return rc;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
