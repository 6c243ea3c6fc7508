// Output packing: Imagefloat::getScanline (reference rtengine/imagefloat.cc L125-169) for every row of a developed frame -- planar float
// RGB in [0, 65535] to interleaved rows in the wire format the reference hands to its image writers:
//   bps 16 integer   (unsigned short) CLIP(v): clamp to [0, 65535] (NaN -> 0), then truncate
//   bps  8 integer   uint16ToUint8Rounded of that: ((i + 128) - ((i + 128) >> 8)) >> 8          (rt_math.h L144-147)
//   bps 32 float     v / 65535.f
//   bps 16 float     DNG_FloatToHalf(v / 65535.f) (halffloat.h L9-47): ties away from zero on the 13 dropped bits, denormals below 2^-14,
//                    signed zero below 2^-25, NaN keeps its top mantissa bits
// On the device this turns the 12 B/px the batch queue sends back over PCIe into 6 (16-bit) or 3 (8-bit).
#include "ctx.h"

namespace {

__device__ __forceinline__ float clipf(float a)
{   // CLIP = max(lo, min(a, hi)) with rtengine's comparison forms: a NaN comes out as 0
    const float m = 65535.f < a ? 65535.f : a;
    return 0.f < m ? m : 0.f;
}
__device__ __forceinline__ unsigned short float_to_half(float f)
{
    const unsigned u = __float_as_uint(f);
    const unsigned sign = (u >> 16) & 0x8000u;
    const int e = (int)((u >> 23) & 0xffu) - 112;
    unsigned m = u & 0x007fffffu;
    if (e <= 0) {
        if (e < -10) return (unsigned short)sign;
        m = (m | 0x00800000u) >> (1 - e);
        if (m & 0x1000u) m += 0x2000u;
        return (unsigned short)(sign | (m >> 13));
    }
    if (e == 143) return (unsigned short)(sign | 0x7c00u | (m >> 13));
    int ee = e;
    if (m & 0x1000u) {
        m += 0x2000u;
        if (m & 0x00800000u) { m = 0; ee += 1; }
    }
    if (ee > 30) return (unsigned short)(sign | 0x7c00u);
    return (unsigned short)(sign | ((unsigned)ee << 10) | (m >> 13));
}

// v / 65535.f as an x86 SSE division gives it: a NaN operand comes back quieted with its payload (the GPU would return the canonical NaN,
// and DNG_FloatToHalf keeps the top mantissa bits)
__device__ __forceinline__ float div65535(float v)
{
    return (v != v) ? __uint_as_float(__float_as_uint(v) | 0x00400000u) : v / 65535.f;
}

struct PackArgs { const float *r, *g, *b; size_t ip; int W, H; unsigned char* out; size_t stride; };

// MODE 0: 16-bit integer, 1: 8-bit integer, 2: float32, 3: half.  One thread per pixel; a warp covers 32 consecutive pixels of a row.
template <int MODE>
__global__ void __launch_bounds__(256) k_scanlines(PackArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= a.W) return;
    for (int y = blockIdx.y; y < a.H; y += gridDim.y) {
        const size_t i = (size_t)y * a.ip + x;
        const float v[3] = {a.r[i], a.g[i], a.b[i]};
        unsigned char* row = a.out + (size_t)y * a.stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t ix = (size_t)x * 3 + c;
            if (MODE == 0) reinterpret_cast<unsigned short*>(row)[ix] = (unsigned short)clipf(v[c]);
            else if (MODE == 1) { const unsigned k = (unsigned short)clipf(v[c]); row[ix] = (unsigned char)(((k + 128) - ((k + 128) >> 8)) >> 8); }
            else if (MODE == 2) reinterpret_cast<float*>(row)[ix] = div65535(v[c]);
            else reinterpret_cast<unsigned short*>(row)[ix] = float_to_half(div65535(v[c]));
        }
    }
}

}  // namespace

int art_scanline_mode(int bps, int is_float)
{
    if (!is_float && bps == 16) return 0;
    if (!is_float && bps == 8) return 1;
    if (is_float && bps == 32) return 2;
    if (is_float && bps == 16) return 3;
    return -1;
}

int art_scanlines_dev(art_hp_ctx* ctx, int W, int H, const float* r, const float* g, const float* b, size_t ip, int bps, int is_float,
                      void* out, size_t stride_bytes)
{
    const int mode = art_scanline_mode(bps, is_float);
    if (mode < 0) return ctx->fail(ART_HP_ERR_INVALID, "bps %d / isFloat %d is not a format of Imagefloat::getScanline", bps, is_float);
    const size_t sample = (size_t)bps / 8;
    if (stride_bytes < (size_t)W * 3 * sample || (stride_bytes % sample) != 0 || (reinterpret_cast<uintptr_t>(out) % sample) != 0)
        return ctx->fail(ART_HP_ERR_INVALID, "output rows need %zu bytes and %zu-byte alignment (stride %zu)", (size_t)W * 3 * sample, sample, stride_bytes);
    PackArgs a{r, g, b, ip, W, H, (unsigned char*)out, stride_bytes};
    const dim3 grid((W + 255) / 256, std::min(H, 148 * 8));
    art_prof_begin(ctx, "k_scanlines");
    if (mode == 0) k_scanlines<0><<<grid, 256, 0, ctx->stream>>>(a);
    else if (mode == 1) k_scanlines<1><<<grid, 256, 0, ctx->stream>>>(a);
    else if (mode == 2) k_scanlines<2><<<grid, 256, 0, ctx->stream>>>(a);
    else k_scanlines<3><<<grid, 256, 0, ctx->stream>>>(a);
    art_prof_end(ctx);
    ctx->launches++;
    ART_CUDA(ctx, cudaGetLastError());
    return ART_HP_OK;
}
