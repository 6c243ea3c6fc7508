/*
 * oracle/pack_port.c -- CPU restatement of the reference's output packing (the wire format of a developed frame).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Restates Imagefloat::getScanline (reference rtengine/imagefloat.cc L125-169): planar float RGB in [0, 65535] to interleaved rows of
 *   bps 16, integer   uint16 = (unsigned short) CLIP(v): clamp to [0, 65535], then truncate (no rounding)
 *   bps 8,  integer   uint8 = uint16ToUint8Rounded((uint16) CLIP(v)) = ((i + 128) - ((i + 128) >> 8)) >> 8      (rt_math.h L144-147)
 *   bps 32, float     v / 65535.f
 *   bps 16, float     DNG_FloatToHalf(v / 65535.f) (halffloat.h L9-47): round to nearest with ties away from zero on the 13 dropped
 *                     bits, denormals below 2^-14, flush to signed zero below 2^-25 (exponent < -10), NaN keeps its top mantissa bits
 * Pinned bit-exact against the reference's own function compiled in place (oracle/_ref) in tests/test_oracle_pack.py.
 */
#include <stdint.h>
#include <string.h>

static inline float clipf(float a)
{   /* CLIP = LIM(a, 0, MAXVAL) = max(lo, min(a, hi)), rt_math.h L55-58, L73-76, L84-100: a NaN comes out as 0 */
    const float m = 65535.f < a ? 65535.f : a;      /* min(a, hi) = hi < a ? hi : a */
    return 0.f < m ? m : 0.f;                       /* max(lo, m) = lo < m ? m : lo */
}

uint16_t artoracle_float_to_half(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const int32_t e = (int32_t)((u >> 23) & 0xffu) - 112;
    uint32_t m = u & 0x007fffffu;
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        m = (m | 0x00800000u) >> (1 - e);
        if (m & 0x1000u) m += 0x2000u;
        return (uint16_t)(sign | (m >> 13));
    }
    if (e == 143) return (uint16_t)(sign | 0x7c00u | (m >> 13));       /* inf (m == 0) / NaN */
    int32_t ee = e;
    if (m & 0x1000u) {
        m += 0x2000u;
        if (m & 0x00800000u) { m = 0; ee += 1; }
    }
    if (ee > 30) return (uint16_t)(sign | 0x7c00u);
    return (uint16_t)(sign | ((uint32_t)ee << 10) | (m >> 13));
}

/* out: H rows of 3 W samples, sample size bps / 8 bytes */
int artoracle_scanlines(const float* r, const float* g, const float* b, int W, int H, int bps, int is_float, void* out)
{
    const float* P[3] = {r, g, b};
    for (int row = 0; row < H; ++row)
        for (int i = 0; i < W; ++i)
            for (int c = 0; c < 3; ++c) {
                const float v = P[c][(size_t)row * W + i];
                const size_t ix = ((size_t)row * W + i) * 3 + c;
                if (is_float && bps == 32) ((float*)out)[ix] = v / 65535.f;
                else if (is_float && bps == 16) ((uint16_t*)out)[ix] = artoracle_float_to_half(v / 65535.f);
                else if (!is_float && bps == 16) ((uint16_t*)out)[ix] = (uint16_t)clipf(v);
                else if (!is_float && bps == 8) { const uint16_t k = (uint16_t)clipf(v); ((uint8_t*)out)[ix] = (uint8_t)(((k + 128) - ((k + 128) >> 8)) >> 8); }
                else return 1;
            }
    return 0;
}
