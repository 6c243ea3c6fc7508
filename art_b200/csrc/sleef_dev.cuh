// sleef routines the hot path calls, for sm_100a.  Scalar forms follow reference rtengine/sleef.h, vector forms
// rtengine/sleefsseavx.h: they differ in the last bits and the reference uses one or the other depending on
// whether a sample sits in a 4-wide SSE group, so both exist here.  Compile with -fmad=false.
#pragma once
#include <cuda_runtime.h>

namespace sleef {

__device__ __forceinline__ float ldexpk4(float x, int q)
{   // vldexpf, sleefsseavx.h L987-996
    int m = q >> 31;
    m = (((m + q) >> 6) - m) << 4;
    q = q - (m << 2);
    float u = __int_as_float((m + 0x7f) << 23);
    x = x * u; x = x * u; x = x * u; x = x * u;
    u = __int_as_float((q + 0x7f) << 23);
    return x * u;
}
__device__ __forceinline__ float ldexpk2(float x, int q)
{   // ldexpkf, sleef.h L953-964
    int m = q >> 31;
    m = (((m + q) >> 6) - m) << 4;
    q = q - (m << 2);
    float u = __int_as_float((m + 0x7f) << 23);
    u = u * u;
    x = x * u * u;
    u = __int_as_float((q + 0x7f) << 23);
    return x * u;
}
constexpr float L2U = 0.693145751953125f, L2L = 1.428606765330187045e-06f;
constexpr float R_LN2 = 1.442695040888963407359924681001892137426645954152985934135449406931f;

__device__ __forceinline__ float exp_poly(float s)
{
    float u = 0.00136324646882712841033936f;
    u = u * s + 0.00836596917361021041870117f;
    u = u * s + 0.0416710823774337768554688f;
    u = u * s + 0.166665524244308471679688f;
    u = u * s + 0.499999850988388061523438f;
    return u;
}
__device__ __forceinline__ float xexpf_scalar(float d)
{   // sleef.h L1247-1266
    if (d <= -104.0f) return 0.0f;
    const int q = __float2int_rn(d * R_LN2);
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = exp_poly(s);
    u = s * (s * u + 1.f) + 1.f;
    return ldexpk2(u, q);
}
__device__ __forceinline__ float xexpf_vector(float d)
{   // sleefsseavx.h L1326-1345
    const int q = __float2int_rn(d * R_LN2);
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = exp_poly(s);
    u = 1.0f + ((s * s) * u + s);
    u = ldexpk4(u, q);
    return (-104.f > d) ? 0.f : u;
}


__device__ __forceinline__ int ilogbp1f(float d)
{   // sleef.h L945-951
    const bool m = d < 5.421010862427522E-20f;
    d = m ? 1.8446744073709552E19f * d : d;
    int q = (__float_as_int(d) >> 23) & 0xff;
    q = m ? q - (64 + 0x7e) : q - 0x7e;
    return q;
}
__device__ __forceinline__ float xlogf_scalar(float d)
{   // sleef.h L1197-1220
    const int e = ilogbp1f(d * 0.7071f);
    const float m = ldexpk2(d, -e);
    float x = (m - 1.0f) / (m + 1.0f);
    const float x2 = x * x;
    float t = 0.2371599674224853515625f;
    t = t * x2 + 0.285279005765914916992188f;
    t = t * x2 + 0.400005519390106201171875f;
    t = t * x2 + 0.666666567325592041015625f;
    t = t * x2 + 2.0f;
    x = x * t + 0.693147180559945286226764f * (float)e;
    if (d == __int_as_float(0x7f800000)) x = __int_as_float(0x7f800000);
    if (d < 0) x = __int_as_float(0x7fc00000);
    if (d == 0) x = __int_as_float(0xff800000);
    return x;
}
// pow_F(a, b) = xexpf(b * xlogf(a)), sleef.h L29; xlin2log sleef.h L1303-1307
__device__ __forceinline__ float pow_F_scalar(float a, float b) { return xexpf_scalar(b * xlogf_scalar(a)); }
__device__ __forceinline__ float xlin2log_scalar(float x, float base) { return xlogf_scalar(x * (base - 1.f) + 1.f) / xlogf_scalar(base); }

}  // namespace sleef
