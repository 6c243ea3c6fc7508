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
// vector forms the reference uses when it fills LUTs: sleefsseavx.h xlogf L1232-1255, xlogfNoCheck L1306-1324,
// xexpfNoCheck L1347-1363
__device__ __forceinline__ float xlogf_nocheck(float d)
{
    const int e = ilogbp1f(d * 0.7071f);
    const float m = ldexpk4(d, -e);
    float x = (-1.0f + m) / (1.0f + m);
    const float x2 = x * x;
    float t = 0.2371599674224853515625f;
    t = t * x2 + 0.285279005765914916992188f;
    t = t * x2 + 0.400005519390106201171875f;
    t = t * x2 + 0.666666567325592041015625f;
    t = t * x2 + 2.0f;
    return x * t + 0.693147180559945286226764f * (float)e;
}
__device__ __forceinline__ float xlogf_vector(float d)
{
    float x = xlogf_nocheck(d);
    if (d == __int_as_float(0x7f800000)) x = __int_as_float(0x7f800000);
    if (0.f > d) x = __int_as_float(0x7fc00000);
    if (d == 0.f) x = __int_as_float(0xff800000);
    return x;
}
__device__ __forceinline__ float xexpf_nocheck(float d)
{
    const int q = __float2int_rn(d * R_LN2);
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = exp_poly(s);
    u = 1.0f + ((s * s) * u + s);
    return ldexpk4(u, q);
}
__device__ __forceinline__ float xcbrtf_scalar(float d)
{   // sleef.h L966-991
    float x, y, q = 1.0f;
    int e = ilogbp1f(d);
    d = ldexpk2(d, -e);
    const int r = (e + 6144) % 3;
    q = (r == 1) ? 1.2599210498948731647672106f : q;
    q = (r == 2) ? 1.5874010519681994747517056f : q;
    q = ldexpk2(q, (e + 6144) / 3 - 2048);
    q = __int_as_float(__float_as_int(q) ^ (__float_as_int(d) & 0x80000000));
    d = fabsf(d);
    x = -0.601564466953277587890625f;
    x = x * d + 2.8208892345428466796875f;
    x = x * d + -5.532182216644287109375f;
    x = x * d + 5.898262500762939453125f;
    x = x * d + -3.8095417022705078125f;
    x = x * d + 2.2241256237030029296875f;
    y = d * x * x;
    y = (y - (2.0f / 3.0f) * y * (y * x - 1.0f)) * q;
    return y;
}

// pow_F(a, b) = xexpf(b * xlogf(a)), sleef.h L29; xlin2log sleef.h L1303-1307
__device__ __forceinline__ float pow_F_scalar(float a, float b) { return xexpf_scalar(b * xlogf_scalar(a)); }
__device__ __forceinline__ float xlin2log_scalar(float x, float base) { return xlogf_scalar(x * (base - 1.f) + 1.f) / xlogf_scalar(base); }

// xlog2lin, sleef.h L1309-1313
__device__ __forceinline__ float xlog2lin_scalar(float x, float base) { return (pow_F_scalar(base, x) - 1.f) / (base - 1.f); }

__device__ __forceinline__ float mulsignf(float x, float y) { return __int_as_float(__float_as_int(x) ^ (__float_as_int(y) & 0x80000000)); }
__device__ __forceinline__ float atan2kf(float y, float x)
{   // sleef.h L1155-1177
    float s, t, u, q = 0.f;
    if (x < 0) { x = -x; q = -2.f; }
    if (y > x) { t = x; x = y; y = -t; q += 1.f; }
    s = y / x;
    t = s * s;
    u = 0.00282363896258175373077393f;
    u = u * t + -0.0159569028764963150024414f;
    u = u * t + 0.0425049886107444763183594f;
    u = u * t + -0.0748900920152664184570312f;
    u = u * t + 0.106347933411598205566406f;
    u = u * t + -0.142027363181114196777344f;
    u = u * t + 0.199926957488059997558594f;
    u = u * t + -0.333331018686294555664062f;
    t = u * t;
    t = t * s + s;
    return q * (float)(3.14159265358979323846 / 2.0) + t;
}
__device__ __forceinline__ float xatan2f(float y, float x)
{   // sleef.h L1179-1188
    const float PI_F = (float)3.14159265358979323846;
    float r = atan2kf(fabsf(y), x);
    r = mulsignf(r, x);
    if (isinf(x) || x == 0) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .5f)) : 0);
    if (isinf(y)) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .25f)) : 0);
    if (y == 0) r = (copysignf(1.f, x) == -1 ? PI_F : 0);
    return (x != x) || (y != y) ? __int_as_float(0x7fc00000) : mulsignf(r, y);
}
__device__ __forceinline__ void xsincosf(float d, float& sn, float& cs)
{   // sleef.h L1048-1052 -> sleefsseavx.h L1051-1101 (an SSE2 build routes the scalar call through the vector form)
    const float A = 0.78515625f * 2, B = 0.00024127960205078125f * 2, C = 6.3329935073852539062e-07f * 2, D = 4.9604681473525147339e-10f * 2;
    const int q = __float2int_rn(d * (float)(2.0 / 3.14159265358979323846));
    float u = (float)q, s = d, t, rx, ry;
    s = u * -A + s; s = u * -B + s; s = u * -C + s; s = u * -D + s;
    t = s;
    s = s * s;
    u = -0.000195169282960705459117889f;
    u = u * s + 0.00833215750753879547119141f;
    u = u * s + -0.166666537523269653320312f;
    u = (u * s) * t;
    rx = t + u;
    u = -2.71811842367242206819355e-07f;
    u = u * s + 2.47990446951007470488548e-05f;
    u = u * s + -0.00138888787478208541870117f;
    u = u * s + 0.0416666641831398010253906f;
    u = u * s + -0.5f;
    ry = 1.f + s * u;
    float x = (q & 1) == 0 ? rx : ry, y = (q & 1) == 0 ? ry : rx;
    if ((q & 2) == 2) x = -x;
    if (((q + 1) & 2) == 2) y = -y;
    if (isinf(d)) { x = __int_as_float(0x7fc00000); y = x; }
    sn = x; cs = y;
}

}  // namespace sleef
