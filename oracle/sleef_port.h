/*
 * oracle/sleef_port.h -- the few sleef routines the hot path calls, restated.  TEST INFRASTRUCTURE ONLY.
 * Scalar forms follow reference rtengine/sleef.h, vector forms rtengine/sleefsseavx.h (they differ in the last
 * bits, and the reference uses one or the other depending on whether a sample sits in a 4-wide SSE group).
 * Compile with -ffp-contract=off.
 */
#ifndef ART_ORACLE_SLEEF_PORT_H
#define ART_ORACLE_SLEEF_PORT_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

static inline float ldexpk(float x, int q)
{   /* ldexpkf (sleef.h L953-964) == vldexpf (sleefsseavx.h L987-996): x * u^4 * 2^q' */
    int m = q >> 31;
    m = (((m + q) >> 6) - m) << 4;
    q = q - (m << 2);
    float u = i2f((int32_t)(m + 0x7f) << 23);
    /* scalar: u = u*u; x = x*u*u  -- vector: x = x*u*u*u*u.  Powers of two: both exact unless they over/underflow,
     * which the shrinkage arguments never reach; we follow the caller's form below */
    x = x * u; x = x * u; x = x * u; x = x * u;
    u = i2f((int32_t)(q + 0x7f) << 23);
    return x * u;
}
static inline float ldexpk_scalar(float x, int q)
{
    int m = q >> 31;
    m = (((m + q) >> 6) - m) << 4;
    q = q - (m << 2);
    float u = i2f((int32_t)(m + 0x7f) << 23);
    u = u * u;
    x = x * u * u;
    u = i2f((int32_t)(q + 0x7f) << 23);
    return x * u;
}

#define L2U 0.693145751953125f
#define L2L 1.428606765330187045e-06f
#define R_LN2 1.442695040888963407359924681001892137426645954152985934135449406931f

static inline float xexpf_scalar(float d)
{   /* sleef.h L1247-1266 */
    if (d <= -104.0f) return 0.0f;
    const int q = (int)lrintf(d * R_LN2);          /* _mm_cvt_ss2si: round to nearest even */
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = 0.00136324646882712841033936f;
    u = u * s + 0.00836596917361021041870117f;
    u = u * s + 0.0416710823774337768554688f;
    u = u * s + 0.166665524244308471679688f;
    u = u * s + 0.499999850988388061523438f;
    u = s * (s * u + 1.f) + 1.f;
    return ldexpk_scalar(u, q);
}
static inline float xexpf_vector(float d)
{   /* sleefsseavx.h L1326-1345 */
    const int q = (int)lrintf(d * R_LN2);
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = 0.00136324646882712841033936f;
    u = u * s + 0.00836596917361021041870117f;
    u = u * s + 0.0416710823774337768554688f;
    u = u * s + 0.166665524244308471679688f;
    u = u * s + 0.499999850988388061523438f;
    u = 1.0f + ((s * s) * u + s);
    u = ldexpk(u, q);
    return (-104.f > d) ? 0.f : u;
}



static inline int ilogbp1f_(float d)
{   /* sleef.h L945-951 */
    const int m = d < 5.421010862427522E-20f;
    d = m ? 1.8446744073709552E19f * d : d;
    int q = (f2i(d) >> 23) & 0xff;
    q = m ? q - (64 + 0x7e) : q - 0x7e;
    return q;
}

static inline float xlogf_scalar(float d)
{   /* sleef.h L1197-1220 */
    const int e = ilogbp1f_(d * 0.7071f);
    const float m = ldexpk_scalar(d, -e);
    float x = (m - 1.0f) / (m + 1.0f);
    const float x2 = x * x;
    float t = 0.2371599674224853515625f;
    t = t * x2 + 0.285279005765914916992188f;
    t = t * x2 + 0.400005519390106201171875f;
    t = t * x2 + 0.666666567325592041015625f;
    t = t * x2 + 2.0f;
    x = x * t + 0.693147180559945286226764f * e;
    if (d == INFINITY) x = INFINITY;
    if (d < 0) x = NAN;
    if (d == 0) x = -INFINITY;
    return x;
}

/* vector forms used when the reference fills LUTs: sleefsseavx.h xlogf L1232-1255, xlogfNoCheck L1306-1324,
 * xexpfNoCheck L1347-1363 */
static inline float xlogf_vcore(float d)
{
    const int e = ilogbp1f_(d * 0.7071f);
    const float m = ldexpk(d, -e);
    float x = (-1.0f + m) / (1.0f + m);
    const float x2 = x * x;
    float t = 0.2371599674224853515625f;
    t = t * x2 + 0.285279005765914916992188f;
    t = t * x2 + 0.400005519390106201171875f;
    t = t * x2 + 0.666666567325592041015625f;
    t = t * x2 + 2.0f;
    return x * t + 0.693147180559945286226764f * (float)e;
}
static inline float xlogf_nocheck(float d) { return xlogf_vcore(d); }
static inline float xlogf_vector(float d)
{
    float x = xlogf_vcore(d);
    if (d == INFINITY) x = INFINITY;
    if (0.f > d) x = NAN;
    if (d == 0.f) x = -INFINITY;
    return x;
}
static inline float xexpf_nocheck(float d)
{
    const int q = (int)lrintf(d * R_LN2);
    float s = (float)q * -L2U + d;
    s = (float)q * -L2L + s;
    float u = 0.00136324646882712841033936f;
    u = u * s + 0.00836596917361021041870117f;
    u = u * s + 0.0416710823774337768554688f;
    u = u * s + 0.166665524244308471679688f;
    u = u * s + 0.499999850988388061523438f;
    u = 1.0f + ((s * s) * u + s);
    return ldexpk(u, q);
}

/* xcbrtf, sleef.h L966-991 */
static inline float xcbrtf_scalar(float d)
{
    float x, y, q = 1.0f;
    int e, r;
    e = ilogbp1f_(d);
    d = ldexpk_scalar(d, -e);
    r = (e + 6144) % 3;
    q = (r == 1) ? 1.2599210498948731647672106f : q;
    q = (r == 2) ? 1.5874010519681994747517056f : q;
    q = ldexpk_scalar(q, (e + 6144) / 3 - 2048);
    q = i2f(f2i(q) ^ (f2i(d) & (int32_t)0x80000000));      /* mulsignf */
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

/* pow_F(a, b) = xexpf(b * xlogf(a)), sleef.h L29; xlin2log sleef.h L1303-1307 */
static inline float pow_F_scalar(float a, float b) { return xexpf_scalar(b * xlogf_scalar(a)); }
static inline float xlin2log_scalar(float x, float base) { return xlogf_scalar(x * (base - 1.f) + 1.f) / xlogf_scalar(base); }
#endif
