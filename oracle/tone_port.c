/*
 * oracle/tone_port.c -- CPU restatement of the reference's default tone-curve path.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates, per pixel,
 *   NeutralToneCurve::BatchApply  (reference rtengine/curves.cc L891-1037; ApplyState constructor L854-888) as apply_tc calls it for
 *                                 ToneCurveParams::TcMode::NEUTRAL -- the default curve mode (procparams.cc L1585) -- with a LINEAR base
 *                                 curve (iptonecurve.cc L89-99): gamut compression towards the output profile, luminance-preserving
 *                                 rescale, the curve (curves::setLutVal, curves.h L224-231: the 65536-entry LUT ToneCurve::Set fills,
 *                                 or Curve::getVal above it), hue twist and desaturation in JzCzhz;
 *   apply_satcurve                (rtengine/iptonecurve.cc L398-441) at white point 1 (the LUT branch) with an identity second curve;
 *   Color::rgb2jzczhz / jzczhz2rgb (rtengine/color.h L1764-1804, color.cc L6690-6742, PQ / PQ_inv L67-85, the Bradford matrices
 *                                 L37-64) and sleef's xatan2f / xsincosf / xexpf (rtengine/sleef.h).
 * The curves themselves stay host objects in the reference (DiagonalCurve / FlatCurve, rtengine/diagonalcurves.cc, flatcurves.cc):
 * what arrives here is what they evaluate to -- the LUT, and for samples above it the chain of Curve::getVal stages the
 * reference's DoubleCurve composes (iptonecurve.cc L525-551, L652-658): an identity, a Catmull-Rom polyline searched as
 * DiagonalCurve::getVal does (diagonalcurves.cc L511-522: nearest polyline point), or ContrastCurve (iptonecurve.cc L104-121).
 * Pinned against the reference's own functions compiled in place (oracle/_ref) in tests/test_oracle_tone.py.
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "sleef_port.h"

typedef struct artoracle_curve_stage {
    int kind;                 /* 0 identity, 1 Catmull-Rom polyline, 2 ContrastCurve */
    const double* poly_x; const double* poly_y; int n;
    double a, b, w;
} artoracle_curve_stage;

static float PQ_TAB[65536 + 1], PQI_TAB[65536 + 1];
static int tabs_ready = 0;

static float pq_(float X)
{   /* color.cc L67-74 */
    X = X < 1e-10f ? 1e-10f : X;
    const float XX = powf(X * 1e-4f, 0.1593017578125f);
    return powf((0.8359375f + 18.8515625f * XX) / (1 + 18.6875f * XX), 134.034375f);
}
static float pq_inv_(float X)
{   /* color.cc L77-84 */
    X = X < 1e-10f ? 1e-10f : X;
    const float XX = powf(X, 7.460772656268214e-03f);
    return 1e4f * powf((0.8359375f - XX) / (18.6875f * XX - 18.8515625f), 6.277394636015326f);
}
static void init_tabs(void)
{
#pragma omp critical(artoracle_tone_tabs)
    if (!tabs_ready) {
        for (int i = 0; i < 65536; ++i) { PQ_TAB[i] = pq_((float)i / 65535.f); PQI_TAB[i] = pq_inv_((float)i / 65535.f); }
        PQ_TAB[65536] = PQ_TAB[65535]; PQI_TAB[65536] = PQI_TAB[65535];
        tabs_ready = 1;
    }
}
/* LUT<float>::operator[](float), LUT.h L437-459 */
static inline float lutf(const float* data, int size, int clip, float index)
{
    int idx = (int)index;
    if (index < 0.f || index != index) {
        if (clip & 1) return data[0];
        idx = 0;
    } else if (index > (float)(size - 2)) {
        if (clip & 2) return data[size - 1];
        idx = size - 2;
    }
    const float diff = index - (float)idx;
    const float p1 = data[idx];
    const float p2 = data[idx + 1] - p1;
    return p1 + p2 * diff;
}
static inline float get_pq(float x) { return (x >= 0.f && x <= 1.f) ? lutf(PQ_TAB, 65536, 0, x * 65535.f) : pq_(x); }
static inline float get_pq_inv(float x) { return (x >= 0.f && x <= 1.f) ? lutf(PQI_TAB, 65536, 0, x * 65535.f) : pq_inv_(x); }

static inline float mulsignf__(float x, float y)
{
    union { float f; uint32_t u; } a, b;
    a.f = x; b.f = y;
    a.u ^= (b.u & 0x80000000u);
    return a.f;
}
static float atan2kf__(float y, float x)
{   /* sleef.h L1155-1177 */
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
static float xatan2f__(float y, float x)
{   /* sleef.h L1179-1188 */
    const float PI_F = (float)3.14159265358979323846;
    float r = atan2kf__(fabsf(y), x);
    r = mulsignf__(r, x);
    if (isinf(x) || x == 0) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .5f)) : 0);
    if (isinf(y)) r = PI_F / 2 - (isinf(x) ? (copysignf(1.f, x) * (float)(PI_F * .25f)) : 0);
    if (y == 0) r = (copysignf(1.f, x) == -1 ? PI_F : 0);
    return (x != x) || (y != y) ? NAN : mulsignf__(r, y);
}
static void xsincosf__(float d, float* sn, float* cs)
{   /* sleef.h L1048-1052 -> sleefsseavx.h L1051-1101 (an SSE2 build takes the vector form for the scalar call) */
    const float A = 0.78515625f * 2, B = 0.00024127960205078125f * 2, C = 6.3329935073852539062e-07f * 2, D = 4.9604681473525147339e-10f * 2;
    const int q = (int)lrintf(d * (float)(2.0 / 3.14159265358979323846));      /* cvtps2dq: nearest even */
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
    if (isinf(d)) { x = NAN; y = NAN; }
    *sn = x; *cs = y;
}

static inline void mat3(const float* m, float a, float b, float c, float* x, float* y, float* z)
{   /* Color::rgbxyz / xyz2rgb (color.cc L833-838) and dot_product(Mat33, Vec3) (linalgebra.h L227-239): left to right */
    *x = m[0] * a + m[1] * b + m[2] * c;
    *y = m[3] * a + m[4] * b + m[5] * c;
    *z = m[6] * a + m[7] * b + m[8] * c;
}
static const float D50_D65[9] = {0.9555766f, -0.0230393f, 0.0631636f, -0.0282895f, 1.0099416f, 0.0210077f, 0.0122982f, -0.0204830f, 1.3299098f};
static const float D65_D50[9] = {1.0478112f, 0.0228866f, -0.0501270f, 0.0295424f, 0.9904844f, -0.0170491f, -0.0092345f, 0.0150436f, 0.7521316f};

static void xyz2jzazbz(float X, float Y, float Z, float* Jz, float* az, float* bz)
{   /* color.cc L6706-6722 */
    float x, y, z;
    mat3(D50_D65, X, Y, Z, &x, &y, &z);
    const float Lp = get_pq(0.674207838f * x + 0.382799340f * y - 0.047570458f * z);
    const float Mp = get_pq(0.149284160f * x + 0.739628340f * y + 0.083327300f * z);
    const float Sp = get_pq(0.070941080f * x + 0.174768000f * y + 0.670970020f * z);
    const float Iz = 0.5f * (Lp + Mp);
    *az = 3.524000f * Lp - 4.066708f * Mp + 0.542708f * Sp;
    *bz = 0.199076f * Lp + 1.096799f * Mp - 1.295875f * Sp;
    *Jz = (0.44f * Iz) / (1.f - 0.56f * Iz) - 1.6295499532821566e-11f;
}
static void jzazbz2xyz(float Jz, float az, float bz, float* X, float* Y, float* Z)
{   /* color.cc L6724-6742 */
    Jz = Jz + 1.6295499532821566e-11f;
    const float Iz = Jz / (0.44f + 0.56f * Jz);
    const float L = get_pq_inv(Iz + 1.386050432715393e-1f * az + 5.804731615611869e-2f * bz);
    const float M = get_pq_inv(Iz - 1.386050432715393e-1f * az - 5.804731615611891e-2f * bz);
    const float S = get_pq_inv(Iz - 9.601924202631895e-2f * az - 8.118918960560390e-1f * bz);
    const float x = +1.661373055774069e+00f * L - 9.145230923250668e-01f * M + 2.313620767186147e-01f * S;
    const float y = -3.250758740427037e-01f * L + 1.571847038366936e+00f * M - 2.182538318672940e-01f * S;
    const float z = -9.098281098284756e-02f * L - 3.127282905230740e-01f * M + 1.522766561305260e+00f * S;
    mat3(D65_D50, x, y, z, X, Y, Z);
}
static void rgb2jzczhz(float R, float G, float B, float* Jz, float* cz, float* hz, const float* ws)
{   /* color.h L1791-1796: rgbxyz, xyz2jzazbz, jzazbz2jzch = yuv2hsl(bz, az, h, c) (color.cc L6691-6695) */
    float X, Y, Z, az, bz;
    mat3(ws, R, G, B, &X, &Y, &Z);
    xyz2jzazbz(X, Y, Z, Jz, &az, &bz);
    *cz = sqrtf(bz * bz + az * az);
    *hz = xatan2f__(bz, az);
}
static void jzczhz2rgb(float Jz, float cz, float hz, float* R, float* G, float* B, const float* iws)
{   /* color.h L1799-1804: jzch2jzazbz = hsl2yuv(h, c, bz, az) (color.cc L6698-6703) */
    float sn, cs, X, Y, Z;
    xsincosf__(hz, &sn, &cs);
    const float bz = cz * sn, az = cz * cs;
    jzazbz2xyz(Jz, az, bz, &X, &Y, &Z);
    mat3(iws, X, Y, Z, R, G, B);
}

static inline void clip_tone(float* r, float* g, float* b, float L)
{   /* color.cc L6650-6658 */
    const float r_ = *r > L ? L : *r;
    const float b_ = *b > L ? L : *b;
    const float g_ = b_ + ((r_ - b_) * (*g - *b) / (*r - *b));
    *r = r_; *g = g_; *b = b_;
}
static void filmlike_clip(float* r, float* g, float* b, float L)
{   /* color.cc L6662-6688 */
    if (*r >= *g) {
        if (*g > *b) clip_tone(r, g, b, L);
        else if (*b > *r) clip_tone(b, r, g, L);
        else if (*b > *g) clip_tone(r, b, g, L);
        else { *r = *r > L ? L : *r; *g = *g > L ? L : *g; *b = *g; }
    } else {
        if (*r >= *b) clip_tone(g, r, b, L);
        else if (*b > *g) clip_tone(b, g, r, L);
        else clip_tone(g, b, r, L);
    }
}

/* Curve::getVal of the composed curve (DoubleCurve: c2(c1(t)), iptonecurve.cc L530-533), stage by stage */
static double curve_eval(const artoracle_curve_stage* st, int n, double t)
{
    for (int s = 0; s < n; ++s) {
        if (st[s].kind == 1) {          /* DiagonalCurve::getVal, DCT_CatmullRom: std::lower_bound + nearest neighbour, CLIPD */
            const double* px = st[s].poly_x; const int np = st[s].n;
            int lo = 0, hi = np;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (px[mid] < t) lo = mid + 1; else hi = mid; }
            if (lo == np) { t = st[s].poly_y[np - 1]; continue; }
            int d = lo;
            if (lo + 1 < np && t - px[lo] > px[lo + 1] - t) ++d;
            const double v = st[s].poly_y[d];
            t = v < 0.0 ? 0.0 : v;      /* std::max(d, 0.0) */
        } else if (st[s].kind == 2) {   /* ContrastCurve::getVal: lin2log(pow(LIM(x, 0, w) / w, a), b) * w */
            const double w = st[s].w;
            double x = w < t ? w : t;    /* min(val, high) = high < val ? high : val */
            x = x < 0.0 ? 0.0 : x;       /* max(low, .) */
            const double p = pow(x / w, st[s].a);
            t = log(p * (st[s].b - 1.0) + 1.0) / log(st[s].b) * w;
        }
    }
    return t;
}

static inline float lim01(float a) { const float m = 1.f < a ? 1.f : a; return 0.f < m ? m : 0.f; }     /* max(T(0), min(a, T(1))) */
static inline float gaussf(float x, float b, float c) { return xexpf_scalar(-((x - b) * (x - b)) / (2 * (c * c))); }

typedef struct { float rhue, bhue, yhue, rrange, brange, yrange; } hue_consts;
static hue_consts hues(void)
{   /* curves.cc L877-887 */
    static const float rec2020[9] = {0.6734241f, 0.1656411f, 0.1251286f, 0.2790177f, 0.6753402f, 0.0456377f, -0.0019300f, 0.0299784f, 0.7973330f};
    hue_consts h; float j, c, ohue;
    rgb2jzczhz(1, 0, 0, &j, &c, &h.rhue, rec2020);
    rgb2jzczhz(0, 0, 1, &j, &c, &h.bhue, rec2020);
    rgb2jzczhz(1, 1, 0, &j, &c, &h.yhue, rec2020);
    rgb2jzczhz(1, 0.5f, 0, &j, &c, &ohue, rec2020);
    h.yrange = fabsf(ohue - h.yhue) * 0.8f;
    h.rrange = fabsf(ohue - h.rhue);
    h.brange = h.rrange;
    return h;
}
void artoracle_tone_hues(float* out6)
{
    init_tabs();
    const hue_consts h = hues();
    out6[0] = h.rhue; out6[1] = h.bhue; out6[2] = h.yhue; out6[3] = h.rrange; out6[4] = h.brange; out6[5] = h.yrange;
}

/* NeutralToneCurve::BatchApply over a frame.  lut: ToneCurve::lutToneCurve (65536); whitecoeff: the white point; stages: the
 * curve above the LUT (NULL / 0 = `!curve`, the LUT serves every sample); ws / iws: working-space matrices as float[9];
 * to_out / to_work: ApplyState's gamut-compression matrices (NULL = identity, no matrix for the output profile) */
int artoracle_tone_neutral(float* R, float* G, float* B, int W, int H, const float* lut, float whitecoeff, const artoracle_curve_stage* stages, int nstages,
                           const float* ws, const float* iws, const float* to_out, const float* to_work)
{
    init_tabs();
    static const float IDENT[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!to_out) to_out = IDENT;
    if (!to_work) to_work = IDENT;
    const hue_consts hc = hues();
    const float whitept = 65535.f * whitecoeff;
    const float Lmax = whitept;
    static const float dl[3] = {1.1f, 1.2f, 1.5f}, th[3] = {0.85f, 0.75f, 0.95f};
    float s[3];
    for (int i = 0; i < 3; ++i) s[i] = (1.f - th[i]) / sqrtf(dl[i] - 1.f);
    const float PI_F_180 = (float)(3.14159265358979323846 / 180.0);
    const size_t n = (size_t)W * H;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        float rgb[3], jch[3];
        rgb[0] = R[i] / 65535.f; rgb[0] = rgb[0] < 0.f ? 0.f : rgb[0];       /* std::max(x, 0.f) = (x < 0) ? 0 : x */
        rgb[1] = G[i] / 65535.f; rgb[1] = rgb[1] < 0.f ? 0.f : rgb[1];
        rgb[2] = B[i] / 65535.f; rgb[2] = rgb[2] < 0.f ? 0.f : rgb[2];
        rgb2jzczhz(rgb[0], rgb[1], rgb[2], &jch[0], &jch[1], &jch[2], ws);
        const float ilum = jch[0];
        float hue = jch[2];
        const float iY = (rgb[0] + rgb[1] + rgb[2]) / 3.f;
        {
            float a, b, c;
            a = 0.f; a += to_out[0] * rgb[0]; a += to_out[1] * rgb[1]; a += to_out[2] * rgb[2];
            b = 0.f; b += to_out[3] * rgb[0]; b += to_out[4] * rgb[1]; b += to_out[5] * rgb[2];
            c = 0.f; c += to_out[6] * rgb[0]; c += to_out[7] * rgb[1]; c += to_out[8] * rgb[2];
            rgb[0] = a; rgb[1] = b; rgb[2] = c;
        }
        float ac = rgb[0] < rgb[1] ? rgb[1] : rgb[0];      /* max(max(a, b), max(c)) */
        ac = ac < rgb[2] ? rgb[2] : ac;
        float d[3] = {0.f, 0.f, 0.f};
        const float aac = fabsf(ac);
        if (ac != 0.f) { d[0] = (ac - rgb[0]) / aac; d[1] = (ac - rgb[1]) / aac; d[2] = (ac - rgb[2]) / aac; }
        float cd[3];
        for (int k = 0; k < 3; ++k)
            cd[k] = d[k] < th[k] ? d[k] : s[k] * sqrtf(d[k] - th[k] + (s[k] * s[k]) / 4.0f) - s[k] * sqrtf((s[k] * s[k]) / 4.0f) + th[k];
        rgb[0] = ac - cd[0] * aac; rgb[1] = ac - cd[1] * aac; rgb[2] = ac - cd[2] * aac;
        {
            float a, b, c;
            a = 0.f; a += to_work[0] * rgb[0]; a += to_work[1] * rgb[1]; a += to_work[2] * rgb[2];
            b = 0.f; b += to_work[3] * rgb[0]; b += to_work[4] * rgb[1]; b += to_work[5] * rgb[2];
            c = 0.f; c += to_work[6] * rgb[0]; c += to_work[7] * rgb[1]; c += to_work[8] * rgb[2];
            rgb[0] = a; rgb[1] = b; rgb[2] = c;
        }
        const float oY = (rgb[0] + rgb[1] + rgb[2]) / 3.f;
        if (oY > 0.f) {
            const float f = iY / oY;
            rgb[0] *= f; rgb[1] *= f; rgb[2] *= f;
            filmlike_clip(&rgb[0], &rgb[1], &rgb[2], Lmax);
        }
        for (int j = 0; j < 3; ++j) {      /* curves::setLutVal */
            float nt = rgb[j] * 65535.f;
            if (nt <= 65535.f || !nstages) nt = lutf(lut, 65536, 3, nt < 0.f ? 0.f : nt);
            else nt = (float)(curve_eval(stages, nstages, nt / 65535.f) * 65535.f);
            rgb[j] = nt / 65535.f;
        }
        rgb2jzczhz(rgb[0], rgb[1], rgb[2], &jch[0], &jch[1], &jch[2], ws);
        float hue_shift = 15.f * PI_F_180 * gaussf(hue, hc.rhue, hc.rrange);
        hue_shift += -5.f * PI_F_180 * gaussf(hue, hc.bhue, hc.brange);
        hue_shift *= lim01((rgb[0] + rgb[1] + rgb[2]) / (3.f * whitecoeff));
        hue += hue_shift;
        float sat = jch[1];
        {
            const float olum = jch[0];
            float ccf = ilum > 1e-5f ? (1.f - (lim01((olum / ilum) - 1.f) * 0.2f)) : 1.f;
            ccf = lim01(ccf + 0.5f * gaussf(hue, hc.yhue, hc.yrange));
            sat *= ccf;
        }
        jzczhz2rgb(jch[0], sat, hue, &rgb[0], &rgb[1], &rgb[2], iws);
        float v;
        v = rgb[0] * 65535.f; v = whitept < v ? whitept : v; R[i] = 0.f < v ? v : 0.f;      /* LIM = max(low, min(val, high)) */
        v = rgb[1] * 65535.f; v = whitept < v ? whitept : v; G[i] = 0.f < v ? v : 0.f;
        v = rgb[2] * 65535.f; v = whitept < v ? whitept : v; B[i] = 0.f < v ? v : 0.f;
    }
    return 0;
}

/* apply_satcurve at white point 1 with an identity second curve: sat = satcurve_lut's table (65536 entries, LUT_CLIP_BELOW) */
int artoracle_tone_satcurve(float* R, float* G, float* B, int W, int H, const float* satlut, const float* ws, const float* iws)
{
    init_tabs();
    const size_t n = (size_t)W * H;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        float X, Y, Z, Jz, az, bz;
        mat3(ws, R[i] / 65535.f, G[i] / 65535.f, B[i] / 65535.f, &X, &Y, &Z);
        xyz2jzazbz(X, Y, Z, &Jz, &az, &bz);
        float cz = sqrtf(bz * bz + az * az);
        const float hz = xatan2f__(bz, az);
        const float s = lutf(satlut, 65536, 1, Y * 65535.f);
        cz *= s;
        float r, g, b;
        jzczhz2rgb(Jz, cz, hz, &r, &g, &b, iws);
        R[i] = r * 65535.f; G[i] = g * 65535.f; B[i] = b * 65535.f;
    }
    return 0;
}

/* sleef's xatan2f / xsincosf for the other ports (hsl_port.c) */
float artoracle_xatan2f(float y, float x) { return xatan2f__(y, x); }
void artoracle_xsincosf(float d, float* sn, float* cs) { xsincosf__(d, sn, cs); }
