"""oracle/_ref, tone-curve translation unit (used by build_ref.py).  TEST INFRASTRUCTURE ONLY.

Hosts the reference's DEFAULT tone-curve path -- ToneCurveParams::TcMode::NEUTRAL, procparams.cc L1585 -- compiled from the
sources where they lie:
  curves.h            class Curve, DiagonalCurve, FlatCurve, curves::setLutVal
  curves.cc           Curve::{Curve, AddPolygons, fillDyByDx, fillHash, ...}, ToneCurve::Set, NeutralToneCurve::BatchApply
  diagonalcurves.cc   everything (DiagonalCurve)
  flatcurves.cc       everything (FlatCurve)
  iptonecurve.cc      ContrastCurve, expand_range, satcurve_lut, SatCurveRemap, apply_satcurve, DoubleCurve and the curve
                      assembly of ImProcFunctions::toneCurve (the `expand` / `adjust` lambdas, L601-650)
  color.cc / color.h  XYZ_D50_to_D65, XYZ_D65_to_D50, PQ, PQ_inv, xyz2jzazbz, jzazbz2xyz, yuv2hsl, hsl2yuv, filmlike_clip, the
                      rgb2jzczhz / jzczhz2rgb family, rgbxyz, xyz2rgb
Written here (not reference code): the Color table initialisation restating color.cc L236-256, L322-326, the ApplyState
constructor body without ICCStore (matrices come in as arguments) and the C wrappers.
"""
import os
import re

from build_ref import RT, cut_function

SHIM_TONE_TU = r"""
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <omp.h>
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "linalgebra.h"
#include "iccmatrices.h"
#include "array2D.h"
namespace rtengine { void guidedFilter(const array2D<float> &guide, const array2D<float> &src, array2D<float> &dst, int r, float epsilon, bool multithread, int subsampling=0);   // guidedfilter.h L27; defined in the guided translation unit
                     void guidedFilterLog(float base, array2D<float> &chan, int r, float eps, bool multithread, int subsampling=0); }

enum DiagonalCurveType { DCT_Empty = -1, DCT_Linear, DCT_Spline, DCT_Parametric, DCT_NURBS, DCT_CatmullRom, DCT_Unchanged };   // rtgui/mydiagonalcurve.h L31-40
enum FlatCurveType { FCT_Empty = -1, FCT_Linear, FCT_MinMaxCPoints, FCT_Unchanged };                                           // rtgui/myflatcurve.h L29-36
#define CURVES_MIN_POLY_POINTS 1000                                                                                            // rtgui/mycurve.h

namespace artref_tone {
using namespace rtengine;
typedef const float (*TMatrix)[3];      // iccstore.h L38
struct Settings { int verbose; };
static const Settings settings_ = {0};
static const Settings* settings = &settings_;

namespace {
typedef Vec3f A3;
#include "tone_color_anon.inc"
}

class Color {
public:
    constexpr static double sRGBGammaCurve = 2.4;
    static LUTf gamma2curve, igammatab_srgb, gammatab_srgb, jzazbz_pq_, jzazbz_pq_inv_;
#include "tone_color_h.inc"
#include "tone_color_yuv.inc"
#include "tone_color_gamma.inc"
    static void init()
    {   // color.cc L236-256, L322-326
        if (gammatab_srgb) return;
        const int maxindex = 65536;
        igammatab_srgb(maxindex, 0); gammatab_srgb(maxindex, 0); jzazbz_pq_(maxindex, 0); jzazbz_pq_inv_(maxindex, 0);
        for (int i = 0; i < maxindex; i++) gammatab_srgb[i] = gamma2(i / 65535.0);
        gammatab_srgb *= 65535.f;
        gamma2curve.share(gammatab_srgb, LUT_CLIP_BELOW | LUT_CLIP_ABOVE);
        for (int i = 0; i < maxindex; i++) igammatab_srgb[i] = igamma2(i / 65535.0);
        igammatab_srgb *= 65535.f;
        for (int i = 0; i < maxindex; ++i) { jzazbz_pq_[i] = PQ(float(i) / 65535.f); jzazbz_pq_inv_[i] = PQ_inv(float(i) / 65535.f); }
    }
    static void rgbxyz (float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb[3][3]);
    static void xyz2rgb (float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz[3][3]);
    static void filmlike_clip(float *r, float *g, float *b, float Lmax);
    static void yuv2hsl(float u, float v, float &h, float &s);
    static void hsl2yuv(float h, float s, float &u, float &v);
    static void xyz2jzazbz(float X, float Y, float Z, float &Jz, float &az, float &bz);
    static void jzazbz2xyz(float Jz, float az, float bz, float &X, float &Y, float &Z);
};
LUTf Color::gamma2curve, Color::igammatab_srgb, Color::gammatab_srgb, Color::jzazbz_pq_, Color::jzazbz_pq_inv_;
#include "tone_color_cc.inc"

#include "tone_curve_classes.inc"
#include "tone_curve_base.inc"
namespace { inline double CLIPD(double d) { return std::max(d, 0.0); } }
#include "tone_diag.inc"
#include "tone_flat.inc"

namespace curves {
#include "tone_setlutval.inc"
}
class ToneCurve { public: LUTf lutToneCurve; float whitecoeff; float whitept; const Curve* curve; ToneCurve() : whitecoeff(1.f), whitept(65535.f), curve(nullptr) {}
    void Set(const Curve &pCurve, float whitecoeff=1.f); };
#include "tone_toneset.inc"
class NeutralToneCurve: public ToneCurve {
public:
    struct ApplyState {
        const Curve *basecurve;
        float ws[3][3];
        float iws[3][3];
        Mat33<float> to_work;
        Mat33<float> to_out;
        float rhue, bhue, yhue, rrange, brange, yrange;
        // the constructor of curves.cc L854-888 with the matrices ICCStore would return passed in; om = nullptr is the
        // "no matrix for this output profile" branch (identity)
        ApplyState(const float work[3][3], const float iwork[3][3], const float (*om_)[3], const Curve* base)
        {
            basecurve = base;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { ws[i][j] = work[i][j]; iws[i][j] = iwork[i][j]; }
            if (om_) {
                Mat33<float> om(om_);
                auto iom = inverse(om);
                to_out = dot_product(iom, Mat33<float>(work));
                to_work = dot_product(Mat33<float>(iwork), om);
            } else { to_out = identity<float>(); to_work = identity<float>(); }
            float j, c;
            const auto hws = xyz_rec2020;
            Color::rgb2jzczhz(1, 0, 0, j, c, rhue, hws);
            Color::rgb2jzczhz(0, 0, 1, j, c, bhue, hws);
            Color::rgb2jzczhz(1, 1, 0, j, c, yhue, hws);
            float ohue;
            Color::rgb2jzczhz(1, 0.5, 0, j, c, ohue, hws);
            yrange = std::abs(ohue - yhue) * 0.8f;
            rrange = std::abs(ohue - rhue);
            brange = rrange;
        }
    };
    void BatchApply(const size_t start, const size_t end, float *r, float *g, float *b, const ApplyState &state) const;
};
#include "tone_batchapply.inc"

struct Plane { float* base; int W; float& operator()(int y, int x) { return base[(size_t)y * W + x]; } };
struct Imagefloat { int width, height; Plane r, g, b; int getWidth() const { return width; } int getHeight() const { return height; } };
struct WsHolder { const float (*ws)[3]; const float (*iws)[3]; };
static WsHolder g_ws;
struct ICCStore { static ICCStore* getInstance() { static ICCStore s; return &s; }
    TMatrix workingSpaceMatrix(const std::string&) { return g_ws.ws; } TMatrix workingSpaceInverseMatrix(const std::string&) { return g_ws.iws; } };
namespace Glib { typedef std::string ustring; }

namespace {
#include "tone_iptonecurve.inc"
}

// the curve assembly of ImProcFunctions::toneCurve for one (possibly contrast-prefixed) curve pair, iptonecurve.cc L596-650
struct Built {
    std::unique_ptr<Curve> ccurve; std::unique_ptr<DiagonalCurve> t1, t2; std::unique_ptr<DoubleCurve> d, dc; Curve* tcurve;
};
static void build(Built& B, const double* c1, int n1, const double* c2, int n2, int contrast, float whitept, double scale, double pivot_gray)
{
    struct { struct { std::vector<double> curve, curve2; } toneCurve; } P; auto* params = &P;
    P.toneCurve.curve.assign(c1, c1 + n1); P.toneCurve.curve2.assign(c2, c2 + n2);
    if (contrast) {        // get_contrast_curve, L335-349 (pivot_gray = logenc.enabled ? targetGray / 100 : 0.18)
        const double pivot = pivot_gray / whitept;
        const double c = std::pow(std::abs(contrast) / 100.0, 1.5) * 16.0;
        const double b = contrast > 0 ? (1 + c) : 1.0 / (1 + c);
        const double a = std::log((std::exp(std::log(b) * pivot) - 1) / (b - 1)) / std::log(pivot);
        B.ccurve.reset(new ContrastCurve(a, b, whitept));
    }
#include "tone_assembly.inc"
    B.t2.reset(new DiagonalCurve(adjust(params->toneCurve.curve2), CURVES_MIN_POLY_POINTS / max(int(scale), 1)));
    B.t1.reset(new DiagonalCurve(adjust(params->toneCurve.curve), CURVES_MIN_POLY_POINTS / max(int(scale), 1)));
    B.d.reset(new DoubleCurve(*B.t1, *B.t2));
    B.tcurve = B.d.get();
    if (B.ccurve) { B.dc.reset(new DoubleCurve(*B.ccurve, *B.d)); B.tcurve = B.dc.get(); }
}

struct PolyPeek : public DiagonalCurve { using DiagonalCurve::poly_x; using DiagonalCurve::poly_y; using DiagonalCurve::kind; };

extern "C" {
// LUT of ToneCurve::Set for the single-curve case (curveMode == curveMode2): lut[65536]; which = 0 the whole chain, 1 curve 1
// alone, 2 curve 2 alone, 3 the contrast curve alone (the two-curve mode's separate passes); returns 1 when that curve is
// the identity (the reference then skips the pass), else 0
int artref_tone_build_lut(const double* c1, int n1, const double* c2, int n2, int contrast, float whitept, double scale, int which, float* lut)
{
    Color::init();
    Built B; build(B, c1, n1, c2, n2, contrast, whitept, scale, 0.18);
    const Curve* c = which == 0 ? B.tcurve : which == 1 ? (Curve*)B.t1.get() : which == 2 ? (Curve*)B.t2.get() : (Curve*)B.ccurve.get();
    if (!c) return 1;
    ToneCurve tc; tc.Set(*c, whitept);
    for (int i = 0; i < 65536; ++i) lut[i] = tc.lutToneCurve[i];
    return c->isIdentity() ? 1 : 0;
}
// the polyline DiagonalCurve::getVal(DCT_CatmullRom) searches (which = 1 | 2); returns the point count (0 = identity curve), fills up to cap
int artref_tone_polyline(const double* c1, int n1, const double* c2, int n2, float whitept, double scale, int which, double* px, double* py, int cap)
{
    Color::init();
    Built B; build(B, c1, n1, c2, n2, 0, whitept, scale, 0.18);
    const PolyPeek* p = static_cast<const PolyPeek*>(which == 1 ? B.t1.get() : B.t2.get());
    if (p->kind == DCT_Empty) return 0;
    const int n = (int)p->poly_x.size();
    for (int i = 0; i < n && i < cap; ++i) { px[i] = p->poly_x[i]; py[i] = p->poly_y[i]; }
    return n;
}
// contrast curve parameters a, b of get_contrast_curve
int artref_tone_contrast_ab(int contrast, float whitept, double* ab)
{
    const double pivot = 0.18 / whitept;
    const double c = std::pow(std::abs(contrast) / 100.0, 1.5) * 16.0;
    const double b = contrast > 0 ? (1 + c) : 1.0 / (1 + c);
    ab[0] = std::log((std::exp(std::log(b) * pivot) - 1) / (b - 1)) / std::log(pivot); ab[1] = b;
    return 0;
}
double artref_tone_curve_eval(const double* c1, int n1, const double* c2, int n2, int contrast, float whitept, double scale, double t)
{
    Color::init();
    Built B; build(B, c1, n1, c2, n2, contrast, whitept, scale, 0.18);
    return B.tcurve->getVal(t);
}
// apply_tc(NEUTRAL) over the whole single-curve chain: R, G, B in place.  om9: the output profile's matrix or NULL
int artref_tone_neutral(float* R, float* G, float* B_, int W, int H, const double* c1, int n1, const double* c2, int n2, int contrast, float whitept, double scale,
                        const float* ws9, const float* iws9, const float* om9)
{
    Color::init();
    Built B; build(B, c1, n1, c2, n2, contrast, whitept, scale, 0.18);
    NeutralToneCurve tc; tc.Set(*B.tcurve, whitept);
    NeutralToneCurve::ApplyState state((const float (*)[3])ws9, (const float (*)[3])iws9, (const float (*)[3])om9, nullptr);
#pragma omp parallel for
    for (int y = 0; y < H; ++y) tc.BatchApply(0, W, R + (size_t)y * W, G + (size_t)y * W, B_ + (size_t)y * W, state);
    return 0;
}
// satcurve_lut: the 65536-entry table apply_satcurve reads at white point 1
int artref_tone_satlut(const double* sc, int ns, double scale, float* lut)
{
    Color::init();
    std::vector<double> pts(sc, sc + ns);
    const FlatCurve satlcurve(pts, false, CURVES_MIN_POLY_POINTS / max(int(scale), 1));
    if (satlcurve.isIdentity()) return 1;
    LUTf sat; satcurve_lut(satlcurve, sat, 1.f);
    for (int i = 0; i < 65536; ++i) lut[i] = sat[i];
    return 0;
}
// apply_satcurve (white point 1 or not; saturation2 = sc2)
int artref_tone_satcurve(float* R, float* G, float* B_, int W, int H, const double* sc, int ns, const double* sc2, int ns2, float whitept, double scale,
                         const float* ws9, const float* iws9)
{
    Color::init();
    std::vector<double> pts(sc, sc + ns), pts2(sc2, sc2 + ns2);
    const FlatCurve satlcurve(pts, false, CURVES_MIN_POLY_POINTS / max(int(scale), 1));
    const DiagonalCurve satccurve(pts2);
    if (satlcurve.isIdentity() && satccurve.isIdentity()) return 1;
    g_ws.ws = (const float (*)[3])ws9; g_ws.iws = (const float (*)[3])iws9;
    Imagefloat im{W, H, {R, W}, {G, W}, {B_, W}};
    apply_satcurve(&im, satlcurve, satccurve, "", whitept, true);
    return 0;
}
// the polyline FlatCurve::getVal searches (flatcurves.cc L339-365): returns the point count (0 = FCT_Empty, the identity), fills up to cap
struct FlatPeek : public FlatCurve { using FlatCurve::poly_x; using FlatCurve::poly_y; using FlatCurve::dyByDx; };
int artref_flat_polyline(const double* pts, int npts, int periodic, int poly_pn, double* px, double* py, double* dy, int cap)
{
    std::vector<double> v(pts, pts + npts);
    const FlatCurve c(v, periodic != 0, poly_pn);
    const FlatPeek* p = static_cast<const FlatPeek*>(&c);
    if (c.isIdentity()) return 0;
    const int n = (int)p->poly_x.size();
    for (int i = 0; i < n && i < cap; ++i) { px[i] = p->poly_x[i]; py[i] = p->poly_y[i]; if (i < (int)p->dyByDx.size()) dy[i] = p->dyByDx[i]; }
    return n;
}
double artref_flat_getval(const double* pts, int npts, int periodic, int poly_pn, double t)
{
    std::vector<double> v(pts, pts + npts);
    const FlatCurve c(v, periodic != 0, poly_pn);
    return c.getVal(t);
}
}   // extern "C"

// ---- ImProcFunctions::hslEqualizer (iphsl.cc L29-221) over a stand-in Imagefloat whose setMode / normalize members are the
// reference's own loops (imagefloat.cc rgb_to_yuv L700-725, yuv_to_rgb L779-803, multiply L396-425)
namespace hsl {
struct Plane { float* base; int stride; float** ptrs; float& operator()(int y, int x) { return ptrs[y][x]; } };
class Imagefloat {
public:
    enum class Mode { RGB, YUV };
    int width, height; Plane r, g, b;
    float ws_[3][3]; vfloat vws_[3][3];
    Imagefloat(int w, int h, const float* R, const float* G, const float* B, const double* ws) : width(w), height(h)
    {
        const int st = (w + 3) / 4 * 4;
        Plane* pl[3] = {&r, &g, &b}; const float* src[3] = {R, G, B};
        for (int c = 0; c < 3; ++c) {
            void* p = nullptr; if (posix_memalign(&p, 64, sizeof(float) * (size_t)st * h)) abort();
            pl[c]->base = (float*)p; pl[c]->stride = st; pl[c]->ptrs = new float*[h];
            for (int y = 0; y < h; ++y) { pl[c]->ptrs[y] = pl[c]->base + (size_t)y * st; memcpy(pl[c]->ptrs[y], src[c] + (size_t)y * w, sizeof(float) * w); }
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { ws_[i][j] = float(ws[3 * i + j]); vws_[i][j] = F2V(float(ws[3 * i + j])); }
    }
    ~Imagefloat() { Plane* pl[3] = {&r, &g, &b}; for (int c = 0; c < 3; ++c) { free(pl[c]->base); delete[] pl[c]->ptrs; } }
    void store(float* R, float* G, float* B) { Plane* pl[3] = {&r, &g, &b}; float* dst[3] = {R, G, B};
        for (int c = 0; c < 3; ++c) for (int y = 0; y < height; ++y) memcpy(dst[c] + (size_t)y * width, pl[c]->ptrs[y], sizeof(float) * width); }
    int getWidth() const { return width; }
    int getHeight() const { return height; }
    void get_ws() {}
    void multiply(float factor, bool multithread);
    void normalizeFloatTo1(bool multithread);
    void normalizeFloatTo65535(bool multithread);
    void rgb_to_yuv(bool multithread);
    void yuv_to_rgb(bool multithread);
    void setMode(Mode m, bool multithread) { if (m == Mode::YUV) rgb_to_yuv(multithread); else yuv_to_rgb(multithread); }    // imagefloat.cc L620-650 for these two modes
};
#include "hsl_imagefloat.inc"
struct Whatever { float& v(int, int) { static float f; return f; } void fill(float) {} };
struct HslParams { std::vector<double> hCurve, sCurve, lCurve; int smoothing; bool enabled; };
struct Params { HslParams hsl; };

extern "C" int artref_hsl_equalizer(float* R, float* G, float* B, int W_, int H_, const double* ws9, const double* hc, int nh, const double* sc, int ns,
                                    const double* lc, int nl, int smoothing, double scale_)
{
    Color::init();
    Params P; const Params* params = &P;
    P.hsl.hCurve.assign(hc, hc + nh); P.hsl.sCurve.assign(sc, sc + ns); P.hsl.lCurve.assign(lc, lc + nl); P.hsl.smoothing = smoothing; P.hsl.enabled = true;
    Imagefloat im(W_, H_, R, G, B, ws9); Imagefloat* img = &im;
    const double scale = scale_;
    const bool multiThread = true;
    Whatever* editWhatever = nullptr;
#include "hsl_body.inc"
    img->setMode(Imagefloat::Mode::RGB, multiThread);       // what the next stage's setMode(RGB) does to the YUV image hslEqualizer leaves
    im.store(R, G, B);
    return 0;
}
// ImProcFunctions::blackAndWhite's two pixel loops (ipbw.cc L283-312, L343-362) with the mode changes of its colour cast (setMode(YUV) before the
// second loop, the next stage's setMode(RGB) after it); the mixer constants and the five tables are the caller's (host code of the reference)
extern "C" int artref_bw(float* R, float* G, float* B, int W_, int H_, const double* ws9, float bwr, float bwg, float bwb, float kcorec,
                         const float* gr, const float* gg, const float* gb, const float* ul, const float* vl)
{
    Imagefloat im(W_, H_, R, G, B, ws9); Imagefloat* img = &im;
    const int W = W_, H = H_;
    const bool multiThread = true;
    const bool hasgammabw = gr != nullptr;
    LUTf gamma_r, gamma_g, gamma_b;
    if (hasgammabw) {
        gamma_r(65536); gamma_g(65536); gamma_b(65536);
        for (int i = 0; i < 65536; ++i) { gamma_r[i] = gr[i]; gamma_g[i] = gg[i]; gamma_b[i] = gb[i]; }
    }
    vfloat bwr_v = F2V(bwr), bwg_v = F2V(bwg), bwb_v = F2V(bwb), kcorec_v = F2V(kcorec);
#pragma omp parallel for if (multiThread)
#include "bw_loop1.inc"
    if (ul) {
        img->setMode(Imagefloat::Mode::YUV, multiThread);
        LUTf ulut(65536), vlut(65536);
        for (int i = 0; i < 65536; ++i) { ulut[i] = ul[i]; vlut[i] = vl[i]; }
#pragma omp parallel for if (multiThread)
#include "bw_loop2.inc"
        img->setMode(Imagefloat::Mode::RGB, multiThread);
    }
    im.store(R, G, B);
    return 0;
}
}   // namespace hsl
// ---- ImProcFunctions::toneEqualizer (iptoneequalizer.cc L68-371): tone_eq() verbatim over stubs for what its colour-map preview branch names (lcms2, never executed here)
namespace toneeq {
typedef void* cmsHPROFILE; typedef void* cmsHTRANSFORM;
enum { TYPE_RGB_FLT = 0, INTENT_RELATIVE_COLORIMETRIC = 0, cmsFLAGS_NOOPTIMIZE = 0, cmsFLAGS_NOCACHE = 0 };
static cmsHTRANSFORM cmsCreateTransform(cmsHPROFILE, int, cmsHPROFILE, int, int, int) { abort(); }
static void cmsDoTransform(cmsHTRANSFORM, void*, void*, int) { abort(); }
static void cmsDeleteTransform(cmsHTRANSFORM) {}
struct Mutex { void lock() {} void unlock() {} };
static Mutex lcmsMutex_, *lcmsMutex = &lcmsMutex_;
static float g_wsm[3][3];
struct ICCStore {
    static ICCStore* getInstance() { static ICCStore s; return &s; }
    TMatrix workingSpaceMatrix(const std::string&) const { return g_wsm; }
    cmsHPROFILE getsRGBProfile() const { return nullptr; }
    cmsHPROFILE workingSpace(const std::string&) const { return nullptr; }
};
struct ToneEqualizerParams { bool enabled; std::array<int, 5> bands; int regularization; bool show_colormap; double pivot; };    // procparams.h L848-853
namespace {
static const std::vector<std::array<float, 3>> colormap;      // iptoneequalizer.cc L53-66: the preview colour map, not used here
#include "toneeq_body.inc"
}
// ImProcFunctions::toneEqualizer (L343-371) on three dense planes: multiply(gain), tone_eq, multiply(1 / gain)
extern "C" int artref_tone_equalizer(float* R_, float* G_, float* B_, int W, int H, const double* ws9, const int* bands5, int regularization, double pivot, double scale)
{
    for (int i = 0; i < 9; ++i) (&g_wsm[0][0])[i] = (float)ws9[i];
    ToneEqualizerParams pp; pp.enabled = true; for (int i = 0; i < 5; ++i) pp.bands[i] = bands5[i]; pp.regularization = regularization; pp.show_colormap = false; pp.pivot = pivot;
    const float gain = 1.f / 65535.f * std::pow(2.f, -pp.pivot);
    // rows 16-byte aligned like the reference's Imagefloat allocation (iimage.h L653-673): tone_eq's vector loop uses aligned loads / stores
    const int st = (W + 3) / 4 * 4;
    float* in[3] = {R_, G_, B_};
    float* buf[3]; float** rows[3];
    for (int c = 0; c < 3; ++c) {
        void* p = nullptr; if (posix_memalign(&p, 64, sizeof(float) * (size_t)st * H)) abort();
        buf[c] = (float*)p; rows[c] = new float*[H];
        for (int i = 0; i < H; ++i) { rows[c][i] = buf[c] + (size_t)i * st; for (int x = 0; x < W; ++x) rows[c][i][x] = in[c][(size_t)i * W + x] * gain; }      // Imagefloat::multiply(gain)
    }
    {
        array2D<float> R(W, H, rows[0], ARRAY2D_BYREFERENCE), G(W, H, rows[1], ARRAY2D_BYREFERENCE), B(W, H, rows[2], ARRAY2D_BYREFERENCE);
        tone_eq(R, G, B, pp, "", scale, true, false, nullptr);
    }
    const float back = 1.f / gain;
    for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < H; ++i) for (int x = 0; x < W; ++x) in[c][(size_t)i * W + x] = rows[c][i][x] * back;
        free(buf[c]); delete[] rows[c];
    }
    return 0;
}
}   // namespace toneeq

// ---- ImProcFunctions::softLight (ipsoftlight.cc L29-81): the reference's sl() and its table + apply loop
namespace softlight {
namespace {
#include "softlight_sl.inc"
}
struct SoftLightParams { bool enabled; int strength; };
struct Params { SoftLightParams softlight; };
extern "C" int artref_softlight(float* R, float* G, float* B_, int W, int H, int strength, float* lut_out)
{
    Color::init();
    Params P; P.softlight.enabled = true; P.softlight.strength = strength; const Params* params = &P;
    Imagefloat im{W, H, {R, W}, {G, W}, {B_, W}}; Imagefloat* rgb = &im;
#include "softlight_body.inc"
    if (lut_out) for (int i = 0; i < 65536; ++i) lut_out[i] = f[i];
    return 0;
}
}   // namespace softlight
extern "C" {
int artref_tone_tables(float* pq, float* pq_inv, float* gamma2curve)
{
    Color::init();
    for (int i = 0; i < 65536; ++i) { if (pq) pq[i] = Color::jzazbz_pq_[i]; if (pq_inv) pq_inv[i] = Color::jzazbz_pq_inv_[i]; if (gamma2curve) gamma2curve[i] = Color::gamma2curve[i]; }
    return 0;
}
}
}  // namespace artref_tone
"""


def between(text, start_regex, end_regex):
    m0 = re.search(start_regex, text, flags=re.M)
    if not m0:
        raise RuntimeError("anchor %r not found" % start_regex)
    m1 = re.search(end_regex, text[m0.start():], flags=re.M)
    if not m1:
        raise RuntimeError("anchor %r not found" % end_regex)
    return text[m0.start(): m0.start() + m1.start()]


def extract(sub):
    rd = lambda p: open(p, encoding="utf-8", errors="replace").read()
    cc, ch = os.path.join(RT, "color.cc"), os.path.join(RT, "color.h")
    cvh, cvc = os.path.join(RT, "curves.h"), os.path.join(RT, "curves.cc")
    ipt = os.path.join(RT, "iptonecurve.cc")
    w = lambda name, text: open(os.path.join(sub, name), "w").write(text)
    w("tone_color_anon.inc", "\n".join([
        cut_function(cc, r"^void XYZ_D50_to_D65\(float &X, float &Y, float &Z\)"),
        cut_function(cc, r"^void XYZ_D65_to_D50\(float &X, float &Y, float &Z\)"),
        cut_function(cc, r"^float PQ\(float X\)"), cut_function(cc, r"^float PQ_inv\(float X\)"),
        cut_function(cc, r"^inline void filmlike_clip_rgb_tone\(float \*r, float \*g, float \*b, const float L\)")]))
    htext = rd(ch)
    w("tone_color_h.inc", "\n".join([
        cut_function(ch, r"static inline double gamma2\(double x\)"), cut_function(ch, r"static inline double igamma2\(double x\)"),
        between(htext, r"^    template <class T>\n    static void rgb2jzazbz\(", r"^    static void xyz2oklab\(")]))
    w("tone_color_cc.inc", "\n".join([
        cut_function(cc, r"^void Color::rgbxyz \(float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb"),
        cut_function(cc, r"^void Color::xyz2rgb \(float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz"),
        cut_function(cc, r"^void Color::filmlike_clip\(float \*r, float \*g, float \*b, float Lmax\)"),
        cut_function(cc, r"^void Color::yuv2hsl\(float u, float v, float &h, float &s\)"),
        cut_function(cc, r"^void Color::hsl2yuv\(float h, float s, float &u, float &v\)"),
        cut_function(cc, r"^void Color::xyz2jzazbz\(float X, float Y, float Z, float &Jz, float &az, float &bz\)"),
        cut_function(cc, r"^void Color::jzazbz2xyz\(float Jz, float az, float bz, float &X, float &Y, float &Z\)")]))
    w("tone_curve_classes.inc", between(rd(cvh), r"^class Curve \{", r"^namespace curves \{"))
    w("tone_curve_base.inc", between(rd(cvc), r"^Curve::Curve \(\)", r"^void ToneCurve::Reset\(\)"))
    w("tone_diag.inc", between(rd(os.path.join(RT, "diagonalcurves.cc")), r"^DiagonalCurve::DiagonalCurve\(", r"^\} // namespace rtengine"))
    ftext = rd(os.path.join(RT, "flatcurves.cc"))
    w("tone_flat.inc", ftext[re.search(r"^FlatCurve::FlatCurve \(", ftext, flags=re.M).start(): ftext.rindex("}")])
    w("tone_setlutval.inc", cut_function(cvh, r"^inline void setLutVal\(const LUTf &lut, const Curve \*curve, float &val\)"))
    w("tone_toneset.inc", cut_function(cvc, r"^void ToneCurve::Set\(const Curve &pCurve, float whitecoeff\)"))
    w("tone_batchapply.inc", cut_function(cvc, r"^void NeutralToneCurve::BatchApply\("))
    w("tone_iptonecurve.inc", "\n".join([
        cut_function(ipt, r"^class ContrastCurve: public Curve") + ";",
        cut_function(ipt, r"^float expand_range\(float whitept, float x\)"),
        cut_function(ipt, r"^void satcurve_lut\(const FlatCurve &curve, LUTf &sat, float whitept\)"),
        cut_function(ipt, r"^class SatCurveRemap") + ";",
        cut_function(ipt, r"^void apply_satcurve\(Imagefloat \*rgb, const FlatCurve &curve"),
        cut_function(ipt, r"^class DoubleCurve: public Curve") + ";"]))
    w("tone_assembly.inc", between(rd(ipt), r"^        const auto expand =", r"^        DiagonalCurve tcurve2\("))
    w("tone_color_yuv.inc", "\n".join([
        "template <class T>\n" + cut_function(ch, r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"),
        "static " + cut_function(ch, r"vfloat rgbLuminance\(vfloat r, vfloat g, vfloat b, const vfloat workingspace\[3\]\[3\]\)"),
        "template <class T>\n" + cut_function(ch, r"static void rgb2yuv\(float r, float g, float b, float &Y, float &u, float &v, const T workingspace\[3\]\[3\]\)"),
        "template <class T>\n" + cut_function(ch, r"static void yuv2rgb\(float Y, float u, float v, float &r, float &g, float &b, const T workingspace\[3\]\[3\]\)"),
        cut_function(ch, r"static void rgb2yuv\(vfloat r, vfloat g, vfloat b, vfloat &Y, vfloat &u, vfloat &v, const vfloat workingspace\[3\]\[3\]\)"),
        cut_function(ch, r"static void yuv2rgb\(vfloat Y, vfloat u, vfloat v, vfloat &r, vfloat &g, vfloat &b, const vfloat workingspace\[3\]\[3\]\)")]))
    imf = os.path.join(RT, "imagefloat.cc")
    w("hsl_imagefloat.inc", "\n".join([
        cut_function(imf, r"^void Imagefloat::multiply\(float factor, bool multithread\)"),
        cut_function(imf, r"^void Imagefloat::normalizeFloatTo1\(bool multithread\)"),
        cut_function(imf, r"^void Imagefloat::normalizeFloatTo65535\(bool multithread\)"),
        cut_function(imf, r"^void Imagefloat::rgb_to_yuv\(bool multithread\)"),
        cut_function(imf, r"^void Imagefloat::yuv_to_rgb\(bool multithread\)")]))
    w("hsl_body.inc", between(rd(os.path.join(RT, "iphsl.cc")), r"^    img->setMode\(Imagefloat::Mode::YUV, multiThread\);", r"^\} // namespace rtengine")
      .rstrip().rstrip("}"))
    w("tone_color_gamma.inc", "\n".join([
        cut_function(ch, r"static inline float  gamma_srgb       \(float x\)"),
        cut_function(ch, r"static inline float  igamma_srgb      \(float x\)")]))
    from build_ref import cut_block
    ipbw = os.path.join(RT, "ipbw.cc")
    w("bw_loop1.inc", cut_block(ipbw, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < W-3; x \+= 4\) \{\s*vfloat r = LVF\(img->r\(y, x\)\);)"))
    w("bw_loop2.inc", cut_block(ipbw, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < W - 3; x \+= 4\) \{\s*vfloat yv = LVF\(img->g\(y, x\)\);)"))
    w("toneeq_body.inc", cut_function(os.path.join(RT, "iptoneequalizer.cc"), r"^void tone_eq\(array2D<float> &R, array2D<float> &G, array2D<float> &B, const ToneEqualizerParams &pp")
      .replace("const Glib::ustring &workingProfile", "const std::string &workingProfile"))
    ipsl = os.path.join(RT, "ipsoftlight.cc")
    w("softlight_sl.inc", cut_function(ipsl, r"^inline float sl\(float blend, float x\)"))
    w("softlight_body.inc", between(rd(ipsl), r"^    const float blend = params->softlight\.strength / 100\.f;", r"^\}\n\n\} // namespace rtengine"))
    w("shim_tone.cc", SHIM_TONE_TU)
    return os.path.join(sub, "shim_tone.cc")
