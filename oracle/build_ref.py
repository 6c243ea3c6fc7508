#!/usr/bin/env python3
"""Build oracle/_ref/libartref*.so from the reference sources WHERE THEY LIE.

TEST INFRASTRUCTURE ONLY -- nothing in art_b200/ may import, link or execute this.

The reference (artpixls/ART, /root/reference) cannot be built whole here (glibmm,
gtkmm, lcms2, fftw3f, exiv2 ... are absent), and its hot-path
translation units cannot be compiled unmodified for the same reason.  What does
work is compiling the reference's own *function bodies*: this script locates each
hot-path function by its signature inside /root/reference/rtengine/*.cc, cuts the
body out by brace matching, writes it to oracle/_ref/src/ (git-ignored scratch,
never committed) and compiles it inside a ~60-line shim class that supplies only
the members the body touches (W, H, FC(), rawData, red/green/blue, initialGain,
border, plistener).  The reference's own headers (array2D.h, rt_math.h, sleef.h,
opthelper.h, helpersse2.h, median.h ...) are included directly from
/root/reference/rtengine.

Two libraries are produced:
  libartref.so      the bodies exactly as they stand in the reference
  libartref_det.so  the same plus the "zero the per-thread scratch at the start
                    of every tile" patch (a memset inserted at extraction time).
                    The stock code reuses calloc'ed per-thread scratch across
                    tiles, which makes a few pixels depend on the OpenMP
                    schedule; the _det build is the
                    canonical, schedule-independent oracle and the tests report
                    the difference between the two as the reference's own
                    self-noise floor.

Flags: g++ -std=c++11 -O3 -fopenmp -ffp-contract=off, default x86-64 => the
__SSE2__ code paths, which is what every x86-64 build of ART executes.

If /root/reference is absent (the GPU box) this script is a no-op: the prebuilt
.so files travel with the repo snapshot.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ART_REFERENCE", "/root/reference")
RT = os.path.join(REF, "rtengine")
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(OUT, "src")


def cut_function(path, signature_regex):
    """Return the text of the function whose header matches signature_regex
    (from the header through the matching closing brace)."""
    text = open(path, encoding="utf-8", errors="replace").read()
    m = re.search(signature_regex, text, flags=re.M)
    if not m:
        raise RuntimeError("signature %r not found in %s" % (signature_regex, path))
    i = text.index("{", m.end() - 1)
    depth = 0
    j = i
    in_line_comment = in_block_comment = False
    in_str = None
    while j < len(text):
        c = text[j]
        nxt = text[j + 1] if j + 1 < len(text) else ""
        if in_line_comment:
            if c == "\n":
                in_line_comment = False
        elif in_block_comment:
            if c == "*" and nxt == "/":
                in_block_comment = False
                j += 1
        elif in_str:
            if c == "\\":
                j += 1
            elif c == in_str:
                in_str = None
        elif c == "/" and nxt == "/":
            in_line_comment = True
        elif c == "/" and nxt == "*":
            in_block_comment = True
        elif c in "\"'":
            in_str = c
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[m.start(): j + 1]
        j += 1
    raise RuntimeError("unbalanced braces after %r in %s" % (signature_regex, path))


def cut_block(path, start_regex):
    """Return the statement that starts at start_regex and runs through the matching close of its first '{'."""
    text = open(path, encoding="utf-8", errors="replace").read()
    m = re.search(start_regex, text)
    if not m:
        raise RuntimeError("anchor %r not found in %s" % (start_regex, path))
    i = text.index("{", m.end() - 1)
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[m.start(): j + 1]
    raise RuntimeError("unbalanced braces")


def insert_before(body, anchor_regex, patch):
    m = re.search(anchor_regex, body)
    if not m:
        raise RuntimeError("patch anchor %r not found" % anchor_regex)
    return body[: m.start()] + patch + body[m.start():]


def insert_after(body, anchor_regex, patch):
    m = re.search(anchor_regex, body)
    if not m:
        raise RuntimeError("patch anchor %r not found" % anchor_regex)
    return body[: m.end()] + patch + body[m.end():]


SHIM_GLIBMM = r"""
// stand-in for <glibmm.h>: only Glib::ustring::compose is touched by the bodies
#pragma once
#include <string>
namespace Glib {
struct ustring : public std::string {
    ustring() {}
    ustring(const char* s) : std::string(s) {}
    ustring(const std::string& s) : std::string(s) {}
    template <class... A> static ustring compose(const ustring& f, A&&...) { return f; }
};
}
"""

SHIM_TU = r"""
// Shim translation unit: hosts reference function bodies cut out of
// /root/reference/rtengine at build time (see oracle/build_ref.py).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <iostream>
#include <memory>
#include <algorithm>
#include <omp.h>
#include "glibmm.h"
#include "array2D.h"
#include "rt_math.h"
#include "sleef.h"
#include "opthelper.h"
#include "median.h"

#define M(x) Glib::ustring(x)
#define BENCHFUN

namespace {
unsigned fc(const unsigned int cfa[2][2], int r, int c) { return cfa[r & 1][c & 1]; }
}

namespace rtengine {

struct ProgressListener {
    void setProgressStr(const Glib::ustring&) {}
    void setProgress(double) {}
};
struct StopWatch { StopWatch(const char*) {} };

struct RawImageSource {
    int W, H;
    unsigned filters;
    int border;
    double initialGain;
    ProgressListener* plistener;
    array2D<float> rawData, red, green, blue;

    RawImageSource(int w, int h, unsigned f, float** raw, float** r, float** g, float** b)
        : W(w), H(h), filters(f), border(4), initialGain(1.0), plistener(nullptr),
          rawData(w, h, raw, ARRAY2D_BYREFERENCE), red(w, h, r, ARRAY2D_BYREFERENCE),
          green(w, h, g, ARRAY2D_BYREFERENCE), blue(w, h, b, ARRAY2D_BYREFERENCE) {}

    unsigned FC(int row, int col) const
    {
        return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3);
    }
    void igv_interpolate(int, int) {}
    void rcd_demosaic();
    void amaze_demosaic_RT(int winx, int winy, int winw, int winh, const array2D<float> &rawData,
                           array2D<float> &red, array2D<float> &green, array2D<float> &blue);
    void border_interpolate2(int winw, int winh, int lborders, const array2D<float> &rawData,
                             array2D<float> &red, array2D<float> &green, array2D<float> &blue);
};

#include "rcd_body.inc"
#include "amaze_body.inc"
#include "border_body.inc"

// ---- colorSpaceConversion_ matrix branch (rawimagesource.cc L3197-3211): the reference's own loop
struct ImShim {
    int W, H; float *R, *G, *B; long s;
    int getHeight() const { return H; }
    int getWidth() const { return W; }
    float& r(int i, int j) { return R[(long)i * s + j]; }
    float& g(int i, int j) { return G[(long)i * s + j]; }
    float& b(int i, int j) { return B[(long)i * s + j]; }
};
// ---- scaleColors, Bayer branch (rawimagesource.cc L2742-2761): the reference's own row/col loop
struct RiShim { float get_optical_black(int, int) const { return 0.f; } };
void ref_scalecolors_bayer(int winx, int winy, int winw, int winh, unsigned filters, float** rawData,
                           const float* cblacksom, const float* scale_mul, float* tmpchmax)
{
    RiShim ri_; RiShim* ri = &ri_;
    const bool dyn_row_noise = false;
    auto FC = [filters](int row, int col) { return (int)((filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1)) & 3); };
    using std::max;
#include "scalecolors_loop.inc"
}

// ---- scaleColors, X-Trans branch (rawimagesource.cc L2807-2815): the reference's own row/col loop
struct RiXtShim { const int* xt; int XTRANSFC(int row, int col) const { return xt[(row % 6) * 6 + (col % 6)]; } };      // rawimage.h: xtrans[(row) % 6][(col) % 6]
void ref_scalecolors_xtrans(int winx, int winy, int winw, int winh, const int* xtrans36, float** rawData,
                            const float* cblacksom, const float* scale_mul, float* tmpchmax)
{
    RiXtShim ri_{xtrans36}; RiXtShim* ri = &ri_;
    using std::max;
#include "scalecolors_xtrans_loop.inc"
}

void ref_matrix_convert(ImShim* im, double mat[3][3], bool multithread)
{
    (void)multithread;
#include "csconv_loop.inc"
}

} // namespace rtengine

namespace {
struct Rows {
    float **raw, **r, **g, **b;
    Rows(int H, const float* raw_, long rs, float* r_, float* g_, float* b_, long os)
    {
        raw = new float*[H]; r = new float*[H]; g = new float*[H]; b = new float*[H];
        for (int i = 0; i < H; ++i) {
            raw[i] = const_cast<float*>(raw_) + (long)i * rs;
            r[i] = r_ + (long)i * os; g[i] = g_ + (long)i * os; b[i] = b_ + (long)i * os;
        }
    }
    ~Rows() { delete[] raw; delete[] r; delete[] g; delete[] b; }
};
}

extern "C" {

// strides in floats.  nthreads<=0 => leave OpenMP default.
int artref_rcd(int W, int H, unsigned filters, const float* raw, long raw_stride,
               float* r, float* g, float* b, long out_stride, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Rows rows(H, raw, raw_stride, r, g, b, out_stride);
    rtengine::RawImageSource src(W, H, filters, rows.raw, rows.r, rows.g, rows.b);
    src.rcd_demosaic();
    return 0;
}

int artref_amaze(int W, int H, unsigned filters, const float* raw, long raw_stride,
                 float* r, float* g, float* b, long out_stride,
                 double initialGain, int border, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    Rows rows(H, raw, raw_stride, r, g, b, out_stride);
    rtengine::RawImageSource src(W, H, filters, rows.raw, rows.r, rows.g, rows.b);
    src.initialGain = initialGain;
    src.border = border;
    src.amaze_demosaic_RT(0, 0, W, H, src.rawData, src.red, src.green, src.blue);
    return 0;
}

int artref_border_interpolate2(int W, int H, unsigned filters, int lborders, const float* raw, long raw_stride,
                               float* r, float* g, float* b, long out_stride)
{
    Rows rows(H, raw, raw_stride, r, g, b, out_stride);
    rtengine::RawImageSource src(W, H, filters, rows.raw, rows.r, rows.g, rows.b);
    src.border_interpolate2(W, H, lborders, src.rawData, src.red, src.green, src.blue);
    return 0;
}

// getImage gain/clip (rawimagesource.cc L957-971 at skip == 1) with the reference's own CLIP (rt_math.h),
// then the colorSpaceConversion_ matrix loop cut from the reference.  mat == NULL skips the matrix.
int artref_scale_convert(int W, int H, float* r, float* g, float* b, long stride,
                         const float* mul, int doClip, const double* mat)
{
    const float rm = mul[0], gm = mul[1], bm = mul[2];
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float rtot = r[(long)i * stride + j], gtot = g[(long)i * stride + j], btot = b[(long)i * stride + j];
            rtot *= rm; gtot *= gm; btot *= bm;
            if (doClip) { rtot = rtengine::CLIP(rtot); gtot = rtengine::CLIP(gtot); btot = rtengine::CLIP(btot); }
            r[(long)i * stride + j] = rtot; g[(long)i * stride + j] = gtot; b[(long)i * stride + j] = btot;
        }
    if (mat) {
        double m[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = mat[3 * i + j];
        rtengine::ImShim im{W, H, r, g, b, stride};
        rtengine::ref_matrix_convert(&im, m, true);
    }
    return 0;
}

int artref_scale_colors_bayer(int W, int H, unsigned filters, float* raw, long stride,
                              const float* cblacksom, const float* scale_mul, float* chmax)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (long)i * stride;
    float t[3] = {0.f, 0.f, 0.f};
    rtengine::ref_scalecolors_bayer(0, 0, W, H, filters, rows, cblacksom, scale_mul, t);
    chmax[0] = t[0]; chmax[1] = t[1]; chmax[2] = t[2];
    delete[] rows;
    return 0;
}

int artref_scale_colors_xtrans(int W, int H, const int* xtrans36, float* raw, long stride,
                               const float* cblacksom, const float* scale_mul, float* chmax)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (long)i * stride;
    float t[3] = {0.f, 0.f, 0.f};
    rtengine::ref_scalecolors_xtrans(0, 0, W, H, xtrans36, rows, cblacksom, scale_mul, t);
    chmax[0] = t[0]; chmax[1] = t[1]; chmax[2] = t[2];
    delete[] rows;
    return 0;
}

int artref_is_deterministic_build(void)
{
#ifdef ARTREF_DET
    return 1;
#else
    return 0;
#endif
}

int artref_max_threads(void) { return omp_get_max_threads(); }

} // extern "C"
"""


SHIM_GAUSS_TU = r"""
// Shim TU hosting the reference's gauss.cc from its first namespace to the end of the file
// (its #include lines are replaced: boxblur.h drags StopWatch.h -> settings.h -> procparams.h -> lcms2.h).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "rt_math.h"
#include "opthelper.h"
#include "alignedbuffer.h"
#include "gauss.h"
namespace rtengine {
// only reached through the `buffer != nullptr` variant (retinex), which the hot path never uses
template <class T, class A> void boxblur(T** src, A** dst, T* buffer, int radx, int rady, int W, int H) { std::abort(); }
}
#include "gauss_body.inc"

extern "C" int artref_gauss(const float* src, long sstride, float* dst, long dstride, int W, int H, double sigma, int inplace)
{
    float** d = new float*[H];
    float** s = inplace ? d : new float*[H];      // the reference tests `src != dst` on the row tables themselves
    for (int i = 0; i < H; ++i) { d[i] = dst + (long)i * dstride; if (!inplace) s[i] = const_cast<float*>(src) + (long)i * sstride; }
#pragma omp parallel
    gaussianBlur(s, d, W, H, sigma);
    if (!inplace) delete[] s;
    delete[] d;
    return 0;
}

// gaussianBlur(src, dst, W, H, sigma, nullptr, GAUSS_MULT (type 1) / GAUSS_DIV (type 2), divb); contiguous planes, src != dst.
// GAUSS_MULT blurs src in place on the way (gauss.cc L1496).
extern "C" int artref_gauss_ex(float* src, float* dst, float* divb, int W, int H, double sigma, int type)
{
    float** s = new float*[H];
    float** d = new float*[H];
    float** v = new float*[H];
    for (int i = 0; i < H; ++i) { s[i] = src + (long)i * W; d[i] = dst + (long)i * W; v[i] = divb ? divb + (long)i * W : nullptr; }
#pragma omp parallel
    gaussianBlur(s, d, W, H, sigma, nullptr, type == 1 ? GAUSS_MULT : GAUSS_DIV, v);
    delete[] s; delete[] d; delete[] v;
    return 0;
}
"""


SHIM_RESIZE_TU = r"""
// Shim TU hosting the reference's Lanczos resampler: Lanc and ImProcFunctions::Lanczos cut from ipresize.cc.
// Written here (not reference code): the Imagefloat / ImProcFunctions stand-ins (setMode / assignMode are no-ops: the Lab round
// trip around the resampler is not part of this unit) and the wrapper.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <omp.h>
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "alignedbuffer.h"
namespace rtengine {
struct ResizePlane { float** ptrs; };
class Imagefloat {
public:
    enum class Mode { RGB, XYZ, YUV, LAB };
    int W, H; ResizePlane r, g, b;
    int getWidth() const { return W; }
    int getHeight() const { return H; }
    Mode mode() const { return Mode::LAB; }
    void setMode(Mode, bool) {}
    void assignMode(Mode) {}
};
class ImProcFunctions { public: bool multiThread; void Lanczos(Imagefloat *src, Imagefloat *dst, float scale); };
namespace {
#include "resize_lanc.inc"
}
#include "resize_lanczos.inc"
}
extern "C" int artref_lanczos(const float* s0, const float* s1, const float* s2, int sW, int sH, float* d0, float* d1, float* d2, int dW, int dH, float scale)
{
    using namespace rtengine;
    auto rows = [](const float* p, int W, int H) { float** t = new float*[H]; for (int i = 0; i < H; ++i) t[i] = const_cast<float*>(p) + (size_t)i * W; return t; };
    // the reference reads g / r / b as L / a / b; plane k of the wrapper goes to the same slot on both sides
    Imagefloat src{sW, sH, {rows(s1, sW, sH)}, {rows(s0, sW, sH)}, {rows(s2, sW, sH)}};
    Imagefloat dst{dW, dH, {rows(d1, dW, dH)}, {rows(d0, dW, dH)}, {rows(d2, dW, dH)}};
    ImProcFunctions ipf; ipf.multiThread = true;
    ipf.Lanczos(&src, &dst, scale);
    delete[] src.r.ptrs; delete[] src.g.ptrs; delete[] src.b.ptrs; delete[] dst.r.ptrs; delete[] dst.g.ptrs; delete[] dst.b.ptrs;
    return 0;
}
"""


SHIM_GREENEQ_TU = r"""
// Shim TU hosting the reference's green equilibration: RawImageSource::green_equilibrate_global and ::green_equilibrate cut from
// green_equil_RT.cc.  Written here (not reference code): the RawImageSource stand-in (W, H, border, FC, the threshold functor's
// interface as rawimagesource.h L205-212 declares it) and the wrappers.
#include <math.h>
#include <cmath>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "array2D.h"
#include "rt_math.h"
#include "opthelper.h"
#define RawImageSource RawImageSourceGreenEq       /* other shim TUs define their own stand-in of this name: keep the inline members apart */
namespace rtengine {
struct RawImageSource {
    int W, H, border; unsigned filters;
    class GreenEqulibrateThreshold {
    public:
        explicit GreenEqulibrateThreshold(float thresh): thresh_(thresh) {}
        virtual ~GreenEqulibrateThreshold() {}
        virtual float operator()(int row, int column) const { return thresh_; }
    protected:
        const float thresh_;
    };
    unsigned FC(int row, int col) const { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
    void green_equilibrate_global(array2D<float> &rawData);
    void green_equilibrate(const GreenEqulibrateThreshold &greenthresh, array2D<float> &rawData);
};
#include "greeneq_body.inc"
}
namespace {
struct MapThreshold : rtengine::RawImageSource::GreenEqulibrateThreshold {
    const float* map; int W;
    MapThreshold(const float* m, int w) : GreenEqulibrateThreshold(0.f), map(m), W(w) {}
    float operator()(int row, int column) const override { return map[(size_t)row * W + column]; }
};
}
extern "C" int artref_green_equilibrate_global(float* raw, int W, int H, unsigned filters, int border)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (size_t)i * W;
    {
        rtengine::array2D<float> rd(W, H, rows, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImageSource s{W, H, border, filters};
        s.green_equilibrate_global(rd);
    }
    delete[] rows;
    return 0;
}
extern "C" int artref_green_equilibrate(float* raw, int W, int H, unsigned filters, float thresh, const float* thresh_map)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (size_t)i * W;
    {
        rtengine::array2D<float> rd(W, H, rows, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImageSource s{W, H, 4, filters};
        if (thresh_map) { MapThreshold t(thresh_map, W); s.green_equilibrate(t, rd); }
        else { rtengine::RawImageSource::GreenEqulibrateThreshold t(thresh); s.green_equilibrate(t, rd); }
    }
    delete[] rows;
    return 0;
}
"""


SHIM_BADPIX_TU = r"""
// Shim TU hosting the reference's hot / dead pixel filter: the helpers of badpixels.cc's anonymous namespace, RawImageSource::findHotDeadPixels
// and ::interpolateBadPixelsBayer cut from badpixels.cc, over the reference's own median.h and pixelsmap.h.  Written here (not reference
// code): the RawImageSource / RawImage stand-ins and the wrappers (byte maps in and out).
#include <math.h>
#include <cmath>
#include <array>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "array2D.h"
#include "rt_math.h"
#include "opthelper.h"
#include "median.h"
#include "pixelsmap.h"
#define BENCHFUN
#define RawImageSource RawImageSourceBadPix       /* other shim TUs define their own stand-in of this name: keep the inline members apart */
namespace rtengine {
enum { ST_BAYER_ = 1, ST_FUJI_XTRANS = 2 };
struct RawImageBadPix {
    int sensor; const int* xt;
    int getSensorType() const { return sensor; }
    unsigned XTRANSFC(int row, int col) const { return (unsigned)xt[(row % 6) * 6 + (col % 6)]; }     // rawimage.h: xtrans[(row) % 6][(col) % 6]
};
struct RawImageSource {
    int W, H; unsigned filters; RawImageBadPix* ri; array2D<float>& rawData;
    unsigned FC(int row, int col) const { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
    int interpolateBadPixelsBayer(const PixelsMap &bitmapBads, array2D<float> &rawData);
    int interpolateBadPixelsXtrans(const PixelsMap &bitmapBads);
    int findHotDeadPixels(PixelsMap &bpMap, const float thresh, const bool findHotPixels, const bool findDeadPixels) const;
};
#include "badpix_body.inc"
}
extern "C" int artref_find_hot_dead(float* raw, int W, int H, const int* xtrans36, float thresh, int hot, int dead, unsigned char* map, int nthreads)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (size_t)i * W;
    int n = 0;
    {
        rtengine::array2D<float> rd(W, H, rows, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImageBadPix ri{xtrans36 ? (int)rtengine::ST_FUJI_XTRANS : (int)rtengine::ST_BAYER_, xtrans36};
        rtengine::RawImageSource s{W, H, 0u, &ri, rd};
        rtengine::PixelsMap pm(W, H);
        const int old = omp_get_max_threads();
        if (nthreads > 0) omp_set_num_threads(nthreads);
        n = s.findHotDeadPixels(pm, thresh, hot != 0, dead != 0);
        omp_set_num_threads(old);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (pm.get(x, y)) map[(size_t)y * W + x] = 1;
    }
    delete[] rows;
    return n;
}
extern "C" int artref_interpolate_bad_bayer(float* raw, int W, int H, unsigned filters, const unsigned char* map)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (size_t)i * W;
    int n = 0;
    {
        rtengine::array2D<float> rd(W, H, rows, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImageBadPix ri{(int)rtengine::ST_BAYER_, nullptr};
        rtengine::RawImageSource s{W, H, filters, &ri, rd};
        rtengine::PixelsMap pm(W, H);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (map[(size_t)y * W + x]) pm.set(x, y);
        n = s.interpolateBadPixelsBayer(pm, rd);
    }
    delete[] rows;
    return n;
}
// nthreads = 1: the raster order.  The stock function builds its "virtual pixel" (and reads the distance-2 pixel) from neighbours it does not check
// against the map and that its own parallel loop may already have rewritten: with more than one thread the result depends on the schedule.
extern "C" int artref_interpolate_bad_xtrans(float* raw, int W, int H, const int* xtrans36, const unsigned char* map, int nthreads)
{
    float** rows = new float*[H];
    for (int i = 0; i < H; ++i) rows[i] = raw + (size_t)i * W;
    int n = 0;
    {
        rtengine::array2D<float> rd(W, H, rows, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImageBadPix ri{(int)rtengine::ST_FUJI_XTRANS, xtrans36};
        rtengine::RawImageSource s{W, H, 0u, &ri, rd};
        rtengine::PixelsMap pm(W, H);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (map[(size_t)y * W + x]) pm.set(x, y);
        const int old = omp_get_max_threads();
        if (nthreads > 0) omp_set_num_threads(nthreads);
        n = s.interpolateBadPixelsXtrans(pm);
        omp_set_num_threads(old);
    }
    delete[] rows;
    return n;
}
"""


SHIM_PACK_TU = r"""
// Shim TU hosting the reference's output packing: Imagefloat::getScanline cut from imagefloat.cc, over the reference's own rt_math.h
// and halffloat.h.  Written here (not reference code): the Imagefloat stand-in (planar accessors) and the wrapper.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "rt_math.h"
#include "halffloat.h"
namespace rtengine {
class Imagefloat {
public:
    const float *R, *G, *B; int width, height; const void* data;
    float r(int row, int col) const { return R[(size_t)row * width + col]; }
    float g(int row, int col) const { return G[(size_t)row * width + col]; }
    float b(int row, int col) const { return B[(size_t)row * width + col]; }
    void getScanline (int row, unsigned char* buffer, int bps, bool isFloat) const;
};
#include "pack_getscanline.inc"
}
extern "C" int artref_scanlines(const float* r, const float* g, const float* b, int W, int H, int bps, int is_float, void* out)
{
    rtengine::Imagefloat im{r, g, b, W, H, r};
    const size_t rowbytes = (size_t)W * 3 * (bps / 8);
    for (int row = 0; row < H; ++row) im.getScanline(row, (unsigned char*)out + rowbytes * row, bps, is_float != 0);
    return 0;
}
extern "C" unsigned short artref_float_to_half(float f) { return rtengine::DNG_FloatToHalf(f); }
// the whole float range in one call: number of bit patterns in [lo, hi) where `fn` (the port under test) disagrees with DNG_FloatToHalf
extern "C" long long artref_float_to_half_mismatches(unsigned lo, unsigned hi, unsigned short (*fn)(float))
{
    long long bad = 0;
#pragma omp parallel for reduction(+: bad) schedule(static)
    for (long long u = lo; u < (long long)hi; ++u) {
        uint32_t v = (uint32_t)u; float f; memcpy(&f, &v, 4);
        if (rtengine::DNG_FloatToHalf(f) != fn(f)) ++bad;
    }
    return bad;
}
"""


SHIM_BILINEAR_TU = r"""
// Shim TU hosting the reference's blended bilinear demosaic: RawImageSource::bayer_bilinear_demosaic(blend, ...) cut from
// bayer_bilinear_demosaic.cc.  Written here (not reference code): the RawImageSource stand-in and the wrapper.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <omp.h>
#include "glibmm.h"
#include "array2D.h"
#include "rt_math.h"
#define M(x) Glib::ustring(x)
#define RawImageSource RawImageSourceBilinear       /* other shim TUs define their own stand-in of this name: keep the inline members apart */
namespace rtengine {
struct BlProgress { void setProgressStr(const Glib::ustring&) {} void setProgress(double) {} };
struct RawImageSource {
    int W, H; unsigned filters; BlProgress* plistener;
    unsigned FC(int row, int col) const { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
    void bayer_bilinear_demosaic(const float* const * blend, const array2D<float> &rawData, array2D<float> &red, array2D<float> &green, array2D<float> &blue);
};
}
using namespace rtengine;
#include "bilinear_body.inc"
extern "C" int artref_bilinear_blend(const float* raw, const float* blend, int W, int H, unsigned filters, float* red, float* green, float* blue)
{
    auto rows = [&](const float* p) { float** t = new float*[H]; for (int i = 0; i < H; ++i) t[i] = const_cast<float*>(p) + (size_t)i * W; return t; };
    float **rr = rows(raw), **bl = rows(blend), **r = rows(red), **g = rows(green), **b = rows(blue);
    {
        array2D<float> rd(W, H, rr, ARRAY2D_BYREFERENCE), R(W, H, r, ARRAY2D_BYREFERENCE), G(W, H, g, ARRAY2D_BYREFERENCE), B(W, H, b, ARRAY2D_BYREFERENCE);
        RawImageSource s{W, H, filters, nullptr};
        s.bayer_bilinear_demosaic(bl, rd, R, G, B);
    }
    delete[] rr; delete[] bl; delete[] r; delete[] g; delete[] b;
    return 0;
}
"""


SHIM_VNG4_TU = r"""
// Shim TU hosting the reference's VNG4 demosaic: vng4interpolate_row_redblue and RawImageSource::vng4_demosaic cut from
// vng4_demosaic_RT.cc, border_interpolate2 from demosaic_algos.cc.  Written here (not reference code): the RawImage / RawImageSource /
// RAWParams stand-ins and the wrapper.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <omp.h>
#include "glibmm.h"
#include "array2D.h"
#include "rt_math.h"
#include "opthelper.h"
#define M(x) Glib::ustring(x)
#define BENCHFUN
#define RawImageSource RawImageSourceVNG4       /* other shim TUs define their own stand-in of this name: keep the inline members apart */
namespace rtengine {
struct RAWParams { struct BayerSensor { enum class Method { VNG4 }; static Glib::ustring getMethodString(Method) { return Glib::ustring("vng4"); } }; };
struct VngProgress { void setProgressStr(const Glib::ustring&) {} void setProgress(double) {} };
struct RawImage {
    unsigned filters, prefilters;
    unsigned FC(unsigned row, unsigned col) const { return (filters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3); }
    bool ISGREEN(unsigned row, unsigned col) const { return FC(row, col) == 1; }
    bool ISBLUE(unsigned row, unsigned col) const { return FC(row, col) == 2; }
};
struct RawImageSource {
    int W, H; RawImage* ri; VngProgress* plistener;
    unsigned FC(int row, int col) const { return ri->FC(row, col); }
    void vng4_demosaic (const array2D<float> &rawData, array2D<float> &red, array2D<float> &green, array2D<float> &blue);
    void border_interpolate2(int winw, int winh, int lborders, const array2D<float> &rawData, array2D<float> &red, array2D<float> &green, array2D<float> &blue);
};
#include "border_body.inc"
}
namespace {
using namespace rtengine;
#include "vng4_rowrb.inc"
}
namespace rtengine {
#define fc(row,col) (prefilters >> ((((row) << 1 & 14) + ((col) & 1)) << 1) & 3)
#include "vng4_body.inc"
#undef fc
}
extern "C" int artref_vng4(int W, int H, unsigned prefilters, const float* raw, float* red, float* green, float* blue, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    auto rows = [&](const float* p) { float** t = new float*[H]; for (int i = 0; i < H; ++i) t[i] = const_cast<float*>(p) + (size_t)i * W; return t; };
    float **rr = rows(raw), **r = rows(red), **g = rows(green), **b = rows(blue);
    {
        rtengine::array2D<float> rd(W, H, rr, rtengine::ARRAY2D_BYREFERENCE), R(W, H, r, rtengine::ARRAY2D_BYREFERENCE), G(W, H, g, rtengine::ARRAY2D_BYREFERENCE), B(W, H, b, rtengine::ARRAY2D_BYREFERENCE);
        rtengine::RawImage ri{prefilters & ~((prefilters & 0x55555555u) << 1), prefilters};
        rtengine::RawImageSource s{W, H, &ri, nullptr};
        s.vng4_demosaic(rd, R, G, B);
    }
    delete[] rr; delete[] r; delete[] g; delete[] b;
    return 0;
}
"""


SHIM_HLBLEND_TU = r"""
// Shim TU hosting the reference's "Blend" highlight reconstruction: RawImageSource::HLRecovery_blend cut from rawimagesource.cc.
// Written here (not reference code): the class stand-in and the wrapper.
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include "rt_math.h"
#define RawImageSource RawImageSourceHL       /* other shim TUs define their own stand-in of this name */
namespace rtengine {
struct RawImageSource { static void HLRecovery_blend(float* rin, float* gin, float* bin, int width, float maxval, float* hlmax); };
#include "hlblend_body.inc"
}
extern "C" int artref_hl_blend(float* rin, float* gin, float* bin, int width, float maxval, const float* hlmax)
{
    float h[3] = {hlmax[0], hlmax[1], hlmax[2]};
    rtengine::RawImageSource::HLRecovery_blend(rin, gin, bin, width, maxval, h);
    return 0;
}
"""


SHIM_GETIMAGE_TU = r"""
// Shim TU hosting the geometry half of RawImageSource::getImage: rotateLine (the coarse rotation of transLineStandard) cut from rawimagesource.cc,
// over a stand-in of PlanarPtr<float>.  Written here (not reference code): the stand-in, the line loop around it (getImage L943-1025 at skip == 1
// with the reference's own CLIP and, through artref_hl_blend, its own HLRecovery_blend) and the two mirror loops of PlanarRGBData::hflip / vflip
// (iimage.h L868-915: swap column j with width - 1 - j, row i with height - 1 - i).
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include "rt_math.h"
#define TR_NONE 0
#define TR_R90 1
#define TR_R180 2
#define TR_R270 3
#define TR_VFLIP 4
#define TR_HFLIP 8
#define TR_ROT 3
namespace rtengine {
template <class T> struct PlanarPtr {
    T* base; long stride;
    T& operator()(int row, int col) { return base[(long)row * stride + col]; }
};
}
namespace {
#include "getimage_rotateline.inc"
}
extern "C" int artref_hl_blend(float* rin, float* gin, float* bin, int width, float maxval, const float* hlmax);
extern "C" int artref_getimage(int W, int H, const float* r, const float* g, const float* b, long stride, const float* mul, int doClip,
                               int doHr, const float* hlmax, int tran, float* outr, float* outg, float* outb, long ostride)
{
    const bool swap = (tran & TR_ROT) == TR_R90 || (tran & TR_ROT) == TR_R270;
    const int ow = swap ? H : W, oh = swap ? W : H;
    rtengine::PlanarPtr<float> pr{outr, ostride}, pg{outg, ostride}, pb{outb, ostride};
    std::vector<float> lr(W), lg(W), lb(W);
    const float rm = mul[0], gm = mul[1], bm = mul[2];
    for (int i = 0; i < H; ++i) {
        for (int j = 0; j < W; ++j) {
            float rtot = r[(long)i * stride + j], gtot = g[(long)i * stride + j], btot = b[(long)i * stride + j];
            rtot *= rm; gtot *= gm; btot *= bm;
            if (doClip) { rtot = rtengine::CLIP(rtot); gtot = rtengine::CLIP(gtot); btot = rtengine::CLIP(btot); }
            lr[j] = rtot; lg[j] = gtot; lb[j] = btot;
        }
        if (doHr) artref_hl_blend(lr.data(), lg.data(), lb.data(), W, 65535.0f, hlmax);
        rotateLine(lr.data(), pr, tran, i, W, H);
        rotateLine(lg.data(), pg, tran, i, W, H);
        rotateLine(lb.data(), pb, tran, i, W, H);
    }
    float* planes[3] = {outr, outg, outb};
    for (int c = 0; c < 3; ++c) {
        float* v = planes[c];
        if (tran & TR_HFLIP)
            for (int i = 0; i < oh; i++)
                for (int j = 0; j < ow / 2; j++) std::swap(v[(long)i * ostride + j], v[(long)i * ostride + ow - 1 - j]);
        if (tran & TR_VFLIP)
            for (int i = 0; i < oh / 2; i++)
                for (int j = 0; j < ow; j++) std::swap(v[(long)i * ostride + j], v[(long)(oh - 1 - i) * ostride + j]);
    }
    return 0;
}

// ---- the preview form: RawImageSource::transformRect (cut from rawimagesource.cc) and getImage's line loop at any skip, its skip x skip box sum
// cut from rawimagesource.cc (the standard-CCD branch, L949-981).  Written here: the stand-ins and the loops around the two cuts.
namespace rtengine {
struct PreviewPropsShim { int x, y, w, h, skip; int getX() const { return x; } int getY() const { return y; } int getWidth() const { return w; }
    int getHeight() const { return h; } int getSkip() const { return skip; } };
#define PreviewProps PreviewPropsShim
struct RiGetImage { int get_FujiWidth() const { return 0; } };
struct RawImageSourceGetImage {
    int W, H, border; bool d1x, fuji; RiGetImage* ri;
    void transformRect(const PreviewProps &pp, int tran, int &ssx1, int &ssy1, int &width, int &height, int &fw);
};
#define RawImageSource RawImageSourceGetImage
#include "getimage_transformrect.inc"
#undef RawImageSource
}
extern "C" int artref_transform_rect(int W, int H, int border, int x, int y, int w, int h, int skip, int tran, int* out4)
{
    rtengine::RiGetImage ri;
    rtengine::RawImageSourceGetImage s{W, H, border, false, false, &ri};
    rtengine::PreviewPropsShim pp{x, y, w, h, skip};
    int fw = 0;
    s.transformRect(pp, tran, out4[0], out4[1], out4[2], out4[3], fw);
    return 0;
}
// source planes W x H (the demosaiced frame); sx1 / sy1 / imwidth / imheight from transformRect; output imwidth x imheight (turned for the quarter turns)
extern "C" int artref_getimage_pp(int W, int H, const float* r, const float* g, const float* b, long stride, const float* mul, int doClip,
                                  int doHr, const float* hlmax, int tran, int sx1, int sy1, int imwidth, int imheight, int skip,
                                  float* outr, float* outg, float* outb, long ostride)
{
    using rtengine::CLIP;
    const bool swap = (tran & TR_ROT) == TR_R90 || (tran & TR_ROT) == TR_R270;
    const int ow = swap ? imheight : imwidth, oh = swap ? imwidth : imheight;
    rtengine::PlanarPtr<float> pr{outr, ostride}, pg{outg, ostride}, pb{outb, ostride};
    std::vector<float> line_red(imwidth), line_grn(imwidth), line_blue(imwidth);
    std::vector<const float*> red(H), green(H), blue(H);
    for (int i = 0; i < H; ++i) { red[i] = r + (long)i * stride; green[i] = g + (long)i * stride; blue[i] = b + (long)i * stride; }
    const float rm = mul[0], gm = mul[1], bm = mul[2];
    const int maxx = W, maxy = H;
    for (int ix = 0; ix < imheight; ix++) {
        int i = sy1 + skip * ix;
        i = std::min(i, maxy - skip); // avoid trouble
#include "getimage_boxsum.inc"
        if (doHr) artref_hl_blend(line_red.data(), line_grn.data(), line_blue.data(), imwidth, 65535.0f, hlmax);
        rotateLine(line_red.data(), pr, tran, ix, imwidth, imheight);
        rotateLine(line_grn.data(), pg, tran, ix, imwidth, imheight);
        rotateLine(line_blue.data(), pb, tran, ix, imwidth, imheight);
    }
    float* planes[3] = {outr, outg, outb};
    for (int c = 0; c < 3; ++c) {
        float* v = planes[c];
        if (tran & TR_HFLIP)
            for (int i = 0; i < oh; i++)
                for (int j = 0; j < ow / 2; j++) std::swap(v[(long)i * ostride + j], v[(long)i * ostride + ow - 1 - j]);
        if (tran & TR_VFLIP)
            for (int i = 0; i < oh / 2; i++)
                for (int j = 0; j < ow; j++) std::swap(v[(long)i * ostride + j], v[(long)(oh - 1 - i) * ostride + j]);
    }
    return 0;
}
"""


SHIM_GUIDED_TU = r"""
// Shim TU hosting the reference's boxblur.h body (its include block is replaced: StopWatch.h drags
// settings.h -> procparams.h -> lcms2.h) and guidedFilter + calculate_subsampling cut from guidedfilter.cc.
#include <assert.h>
#include <memory>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include "alignedbuffer.h"
#include "rt_math.h"
#include "opthelper.h"
#include "array2D.h"
#include "sleef.h"
#define BENCHFUN
#include "boxblur_body.inc"
#include "rescale.h"
#define DEBUG_DUMP(arr)
namespace rtengine {
namespace {
#include "guided_subsampling.inc"
}
#include "guided_body.inc"
// guidedFilterLog (guidedfilter.cc L243-263), guided_smoothing (ipsmoothing.cc L334-409) over the Color members it calls (color.h)
void guidedFilterLog(const array2D<float> &guide, float base, array2D<float> &chan, int r, float eps, bool multithread, int subsampling=0);    // guidedfilter.h
void guidedFilterLog(float base, array2D<float> &chan, int r, float eps, bool multithread, int subsampling=0);
#include "guided_log.inc"
#include "guided_log2.inc"
typedef const float (*TMatrix)[3];      // iccstore.h L38
class Color { public:
template <class T>
#include "gs_rgblum.inc"
template <class T>
#include "gs_rgb2yuv.inc"
template <class T>
#include "gs_yuv2rgb.inc"
};
namespace {
enum class Channel { L, C, LC };        // ipsmoothing.cc L327-331
#include "guided_smoothing.inc"
}
// denoise::denoiseGuidedSmoothing (ipsmoothing.cc L875-897) with Imagefloat::normalizeFloatTo1 / To65535 (imagefloat.cc L396-438: x *= factor)
void artref_gs(float* R, float* G, float* B, int W, int H, const float ws[3][3], int guidedChromaRadius, double scale)
{
    if (guidedChromaRadius == 0) return;
    const size_t n = (size_t)W * H;
    const float down = 1.f / 65535.f, up = 65535.f;
    for (size_t k = 0; k < n; ++k) { R[k] *= down; G[k] *= down; B[k] *= down; }
    float **r = new float*[H], **g = new float*[H], **b = new float*[H];
    for (int i = 0; i < H; ++i) { r[i] = R + (size_t)i * W; g[i] = G + (size_t)i * W; b[i] = B + (size_t)i * W; }
    {
        array2D<float> aR(W, H, r, ARRAY2D_BYREFERENCE), aG(W, H, g, ARRAY2D_BYREFERENCE), aB(W, H, b, ARRAY2D_BYREFERENCE);
        TMatrix wsm = ws, iws = ws;
        guided_smoothing(aR, aG, aB, wsm, iws, Channel::C, guidedChromaRadius, 0.001f, scale, true);
    }
    delete[] r; delete[] g; delete[] b;
    for (size_t k = 0; k < n; ++k) { R[k] *= up; G[k] *= up; B[k] *= up; }
}
}
extern "C" int artref_denoise_guided_smoothing(float* R, float* G, float* B, int W, int H, const double* ws9, int guidedChromaRadius, double scale)
{
    float ws[3][3]; for (int i = 0; i < 9; ++i) (&ws[0][0])[i] = (float)ws9[i];
    rtengine::artref_gs(R, G, B, W, H, ws, guidedChromaRadius, scale);
    return 0;
}

namespace {
struct Tab {
    float** p;
    Tab(float* base, long stride, int H) { p = new float*[H]; for (int i = 0; i < H; ++i) p[i] = base + (long)i * stride; }
    ~Tab() { delete[] p; }
};
}

extern "C" int artref_boxblur(const float* src, long sstride, float* dst, long dstride, int W, int H, int radius, int inplace)
{
    Tab d(dst, dstride, H);
    if (inplace) { rtengine::boxblur(d.p, d.p, radius, W, H, true); return 0; }
    Tab s(const_cast<float*>(src), sstride, H);
    rtengine::boxblur(s.p, d.p, radius, W, H, true);
    return 0;
}

extern "C" int artref_guided_filter(const float* guide, const float* src, float* dst, long stride, int W, int H,
                                    int r, float epsilon, int subsampling)
{
    Tab g(const_cast<float*>(guide), stride, H), s(const_cast<float*>(src), stride, H), d(dst, stride, H);
    rtengine::array2D<float> G(W, H, g.p, rtengine::ARRAY2D_BYREFERENCE), S(W, H, s.p, rtengine::ARRAY2D_BYREFERENCE), D(W, H, d.p, rtengine::ARRAY2D_BYREFERENCE);
    rtengine::guidedFilter(G, S, D, r, epsilon, true, subsampling);
    return 0;
}
"""


SHIM_WAVELET_TU = r"""
// The reference's wavelet headers compile standalone: included unmodified, straight from /root/reference.
#include <algorithm>
#include <cstring>
#include "cplx_wavelet_dec.h"
#include "cplx_wavelet_dec.cc"
using rtengine::wavelet_decomposition;
extern "C" {
void* artref_wavelet_new(float* src, int W, int H, int maxlvl, int subsamp, int nthreads)
{ return new wavelet_decomposition(src, W, H, maxlvl, subsamp, 1, nthreads > 0 ? nthreads : 1, 6); }
int artref_wavelet_maxlevel(void* w) { return ((wavelet_decomposition*)w)->maxlevel(); }
int artref_wavelet_level_W(void* w, int l) { return ((wavelet_decomposition*)w)->level_W(l); }
int artref_wavelet_level_H(void* w, int l) { return ((wavelet_decomposition*)w)->level_H(l); }
int artref_wavelet_level_stride(void* w, int l) { return ((wavelet_decomposition*)w)->level_stride(l); }
float* artref_wavelet_band(void* w, int l, int dir)
{ wavelet_decomposition* d = (wavelet_decomposition*)w; return dir == 0 ? d->coeff0 : d->level_coeffs(l)[dir]; }
void artref_wavelet_reconstruct(void* w, float* dst, float blend) { ((wavelet_decomposition*)w)->reconstruct(dst, blend); }
void artref_wavelet_delete(void* w) { delete (wavelet_decomposition*)w; }
}
"""


SHIM_SHRINK_TU = r"""
// Shim TU hosting the reference's wavelet shrinkage: MadRgb, ShrinkAllL, ShrinkAllAB, WaveletDenoiseAllL,
// WaveletDenoiseAllAB cut from FTblockDN.cc, over the reference's own wavelet headers, boxblur.h body and sleef.
#include <assert.h>
#include <memory>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include "alignedbuffer.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "cplx_wavelet_dec.h"
#define BENCHFUN
#include "boxblur_body.inc"
namespace rtengine {
namespace {
int denoiseNestedLevels = 1;
#include "shrink_body.inc"
}
}
using rtengine::wavelet_decomposition;
extern "C" {
float artref_madrgb(float* data, int n) { return rtengine::MadRgb(data, n); }
int artref_wavelet_denoise_L(void* wL, float* noisevarlum, float* madL /*[8][3]*/, double scale)
{ return rtengine::WaveletDenoiseAllL(scale, *(wavelet_decomposition*)wL, noisevarlum, (float(*)[3])madL, nullptr, 0) ? 0 : 1; }
int artref_wavelet_denoise_AB(void* wL, void* wab, float* noisevarchrom, float* madL, float noisevar_ab, int useNoiseCCurve, int autoch, double scale)
{ return rtengine::WaveletDenoiseAllAB(scale, *(wavelet_decomposition*)wL, *(wavelet_decomposition*)wab, noisevarchrom, (float(*)[3])madL, noisevar_ab, useNoiseCCurve != 0, autoch != 0) ? 0 : 1; }
}
"""

SHIM_NLMEANS_TU = r"""
// Shim TU hosting the reference's NL-means: NLMeans cut from nlmeans.cc, laplacian + detail_mask cut from
// FTblockDN.cc, over the reference's own array2D.h / LUT.h / rescale.h / sleef.h and the gaussianBlur of shim_gauss.cc.
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <omp.h>
#include "array2D.h"
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "gauss.h"
#include "rescale.h"
#define BENCHFUN
#include "boxblur_body.inc"
namespace rtengine { namespace denoise {
enum class BlurType { OFF, BOX, GAUSS };
void detail_mask(const array2D<float> &src, array2D<float> &mask, float scaling, float threshold, float ceiling, float factor, BlurType blur, float blur_radius, bool multithread);
// laplacian + detail_mask are defined once, in shim_denoise.cc (the whole FTblockDN.cc body)
#include "nlmeans_body.inc"
}}
namespace {
using rtengine::array2D; using rtengine::ARRAY2D_BYREFERENCE;
struct Rows { float** r; Rows(float* p, int W, int H) : r(new float*[H]) { for (int i = 0; i < H; ++i) r[i] = p + (size_t)i * W; } ~Rows() { delete[] r; } };
}
extern "C" {
int artref_detail_mask(float* src, float* mask, int W, int H, float scaling, float threshold, float ceiling, float factor, int blur_type, float blur)
{
    Rows rs(src, W, H);
    array2D<float> s(W, H, rs.r, ARRAY2D_BYREFERENCE), m;
    rtengine::denoise::detail_mask(s, m, scaling, threshold, ceiling, factor, (rtengine::denoise::BlurType)blur_type, blur, true);
    for (int y = 0; y < H; ++y) memcpy(mask + (size_t)y * W, m[y], sizeof(float) * W);
    return 0;
}
int artref_nlmeans(float* img, int W, int H, float normcoeff, int strength, int detail_thresh, float scale)
{
    Rows rs(img, W, H);
    array2D<float> a(W, H, rs.r, ARRAY2D_BYREFERENCE);
    rtengine::denoise::NLMeans(a, normcoeff, strength, detail_thresh, scale, true);
    return 0;
}
}
"""

SHIM_DENOISE_TU = r"""
// Shim TU hosting the reference's FTblockDN.cc from its `#define TS 64` to the end of the file (RGB_denoise and
// everything it calls), with the #include block replaced by the minimal stand-ins below: the real headers drag
// glibmm / lcms2 / fftw3, none of which exist in this image.  Stand-ins written here (not reference code):
// Imagefloat, LabImage, ProcParams/DenoiseParams, ImProcData, NoiseCurve, ICCStore, Settings/Options, MyMutex, MyTime,
// and <fftw3.h> = oracle/dct_standin.h.  Color's members used by RGB_denoise are cut from color.h / color.cc.
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>
#include <omp.h>
#include "array2D.h"
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "alignedbuffer.h"
#include "median.h"
#include "gauss.h"
#include "rescale.h"
#include "cplx_wavelet_dec.h"
#include "dct_standin.h"
#define BENCHFUN
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))   /* glib's gmacros.h forms */
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#include "boxblur_body.inc"

namespace rtengine {

typedef const float (*TMatrix)[3];      // iccstore.h L38

struct Settings { bool verbose; };
static const Settings artref_settings = {false};
const Settings* settings = &artref_settings;
struct Options { int rgbDenoiseThreadLimit; };
static Options options = {1};                 // one thread inside detail_recovery: its overlap-add races otherwise

class MyMutex { public: class MyLock { public: explicit MyLock(MyMutex&) {} }; };
static MyMutex artref_mutex;
MyMutex* fftwMutex = &artref_mutex;
class MyTime { public: void set() {} int etime(const MyTime&) { return 0; } };

class Imagefloat {
public:
    int W, H; float *R, *G, *B; bool own;
    Imagefloat(int w, int h) : W(w), H(h), R(new float[(size_t)w * h]), G(new float[(size_t)w * h]), B(new float[(size_t)w * h]), own(true) {}
    Imagefloat(int w, int h, float* r_, float* g_, float* b_) : W(w), H(h), R(r_), G(g_), B(b_), own(false) {}
    ~Imagefloat() { if (own) { delete[] R; delete[] G; delete[] B; } }
    int getWidth() const { return W; }
    int getHeight() const { return H; }
    float& r(int i, int j) { return R[(size_t)i * W + j]; }
    float& g(int i, int j) { return G[(size_t)i * W + j]; }
    float& b(int i, int j) { return B[(size_t)i * W + j]; }
    void copyData(Imagefloat* d) { memcpy(d->R, R, sizeof(float) * (size_t)W * H); memcpy(d->G, G, sizeof(float) * (size_t)W * H); memcpy(d->B, B, sizeof(float) * (size_t)W * H); }
};

class LabImage {
public:
    int W, H; float* data; float **L, **a, **b;
    LabImage(int w, int h) : W(w), H(h), data(new float[(size_t)w * h * 3]), L(new float*[h]), a(new float*[h]), b(new float*[h])
    { for (int i = 0; i < h; ++i) { L[i] = data + (size_t)i * w; a[i] = data + (size_t)(h + i) * w; b[i] = data + (size_t)(2 * h + i) * w; } }
    ~LabImage() { delete[] data; delete[] L; delete[] a; delete[] b; }
};

namespace procparams {
struct DenoiseParams {
    enum class ChrominanceMethod { MANUAL, AUTOMATIC };
    enum class ColorSpace { RGB, LAB };
    bool enabled; ColorSpace colorSpace; bool aggressive; double gamma; double luminance; double luminanceDetail; int luminanceDetailThreshold;
    ChrominanceMethod chrominanceMethod; double chrominance; double chrominanceRedGreen; double chrominanceBlueYellow;
};
struct ICMParams { std::string workingProfile; };
struct ProcParams { ICMParams icm; DenoiseParams denoise; };
}
using procparams::ProcParams;
struct ImProcData { const ProcParams* params; double scale; bool multiThread; };

static float artref_wp[3][3], artref_wpi[3][3];     // iccmatrices.h holds float constants
class ICCStore {
public:
    static ICCStore* getInstance() { static ICCStore s; return &s; }
    TMatrix workingSpaceMatrix(const std::string&) const { return artref_wp; }
    TMatrix workingSpaceInverseMatrix(const std::string&) const { return artref_wpi; }
};

class Color {
public:
    constexpr static double sRGBGammaCurve = 2.4;
    constexpr static double eps = 216.0 / 24389.0;
    constexpr static double MAXVALD = 65535.0; constexpr static float MAXVALF = 65535.f;
    constexpr static double eps_max = MAXVALF * eps;
    constexpr static double kappa = 24389.0 / 27.0;
    constexpr static float D50x = 0.9642f, D50z = 0.8249f;
    static LUTf cachef, cachefy, denoiseGammaTab, denoiseIGammaTab, igammatab_srgb, gammatab_srgb;
    static void init()
    {
        if (cachef) return;
        cachef(65536, LUT_CLIP_BELOW); cachefy(65536, LUT_CLIP_BELOW);
        int i = 0; const int epsmaxint = eps_max;            // color.cc L205-233
        for (; i <= epsmaxint; i++) { cachef[i] = 327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0); cachefy[i] = 327.68 * (kappa * i / MAXVALF); }
        for (; i < 65536; i++) { cachef[i] = 327.68 * std::cbrt((double)i / MAXVALF); cachefy[i] = 327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0); }
        denoiseGammaTab(65536, 0); denoiseIGammaTab(65536, 0);          // color.cc L188-189, L278-292
        for (int k = 0; k < 65536; k++) { denoiseGammaTab[k] = 65535.0 * gamma55 (k / 65535.0); denoiseIGammaTab[k] = 65535.0 * igamma55 (k / 65535.0); }
    }
#include "color_h_members.inc"
    static float computeXYZ2Lab(float f);
    static float computeXYZ2LabY(float f);
    static void gammaf2lut (LUTf &gammacurve, float gamma, float start, float slope, float divisor, float factor);
    static void rgbxyz (float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb[3][3]);
    static void XYZ2Lab(float X, float Y, float Z, float &L, float &a, float &b);
    static void RGB2L(float *R, float *G, float *B, float *L, const float wp[3][3], int width);
    constexpr static double kappaInv = 27.0 / 24389.0;
    constexpr static double epsilonExpInv3 = 6.0 / 29.0;
    constexpr static float kappaInvf = kappaInv;
    constexpr static float epsilonExpInv3f = epsilonExpInv3;
    constexpr static double epskap = 8.0;
    constexpr static float c1By116 = 1.0 / 116.0;
    constexpr static float c16By116 = 16.0 / 116.0;
    static void xyz2rgb (float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz[3][3]);
    static void Lab2XYZ(float L, float a, float b, float &x, float &y, float &z);
#include "color_h_lab_members.inc"
};
LUTf Color::cachef, Color::cachefy, Color::denoiseGammaTab, Color::denoiseIGammaTab, Color::igammatab_srgb, Color::gammatab_srgb;
#include "color_cc_members.inc"


namespace denoise {
class NoiseCurve {
public:
    LUTf lutNoiseCurve; float sum;
    NoiseCurve() : sum(0.f) {}
    float getSum() const { return sum; }
    float operator[](float index) const { return lutNoiseCurve[index]; }
    operator bool() const { return lutNoiseCurve; }
};
enum class Median { TYPE_3X3_SOFT, TYPE_3X3_STRONG, TYPE_5X5_SOFT, TYPE_5X5_STRONG, TYPE_7X7, TYPE_9X9 };
enum class BlurType { OFF, BOX, GAUSS };
void detail_mask(const array2D<float> &src, array2D<float> &mask, float scaling, float threshold, float ceiling, float factor, BlurType blur, float blur_radius, bool multithread);
void Tile_calc(int tilesize, int overlap, int kall, int imwidth, int imheight, int &numtiles_W, int &numtiles_H, int &tilewidth, int &tileheight, int &tileWskip, int &tileHskip);
}
}  // namespace rtengine

#include "ftblockdn_body.inc"

// calcautodn_info, RGB_denoise_infoGamCurve, RGB_denoise_info cut from ipdenoise.cc (the AUTOMATIC chroma estimator)
namespace rtengine { namespace {
#include "ipdenoise_info.inc"
} }

using namespace rtengine;
extern "C" {
// RGB_denoise_info over one crop (isRAW); out[15] = chaut, Nb, redaut, blueaut, maxredaut, maxblueaut, minredaut, minblueaut, chromina,
// sigma, lumema, sigma_L, redyel, skinc, nsknc, read as the initial values and written back
int artref_denoise_info(const float* r, const float* g, const float* b, int W, int H, const float* cr, const float* cg, const float* cb,
                        double gamma, int aggressive, double scale, double expcomp, const double* wp, float* out)
{
    Color::init();
    for (int i = 0; i < 9; ++i) { (&artref_wp[0][0])[i] = (float)wp[i]; }
    procparams::DenoiseParams dn;
    dn.enabled = true; dn.colorSpace = procparams::DenoiseParams::ColorSpace::RGB; dn.aggressive = aggressive != 0;
    dn.chrominanceMethod = procparams::DenoiseParams::ChrominanceMethod::AUTOMATIC; dn.gamma = gamma;
    dn.luminance = 0; dn.luminanceDetail = 0; dn.luminanceDetailThreshold = 0; dn.chrominance = 15; dn.chrominanceRedGreen = 0; dn.chrominanceBlueYellow = 0;
    ProcParams pp; pp.icm.workingProfile = "ProPhoto"; pp.denoise = dn;
    ImProcData im = {&pp, scale, true};
    Imagefloat src(W, H, const_cast<float*>(r), const_cast<float*>(g), const_cast<float*>(b));
    Imagefloat provicalc((W + 1) / 2, (H + 1) / 2, const_cast<float*>(cr), const_cast<float*>(cg), const_cast<float*>(cb));
    LUTf gamcurve(65536, 0);
    float gam, gamthresh, gamslope;
    RGB_denoise_infoGamCurve(dn, true, gamcurve, gam, gamthresh, gamslope);
    float chaut = out[0], redaut = out[2], blueaut = out[3], maxredaut = out[4], maxblueaut = out[5], minredaut = out[6], minblueaut = out[7],
          chromina = out[8], sigma = out[9], lumema = out[10], sigma_L = out[11], redyel = out[12], skinc = out[13], nsknc = out[14];
    int nb = (int)out[1];
    RGB_denoise_info(im, &src, &provicalc, true, gamcurve, gam, gamthresh, gamslope, dn, expcomp, chaut, nb, redaut, blueaut, maxredaut, maxblueaut,
                     minredaut, minblueaut, chromina, sigma, lumema, sigma_L, redyel, skinc, nsknc);
    out[0] = chaut; out[1] = (float)nb; out[2] = redaut; out[3] = blueaut; out[4] = maxredaut; out[5] = maxblueaut; out[6] = minredaut; out[7] = minblueaut;
    out[8] = chromina; out[9] = sigma; out[10] = lumema; out[11] = sigma_L; out[12] = redyel; out[13] = skinc; out[14] = nsknc;
    return 0;
}
// Color::RGB2L row by row, as dual_demosaic_RT.cc L101-106 calls it
int artref_rgb2l(const float* R, const float* G, const float* B, float* L, int W, int H, const float* wp9)
{
    Color::init();
    float wp[3][3]; for (int i = 0; i < 9; ++i) (&wp[0][0])[i] = wp9[i];
    for (int i = 0; i < H; ++i)
        Color::RGB2L(const_cast<float*>(R) + (size_t)i * W, const_cast<float*>(G) + (size_t)i * W, const_cast<float*>(B) + (size_t)i * W, L + (size_t)i * W, wp, W);
    return 0;
}
// the nine-crop combination of ImProcFunctions::denoiseComputeParams (ipdenoise.cc, from `float chM = 0.f;` to the store assignments), cut in place
int artref_denoise_auto_params(const float* stats, int isRAW_, int aggressive, float* out3)
{
    struct Src { bool raw; bool isRAW() const { return raw; } } src_{isRAW_ != 0};
    Src* imgsrc = &src_;
    struct { float ch_M[9], max_r[9], max_b[9]; double chrominance, chrominanceRedGreen, chrominanceBlueYellow; } store;
    ProcParams pp; pp.denoise.aggressive = aggressive != 0;
    const ProcParams* params = &pp;
    float min_b[9], min_r[9], lumL[9], chromC[9], ry[9], sk[9], pcsk[9];
    int Nb[9];
    const float pondcorrec = 1.0f;
    for (int k = 0; k < 9; ++k) {
        const float* s = stats + 15 * k;
        Nb[k] = (int)s[1]; store.ch_M[k] = pondcorrec * s[0]; store.max_r[k] = pondcorrec * s[4]; store.max_b[k] = pondcorrec * s[5];
        min_r[k] = pondcorrec * s[6]; min_b[k] = pondcorrec * s[7]; lumL[k] = s[10]; chromC[k] = s[8]; ry[k] = s[12]; sk[k] = s[13]; pcsk[k] = s[14];
    }
    float autoNR = 10, autoNRmax = 40, lowdenoise = 1.f;
    int levaut = 0;
#include "ipdenoise_combine.inc"
    out3[0] = (float)store.chrominance; out3[1] = (float)store.chrominanceRedGreen; out3[2] = (float)store.chrominanceBlueYellow;
    return 0;
}
// timing only: 0 = the reference's default (all OpenMP threads inside detail_recovery, whose overlap-add then races); parity tests keep 1
void artref_set_denoise_thread_limit(int n) { rtengine::options.rgbDenoiseThreadLimit = n; }
// p: luminance, luminanceDetail, luminanceDetailThreshold, chrominance, chrominanceRedGreen, chrominanceBlueYellow, gamma, scale
// ccurve: 501-entry NoiseCurve LUT (or null = curve not set), calclum: 3 planes of ((H+1)/2) x ((W+1)/2) (or null)
int artref_rgb_denoise_ex2(float* r, float* g, float* b, int W, int H, const double* p, const double* wp, const double* wpi,
                       const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* nresi_highresi, int aggressive, int lab_mode)
{
    Color::init();
    for (int i = 0; i < 9; ++i) { (&artref_wp[0][0])[i] = (float)((const double*)wp)[i]; (&artref_wpi[0][0])[i] = (float)((const double*)wpi)[i]; }
    procparams::DenoiseParams dn;
    dn.enabled = true; dn.colorSpace = lab_mode ? procparams::DenoiseParams::ColorSpace::LAB : procparams::DenoiseParams::ColorSpace::RGB; dn.aggressive = aggressive != 0;
    dn.luminance = p[0]; dn.luminanceDetail = p[1]; dn.luminanceDetailThreshold = (int)p[2];
    dn.chrominanceMethod = procparams::DenoiseParams::ChrominanceMethod::MANUAL;
    dn.chrominance = p[3]; dn.chrominanceRedGreen = p[4]; dn.chrominanceBlueYellow = p[5]; dn.gamma = p[6];
    ProcParams pp; pp.icm.workingProfile = "ProPhoto";
    ImProcData im = {&pp, p[7], true};
    Imagefloat img(W, H, r, g, b);
    denoise::NoiseCurve lc, cc;
    Imagefloat* calclum = nullptr;
    if (ccurve) {
        cc.lutNoiseCurve(501); for (int i = 0; i < 501; ++i) cc.lutNoiseCurve[i] = ccurve[i];
        cc.sum = ccurve_sum;
    }
    if (cl_r) {
        const int w2 = (W + 1) / 2, h2 = (H + 1) / 2;
        calclum = new Imagefloat(w2, h2);
        memcpy(calclum->R, cl_r, sizeof(float) * (size_t)w2 * h2); memcpy(calclum->G, cl_g, sizeof(float) * (size_t)w2 * h2); memcpy(calclum->B, cl_b, sizeof(float) * (size_t)w2 * h2);
    }
    float nresi = 0.f, highresi = 0.f;
    denoise::RGB_denoise(im, 0, &img, &img, calclum, nullptr, nullptr, nullptr, true, dn, 0.0, lc, cc, nresi, highresi);
    if (nresi_highresi) { nresi_highresi[0] = nresi; nresi_highresi[1] = highresi; }
    return 0;
}
int artref_rgb_denoise_ex(float* r, float* g, float* b, int W, int H, const double* p, const double* wp, const double* wpi,
                       const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* nresi_highresi, int aggressive)
{
    return artref_rgb_denoise_ex2(r, g, b, W, H, p, wp, wpi, ccurve, ccurve_sum, cl_r, cl_g, cl_b, nresi_highresi, aggressive, 0);
}
int artref_rgb_denoise(float* r, float* g, float* b, int W, int H, const double* p, const double* wp, const double* wpi,
                       const float* ccurve, float ccurve_sum, const float* cl_r, const float* cl_g, const float* cl_b, float* nresi_highresi)
{
    return artref_rgb_denoise_ex(r, g, b, W, H, p, wp, wpi, ccurve, ccurve_sum, cl_r, cl_g, cl_b, nresi_highresi, 0);
}
}
"""


SHIM_FATTAL_TU = r"""
// Shim TU hosting the reference's tmo_fattal02.cc from `namespace rtengine {` (after its #include block) up to, not
// including, ImProcFunctions::dynamicRangeCompression, (Median_Denoise comes from shim_denoise.cc = FTblockDN.cc whole).  Stand-ins written
// here (not reference code): Imagefloat, ProcParams/FattalToneMappingParams, ICCStore, Settings, MyMutex and <fftw3.h> =
// oracle/dct_standin.h (REDFT00 restated in double: parity unpinned at that boundary, fftw3f is absent here).
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include <array>
#include <iostream>
#include <string>
#include <vector>
#include <omp.h>
#include "array2D.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "median.h"
#include "rescale.h"
#include "dct_standin.h"

namespace rtengine {
typedef const float (*TMatrix)[3];      // iccstore.h L38
struct Settings { bool verbose; };                 // same stand-ins as shim_denoise.cc, which defines the two globals
extern const Settings* settings;
class MyMutex { public: class MyLock { public: explicit MyLock(MyMutex&) {} }; };
extern MyMutex* fftwMutex;

class Imagefloat {
public:
    int W, H; float *R, *G, *B;
    Imagefloat(int w, int h, float* r_, float* g_, float* b_) : W(w), H(h), R(r_), G(g_), B(b_) {}
    int getWidth() const { return W; }
    int getHeight() const { return H; }
    float& r(int i, int j) { return R[(size_t)i * W + j]; }
    float& g(int i, int j) { return G[(size_t)i * W + j]; }
    float& b(int i, int j) { return B[(size_t)i * W + j]; }
};
namespace procparams {
struct FattalToneMappingParams { bool enabled; int threshold; int amount; bool satcontrol; };
struct ICMParams { std::string workingProfile; };
struct ProcParams { FattalToneMappingParams fattal; ICMParams icm; };
}
using procparams::ProcParams;
class ImProcFunctions;
static float artref_fattal_ws[3][3];
class ICCStore {
public:
    static ICCStore* getInstance() { static ICCStore s; return &s; }
    TMatrix workingSpaceMatrix(const std::string&) const { return artref_fattal_ws; }
};
class Color {
public:
#include "fattal_color_members.inc"
};
namespace denoise {
enum class Median { TYPE_3X3_SOFT, TYPE_3X3_STRONG, TYPE_5X5_SOFT, TYPE_5X5_STRONG, TYPE_7X7, TYPE_9X9 };
}
namespace denoise {   // defined in shim_denoise.cc: the reference's FTblockDN.cc, whole
void Median_Denoise(float **src, float **dst, float upperBound, const int width, const int height, const Median medianType, const int iterations, const int numThreads, float **buffer);
void Median_Denoise(float **src, float **dst, const int width, const int height, const Median medianType, const int iterations, const int numThreads, float **buffer);
}
}  // namespace rtengine

#include "fattal_body.inc"
// (fattal_body.inc leaves `namespace rtengine {` open)

extern "C" {
int artref_fattal(float* r, float* g, float* b, int W, int H, int threshold, int amount, int satcontrol, const double* ws)
{
    for (int i = 0; i < 9; ++i) (&artref_fattal_ws[0][0])[i] = (float)((const double*)ws)[i];
    ProcParams pp; pp.fattal.enabled = true; pp.fattal.threshold = threshold; pp.fattal.amount = amount; pp.fattal.satcontrol = satcontrol != 0;
    pp.icm.workingProfile = "ProPhoto";
    Imagefloat img(W, H, r, g, b);
    ToneMapFattal02(&img, nullptr, &pp, true);
    return 0;
}
// the inner operator alone: Y (w x h, contiguous) -> L, the call at tmo_fattal02.cc L1126
int artref_tmo_fattal02(int w, int h, float* Y, float alfa, float beta, float noise, int detail_level)
{
    Array2Df L(w, h);
    memcpy(L.data(), Y, sizeof(float) * (size_t)w * h);
    tmo_fattal02(w, h, L, L, alfa, beta, noise, detail_level, true);
    memcpy(Y, L.data(), sizeof(float) * (size_t)w * h);
    return 0;
}
int artref_median_denoise(float* src, float* dst, float upperBound, int useUpper, int W, int H, int type, int iterations)
{
    float** s = new float*[H]; float** d = new float*[H];
    for (int i = 0; i < H; ++i) { s[i] = src + (size_t)i * W; d[i] = dst + (size_t)i * W; }
    if (useUpper) denoise::Median_Denoise(s, src == dst ? s : d, upperBound, W, H, (denoise::Median)type, iterations, omp_get_num_procs(), nullptr);
    else denoise::Median_Denoise(s, src == dst ? s : d, W, H, (denoise::Median)type, iterations, omp_get_num_procs(), nullptr);
    delete[] s; delete[] d;
    return 0;
}
int artref_find_fast_dim(int dim) { return find_fast_dim(dim); }
void artref_redft00_2d(int n0, int n1, const float* in, float* out) { artdct_redft00_2d(n0, n1, in, out); }
void artref_redft00_1d_naive(int n, const double* x, double* y) { artdct_redft00_1d_naive(n, x, y); }
void artref_redft00_1d(int n, const double* x, double* y)
{
    artdct_cplx* tw = artdct_twiddles(2 * (n - 1));
    artdct_cplx* a = (artdct_cplx*)malloc(sizeof(artdct_cplx) * 2 * n); artdct_cplx* b = (artdct_cplx*)malloc(sizeof(artdct_cplx) * 2 * n);
    artdct_redft00_1d(n, x, 1, y, 1, tw, a, b);
    free(tw); free(a); free(b);
}
}
}  // namespace rtengine
"""


SHIM_CHAIN_TU = r"""
// Shim TU hosting the per-pixel colour / curve chain of ImProcFunctions::process: the pixel loops of expcomp
// (ipexposure.cc), saturationVibrance (ipsaturation.cc), filmlike_clip + Standard / Adobe tone curves (iptonecurve.cc,
// curves.h), rgbCurves (iprgbcurves.cc), lab_adjustments (iplabadjustments.cc), Imagefloat::rgb_to_lab / lab_to_rgb
// (imagefloat.cc) and the Color members they call (color.h / color.cc), all cut from the reference at build time.
// Written here (not reference code): the Imagefloat / Plane stand-in, Color's constants and LUT initialisation
// (restating color.cc L205-233), the wrappers.  Everything sits in its own namespace so that this Color does not collide
// with the one in shim_denoise.cc.
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include <omp.h>
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"

namespace artref_chain {
using namespace rtengine;
typedef const float (*TMatrix)[3];      // iccstore.h L38

class Curve { public: virtual ~Curve() {} virtual double getVal(double) const { abort(); } };

class Color {
public:
    constexpr static double eps = 216.0 / 24389.0;
    constexpr static double MAXVALD = 65535.0; constexpr static float MAXVALF = 65535.f;
    constexpr static double eps_max = MAXVALF * eps;
    constexpr static double kappa = 24389.0 / 27.0;
    constexpr static double kappaInv = 27.0 / 24389.0;
    constexpr static double epsilonExpInv3 = 6.0 / 29.0;
    constexpr static float kappaInvf = kappaInv;
    constexpr static float epsilonExpInv3f = epsilonExpInv3;
    constexpr static float D50x = 0.9642f, D50z = 0.8249f;
    constexpr static double epskap = 8.0;
    constexpr static float c1By116 = 1.0 / 116.0;
    constexpr static float c16By116 = 16.0 / 116.0;
    static LUTf cachef, cachefy;
    static void init()
    {
        if (cachef) return;
        cachef(65536, LUT_CLIP_BELOW); cachefy(65536, LUT_CLIP_BELOW);
        int i = 0; const int epsmaxint = eps_max;            // color.cc L205-233
        for (; i <= epsmaxint; i++) { cachef[i] = 327.68 * ((kappa * i / MAXVALF + 16.0) / 116.0); cachefy[i] = 327.68 * (kappa * i / MAXVALF); }
        for (; i < 65536; i++) { cachef[i] = 327.68 * std::cbrt((double)i / MAXVALF); cachefy[i] = 327.68 * (116.0 * std::cbrt((double)i / MAXVALF) - 16.0); }
    }
#include "chain_color_h.inc"
    static float computeXYZ2Lab(float f);
    static float computeXYZ2LabY(float f);
    static void rgbxyz (float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb[3][3]);
    static void rgbxyz (vfloat r, vfloat g, vfloat b, vfloat &x, vfloat &y, vfloat &z, const vfloat xyz_rgb[3][3]);
    static void xyz2rgb (float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz[3][3]);
    static void xyz2rgb (vfloat x, vfloat y, vfloat z, vfloat &r, vfloat &g, vfloat &b, const vfloat rgb_xyz[3][3]);
    static void XYZ2Lab(float X, float Y, float Z, float &L, float &a, float &b);
    static void XYZ2Lab(vfloat X, vfloat Y, vfloat Z, vfloat &L, vfloat &a, vfloat &b);
    static void Lab2XYZ(float L, float a, float b, float &x, float &y, float &z);
    static void Lab2XYZ(vfloat L, vfloat a, vfloat b, vfloat &x, vfloat &y, vfloat &z);
    static void filmlike_clip(float *r, float *g, float *b, float Lmax);
    static void rgb2hsv(float r, float g, float b, float &h, float &s, float &v);
    static void hsv2rgb (float h, float s, float v, float &r, float &g, float &b);
};
LUTf Color::cachef, Color::cachefy;
#include "chain_color_cc.inc"

// stand-in for PlanarPtr / Imagefloat: rows 16-byte aligned like the reference's allocation (iimage.h L653-673)
struct Plane {
    float* base; int stride; float** ptrs;
    float& operator()(int y, int x) { return base[(size_t)y * stride + x]; }
    float* operator()(int y) { return base + (size_t)y * stride; }
};
class Imagefloat {
public:
    int width, height; Plane r, g, b;
    float ws_[3][3], iws_[3][3]; vfloat vws_[3][3], viws_[3][3];
    Imagefloat(int w, int h, const float* R, const float* G, const float* B, const double* ws, const double* iws) : width(w), height(h)
    {
        const int st = (w + 3) / 4 * 4;
        Plane* pl[3] = {&r, &g, &b}; const float* src[3] = {R, G, B};
        for (int c = 0; c < 3; ++c) {
            void* p = nullptr; if (posix_memalign(&p, 64, sizeof(float) * (size_t)st * h)) abort();
            pl[c]->base = (float*)p; pl[c]->stride = st; pl[c]->ptrs = new float*[h];
            for (int y = 0; y < h; ++y) { pl[c]->ptrs[y] = pl[c]->base + (size_t)y * st; memcpy(pl[c]->ptrs[y], src[c] + (size_t)y * w, sizeof(float) * w); }
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {      // get_ws(), imagefloat.cc L595-617
            ws_[i][j] = float(ws[3 * i + j]); iws_[i][j] = float(iws ? iws[3 * i + j] : 0.0);
            vws_[i][j] = F2V(float(ws[3 * i + j])); viws_[i][j] = F2V(float(iws ? iws[3 * i + j] : 0.0));
        }
    }
    ~Imagefloat() { Plane* pl[3] = {&r, &g, &b}; for (int c = 0; c < 3; ++c) { free(pl[c]->base); delete[] pl[c]->ptrs; } }
    void store(float* R, float* G, float* B) { Plane* pl[3] = {&r, &g, &b}; float* dst[3] = {R, G, B};
        for (int c = 0; c < 3; ++c) for (int y = 0; y < height; ++y) memcpy(dst[c] + (size_t)y * width, pl[c]->ptrs[y], sizeof(float) * width); }
    int getWidth() const { return width; }
    int getHeight() const { return height; }
    void get_ws() {}
    void rgb_to_lab(bool multithread);
    inline void rgb_to_lab(int y, int x, float &L, float &a, float &b);
    void lab_to_rgb(bool multithread);
};
#include "chain_imagefloat.inc"

namespace curves {
#include "chain_setlutval.inc"
}
class ToneCurve { public: LUTf lutToneCurve; float whitecoeff; float whitept; const Curve* curve; ToneCurve() : whitecoeff(1.f), whitept(65535.f), curve(nullptr) {} };
class StandardToneCurve : public ToneCurve { public: void Apply(float& r, float& g, float& b) const; };
class AdobeToneCurve : public ToneCurve { void RGBTone(float& r, float& g, float& b) const; public: void Apply(float& r, float& g, float& b) const; };
class SatAndValueBlendingToneCurve : public ToneCurve { public: void Apply(float& r, float& g, float& b) const; };
class WeightedStdToneCurve : public ToneCurve { float Triangle(float refX, float refY, float X2) const; public: void Apply(float& r, float& g, float& b) const; };
class LuminanceToneCurve : public ToneCurve { public: void Apply(float& r, float& g, float& b, const float ws[3][3]) const; };
#include "chain_tonecurves.inc"

namespace {
#include "chain_iptonecurve.inc"
#include "chain_vibrance.inc"
#include "chain_prophotoblue.inc"
}

extern "C" {
int artref_chain_expcomp(float* R, float* G, float* B, int W, int H, float exp_scale, float black)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    Imagefloat* img = &im;
    const bool multiThread = true;
    vfloat exp_scalev = F2V(exp_scale);
    vfloat blackv = F2V(black);
    float **chan[3] = { img->r.ptrs, img->g.ptrs, img->b.ptrs };
#pragma omp parallel for if (multiThread)
#include "chain_expcomp_loop.inc"
    im.store(R, G, B);
    return 0;
}
int artref_chmixer(float* R, float* G, float* B, int W, int H, const float* m)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    Imagefloat* img = &im;
    const bool multiThread = true;
    const float RR = m[0], RG = m[1], RB = m[2], GR = m[3], GG = m[4], GB = m[5], BR = m[6], BG = m[7], BB = m[8];
    vfloat vRR = F2V(RR), vRG = F2V(RG), vRB = F2V(RB), vGR = F2V(GR), vGG = F2V(GG), vGB = F2V(GB), vBR = F2V(BR), vBG = F2V(BG), vBB = F2V(BB);
#pragma omp parallel for if (multiThread)
#include "chain_chmixer_loop.inc"
    im.store(R, G, B);
    return 0;
}
int artref_chain_saturation(float* R, float* G, float* B, int W, int H, int sat, int vibr, const double* wsd)
{
    Imagefloat im(W, H, R, G, B, wsd, nullptr);
    Imagefloat* rgb = &im;
    const bool multiThread = true;
    float wsm[3][3]; for (int i = 0; i < 9; ++i) (&wsm[0][0])[i] = (float)wsd[i];
    TMatrix ws = wsm;
    const float saturation = 1.f + sat / 100.f;
    const float vibrance = 1.f - vibr / 1000.f;
    const float noise = pow_F(2.f, -16.f);
    const bool vib = vibr;
#pragma omp parallel for if (multiThread)
#include "chain_saturation_loop.inc"
    im.store(R, G, B);
    return 0;
}
// mode 0 = ToneCurveParams::TcMode::STD, 1 = FILMLIKE; lut = ToneCurve::lutToneCurve (65536 entries); whitept = 1
int artref_chain_tonecurve(float* R, float* G, float* B, int W, int H, int mode, const float* lut, float whitept)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    filmlike_clip(&im, whitept, true);
    if (mode == 0) {
        StandardToneCurve c; c.lutToneCurve(65536); for (int i = 0; i < 65536; ++i) c.lutToneCurve[i] = lut[i];
        c.whitecoeff = whitept; c.whitept = 65535.f * whitept;
        apply(c, &im, W, H, true);
    } else {
        AdobeToneCurve c; c.lutToneCurve(65536); for (int i = 0; i < 65536; ++i) c.lutToneCurve[i] = lut[i];
        c.whitecoeff = whitept; c.whitept = 65535.f * whitept;
        apply(c, &im, W, H, true);
    }
    im.store(R, G, B);
    return 0;
}
// apply_tc's other per-pixel classes (iptonecurve.cc L48-85): mode 3 = WEIGHTEDSTD, 4 = SATANDVALBLENDING, 5 = LUMINANCE (ws = the
// float TMatrix of the working space), after the filmlike_clip pass ImProcFunctions::toneCurve runs first (L581-589)
int artref_chain_tonecurve_ex(float* R, float* G, float* B, int W, int H, int mode, const float* lut, float whitept, const double* wsd)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    Imagefloat* rgb = &im;
    filmlike_clip(&im, whitept, true);
    if (mode == 3) {
        WeightedStdToneCurve c; c.lutToneCurve(65536); for (int i = 0; i < 65536; ++i) c.lutToneCurve[i] = lut[i];
        c.whitecoeff = whitept; c.whitept = 65535.f * whitept;
        apply(c, &im, W, H, true);
    } else if (mode == 4) {
        SatAndValueBlendingToneCurve c; c.lutToneCurve(65536); for (int i = 0; i < 65536; ++i) c.lutToneCurve[i] = lut[i];
        c.whitecoeff = whitept; c.whitept = 65535.f * whitept;
        apply(c, &im, W, H, true);
    } else if (mode == 5) {
        LuminanceToneCurve c_; c_.lutToneCurve(65536); for (int i = 0; i < 65536; ++i) c_.lutToneCurve[i] = lut[i];
        c_.whitecoeff = whitept; c_.whitept = 65535.f * whitept;
        const LuminanceToneCurve& c = c_;
        float wsm[3][3]; for (int i = 0; i < 9; ++i) (&wsm[0][0])[i] = (float)wsd[i];
        TMatrix ws = wsm;
        const bool multithread = true;
#pragma omp parallel for if (multithread)
#include "chain_lumtone_loop.inc"
    } else return -1;
    im.store(R, G, B);
    return 0;
}
// proPhotoBlue (improcfun.cc L312-357): STAGE_1's last step when the working profile is ProPhoto
int artref_prophoto_blue(float* R, float* G, float* B, int W, int H)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    proPhotoBlue(&im, true);
    im.store(R, G, B);
    return 0;
}
int artref_chain_rgbcurves(float* R, float* G, float* B, int W, int H, const float* rc, const float* gc, const float* bc)
{
    Imagefloat im(W, H, R, G, B, (const double[9]){1,0,0,0,1,0,0,0,1}, nullptr);
    Imagefloat* img = &im;
    const bool multiThread = true;
    LUTf rCurve, gCurve, bCurve;
    if (rc) { rCurve(65536, 0); for (int i = 0; i < 65536; ++i) rCurve[i] = rc[i]; }
    if (gc) { gCurve(65536, 0); for (int i = 0; i < 65536; ++i) gCurve[i] = gc[i]; }
    if (bc) { bCurve(65536, 0); for (int i = 0; i < 65536; ++i) bCurve[i] = bc[i]; }
    if (rCurve || gCurve || bCurve) {
#pragma omp parallel for if (multiThread)
#include "chain_rgbcurves_loop.inc"
    }
    im.store(R, G, B);
    return 0;
}
// labAdjustments: setMode(LAB), lab_adjustments' pixel loop, setMode(RGB); lcurve 32770, acurve / bcurve 65536 entries
int artref_chain_lab(float* R, float* G, float* B, int W, int H, const float* lc, const float* ac, const float* bc, float chroma,
                     const double* wsd, const double* iwsd)
{
    Color::init();
    Imagefloat im(W, H, R, G, B, wsd, iwsd);
    Imagefloat* img = &im;
    const bool multiThread = true;
    im.rgb_to_lab(true);
    LUTf lcurve, acurve, bcurve;
    lcurve(32770, 0); acurve(65536); bcurve(65536);
    for (int i = 0; i < 32770; ++i) lcurve[i] = lc[i];
    for (int i = 0; i < 65536; ++i) { acurve[i] = ac[i]; bcurve[i] = bc[i]; }
    const vfloat chromav = F2V(chroma);
    const vfloat v32768 = F2V(32768.f);
#pragma omp parallel for if (multiThread)
#include "chain_lab_loop.inc"
    im.lab_to_rgb(true);
    im.store(R, G, B);
    return 0;
}
int artref_chain_rgb2lab(float* R, float* G, float* B, int W, int H, const double* wsd, const double* iwsd, int back)
{
    Color::init();
    Imagefloat im(W, H, R, G, B, wsd, iwsd);
    if (back) im.lab_to_rgb(true); else im.rgb_to_lab(true);
    im.store(R, G, B);
    return 0;
}
}
}  // namespace artref_chain
"""


SHIM_USM_TU = r"""
// Shim TU hosting the reference's unsharp-mask sharpening: apply_gamma, sharpenHaloCtrl, unsharp_mask cut from ipsharpen.cc;
// calcBlendFactor, tileAverage, tileVariance, calcContrastThreshold, buildBlendMask, get_luminance, multiply cut from
// rt_algo.cc; Threshold<T> cut from procparams.h; Color::rgbLuminance cut from color.h; gaussianBlur from shim_gauss.cc.
// bilateral<T, A> and its 21 fixed-kernel variants cut from bilateral2.h (from its ELEM macro to the end of the dispatcher).
// Written here (not reference code): the Imagefloat / SharpeningParams stand-ins and the wrapper, which restates the "usm"
// route of ImProcFunctions::doSharpening (ipsharpen.cc L711-790).
#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <type_traits>
#include <vector>
#include <omp.h>
#include "array2D.h"
#include "jaggedarray.h"
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#include "sleef.h"
#include "gauss.h"
#define BENCHFUN
namespace artref_usm {
using namespace rtengine;
typedef const float (*TMatrix)[3];      // iccstore.h L38
class Color { public:
template <class T>
#include "usm_rgblum.inc"
};
#include "usm_threshold.inc"
;
struct SharpeningParams { double contrast, radius; int amount; Threshold<int> threshold; bool edgesonly; double edges_radius; int edges_tolerance;
                          bool halocontrol; int halocontrol_amount;
                          SharpeningParams() : contrast(20.0), radius(0.5), amount(200), threshold(20, 80, 2000, 1200, false), edgesonly(false), edges_radius(1.9),
                                               edges_tolerance(1800), halocontrol(false), halocontrol_amount(85) {} };
#include "usm_bilateral.inc"
struct Chan { float* base; int W; float& operator()(int y, int x) const { return base[(size_t)y * W + x]; } };
class Imagefloat { public: int width, height; Chan r, g, b; int getWidth() const { return width; } int getHeight() const { return height; } };
namespace {
#include "usm_rtalgo_anon.inc"
}
#include "usm_rtalgo.inc"
namespace {
#include "usm_ipsharpen.inc"
}

extern "C" int artref_usm_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                             const int* thr, int halocontrol, int halocontrol_amount, float* blend_out, int edgesonly, double edges_radius, int edges_tolerance);
extern "C" int artref_usm(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                          const int* thr, int halocontrol, int halocontrol_amount, float* blend_out)
{
    return artref_usm_ex(R, G, B, W, H, wsd, scale, contrast_p, radius, amount, thr, halocontrol, halocontrol_amount, blend_out, 0, 1.9, 1800);
}
extern "C" int artref_usm_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double radius, int amount,
                             const int* thr, int halocontrol, int halocontrol_amount, float* blend_out, int edgesonly, double edges_radius, int edges_tolerance)
{
    if (amount < 1 || W < 8 || H < 8) return 0;                    // doSharpening L716-718
    const bool multiThread = true;
    Imagefloat im{W, H, {R, W}, {G, W}, {B, W}};
    Imagefloat* rgb = &im;
    SharpeningParams sharpenParam;
    sharpenParam.contrast = contrast_p; sharpenParam.radius = radius; sharpenParam.amount = amount;
    sharpenParam.threshold = Threshold<int>(thr[0], thr[1], thr[2], thr[3], false);
    sharpenParam.halocontrol = halocontrol != 0; sharpenParam.halocontrol_amount = halocontrol_amount;
    sharpenParam.edgesonly = edgesonly != 0; sharpenParam.edges_radius = edges_radius; sharpenParam.edges_tolerance = edges_tolerance;
    float wsm[3][3]; for (int i = 0; i < 9; ++i) (&wsm[0][0])[i] = (float)wsd[i];
    TMatrix ws = wsm;
    array2D<float> Y(ARRAY2D_ALIGNED);
    get_luminance(rgb, Y, ws, multiThread);
    float s_scale = std::sqrt(scale);
    float contrast = pow_F(sharpenParam.contrast / 100.f, 1.2f) * s_scale;
    JaggedArray<float> blend(W, H);
    buildBlendMask(Y, blend, W, H, contrast, 1.f, false, 2.f / s_scale, 1.f);
    if (blend_out) for (int y = 0; y < H; ++y) memcpy(blend_out + (size_t)y * W, blend[y], sizeof(float) * W);
    array2D<float> YY(W, H, Y, ARRAY2D_ALIGNED);
    unsharp_mask(YY, blend, W, H, sharpenParam, scale, multiThread);
    multiply(rgb, YY, Y, multiThread);
    return 0;
}

// buildBlendMask(luminance, blend, W, H, contrastThreshold, amount, autoContrast, blur_radius, 1.f) on contiguous planes; the threshold is in / out
extern "C" int artref_blend_mask_ex(const float* lum, float* blend_out, int W, int H, float* contrastThreshold, float amount, int autoContrast, float blur_radius);
extern "C" int artref_blend_mask(const float* lum, float* blend_out, int W, int H, float contrastThreshold, float amount, float blur_radius)
{
    return artref_blend_mask_ex(lum, blend_out, W, H, &contrastThreshold, amount, 0, blur_radius);
}
extern "C" int artref_blend_mask_ex(const float* lum, float* blend_out, int W, int H, float* contrastThreshold, float amount, int autoContrast, float blur_radius)
{
    float** l = new float*[H];
    for (int y = 0; y < H; ++y) l[y] = const_cast<float*>(lum) + (size_t)y * W;
    JaggedArray<float> blend(W, H);
    buildBlendMask(l, blend, W, H, *contrastThreshold, amount, autoContrast != 0, blur_radius, 1.f);
    for (int y = 0; y < H; ++y) memcpy(blend_out + (size_t)y * W, blend[y], sizeof(float) * W);
    delete[] l;
    return 0;
}

// the "rld" route of doSharpening (ipsharpen.cc L747-771 without the corner boost): markImpulse(Y, 2), deconvsharpening(copy of Y), multiply
extern "C" int artref_rld_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double deconvradius,
                             int deconvamount, float* impulse_out, double deconvCornerBoost, int deconvCornerLatitude, int offset_x, int offset_y,
                             int full_width, int full_height);
extern "C" int artref_rld(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double deconvradius,
                          int deconvamount, float* impulse_out)
{
    return artref_rld_ex(R, G, B, W, H, wsd, scale, contrast_p, deconvradius, deconvamount, impulse_out, 0.0, 25, 0, 0, 0, 0);
}
// doSharpening L753-771 with the corner boost: two deconvolutions (sigma, sigma + delta) mixed by CornerBoostMask
extern "C" int artref_rld_ex(float* R, float* G, float* B, int W, int H, const double* wsd, double scale, double contrast_p, double deconvradius,
                             int deconvamount, float* impulse_out, double deconvCornerBoost, int deconvCornerLatitude, int offset_x, int offset_y,
                             int full_width, int full_height)
{
    if (W < 8 || H < 8) return 0;
    const bool multiThread = true;
    Imagefloat im{W, H, {R, W}, {G, W}, {B, W}};
    Imagefloat* rgb = &im;
    float wsm[3][3]; for (int i = 0; i < 9; ++i) (&wsm[0][0])[i] = (float)wsd[i];
    TMatrix ws = wsm;
    array2D<float> Y(ARRAY2D_ALIGNED);
    get_luminance(rgb, Y, ws, multiThread);
    float s_scale = std::sqrt(scale);
    float contrast = pow_F(contrast_p / 100.f, 1.2f) * s_scale;
    JaggedArray<float> blend(W, H);
    buildBlendMask(Y, blend, W, H, contrast, 1.f, false, 2.f / s_scale, 1.f);
    JaggedArray<char> impulse(W, H);
    markImpulse(W, H, Y, impulse, 2.f);
    if (impulse_out) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) impulse_out[(size_t)y * W + x] = impulse[y][x];
    array2D<float> YY(W, H, Y, ARRAY2D_ALIGNED);
    double sigma = deconvradius / scale;
    float amount = deconvamount / 100.f;
    float delta = deconvCornerBoost / scale;
    if (delta > 0.01f) {
        array2D<float> YY2(W, H, Y, ARRAY2D_ALIGNED);
        deconvsharpening(YY, blend, impulse, W, H, sigma, amount, multiThread);
        deconvsharpening(YY2, blend, impulse, W, H, sigma + delta, amount, multiThread);
        int fw = full_width > 0 ? full_width : W;
        int fh = full_height > 0 ? full_height : H;
        CornerBoostMask mask(offset_x, offset_y, fw, fh, deconvCornerLatitude);
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                float blend = mask(x, y);
                YY[y][x] = intp(blend, YY2[y][x], YY[y][x]);
            }
        }
    } else {
        deconvsharpening(YY, blend, impulse, W, H, sigma, amount, multiThread);
    }
    multiply(rgb, YY, Y, multiThread);
    return 0;
}
}  // namespace artref_usm
"""


SHIM_XTRANS_TU = r"""
// Shim TU hosting the reference's X-Trans demosaic: the xyz_rgb / d65_white constants, RawImageSource::cielab,
// xtransborder_interpolate and xtrans_interpolate (Markesteijn) cut from xtrans_demosaic.cc at build time, inside a stand-in
// class holding only the members those bodies touch.  ARTREF_DET clears the per-thread tile buffer at the start of every tile.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <algorithm>
#include <omp.h>
#include "glibmm.h"
#include "array2D.h"
#include "LUT.h"
#include "rt_math.h"
#include "opthelper.h"
#define M(x) Glib::ustring(x)
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
#define LIM(x, lo, hi) rtengine::LIM(x, lo, hi)
#define SQR(x) rtengine::SQR(x)
namespace artref_xt {
using namespace rtengine;
typedef unsigned short ushort;
struct ProgressListener { void setProgressStr(const Glib::ustring&) {} void setProgress(double) {} };
struct StopWatch { StopWatch(const char*) {} };
struct Color { constexpr static double eps = 216.0 / 24389.0; constexpr static double kappa = 24389.0 / 27.0; };   // color.h
struct RawImage {
    int xt[6][6]; float cam[3][4];
    void getXtransMatrix(int m[6][6]) const { memcpy(m, xt, sizeof xt); }
    void getRgbCam(float m[3][4]) const { memcpy(m, cam, sizeof cam); }
};
struct RawImageSource {
    int W, H; RawImage* ri; ProgressListener* plistener;
    array2D<float> rawData, red, green, blue;
    RawImageSource(int w, int h, RawImage* r_, float** raw, float** r, float** g, float** b)
        : W(w), H(h), ri(r_), plistener(nullptr), rawData(w, h, raw, ARRAY2D_BYREFERENCE), red(w, h, r, ARRAY2D_BYREFERENCE),
          green(w, h, g, ARRAY2D_BYREFERENCE), blue(w, h, b, ARRAY2D_BYREFERENCE) {}
    void cielab(const float (*rgb)[3], float* l, float* a, float* b, const int width, const int height, const int labWidth, const float xyz_cam[3][3]);
    void xtransborder_interpolate(int border, array2D<float>& red, array2D<float>& green, array2D<float>& blue);
    void xtrans_interpolate(const int passes, const bool useCieLab);
    void fast_xtrans_interpolate_blend (const float* const * blend, const array2D<float> &rawData, array2D<float> &red, array2D<float> &green, array2D<float> &blue);
};
#include "xtrans_body.inc"
}  // namespace artref_xt

extern "C" int artref_xtrans(int W, int H, const int* xtrans36, const float* rgb_cam12, int passes, int useCieLab, const float* raw,
                             float* r, float* g, float* b, int border_only, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    float **rr = new float*[H], **gr = new float*[H], **br = new float*[H], **wr = new float*[H];
    for (int i = 0; i < H; ++i) { rr[i] = r + (size_t)i * W; gr[i] = g + (size_t)i * W; br[i] = b + (size_t)i * W; wr[i] = const_cast<float*>(raw) + (size_t)i * W; }
    artref_xt::RawImage ri;
    memcpy(ri.xt, xtrans36, sizeof ri.xt); memcpy(ri.cam, rgb_cam12, sizeof ri.cam);
    artref_xt::RawImageSource src(W, H, &ri, wr, rr, gr, br);
    if (border_only) src.xtransborder_interpolate(border_only, src.red, src.green, src.blue);
    else src.xtrans_interpolate(passes, useCieLab != 0);
    delete[] rr; delete[] gr; delete[] br; delete[] wr;
    return 0;
}
// fast_xtrans_interpolate_blend(blend, rawData, red, green, blue): the second half of the dual demosaic on X-Trans (dual_demosaic_RT.cc L151)
extern "C" int artref_xtrans_fast_blend(int W, int H, const int* xtrans36, const float* raw, const float* blend, float* r, float* g, float* b)
{
    float **rr = new float*[H], **gr = new float*[H], **br = new float*[H], **wr = new float*[H], **bl = new float*[H];
    for (int i = 0; i < H; ++i) { rr[i] = r + (size_t)i * W; gr[i] = g + (size_t)i * W; br[i] = b + (size_t)i * W; wr[i] = const_cast<float*>(raw) + (size_t)i * W; bl[i] = const_cast<float*>(blend) + (size_t)i * W; }
    artref_xt::RawImage ri;
    memcpy(ri.xt, xtrans36, sizeof ri.xt); memset(ri.cam, 0, sizeof ri.cam);
    artref_xt::RawImageSource src(W, H, &ri, wr, rr, gr, br);
    src.fast_xtrans_interpolate_blend(bl, src.rawData, src.red, src.green, src.blue);
    delete[] rr; delete[] gr; delete[] br; delete[] wr; delete[] bl;
    return 0;
}
"""


def extract(det):
    sub = os.path.join(SRC, "det" if det else "stock")
    os.makedirs(sub, exist_ok=True)
    rcd = cut_function(os.path.join(RT, "rcd_demosaic.cc"), r"void\s+RawImageSource::rcd_demosaic\s*\(\s*\)")
    amaze = cut_function(os.path.join(RT, "amaze_demosaic_RT.cc"),
                         r"void\s+RawImageSource::amaze_demosaic_RT\s*\([^)]*\)")
    border = cut_function(os.path.join(RT, "demosaic_algos.cc"),
                          r"void\s+RawImageSource::border_interpolate2\s*\([^)]*\)")
    if det:
        # AMaZE: clear the whole per-thread scratch at the top of every tile
        amaze = insert_before(
            amaze, r"memset\(&nyquist\[3 \* tsh\]",
            "memset(data, 0, 14 * sizeof(float) * ts * ts + sizeof(char) * ts * tsh + 18 * cldf * 64);\n                ")
        # RCD: same idea -- VH_Dir outside [4,rows-4)x[4,cols-4) and friends read as 0
        rcd = insert_after(
            rcd, r"const int tilecols = [^;]*;",
            "\n            memset(cfa, 0, sizeof(float) * tileSize * tileSize);"
            "\n            memset(rgb, 0, 3 * sizeof *rgb);"
            "\n            memset(VH_Dir, 0, sizeof(float) * tileSize * tileSize);"
            "\n            memset(PQ_Dir, 0, sizeof(float) * (tileSize * tileSize / 2));"
            "\n            memset(P_CDiff_Hpf, 0, sizeof(float) * (tileSize * tileSize / 2));"
            "\n            memset(Q_CDiff_Hpf, 0, sizeof(float) * (tileSize * tileSize / 2));")
    open(os.path.join(sub, "rcd_body.inc"), "w").write(rcd)
    open(os.path.join(sub, "amaze_body.inc"), "w").write(amaze)
    open(os.path.join(sub, "border_body.inc"), "w").write(border)
    csconv = cut_block(os.path.join(RT, "rawimagesource.cc"),
                       r"for \(int i = 0; i < im->getHeight\(\); i\+\+\)\s*for \(int j = 0; j < im->getWidth\(\); j\+\+\) \{(?=\s*float newr = mat\[0\]\[0\])")
    open(os.path.join(sub, "csconv_loop.inc"), "w").write(csconv)
    scl = cut_block(os.path.join(RT, "rawimagesource.cc"),
                    r"for \(int row = winy; row < winy \+ winh; row \+\+\)\s*\{(?=\s*for \(int col = winx; col < winx \+ winw; col\+\+\) \{\s*const int c  = FC\(row, col\);)")
    open(os.path.join(sub, "scalecolors_loop.inc"), "w").write(scl)
    sclx = cut_block(os.path.join(RT, "rawimagesource.cc"),
                     r"for \(int row = winy; row < winy \+ winh; row \+\+\)\s*\{(?=\s*for \(int col = winx; col < winx \+ winw; col\+\+\) \{\s*const int c = ri->XTRANSFC\(row, col\);)")
    open(os.path.join(sub, "scalecolors_xtrans_loop.inc"), "w").write(sclx)
    open(os.path.join(sub, "glibmm.h"), "w").write(SHIM_GLIBMM)
    open(os.path.join(sub, "shim.cc"), "w").write(SHIM_TU)
    gtext = open(os.path.join(RT, "gauss.cc"), encoding="utf-8", errors="replace").read()
    m = re.search(r"^namespace \{", gtext, flags=re.M)
    open(os.path.join(sub, "gauss_body.inc"), "w").write(gtext[m.start():])
    open(os.path.join(sub, "shim_gauss.cc"), "w").write(SHIM_GAUSS_TU)
    btext = open(os.path.join(RT, "boxblur.h"), encoding="utf-8", errors="replace").read()
    m = re.search(r"^namespace rtengine", btext, flags=re.M)
    e = btext.rindex("#endif")
    open(os.path.join(sub, "boxblur_body.inc"), "w").write(btext[m.start():e])
    gf = os.path.join(RT, "guidedfilter.cc")
    open(os.path.join(sub, "guided_subsampling.inc"), "w").write(cut_function(gf, r"int calculate_subsampling\(int w, int h, int r\)"))
    open(os.path.join(sub, "guided_body.inc"), "w").write(cut_function(gf, r"void guidedFilter\(const array2D<float> &guide[^)]*\)"))
    open(os.path.join(sub, "guided_log.inc"), "w").write(cut_function(gf, r"^void guidedFilterLog\(const array2D<float> &guide, float base[^)]*\)"))
    open(os.path.join(sub, "guided_log2.inc"), "w").write(cut_function(gf, r"^void guidedFilterLog\(float base, array2D<float> &chan[^)]*\)"))
    open(os.path.join(sub, "guided_smoothing.inc"), "w").write(cut_function(os.path.join(RT, "ipsmoothing.cc"), r"^void guided_smoothing\(array2D<float> &R[^)]*\)"))
    chh = os.path.join(RT, "color.h")
    open(os.path.join(sub, "gs_rgblum.inc"), "w").write(cut_function(chh, r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"))
    open(os.path.join(sub, "gs_rgb2yuv.inc"), "w").write(cut_function(chh, r"static void rgb2yuv\(float r, float g, float b, float &Y"))
    open(os.path.join(sub, "gs_yuv2rgb.inc"), "w").write(cut_function(chh, r"static void yuv2rgb\(float Y, float u, float v, float &r"))
    open(os.path.join(sub, "shim_guided.cc"), "w").write(SHIM_GUIDED_TU)
    open(os.path.join(sub, "shim_wavelet.cc"), "w").write(SHIM_WAVELET_TU)
    ft = os.path.join(RT, "FTblockDN.cc")
    parts = [cut_function(ft, r"^float MadRgb\(float \* DataList, const int datalen\)"),
             cut_function(ft, r"^void ShrinkAllL\(double scale,[^)]*\)"),
             cut_function(ft, r"^void ShrinkAllAB\(double scale,[^)]*\)"),
             cut_function(ft, r"^bool WaveletDenoiseAllL\(double scale,[^)]*\)"),
             cut_function(ft, r"^bool WaveletDenoiseAllAB\(double scale,[^)]*\)")]
    open(os.path.join(sub, "shrink_body.inc"), "w").write("\n\n".join(parts))
    open(os.path.join(sub, "shim_shrink.cc"), "w").write(SHIM_SHRINK_TU)
    ft = os.path.join(RT, "FTblockDN.cc")
    open(os.path.join(sub, "laplacian_body.inc"), "w").write(cut_function(ft, r"^void laplacian\(const array2D<float> &src[^)]*\)"))
    open(os.path.join(sub, "detail_mask_body.inc"), "w").write(cut_function(ft, r"^void detail_mask\(const array2D<float> &src[^)]*\)"))
    open(os.path.join(sub, "nlmeans_body.inc"), "w").write(cut_function(os.path.join(RT, "nlmeans.cc"), r"^void NLMeans\(array2D<float> &img[^)]*\)"))
    open(os.path.join(sub, "shim_nlmeans.cc"), "w").write(SHIM_NLMEANS_TU)
    fttext = open(ft, encoding="utf-8", errors="replace").read()
    m = re.search(r"^#define TS 64", fttext, flags=re.M)
    open(os.path.join(sub, "ftblockdn_body.inc"), "w").write(fttext[m.start():])
    ch, cc = os.path.join(RT, "color.h"), os.path.join(RT, "color.cc")
    members = [cut_function(ch, r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"),
               cut_function(ch, r"static void rgb2yuv\(float r, float g, float b, float &Y"),
               cut_function(ch, r"static void yuv2rgb\(float Y, float u, float v, float &r"),
               cut_function(ch, r"static inline float gammaf\s*\(float x, float gamma, float start, float slope\)"),
               cut_function(ch, r"static inline float gammanf\s*\(float x, float gamma\)")]
    labm = ["template <class T>\n" + cut_function(ch, r"static void rgb2lab\(float R, float G, float B, float &l, float &a, float &b, const T ws\[3\]\[3\]\)"),
            "template <class T>\n" + cut_function(ch, r"static void lab2rgb\(float l, float a, float b, float &R, float &G, float &B, const T iws\[3\]\[3\]\)"),
            cut_function(ch, r"static inline float f2xyz\(float f\)"),
            cut_function(ch, r"static inline double gamma55\(double x\)"), cut_function(ch, r"static inline double igamma55\(double x\)")]
    open(os.path.join(sub, "color_h_lab_members.inc"), "w").write("\n".join(labm))
    open(os.path.join(sub, "color_h_members.inc"), "w").write("\n".join(("template <class T>\n" if "workingspace" in t else "") + t for t in members))
    ccm = [cut_function(cc, r"^inline float Color::computeXYZ2Lab\(float f\)"), cut_function(cc, r"^inline float Color::computeXYZ2LabY\(float f\)"),
           cut_function(cc, r"^void Color::gammaf2lut \(LUTf &gammacurve"),
           cut_function(cc, r"^void Color::rgbxyz \(float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb"),
           cut_function(cc, r"^void Color::XYZ2Lab\(float X, float Y, float Z, float &L"),
           cut_function(cc, r"^void Color::xyz2rgb \(float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz"),
           cut_function(cc, r"^void Color::Lab2XYZ\(float L, float a, float b, float &x, float &y, float &z\)"),
           cut_function(cc, r"^void Color::RGB2L\(float \*R, float \*G, float \*B, float \*L, const float wp\[3\]\[3\], int width\)")]
    open(os.path.join(sub, "color_cc_members.inc"), "w").write("\n".join(ccm))
    open(os.path.join(sub, "dct_standin.h"), "w").write(open(os.path.join(HERE, "dct_standin.h")).read())
    ipd = os.path.join(RT, "ipdenoise.cc")
    info = [cut_function(ipd, r"^void calcautodn_info\(const ProcParams \*params[^)]*\)"),
            cut_function(ipd, r"^void RGB_denoise_infoGamCurve\(const procparams::DenoiseParams & dnparams[^)]*\)"),
            cut_function(ipd, r"^void RGB_denoise_info\(ImProcData &im[^)]*\)")]
    open(os.path.join(sub, "ipdenoise_info.inc"), "w").write("\n\n".join(info))
    iptext = open(ipd, encoding="utf-8", errors="replace").read()
    c0 = re.search(r"^        float chM = 0\.f;", iptext, flags=re.M)
    c1 = re.search(r"^        store\.chrominanceBlueYellow = maxb;\n", iptext, flags=re.M)
    open(os.path.join(sub, "ipdenoise_combine.inc"), "w").write(iptext[c0.start():c1.end()])
    open(os.path.join(sub, "shim_denoise.cc"), "w").write(SHIM_DENOISE_TU)
    fat = open(os.path.join(RT, "tmo_fattal02.cc"), encoding="utf-8", errors="replace").read()
    m0 = re.search(r"^namespace rtengine\s*\{", fat, flags=re.M)
    m1 = re.search(r"^void ImProcFunctions::dynamicRangeCompression", fat, flags=re.M)
    body = fat[m0.start():m1.start()]
    # ToneMapFattal02 sits in the anonymous namespace; the wrapper appended by the shim is in the same TU so it can call it
    open(os.path.join(sub, "fattal_body.inc"), "w").write(body)
    open(os.path.join(sub, "fattal_color_members.inc"), "w").write(
        "template <class T>\n" + cut_function(ch, r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"))
    open(os.path.join(sub, "shim_fattal.cc"), "w").write(SHIM_FATTAL_TU)
    # ---- colour / curve chain
    def block_after(path, anchor):
        return cut_block(path, anchor)
    ipx = os.path.join(RT, "ipexposure.cc")
    open(os.path.join(sub, "chain_expcomp_loop.inc"), "w").write(
        block_after(ipx, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < W - 3; x \+= 4\) \{\s*for \(int c = 0; c < 3; \+\+c\))"))
    ipm = os.path.join(RT, "ipchmixer.cc")
    open(os.path.join(sub, "chain_chmixer_loop.inc"), "w").write(
        block_after(ipm, r"for \(int y = 0; y < img->getHeight\(\); \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < img->getWidth\(\)-3; x \+= 4\) \{\s*vfloat r = LVF\(img->r\(y, x\)\);)"))
    ips = os.path.join(RT, "ipsaturation.cc")
    open(os.path.join(sub, "chain_vibrance.inc"), "w").write(cut_function(ips, r"^float apply_vibrance\(float x, float vib\)"))
    open(os.path.join(sub, "chain_saturation_loop.inc"), "w").write(
        block_after(ips, r"for \(int i = 0; i < H; \+\+i\) \{(?=\s*for \(int j = 0; j < W; \+\+j\) \{\s*float &r = rgb->r\(i, j\);)"))
    ipr = os.path.join(RT, "iprgbcurves.cc")
    open(os.path.join(sub, "chain_rgbcurves_loop.inc"), "w").write(
        block_after(ipr, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < W-3; x \+= 4\) \{\s*if \(rCurve\))"))
    ipl = os.path.join(RT, "iplabadjustments.cc")
    open(os.path.join(sub, "chain_lab_loop.inc"), "w").write(
        block_after(ipl, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*int x = 0;\s*#ifdef __SSE2__\s*for \(; x < W-3; x \+= 4\) \{\s*vfloat L = LVF)"))
    ipt = os.path.join(RT, "iptonecurve.cc")
    open(os.path.join(sub, "chain_iptonecurve.inc"), "w").write(
        "template <class Curve>\n" + cut_function(ipt, r"^inline void apply\(const Curve &c, Imagefloat \*rgb, int W, int H, bool multithread\)") + "\n" +
        cut_function(ipt, r"^void filmlike_clip\(Imagefloat \*rgb, float whitept, bool multithread\)"))
    cvh = os.path.join(RT, "curves.h")
    open(os.path.join(sub, "chain_setlutval.inc"), "w").write(cut_function(cvh, r"^inline void setLutVal\(const LUTf &lut, const Curve \*curve, float &val\)"))
    open(os.path.join(sub, "chain_tonecurves.inc"), "w").write("\n".join([
        cut_function(cvh, r"^inline void StandardToneCurve::Apply \(float& r, float& g, float& b\) const"),
        cut_function(cvh, r"^inline void AdobeToneCurve::RGBTone \(float& r, float& g, float& b\) const"),
        cut_function(cvh, r"^inline void AdobeToneCurve::Apply \(float& ir, float& ig, float& ib\) const"),
        cut_function(cvh, r"^inline void LuminanceToneCurve::Apply\(float &ir, float &ig, float &ib, const float ws\[3\]\[3\]\) const"),
        cut_function(cvh, r"^inline float WeightedStdToneCurve::Triangle\(float a, float a1, float b\) const"),
        cut_function(cvh, r"^inline void WeightedStdToneCurve::Apply \(float& ir, float& ig, float& ib\) const"),
        cut_function(cvh, r"^inline void SatAndValueBlendingToneCurve::Apply \(float& ir, float& ig, float& ib\) const")]))
    open(os.path.join(sub, "chain_lumtone_loop.inc"), "w").write(
        block_after(ipt, r"for \(int y = 0; y < H; \+\+y\) \{(?=\s*for \(int x = 0; x < W; \+\+x\) \{\s*c\.Apply\(rgb->r\(y, x\), rgb->g\(y, x\), rgb->b\(y, x\), ws\);)"))
    imf = os.path.join(RT, "imagefloat.cc")
    open(os.path.join(sub, "chain_imagefloat.inc"), "w").write("\n".join([
        cut_function(imf, r"^inline void Imagefloat::rgb_to_lab\(int y, int x, float &L, float &a, float &b\)"),
        cut_function(imf, r"^void Imagefloat::rgb_to_lab\(bool multithread\)"),
        cut_function(imf, r"^void Imagefloat::lab_to_rgb\(bool multithread\)")]))
    hm = [cut_function(ch, r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"),
          cut_function(ch, r"static inline void rgb2hsvtc\(float r, float g, float b, float &h, float &s, float &v\)"),
          cut_function(ch, r"static inline void hsv2rgbdcp \(float h, float s, float v, float &r, float &g, float &b\)"),
          cut_function(ch, r"static inline float f2xyz\(float f\)"),
          cut_function(ch, r"static inline vfloat f2xyz\(vfloat f\)"),
          cut_function(ch, r"static void rgb2lab\(float R, float G, float B, float &l, float &a, float &b, const T ws\[3\]\[3\]\)"),
          cut_function(ch, r"static void lab2rgb\(float l, float a, float b, float &R, float &G, float &B, const T iws\[3\]\[3\]\)"),
          cut_function(ch, r"static void rgb2lab\(vfloat R, vfloat G, vfloat B, vfloat &l, vfloat &a, vfloat &b, const vfloat ws\[3\]\[3\]\)"),
          cut_function(ch, r"static void lab2rgb\(vfloat l, vfloat a, vfloat b, vfloat &R, vfloat &G, vfloat &B, const vfloat iws\[3\]\[3\]\)")]
    open(os.path.join(sub, "chain_color_h.inc"), "w").write("\n".join(("template <class T>\n" if "const T " in t else "") + t for t in hm))
    ccf = [cut_function(cc, r"^inline float Color::computeXYZ2Lab\(float f\)"), cut_function(cc, r"^inline float Color::computeXYZ2LabY\(float f\)"),
           cut_function(cc, r"^void Color::rgbxyz \(float r, float g, float b, float &x, float &y, float &z, const float xyz_rgb"),
           cut_function(cc, r"^void Color::rgbxyz \(vfloat r, vfloat g, vfloat b, vfloat &x, vfloat &y, vfloat &z, const vfloat xyz_rgb"),
           cut_function(cc, r"^void Color::xyz2rgb \(float x, float y, float z, float &r, float &g, float &b, const float rgb_xyz"),
           cut_function(cc, r"^void Color::xyz2rgb \(vfloat x, vfloat y, vfloat z, vfloat &r, vfloat &g, vfloat &b, const vfloat rgb_xyz"),
           cut_function(cc, r"^void Color::XYZ2Lab\(float X, float Y, float Z, float &L"),
           cut_function(cc, r"^void Color::XYZ2Lab\(vfloat X, vfloat Y, vfloat Z, vfloat &L"),
           cut_function(cc, r"^void Color::Lab2XYZ\(float L, float a, float b, float &x, float &y, float &z\)"),
           cut_function(cc, r"^void Color::Lab2XYZ\(vfloat L, vfloat a, vfloat b, vfloat &x, vfloat &y, vfloat &z\)"),
           cut_function(cc, r"^inline void filmlike_clip_rgb_tone\(float \*r, float \*g, float \*b, const float L\)"),
           cut_function(cc, r"^void Color::filmlike_clip\(float \*r, float \*g, float \*b, float Lmax\)"),
           cut_function(cc, r"^void Color::rgb2hsv\(float r, float g, float b, float &h, float &s, float &v\)"),
           cut_function(cc, r"^void Color::hsv2rgb \(float h, float s, float v, float &r, float &g, float &b\)")]
    open(os.path.join(sub, "chain_color_cc.inc"), "w").write("\n".join(ccf))
    open(os.path.join(sub, "chain_prophotoblue.inc"), "w").write(cut_function(os.path.join(RT, "improcfun.cc"), r"^void proPhotoBlue\(Imagefloat \*rgb, bool multiThread\)"))
    open(os.path.join(sub, "shim_chain.cc"), "w").write(SHIM_CHAIN_TU)

    # USM sharpening (ipsharpen.cc, rt_algo.cc)
    ra = os.path.join(RT, "rt_algo.cc")
    ish = os.path.join(RT, "ipsharpen.cc")
    open(os.path.join(sub, "usm_rgblum.inc"), "w").write(
        cut_function(os.path.join(RT, "color.h"), r"static float rgbLuminance\(float r, float g, float b, const T workingspace\[3\]\[3\]\)"))
    open(os.path.join(sub, "usm_threshold.inc"), "w").write(cut_function(os.path.join(RT, "procparams.h"), r"^template<typename T>\nclass Threshold final"))
    anon = [cut_function(ra, r"^float calcBlendFactor\(float val, float threshold\)"),
            cut_function(ra, r"^vfloat calcBlendFactor\(vfloat valv, vfloat thresholdv\)"),
            cut_function(ra, r"^float tileAverage\(float \*\*data[^)]*\)"),
            cut_function(ra, r"^float tileVariance\(float \*\*data[^)]*\)"),
            cut_function(ra, r"^float calcContrastThreshold\(float\*\* luminance[^)]*\)")]
    open(os.path.join(sub, "usm_rtalgo_anon.inc"), "w").write("\n\n".join(anon))
    pub = [cut_function(ra, r"^void markImpulse\(int width, int height, float \*\*const src, char \*\*impulse, float thresh\)"),
           cut_function(ra, r"^void buildBlendMask\(float\*\* luminance[^)]*\)"),
           cut_function(ra, r"^void get_luminance\(const Imagefloat \*src[^)]*\)"),
           cut_function(ra, r"^void multiply\(Imagefloat \*img[^)]*\)")]
    open(os.path.join(sub, "usm_rtalgo.inc"), "w").write("\n\n".join(pub))
    ips = ["template <bool reverse>\n" + cut_function(ish, r"^void apply_gamma\(float \*\*Y[^)]*\)"),
           cut_function(ish, r"^void sharpenHaloCtrl\(float\*\* luminance[^)]*\)"),
           cut_function(ish, r"^void deconvsharpening\(float \*\*luminance, float \*\*blend, char \*\*impulse[^)]*\)"),
           cut_function(ish, r"^class CornerBoostMask ") + ";",
           cut_function(ish, r"^void unsharp_mask\(float \*\*Y[^)]*\)")]
    open(os.path.join(sub, "usm_ipsharpen.inc"), "w").write("\n\n".join(ips))
    btext = open(os.path.join(RT, "bilateral2.h"), encoding="utf-8", errors="replace").read()
    b0 = re.search(r"^#define ELEM\(a,b\)", btext, flags=re.M)
    b1 = re.search(r"^// START OF EXPERIMENTAL CODE", btext, flags=re.M)
    open(os.path.join(sub, "usm_bilateral.inc"), "w").write(btext[b0.start():b1.start()])
    open(os.path.join(sub, "shim_usm.cc"), "w").write(SHIM_USM_TU)

    open(os.path.join(sub, "pack_getscanline.inc"), "w").write(
        cut_function(os.path.join(RT, "imagefloat.cc"), r"^void Imagefloat::getScanline \(int row, unsigned char\* buffer, int bps, bool isFloat\) const"))
    open(os.path.join(sub, "shim_pack.cc"), "w").write(SHIM_PACK_TU)
    open(os.path.join(sub, "bilinear_body.inc"), "w").write(
        cut_function(os.path.join(RT, "bayer_bilinear_demosaic.cc"), r"^void RawImageSource::bayer_bilinear_demosaic\(const float\* const \* blend[^)]*\)"))
    open(os.path.join(sub, "shim_bilinear.cc"), "w").write(SHIM_BILINEAR_TU)
    open(os.path.join(sub, "hlblend_body.inc"), "w").write(
        cut_function(os.path.join(RT, "rawimagesource.cc"), r"^void RawImageSource::HLRecovery_blend\(float\* rin, float\* gin, float\* bin, int width, float maxval, float\* hlmax\)"))
    open(os.path.join(sub, "shim_hlblend.cc"), "w").write(SHIM_HLBLEND_TU)
    open(os.path.join(sub, "getimage_rotateline.inc"), "w").write(
        cut_function(os.path.join(RT, "rawimagesource.cc"), r"^void rotateLine \(const float\* const line, rtengine::PlanarPtr<float> &channel, const int tran, const int i, const int w, const int h\)"))
    open(os.path.join(sub, "getimage_transformrect.inc"), "w").write(
        cut_function(os.path.join(RT, "rawimagesource.cc"), r"^void RawImageSource::transformRect \(const PreviewProps &pp, int tran, int &ssx1, int &ssy1, int &width, int &height, int &fw\)"))
    open(os.path.join(sub, "getimage_boxsum.inc"), "w").write(
        cut_block(os.path.join(RT, "rawimagesource.cc"), r"for \(int j = 0, jx = sx1; j < imwidth; j\+\+, jx \+= skip\) \{(?=\s*jx = std::min\(jx, maxx - skip\); // avoid trouble)"))
    open(os.path.join(sub, "shim_getimage.cc"), "w").write(SHIM_GETIMAGE_TU)
    vg = os.path.join(RT, "vng4_demosaic_RT.cc")
    open(os.path.join(sub, "vng4_rowrb.inc"), "w").write(cut_function(vg, r"^inline void vng4interpolate_row_redblue \(const RawImage \*ri[^)]*\)"))
    open(os.path.join(sub, "vng4_body.inc"), "w").write(cut_function(vg, r"^void RawImageSource::vng4_demosaic \(const array2D<float> &rawData[^)]*\)"))
    open(os.path.join(sub, "shim_vng4.cc"), "w").write(SHIM_VNG4_TU)
    ge = os.path.join(RT, "green_equil_RT.cc")
    open(os.path.join(sub, "greeneq_body.inc"), "w").write(
        cut_function(ge, r"^void RawImageSource::green_equilibrate_global\(array2D<float> &rawData\)") + "\n\n" +
        cut_function(ge, r"^void RawImageSource::green_equilibrate\(const GreenEqulibrateThreshold &thresh, array2D<float> &rawData\)"))
    open(os.path.join(sub, "shim_greeneq.cc"), "w").write(SHIM_GREENEQ_TU)
    bp = os.path.join(RT, "badpixels.cc")
    bptext = open(bp, encoding="utf-8", errors="replace").read()
    a0 = re.search(r"^namespace \{", bptext, flags=re.M)
    a1 = re.search(r"^\} // namespace", bptext, flags=re.M)
    open(os.path.join(sub, "badpix_body.inc"), "w").write(
        bptext[a0.start():a1.end()] + "\n\n" +
        cut_function(bp, r"^int RawImageSource::interpolateBadPixelsBayer\(const PixelsMap &bitmapBads, array2D<float> &rawData\)") + "\n\n" +
        cut_function(bp, r"^int RawImageSource::interpolateBadPixelsXtrans\(const PixelsMap &bitmapBads\)") + "\n\n" +
        cut_function(bp, r"^int RawImageSource::findHotDeadPixels\(PixelsMap &bpMap, const float thresh, const bool findHotPixels, const bool findDeadPixels\) const"))
    open(os.path.join(sub, "shim_badpix.cc"), "w").write(SHIM_BADPIX_TU)
    rz = os.path.join(RT, "ipresize.cc")
    open(os.path.join(sub, "resize_lanc.inc"), "w").write(cut_function(rz, r"^inline float Lanc\(float x, float a\)"))
    open(os.path.join(sub, "resize_lanczos.inc"), "w").write(cut_function(rz, r"^void ImProcFunctions::Lanczos\(Imagefloat \*src, Imagefloat \*dst, float scale\)"))
    open(os.path.join(sub, "shim_resize.cc"), "w").write(SHIM_RESIZE_TU)

    # X-Trans demosaic (xtrans_demosaic.cc): constants + cielab + border + Markesteijn
    xt = os.path.join(RT, "xtrans_demosaic.cc")
    xtext = open(xt, encoding="utf-8", errors="replace").read()
    m0 = re.search(r"^const float xyz_rgb\[3\]\[3\]", xtext, flags=re.M)
    m1 = re.search(r"^void RawImageSource::cielab", xtext, flags=re.M)
    body = [xtext[m0.start():m1.start()],
            cut_function(xt, r"^void RawImageSource::cielab \([^)]*\)"),
            "#define fcol(row,col) xtrans[(row)%6][(col)%6]\n#define isgreen(row,col) (xtrans[(row)%3][(col)%3]&1)\n",
            cut_function(xt, r"^void RawImageSource::xtransborder_interpolate \([^)]*\)"),
            "#define CLIP(x) (x)\n",
            cut_function(xt, r"^void RawImageSource::xtrans_interpolate\(const int passes, const bool useCieLab\)"),
            "#undef CLIP\n",
            cut_function(xt, r"^void RawImageSource::fast_xtrans_interpolate_blend \(const float\* const \* blend[^)]*\)"),
            "#undef fcol\n#undef isgreen\n"]
    if det:
        body[5] = insert_before(body[5], r"int mrow = MIN \(top \+ ts, height - 3\);",
                                "memset(buffer, 0, (ts * ts * (ndir * 4 + 3) + 128) * sizeof(float));\n                ")
    open(os.path.join(sub, "xtrans_body.inc"), "w").write("\n".join(body))
    open(os.path.join(sub, "shim_xtrans.cc"), "w").write(SHIM_XTRANS_TU)
    return sub


def build(det):
    sub = extract(det)
    import build_ref_tone
    build_ref_tone.extract(sub)
    lib = os.path.join(OUT, "libartref_det.so" if det else "libartref.so")
    tus = ["shim.cc", "shim_gauss.cc", "shim_guided.cc", "shim_wavelet.cc", "shim_shrink.cc", "shim_nlmeans.cc", "shim_denoise.cc", "shim_fattal.cc",
           "shim_chain.cc", "shim_usm.cc", "shim_xtrans.cc", "shim_resize.cc", "shim_greeneq.cc", "shim_pack.cc", "shim_bilinear.cc", "shim_vng4.cc",
           "shim_hlblend.cc", "shim_getimage.cc", "shim_tone.cc", "shim_badpix.cc"]
    base = ["g++", "-std=c++11", "-O3", "-fopenmp", "-ffp-contract=off", "-fPIC", "-w", "-I", sub, "-I", RT] + (["-DARTREF_DET"] if det else [])
    from concurrent.futures import ThreadPoolExecutor
    objs = [os.path.join(sub, t[:-3] + ".o") for t in tus]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:      # one object per translation unit, side by side
        list(ex.map(subprocess.check_call, [base + ["-c", os.path.join(sub, t), "-o", o] for t, o in zip(tus, objs)]))
    subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", lib] + objs)
    return lib


def main():
    if not os.path.isdir(RT):
        print("build_ref: %s absent -- keeping prebuilt oracle/_ref (if any)" % RT)
        return 0
    os.makedirs(OUT, exist_ok=True)
    for det in (False, True):
        print("built", build(det))
    return 0


if __name__ == "__main__":
    sys.exit(main())
