"""oracle -- CPU checkers for the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Nothing under art_b200/ does.

Two checkers:
  port : oracle/libartoracle.so, our plain-C restatement (oracle/*_port.c)
  ref  : oracle/_ref/libartref{,_det}.so, the reference's own function bodies
         compiled where they lie (oracle/build_ref.py); `det` adds the
         zero-scratch-per-tile patch that makes the reference schedule-independent.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_c_float_p = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return a.ctypes.data_as(_c_float_p)


def build(force=False):
    """Compile the port (always possible) and, when /root/reference is present, _ref."""
    lib = os.path.join(HERE, "libartoracle.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith("_port.c")]
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libartoracle.so"])
    ref_lib = os.path.join(HERE, "_ref", "libartref_det.so")
    recipes = [os.path.join(HERE, "build_ref.py"), os.path.join(HERE, "build_ref_tone.py")]
    if os.path.isdir("/root/reference/rtengine") and (force or not os.path.exists(ref_lib) or
                                                       any(os.path.getmtime(r) > os.path.getmtime(ref_lib) for r in recipes)):
        subprocess.check_call(["python3", os.path.join(HERE, "build_ref.py")])


class _Port:
    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(HERE, "libartoracle.so"))

    def _planes(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        return raw, H, W, [np.zeros((H, W), np.float32) for _ in range(3)]

    def rcd(self, raw, filters):
        raw, H, W, (r, g, b) = self._planes(raw)
        rc = self.lib.artoracle_rcd(W, H, ctypes.c_uint(filters), _fp(raw), ctypes.c_long(W),
                                    _fp(r), _fp(g), _fp(b), ctypes.c_long(W))
        assert rc == 0
        return r, g, b

    def border_interpolate2(self, raw, filters, bord):
        raw, H, W, (r, g, b) = self._planes(raw)
        self.lib.artoracle_border_interpolate2(W, H, ctypes.c_uint(filters), bord, _fp(raw), ctypes.c_long(W),
                                               _fp(r), _fp(g), _fp(b), ctypes.c_long(W))
        return r, g, b

    def scale_convert(self, planes, mul, do_clip, mat=None):
        return _scale_convert(self.lib, "artoracle_scale_convert", planes, mul, do_clip, mat)

    def scale_colors_bayer(self, raw, filters, black, mul):
        return _scale_colors(self.lib, "artoracle_scale_colors_bayer", raw, filters, black, mul)

    def gauss(self, src, sigma, inplace=False):
        return _gauss(self.lib, "artoracle_gauss", src, sigma, inplace, False)

    def gauss_iir(self, src, dst, divb, sigma, kind):
        return _gauss_iir(self.lib, False, src, dst, divb, sigma, kind)

    def boxblur(self, src, radius, inplace=False):
        return _boxblur(self.lib, "artoracle_boxblur", src, radius, inplace, False)

    def wavelet(self, src, maxlvl, subsamp=1):
        return Wavelet(self.lib, "artoracle", src, maxlvl, subsamp)

    def guided_filter(self, guide, src, r, eps, subsampling=0):
        return _guided(self.lib, "artoracle_guided_filter", guide, src, r, eps, subsampling)

    def amaze(self, raw, filters, initial_gain=1.0, border=4):
        raw, H, W, (r, g, b) = self._planes(raw)
        rc = self.lib.artoracle_amaze(W, H, ctypes.c_uint(filters), _fp(raw), ctypes.c_long(W),
                                      _fp(r), _fp(g), _fp(b), ctypes.c_long(W),
                                      ctypes.c_float(initial_gain), border)
        assert rc == 0
        return r, g, b


def _gauss_iir(lib, ref, src, dst, divb, sigma, kind):
    """GAUSS_MULT ("mult": dst *= blur(src), src blurred horizontally in place) / GAUSS_DIV ("div": dst = divb / blur(src)) of the
    recursive branch; returns (src_after, dst)."""
    s = np.array(src, dtype=np.float32, order="C", copy=True)
    d = np.array(dst, dtype=np.float32, order="C", copy=True)
    H, W = s.shape
    t = 1 if kind == "mult" else 2
    v = None if divb is None else np.ascontiguousarray(divb, dtype=np.float32)
    vp = _fp(v) if v is not None else None
    if ref:
        rc = lib.artref_gauss_ex(_fp(s), _fp(d), vp, W, H, ctypes.c_double(sigma), t)
    else:
        rc = lib.artoracle_gauss_iir(_fp(s), ctypes.c_long(W), _fp(d), ctypes.c_long(W), vp, ctypes.c_long(W), W, H, ctypes.c_double(sigma), t)
    assert rc == 0
    return s, d


def _gauss(lib, fname, src, sigma, inplace, extra_arg):
    src = np.ascontiguousarray(src, dtype=np.float32)
    H, W = src.shape
    if inplace:
        dst = src.copy()
        args = [_fp(dst), ctypes.c_long(W), _fp(dst), ctypes.c_long(W), W, H, ctypes.c_double(sigma)]
    else:
        dst = np.zeros_like(src)
        args = [_fp(src), ctypes.c_long(W), _fp(dst), ctypes.c_long(W), W, H, ctypes.c_double(sigma)]
    if extra_arg:
        args.append(1 if inplace else 0)
    rc = getattr(lib, fname)(*args)
    assert rc == 0
    return dst


def _boxblur(lib, fname, src, radius, inplace, extra_arg):
    src = np.ascontiguousarray(src, dtype=np.float32)
    H, W = src.shape
    dst = src.copy() if inplace else np.zeros_like(src)
    a = dst if inplace else src
    args = [_fp(a), ctypes.c_long(W), _fp(dst), ctypes.c_long(W), W, H, int(radius)]
    if extra_arg:
        args.append(1 if inplace else 0)
    rc = getattr(lib, fname)(*args)
    assert rc == 0
    return dst


def _guided(lib, fname, guide, src, r, eps, subsampling):
    guide = np.ascontiguousarray(guide, dtype=np.float32)
    src = np.ascontiguousarray(src, dtype=np.float32)
    H, W = src.shape
    dst = np.zeros_like(src)
    rc = getattr(lib, fname)(_fp(guide), _fp(src), _fp(dst), ctypes.c_long(W), W, H, int(r), ctypes.c_float(eps), int(subsampling))
    assert rc == 0
    return dst


def _scale_convert(lib, fname, planes, mul, do_clip, mat):
    outs = [np.array(p, dtype=np.float32, order="C", copy=True) for p in planes]
    H, W = outs[0].shape
    m = (ctypes.c_float * 3)(*[float(x) for x in mul])
    mm = None if mat is None else (ctypes.c_double * 9)(*[float(x) for x in np.asarray(mat, np.float64).reshape(9)])
    getattr(lib, fname)(W, H, _fp(outs[0]), _fp(outs[1]), _fp(outs[2]), ctypes.c_long(W), m, int(bool(do_clip)), mm)
    return tuple(outs)


def _scale_colors(lib, fname, raw, filters, black, mul):
    out = np.array(raw, dtype=np.float32, order="C", copy=True)
    H, W = out.shape
    bl = (ctypes.c_float * 4)(*[float(x) for x in black])
    mu = (ctypes.c_float * 4)(*[float(x) for x in mul])
    ch = (ctypes.c_float * 3)()
    getattr(lib, fname)(W, H, ctypes.c_uint(filters), _fp(out), ctypes.c_long(W), bl, mu, ch)
    return out, [float(ch[0]), float(ch[1]), float(ch[2])]


class _Ref:
    def wavelet(self, src, maxlvl, subsamp=1):
        return Wavelet(self.lib, "artref", src, maxlvl, subsamp)

    def boxblur(self, src, radius, inplace=False):
        return _boxblur(self.lib, "artref_boxblur", src, radius, inplace, True)

    def guided_filter(self, guide, src, r, eps, subsampling=0):
        return _guided(self.lib, "artref_guided_filter", guide, src, r, eps, subsampling)

    def gauss(self, src, sigma, inplace=False):
        return _gauss(self.lib, "artref_gauss", src, sigma, inplace, True)

    def gauss_iir(self, src, dst, divb, sigma, kind):
        return _gauss_iir(self.lib, True, src, dst, divb, sigma, kind)

    def scale_colors_bayer(self, raw, filters, black, mul):
        return _scale_colors(self.lib, "artref_scale_colors_bayer", raw, filters, black, mul)

    def scale_convert(self, planes, mul, do_clip, mat=None):
        return _scale_convert(self.lib, "artref_scale_convert", planes, mul, do_clip, mat)

    def __init__(self, det=True):
        path = os.path.join(HERE, "_ref", "libartref_det.so" if det else "libartref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)
        self.det = det

    def _planes(self, raw, out=None):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        if out is None:
            out = [np.zeros((H, W), np.float32) for _ in range(3)]
        return raw, H, W, out

    def max_threads(self):
        return int(self.lib.artref_max_threads())

    def rcd(self, raw, filters, nthreads=0, out=None):
        raw, H, W, (r, g, b) = self._planes(raw, out)
        self.lib.artref_rcd(W, H, ctypes.c_uint(filters), _fp(raw), ctypes.c_long(W),
                            _fp(r), _fp(g), _fp(b), ctypes.c_long(W), nthreads)
        return r, g, b

    def amaze(self, raw, filters, initial_gain=1.0, border=4, nthreads=0, out=None):
        raw, H, W, (r, g, b) = self._planes(raw, out)
        self.lib.artref_amaze(W, H, ctypes.c_uint(filters), _fp(raw), ctypes.c_long(W),
                              _fp(r), _fp(g), _fp(b), ctypes.c_long(W),
                              ctypes.c_double(initial_gain), border, nthreads)
        return r, g, b

    def border_interpolate2(self, raw, filters, bord):
        raw, H, W, (r, g, b) = self._planes(raw)
        self.lib.artref_border_interpolate2(W, H, ctypes.c_uint(filters), bord, _fp(raw), ctypes.c_long(W),
                                            _fp(r), _fp(g), _fp(b), ctypes.c_long(W))
        return r, g, b


class Wavelet:
    """wavelet_decomposition through either checker (prefix 'artref' = reference headers, 'artoracle' = port)."""

    def __init__(self, lib, prefix, src, maxlvl, subsamp=1):
        self.lib, self.p = lib, prefix
        self.src = np.ascontiguousarray(src, dtype=np.float32)      # keep alive
        H, W = self.src.shape
        f = getattr(lib, prefix + "_wavelet_new")
        f.restype = ctypes.c_void_p
        args = [_fp(self.src), W, H, int(maxlvl), int(subsamp)]
        if prefix == "artref":
            args.append(0)
        self.h = ctypes.c_void_p(f(*args))
        self.shape = (H, W)
        for name in ("maxlevel", "level_W", "level_H", "level_stride"):
            getattr(lib, "%s_wavelet_%s" % (prefix, name)).argtypes = [ctypes.c_void_p] + ([ctypes.c_int] if name != "maxlevel" else [])
        getattr(lib, prefix + "_wavelet_band").restype = ctypes.POINTER(ctypes.c_float)
        getattr(lib, prefix + "_wavelet_band").argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        getattr(lib, prefix + "_wavelet_reconstruct").argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_float]
        getattr(lib, prefix + "_wavelet_delete").argtypes = [ctypes.c_void_p]

    def _c(self, name, *a):
        return getattr(self.lib, "%s_wavelet_%s" % (self.p, name))(self.h, *a)

    def maxlevel(self):
        return int(self._c("maxlevel"))

    def dims(self, lvl):
        return int(self._c("level_H", lvl)), int(self._c("level_W", lvl)), int(self._c("level_stride", lvl))

    def band(self, lvl, d):
        """numpy VIEW of subband d (1..3) of level lvl; d == 0: the final lowpass (dims of the last level)."""
        h, w, _ = self.dims(lvl if d else self.maxlevel() - 1)
        ptr = self._c("band", lvl, d)
        return np.ctypeslib.as_array(ptr, shape=(h, w))

    def reconstruct(self, dst=None, blend=1.0):
        H, W = self.shape
        if dst is None:
            dst = np.zeros((H, W), np.float32)
        self._c("reconstruct", _fp(dst), ctypes.c_float(blend))
        return dst

    def close(self):
        if self.h:
            self._c("delete")
            self.h = None


_cache = {}


def port():
    if "port" not in _cache:
        _cache["port"] = _Port()
    return _cache["port"]


def ref(det=True):
    key = "ref_det" if det else "ref"
    if key not in _cache:
        _cache[key] = _Ref(det)
    return _cache[key]


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libartref_det.so")) or os.path.isdir("/root/reference/rtengine")
