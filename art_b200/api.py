"""ctypes binding of include/art_hotpath.h -- one Python method per C entry point."""
import ctypes
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "art_hotpath.h")

BAYER_AMAZE = 0
BAYER_RCD = 1
XTRANS_3PASS = 2     # art_hp_develop only
XTRANS_1PASS = 3

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_float_pp = ctypes.POINTER(_c_float_p)


class HotPathError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("art_hotpath error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    return os.path.join(HERE, "libart_hotpath.so")


def _declared_symbols():
    """Every function name include/art_hotpath.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(art_hp_[a-z0-9_]+)\s*\(", text)))


ABI_SYMBOLS = _declared_symbols()
_lib = None


def load_library(build_if_missing=True):
    """Load libart_hotpath.so; raises if it is absent and cannot be built.  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise FileNotFoundError(path)
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(path)
    vp, i, u, sz, d, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_double, ctypes.c_float
    sig = {
        "art_hp_abi_version": (i, []),
        "art_hp_device_count": (i, []),
        "art_hp_create": (i, [ctypes.POINTER(vp), i]),
        "art_hp_destroy": (None, [vp]),
        "art_hp_last_error": (ctypes.c_char_p, [vp]),
        "art_hp_set_stream": (i, [vp, vp]),
        "art_hp_get_stream": (vp, [vp]),
        "art_hp_sync": (i, [vp]),
        "art_hp_launch_count": (ctypes.c_ulonglong, [vp]),
        "art_hp_profile_enable": (i, [vp, i]),
        "art_hp_profile_collect": (i, [vp]),
        "art_hp_profile_entry": (i, [vp, i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(d), ctypes.POINTER(i)]),
        "art_hp_host_alloc": (vp, [sz]),
        "art_hp_host_free": (None, [vp]),
        "art_hp_demosaic_bayer": (i, [vp, i, i, i, u, vp, vp, vp, vp, d, i]),
        "art_hp_demosaic_bayer_dev": (i, [vp, i, i, i, u, vp, sz, vp, vp, vp, sz, d, i]),
        "art_hp_border_interpolate2_dev": (i, [vp, i, i, u, i, vp, sz, vp, vp, vp, sz]),
        "art_hp_wavelet_decompose_dev": (i, [vp, vp, sz, i, i, i, i, ctypes.POINTER(vp)]),
        "art_hp_wavelet_maxlevel": (i, [vp]),
        "art_hp_wavelet_level_dims": (i, [vp, i, ctypes.POINTER(i), ctypes.POINTER(i), ctypes.POINTER(i)]),
        "art_hp_wavelet_band_dev": (vp, [vp, i, i]),
        "art_hp_wavelet_reconstruct_dev": (i, [vp, vp, sz, ctypes.c_float]),
        "art_hp_wavelet_get_band": (i, [vp, i, i, vp]),
        "art_hp_wavelet_set_band": (i, [vp, i, i, vp]),
        "art_hp_wavelet_destroy": (None, [vp]),
        "art_hp_wavelet_mad_dev": (i, [vp, vp, vp]),
        "art_hp_wavelet_denoise_L_dev": (i, [vp, vp, vp, vp, d]),
        "art_hp_wavelet_denoise_AB_dev": (i, [vp, vp, vp, vp, vp, ctypes.c_float, i, i, d]),
        "art_hp_boxblur": (i, [vp, vp, vp, i, i, i]),
        "art_hp_boxblur_dev": (i, [vp, vp, sz, vp, sz, i, i, i]),
        "art_hp_guided_filter": (i, [vp, i, i, vp, vp, vp, i, ctypes.c_float, i]),
        "art_hp_guided_filter_dev": (i, [vp, i, i, vp, sz, vp, sz, vp, sz, i, ctypes.c_float, i]),
        "art_hp_rgb_denoise": (i, [vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp]),
        "art_hp_rgb_denoise_dev": (i, [vp, vp, vp, vp, sz, i, i, vp, vp, vp, vp, vp, sz, vp]),
        "art_hp_detail_mask": (i, [vp, vp, vp, i, i, f, f, f, f, i, f]),
        "art_hp_detail_mask_dev": (i, [vp, vp, sz, vp, sz, i, i, f, f, f, f, i, f]),
        "art_hp_nlmeans": (i, [vp, vp, i, i, f, i, i, f]),
        "art_hp_nlmeans_dev": (i, [vp, vp, sz, i, i, f, i, i, f]),
        "art_hp_gauss": (i, [vp, vp, vp, i, i, d, i]),
        "art_hp_gauss_dev": (i, [vp, vp, sz, vp, sz, i, i, d, i]),
        "art_hp_scale_colors_bayer": (i, [vp, i, i, u, vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "art_hp_scale_colors_bayer_dev": (i, [vp, i, i, u, vp, sz, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "art_hp_scale_colors_xtrans": (i, [vp, i, i, vp, vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "art_hp_scale_colors_xtrans_dev": (i, [vp, i, i, vp, vp, sz, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "art_hp_scale_convert": (i, [vp, i, i, vp, vp, vp, ctypes.POINTER(ctypes.c_float), i, ctypes.POINTER(d)]),
        "art_hp_scale_convert_dev": (i, [vp, i, i, vp, vp, vp, sz, ctypes.POINTER(ctypes.c_float), i, ctypes.POINTER(d)]),
        "art_hp_develop": (i, [vp, vp, i, i, vp, vp, vp, vp]),
        "art_hp_develop_dev": (i, [vp, vp, i, i, vp, sz, vp, vp, vp, sz]),
        "art_hp_develop_submit": (i, [vp, vp, i, i, vp, vp, vp, vp]),
        "art_hp_denoise_compute_params": (i, [vp, i, i, vp, vp, vp, vp, i, vp, vp, d, i, vp, vp]),
        "art_hp_denoise_compute_params_dev": (i, [vp, i, i, vp, vp, vp, sz, vp, i, vp, vp, d, i, vp, vp]),
        "art_hp_develop_size": (i, [vp, i, i, ctypes.POINTER(i), ctypes.POINTER(i), ctypes.POINTER(i)]),
        "art_hp_green_equilibrate_global": (i, [vp, i, i, u, vp, i]),
        "art_hp_green_equilibrate_global_dev": (i, [vp, i, i, u, vp, sz, i]),
        "art_hp_green_equilibrate": (i, [vp, i, i, u, vp, f, vp]),
        "art_hp_green_equilibrate_dev": (i, [vp, i, i, u, vp, sz, f, vp, sz]),
        "art_hp_band_plan_rows": (i, [vp, i, i, i, i, i, vp]),
        "art_hp_set_allreduce": (i, [vp, vp, vp]),
        "art_hp_comm_unique_id": (i, [vp]),
        "art_hp_comm_init": (i, [vp, vp, i, i]),
        "art_hp_comm_destroy": (i, [vp]),
        "art_hp_develop_band_dev": (i, [vp, vp, i, i, vp, sz, vp, vp, vp, sz, vp]),
        "art_hp_denoise_guided_smoothing": (i, [vp, i, i, vp, vp, vp, vp, i, d]),
        "art_hp_denoise_guided_smoothing_dev": (i, [vp, i, i, vp, vp, vp, sz, vp, i, d]),
        "art_hp_develop_submit_packed": (i, [vp, vp, i, i, vp, i, i, vp, sz]),
        "art_hp_scanlines": (i, [vp, i, i, vp, vp, vp, i, i, vp, sz]),
        "art_hp_demosaic_vng4": (i, [vp, i, i, u, vp, vp, vp, vp]),
        "art_hp_demosaic_vng4_dev": (i, [vp, i, i, u, vp, sz, vp, vp, vp, sz]),
        "art_hp_dual_demosaic_bayer": (i, [vp, i, i, i, i, u, u, vp, vp, vp, vp, d, i, vp, i]),
        "art_hp_dual_demosaic_bayer_dev": (i, [vp, i, i, i, i, u, u, vp, sz, vp, vp, vp, sz, d, i, d, i, vp]),
        "art_hp_dual_demosaic_xtrans": (i, [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, i]),
        "art_hp_dual_demosaic_xtrans_dev": (i, [vp, i, i, i, i, vp, vp, vp, sz, vp, vp, vp, sz, d, i, vp]),
        "art_hp_lab_histogram": (i, [vp, i, i, vp, vp, vp, vp, vp]),
        "art_hp_lab_histogram_dev": (i, [vp, i, i, vp, vp, vp, sz, vp, vp]),
        "art_hp_prophoto_blue": (i, [vp, i, i, vp, vp, vp]),
        "art_hp_prophoto_blue_dev": (i, [vp, i, i, vp, vp, vp, sz]),
        "art_hp_black_and_white": (i, [vp, i, i, vp, vp, vp, vp]),
        "art_hp_black_and_white_dev": (i, [vp, i, i, vp, vp, vp, sz, vp]),
        "art_hp_tone_equalizer": (i, [vp, i, i, vp, vp, vp, vp]),
        "art_hp_tone_equalizer_dev": (i, [vp, i, i, vp, vp, vp, sz, vp]),
        "art_hp_hsl_equalizer": (i, [vp, i, i, vp, vp, vp, vp]),
        "art_hp_hsl_equalizer_dev": (i, [vp, i, i, vp, vp, vp, sz, vp]),
        "art_hp_channel_mixer": (i, [vp, i, i, vp, vp, vp, vp]),
        "art_hp_channel_mixer_dev": (i, [vp, i, i, vp, vp, vp, sz, vp]),
        "art_hp_find_hot_dead_pixels": (i, [vp, i, i, vp, vp, f, i, i, vp, sz, ctypes.POINTER(i)]),
        "art_hp_find_hot_dead_pixels_dev": (i, [vp, i, i, vp, vp, sz, f, i, i, vp, sz, ctypes.POINTER(i)]),
        "art_hp_interpolate_bad_pixels_bayer": (i, [vp, i, i, u, vp, vp, sz, ctypes.POINTER(i)]),
        "art_hp_interpolate_bad_pixels_bayer_dev": (i, [vp, i, i, u, vp, sz, vp, sz, ctypes.POINTER(i)]),
        "art_hp_interpolate_bad_pixels_xtrans": (i, [vp, i, i, vp, vp, vp, sz, ctypes.POINTER(i)]),
        "art_hp_interpolate_bad_pixels_xtrans_dev": (i, [vp, i, i, vp, vp, sz, vp, sz, ctypes.POINTER(i)]),
        "art_hp_resize_lanczos": (i, [vp, i, i, vp, vp, vp, i, i, vp, vp, vp, f]),
        "art_hp_resize_lanczos_dev": (i, [vp, i, i, vp, vp, vp, sz, i, i, vp, vp, vp, sz, f]),
        "art_hp_scanlines_dev": (i, [vp, i, i, vp, vp, vp, sz, i, i, vp, sz]),
        "art_hp_develop_wait": (i, [vp]),
        "art_hp_develop_pending": (i, [vp]),
        "art_hp_fattal": (i, [vp, i, i, vp, vp, vp, i, i, i, ctypes.POINTER(d)]),
        "art_hp_fattal_dev": (i, [vp, i, i, vp, vp, vp, sz, i, i, i, ctypes.POINTER(d)]),
        "art_hp_fattal_fast_dim": (i, [i]),
        "art_hp_demosaic_xtrans": (i, [vp, i, i, i, i, ctypes.POINTER(i), ctypes.POINTER(f), vp, vp, vp, vp]),
        "art_hp_demosaic_xtrans_dev": (i, [vp, i, i, i, i, ctypes.POINTER(i), ctypes.POINTER(f), vp, sz, vp, vp, vp, sz]),
        "art_hp_sharpen_usm": (i, [vp, i, i, vp, vp, vp, vp, ctypes.POINTER(d)]),
        "art_hp_sharpen_usm_dev": (i, [vp, i, i, vp, vp, vp, sz, vp, ctypes.POINTER(d)]),
        "art_hp_color_chain": (i, [vp, i, i, vp, vp, vp, vp]),
        "art_hp_color_chain_dev": (i, [vp, i, i, vp, vp, vp, sz, vp]),
        "art_hp_median_denoise": (i, [vp, vp, vp, i, i, i, i, f]),
        "art_hp_median_denoise_dev": (i, [vp, vp, sz, vp, sz, i, i, i, i, f]),
        "art_hp_redft00_2d": (i, [vp, i, i, vp, vp]),
        "art_hp_band_align": (i, [i, ctypes.POINTER(i), ctypes.POINTER(i)]),
        "art_hp_band_halo": (i, [i]),
        "art_hp_demosaic_bayer_rows_dev": (i, [vp, i, i, i, u, vp, sz, vp, vp, vp, sz, d, i, i, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_row_tables = {}


def row_table(a):
    """H row pointers into a 2-D float32 array (what rtengine::array2D<float> holds)."""
    assert a.dtype == np.float32 and a.ndim == 2 and a.strides[1] == 4
    H = a.shape[0]
    base = a.ctypes.data
    key = (base, H, a.strides[0])
    tbl = _row_tables.get(key)
    if tbl is None:      # a table only depends on (base, H, stride): frames that reuse their buffers reuse their tables
        if len(_row_tables) >= 64:
            _row_tables.clear()
        tbl = (ctypes.c_void_p * H)(*[base + r * a.strides[0] for r in range(H)])
        _row_tables[key] = tbl
    return tbl


class PinnedArray:
    """An (H, W) numpy view (float32 unless `dtype` says otherwise) over art_hp_host_alloc'ed (pinned) memory."""

    def __init__(self, lib, H, W, dtype=np.float32):
        self._lib = lib
        n = H * W * np.dtype(dtype).itemsize
        self.ptr = lib.art_hp_host_alloc(n)
        if not self.ptr:
            raise MemoryError("art_hp_host_alloc(%d)" % n)
        buf = (ctypes.c_ubyte * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(H, W)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.art_hp_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _DenoiseParamsC(ctypes.Structure):
    _fields_ = [("luminance", ctypes.c_double), ("luminanceDetail", ctypes.c_double), ("luminanceDetailThreshold", ctypes.c_int),
                ("chrominance", ctypes.c_double), ("chrominanceRedGreen", ctypes.c_double), ("chrominanceBlueYellow", ctypes.c_double),
                ("gamma", ctypes.c_double), ("scale", ctypes.c_double), ("colorSpace", ctypes.c_int), ("aggressive", ctypes.c_int),
                ("chrominanceMethod", ctypes.c_int), ("noiseCCurve", ctypes.c_void_p), ("noiseCCurveSum", ctypes.c_float),
                ("wprof_inverse", ctypes.POINTER(ctypes.c_double)), ("chrominanceAutoFactor", ctypes.c_double)]


class BandPlan(ctypes.Structure):
    """art_hp_band_plan: the rows of one rank when a frame is split across GPUs (all in rows; see include/art_hotpath.h)."""
    _fields_ = [("own_begin", ctypes.c_int), ("own_end", ctypes.c_int), ("band_begin", ctypes.c_int), ("band_end", ctypes.c_int),
                ("dm_begin", ctypes.c_int), ("dm_end", ctypes.c_int), ("raw_begin", ctypes.c_int), ("raw_end", ctypes.c_int)]

    def __repr__(self):
        return "BandPlan(own=[%d,%d) band=[%d,%d) dm=[%d,%d) raw=[%d,%d))" % (self.own_begin, self.own_end, self.band_begin, self.band_end,
                                                                             self.dm_begin, self.dm_end, self.raw_begin, self.raw_end)


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


def band_plan(params, W, H, own_begin, own_end, halo=200):
    """Rows of a rank that delivers rows [own_begin, own_end) of the developed frame of a W x H raw frame (no GPU needed)."""
    c = params.c_struct()
    plan = BandPlan()
    rc = load_library().art_hp_band_plan_rows(ctypes.byref(c), W, H, int(own_begin), int(own_end), int(halo), ctypes.byref(plan))
    if rc:
        raise HotPathError(rc, "art_hp_band_plan_rows(own=[%d,%d), halo=%d) on a %dx%d frame" % (own_begin, own_end, halo, W, H))
    return plan


class _DevelopParamsC(ctypes.Structure):
    _fields_ = [("method", ctypes.c_int), ("filters", ctypes.c_uint), ("initialGain", ctypes.c_double), ("border", ctypes.c_int),
                ("mul", ctypes.c_float * 3), ("doClip", ctypes.c_int), ("cam2work", ctypes.POINTER(ctypes.c_double)),
                ("denoise", ctypes.POINTER(_DenoiseParamsC)), ("nlStrength", ctypes.c_int), ("nlDetail", ctypes.c_int),
                ("fattal_enabled", ctypes.c_int), ("fattal_threshold", ctypes.c_int), ("fattal_amount", ctypes.c_int),
                ("fattal_satcontrol", ctypes.c_int), ("wprof", ctypes.POINTER(ctypes.c_double)),
                ("sharpen", ctypes.c_void_p), ("chain", ctypes.c_void_p),
                ("xtrans", ctypes.POINTER(ctypes.c_int)), ("rgb_cam", ctypes.POINTER(ctypes.c_float)),
                ("full_frame", ctypes.c_int), ("guidedChromaRadius", ctypes.c_int), ("denoise_expcomp", ctypes.c_double),
                ("tran", ctypes.c_int), ("hr_blend", ctypes.c_int), ("hlmax", ctypes.c_float * 3),
                ("pp_x", ctypes.c_int), ("pp_y", ctypes.c_int), ("pp_width", ctypes.c_int), ("pp_height", ctypes.c_int), ("pp_skip", ctypes.c_int)]


class _SharpenParamsC(ctypes.Structure):
    _fields_ = [("contrast", ctypes.c_double), ("radius", ctypes.c_double), ("amount", ctypes.c_int), ("threshold", ctypes.c_int * 4),
                ("edgesonly", ctypes.c_int), ("halocontrol", ctypes.c_int), ("halocontrol_amount", ctypes.c_int), ("scale", ctypes.c_double),
                ("method", ctypes.c_int), ("deconvradius", ctypes.c_double), ("deconvamount", ctypes.c_int), ("deconvCornerBoost", ctypes.c_double),
                ("deconvCornerLatitude", ctypes.c_int), ("offset_x", ctypes.c_int), ("offset_y", ctypes.c_int), ("full_width", ctypes.c_int),
                ("full_height", ctypes.c_int), ("edges_radius", ctypes.c_double), ("edges_tolerance", ctypes.c_int)]


class SharpenParams:
    """Mirror of procparams::SharpeningParams for methods "usm" and "rld" (defaults: rtengine/procparams.cc L1756-1776)."""

    def __init__(self, contrast=20.0, radius=0.5, amount=200, threshold=(20, 80, 2000, 1200), edgesonly=False, halocontrol=False,
                 halocontrol_amount=85, scale=1.0, method="usm", deconvradius=0.75, deconvamount=100, deconvCornerBoost=0.0,
                 deconvCornerLatitude=25, offset_x=0, offset_y=0, full_width=0, full_height=0, edges_radius=1.9, edges_tolerance=1800):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        return _SharpenParamsC(float(self.contrast), float(self.radius), int(self.amount), (ctypes.c_int * 4)(*[int(t) for t in self.threshold]),
                               int(bool(self.edgesonly)), int(bool(self.halocontrol)), int(self.halocontrol_amount), float(self.scale),
                               {"usm": 0, "rld": 1, "psf": 2}[self.method], float(self.deconvradius), int(self.deconvamount), float(self.deconvCornerBoost),
                               int(self.deconvCornerLatitude), int(self.offset_x), int(self.offset_y), int(self.full_width), int(self.full_height),
                               float(self.edges_radius), int(self.edges_tolerance))


class _CurveStageC(ctypes.Structure):      # art_hp_curve_stage
    _dp = ctypes.POINTER(ctypes.c_double)
    _fields_ = [("kind", ctypes.c_int), ("poly_x", _dp), ("poly_y", _dp), ("n", ctypes.c_int),
                ("a", ctypes.c_double), ("b", ctypes.c_double), ("w", ctypes.c_double)]


class _FlatCurveC(ctypes.Structure):       # art_hp_flat_curve
    _dp = ctypes.POINTER(ctypes.c_double)
    _fields_ = [("n", ctypes.c_int), ("poly_x", _dp), ("poly_y", _dp), ("dy_by_dx", _dp)]


class _HslParamsC(ctypes.Structure):       # art_hp_hsl_params
    _fields_ = [("hcurve", _FlatCurveC), ("scurve", _FlatCurveC), ("lcurve", _FlatCurveC), ("coeff", _FlatCurveC),
                ("smoothing", ctypes.c_int), ("scale", ctypes.c_double), ("ws", ctypes.POINTER(ctypes.c_double))]


class HslParams:
    """art_hp_hsl_params: hcurve / scurve / lcurve / coeff = (poly_x, poly_y, dy_by_dx) of the host-built FlatCurve or None for an
    identity curve; smoothing = params->hsl.smoothing; scale = ImProcFunctions::scale; ws = the working-space matrix."""

    def __init__(self, hcurve=None, scurve=None, lcurve=None, coeff=None, smoothing=0, scale=1.0, ws=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        c = _HslParamsC()
        self._keep = []
        dp = ctypes.POINTER(ctypes.c_double)
        for name in ("hcurve", "scurve", "lcurve", "coeff"):
            cv = getattr(self, name)
            if cv is None or len(cv[0]) == 0:
                continue
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in cv]
            self._keep += arrs
            f = getattr(c, name)
            f.n, f.poly_x, f.poly_y, f.dy_by_dx = int(arrs[0].size), arrs[0].ctypes.data_as(dp), arrs[1].ctypes.data_as(dp), arrs[2].ctypes.data_as(dp)
        c.smoothing, c.scale = int(self.smoothing), float(self.scale)
        if self.ws is not None:
            w = np.ascontiguousarray(self.ws, dtype=np.float64).reshape(9)
            self._keep.append(w)
            c.ws = w.ctypes.data_as(dp)
        return c


class _BwParamsC(ctypes.Structure):        # art_hp_bw_params
    _fp = ctypes.POINTER(ctypes.c_float)
    _fields_ = [("bwr", ctypes.c_float), ("bwg", ctypes.c_float), ("bwb", ctypes.c_float), ("kcorec", ctypes.c_float),
                ("gamma_r", _fp), ("gamma_g", _fp), ("gamma_b", _fp), ("ulut", _fp), ("vlut", _fp), ("ws", ctypes.POINTER(ctypes.c_double))]


class BwParams:
    """art_hp_bw_params: mixer = (bwr, bwg, bwb), kcorec, gamma = three 65536-entry tables or None, cast = (ulut, vlut) or None, ws."""

    def __init__(self, mixer, kcorec=1.0, gamma=None, cast=None, ws=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        c = _BwParamsC()
        fp = ctypes.POINTER(ctypes.c_float)
        self._keep = []

        def tab(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == 65536
            self._keep.append(a)
            return a.ctypes.data_as(fp)
        c.bwr, c.bwg, c.bwb, c.kcorec = float(self.mixer[0]), float(self.mixer[1]), float(self.mixer[2]), float(self.kcorec)
        if self.gamma is not None:
            c.gamma_r, c.gamma_g, c.gamma_b = [tab(t) for t in self.gamma]
        if self.cast is not None:
            c.ulut, c.vlut = [tab(t) for t in self.cast]
        if self.ws is not None:
            w = np.ascontiguousarray(self.ws, dtype=np.float64).reshape(9)
            self._keep.append(w)
            c.ws = w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return c


class _ToneEqParamsC(ctypes.Structure):    # art_hp_toneeq_params
    _fields_ = [("bands", ctypes.c_int * 5), ("regularization", ctypes.c_int), ("pivot", ctypes.c_double), ("scale", ctypes.c_double),
                ("ws", ctypes.POINTER(ctypes.c_double))]


class ToneEqParams:
    """art_hp_toneeq_params: bands = the five sliders (blacks .. whites), regularization, pivot, scale, ws = the working-space matrix."""

    def __init__(self, bands=(0, 0, 0, 0, 0), regularization=0, pivot=0.0, scale=1.0, ws=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        c = _ToneEqParamsC()
        for k in range(5):
            c.bands[k] = int(self.bands[k])
        c.regularization, c.pivot, c.scale = int(self.regularization), float(self.pivot), float(self.scale)
        if self.ws is not None:
            self._w = np.ascontiguousarray(self.ws, dtype=np.float64).reshape(9)
            c.ws = self._w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return c


class _ChainParamsC(ctypes.Structure):
    _fp, _dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    _fields_ = [("exposure_enabled", ctypes.c_int), ("exp_scale", ctypes.c_float), ("black", ctypes.c_float),
                ("saturation_enabled", ctypes.c_int), ("saturation", ctypes.c_int), ("vibrance", ctypes.c_int),
                ("tonecurve_mode", ctypes.c_int), ("tonecurve_lut", _fp),
                ("rcurve", _fp), ("gcurve", _fp), ("bcurve", _fp),
                ("lab_enabled", ctypes.c_int), ("lab_lcurve", _fp), ("lab_acurve", _fp), ("lab_bcurve", _fp), ("lab_chroma", ctypes.c_float),
                ("ws", _dp), ("iws", _dp),
                ("tonecurve_whitept", ctypes.c_float), ("tonecurve_stages", ctypes.POINTER(_CurveStageC)), ("tonecurve_nstages", ctypes.c_int),
                ("neutral_to_out", _fp), ("neutral_to_work", _fp), ("satcurve_lut", _fp), ("softlight_lut", _fp)]


class ChainParams:
    """Parameters of art_hp_color_chain: the per-pixel stages of ImProcFunctions::process (improcfun.cc L567-641).
    exposure = (expcomp EV, black) as in procparams::ExposureParams; saturation = (saturation, vibrance) integers;
    tonecurve = (mode, 65536-entry LUT) with mode 0 STD, 1 FILMLIKE, 2 NEUTRAL (ToneCurveParams::TcMode, the reference default);
    whitept = ToneCurve::whitecoeff; stages = the Curve::getVal chain above the LUT as (kind, poly_x, poly_y, a, b, w) tuples
    (art_hp_curve_stage); to_out / to_work = NeutralToneCurve::ApplyState's 3x3 float matrices or None; satcurve = apply_satcurve's
    65536-entry table; rgbcurves = three LUTs or None each; lab = (L LUT of 32770, a LUT, b LUT, chroma); softlight = softLight's
    65536-entry table or None."""

    def __init__(self, exposure=None, saturation=None, tonecurve=None, rgbcurves=None, lab=None, ws=None, iws=None,
                 whitept=1.0, stages=None, to_out=None, to_work=None, satcurve=None, softlight=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        c = _ChainParamsC()
        self._keep = []
        fp = ctypes.POINTER(ctypes.c_float)

        def lut(a, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == n, "LUT of %d entries expected, got %d" % (n, a.size)
            self._keep.append(a)
            return a.ctypes.data_as(fp)

        def mat(m):
            if m is None:
                return None
            v = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(m, dtype=np.float64).reshape(9)])
            self._keep.append(v)
            return ctypes.cast(v, ctypes.POINTER(ctypes.c_double))

        if self.exposure is not None:
            ev, black = self.exposure
            # ipexposure.cc L33-34: exp_scale = pow(2.f, expcomp), black = params black * 2000.f
            c.exposure_enabled, c.exp_scale, c.black = 1, float(np.float32(2.0) ** np.float32(ev)), float(np.float32(black) * np.float32(2000.0))
        if self.saturation is not None:
            c.saturation_enabled, c.saturation, c.vibrance = 1, int(self.saturation[0]), int(self.saturation[1])
        if self.tonecurve is not None:
            c.tonecurve_mode, c.tonecurve_lut = int(self.tonecurve[0]), lut(self.tonecurve[1], 65536)
        if self.rgbcurves is not None:
            c.rcurve, c.gcurve, c.bcurve = [lut(x, 65536) for x in self.rgbcurves]
        if self.lab is not None:
            c.lab_enabled = 1
            c.lab_lcurve, c.lab_acurve, c.lab_bcurve = lut(self.lab[0], 32770), lut(self.lab[1], 65536), lut(self.lab[2], 65536)
            c.lab_chroma = float(self.lab[3])
        c.ws, c.iws = mat(self.ws), mat(self.iws)
        c.tonecurve_whitept = float(self.whitept)
        if self.stages:
            arr = (_CurveStageC * len(self.stages))()
            dp = ctypes.POINTER(ctypes.c_double)
            for i, (kind, px, py, ca, cb, cw) in enumerate(self.stages):
                arr[i].kind = int(kind)
                if int(kind) == 1:
                    px, py = np.ascontiguousarray(px, np.float64), np.ascontiguousarray(py, np.float64)
                    self._keep += [px, py]
                    arr[i].poly_x, arr[i].poly_y, arr[i].n = px.ctypes.data_as(dp), py.ctypes.data_as(dp), int(px.size)
                arr[i].a, arr[i].b, arr[i].w = float(ca), float(cb), float(cw)
            self._keep.append(arr)
            c.tonecurve_stages, c.tonecurve_nstages = arr, len(self.stages)
        for name, m in (("neutral_to_out", self.to_out), ("neutral_to_work", self.to_work)):
            if m is not None:
                v = np.ascontiguousarray(m, np.float32).reshape(9)
                self._keep.append(v)
                setattr(c, name, v.ctypes.data_as(fp))
        c.satcurve_lut = lut(self.satcurve, 65536)
        c.softlight_lut = lut(self.softlight, 65536)
        return c


class DevelopParams:
    """Parameters of art_hp_develop: the simpleprocess.cc stages on the hot path (demosaic, gains + matrix, denoise, Fattal, sharpening,
    the colour chain).  Like the reference the frame is cropped by the raw border after the demosaic (4 px for Bayer, 7 for X-Trans)
    unless full_frame; guided_chroma_radius / nl_strength are DenoiseParams' smoothing fields (0 unless smoothingEnabled);
    denoise_expcomp = the exposure compensation ImProcFunctions::denoise brackets its stage with when positive.  pp = (x, y, width, height,
    skip): a PreviewProps window of the transformed, cropped full image (the preview path; mul must carry the 1 / skip^2 of getImage)."""

    def __init__(self, method=0, filters=0x94949494, initial_gain=1.0, border=4, mul=(1.0, 1.0, 1.0), do_clip=True, cam2work=None,
                 denoise=None, nl_strength=0, nl_detail=80, fattal=None, wprof=None, sharpen=None, chain=None, xtrans=None, rgb_cam=None,
                 full_frame=False, guided_chroma_radius=0, denoise_expcomp=0.0, tran=0, hr_blend=False, hlmax=(65535.0, 65535.0, 65535.0), pp=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def out_shape(self, H, W):
        """(rows, columns) of the developed planes for an (H, W) raw frame"""
        if self.pp is not None:
            x, y, w, h, skip = [int(v) for v in self.pp]
            return (h + skip - 1) // skip, (w + skip - 1) // skip
        b = 0 if self.full_frame else (7 if self.method in (2, 3) else max(int(self.border), 0))
        return (W - 2 * b, H - 2 * b) if int(self.tran) & 1 else (H - 2 * b, W - 2 * b)       # TR_R90 / TR_R270 turn the frame

    def c_struct(self):
        c = _DevelopParamsC()
        c.method, c.filters, c.initialGain, c.border = int(self.method), int(self.filters), float(self.initial_gain), int(self.border)
        c.mul = (ctypes.c_float * 3)(*[float(x) for x in self.mul])
        c.doClip = int(bool(self.do_clip))
        self._keep = []
        if self.cam2work is not None:
            m = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(self.cam2work, dtype=np.float64).reshape(9)])
            self._keep.append(m)
            c.cam2work = ctypes.cast(m, ctypes.POINTER(ctypes.c_double))
        if self.wprof is not None:
            w = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(self.wprof, dtype=np.float64).reshape(9)])
            self._keep.append(w)
            c.wprof = ctypes.cast(w, ctypes.POINTER(ctypes.c_double))
        if self.denoise is not None:
            d = self.denoise.c_struct()
            self._keep.append(d)
            c.denoise = ctypes.pointer(d)
        c.nlStrength, c.nlDetail = int(self.nl_strength), int(self.nl_detail)
        c.full_frame, c.guidedChromaRadius, c.denoise_expcomp = int(bool(self.full_frame)), int(self.guided_chroma_radius), float(self.denoise_expcomp)
        c.tran, c.hr_blend = int(self.tran), int(bool(self.hr_blend))
        c.hlmax = (ctypes.c_float * 3)(*[float(x) for x in self.hlmax])
        if self.pp is not None:
            c.pp_x, c.pp_y, c.pp_width, c.pp_height, c.pp_skip = [int(v) for v in self.pp]
        if self.fattal is not None:
            thr, amt, sat = self.fattal
            c.fattal_enabled, c.fattal_threshold, c.fattal_amount, c.fattal_satcontrol = 1, int(thr), int(amt), int(bool(sat))
        if self.xtrans is not None:
            xt = (ctypes.c_int * 36)(*[int(v) for v in np.asarray(self.xtrans).reshape(36)])
            cam = (ctypes.c_float * 12)(*[float(v) for v in np.asarray(self.rgb_cam, dtype=np.float32).reshape(12)])
            self._keep += [xt, cam]
            c.xtrans = ctypes.cast(xt, ctypes.POINTER(ctypes.c_int))
            c.rgb_cam = ctypes.cast(cam, ctypes.POINTER(ctypes.c_float))
        if self.sharpen is not None:
            sc = self.sharpen.c_struct()
            self._keep.append(sc)
            c.sharpen = ctypes.addressof(sc)
        if self.chain is not None:
            self._keep.append(self.chain)
            cc = self.chain.c_struct()
            self._keep.append(cc)
            c.chain = ctypes.addressof(cc)
        return c


class DenoiseParams:
    """Mirror of the procparams::DenoiseParams fields RGB_denoise reads (defaults: rtengine/procparams.cc L1901-1918)."""

    def __init__(self, luminance=0.0, luminanceDetail=0.0, luminanceDetailThreshold=0, chrominance=15.0, chrominanceRedGreen=0.0,
                 chrominanceBlueYellow=0.0, gamma=1.7, scale=1.0, colorSpace=0, aggressive=0, chrominanceMethod=0, noiseCCurve=None,
                 wprof_inverse=None, chrominanceAutoFactor=1.0):
        self.__dict__.update(locals())
        del self.__dict__["self"]

    def c_struct(self):
        c = _DenoiseParamsC(self.luminance, self.luminanceDetail, int(self.luminanceDetailThreshold), self.chrominance,
                            self.chrominanceRedGreen, self.chrominanceBlueYellow, self.gamma, self.scale, int(self.colorSpace),
                            int(self.aggressive), int(self.chrominanceMethod), None, 0.0, None, float(self.chrominanceAutoFactor))
        if self.wprof_inverse is not None:      # needed by colorSpace 1 (LAB)
            self._wpi = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(self.wprof_inverse, dtype=np.float64).reshape(9)])
            c.wprof_inverse = ctypes.cast(self._wpi, ctypes.POINTER(ctypes.c_double))
        if self.noiseCCurve is not None:
            self._curve = np.ascontiguousarray(self.noiseCCurve, dtype=np.float32)
            assert self._curve.size == 501
            c.noiseCCurve = self._curve.ctypes.data
            c.noiseCCurveSum = float(np.float32(self._curve.sum(dtype=np.float32)))
        return c


class WaveletDev:
    """Device-resident wavelet_decomposition (mirror of rtengine::wavelet_decomposition's accessors)."""

    def __init__(self, hp, handle, W, H):
        self.hp, self.h, self.W, self.H = hp, handle, W, H

    def maxlevel(self):
        return int(self.hp.lib.art_hp_wavelet_maxlevel(self.h))

    def dims(self, lvl):
        w, h, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.hp._check(self.hp.lib.art_hp_wavelet_level_dims(self.h, lvl, ctypes.byref(w), ctypes.byref(h), ctypes.byref(s)))
        return h.value, w.value, s.value

    def band_ptr(self, lvl, d):
        return int(self.hp.lib.art_hp_wavelet_band_dev(self.h, lvl, d) or 0)

    def band(self, lvl, d):
        """Download subband d (1..3) of level lvl, or the final lowpass (d == 0), as a numpy array."""
        h, w, _ = self.dims(lvl if d else self.maxlevel() - 1)
        out = np.empty((h, w), np.float32)
        self.hp._check(self.hp.lib.art_hp_wavelet_get_band(self.h, lvl, d, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def set_band(self, lvl, d, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        self.hp._check(self.hp.lib.art_hp_wavelet_set_band(self.h, lvl, d, arr.ctypes.data_as(ctypes.c_void_p)))

    def reconstruct_dev(self, d_dst, pitch, blend=1.0):
        self.hp._check(self.hp.lib.art_hp_wavelet_reconstruct_dev(self.h, d_dst, pitch, float(blend)))

    def close(self):
        if self.h:
            self.hp.lib.art_hp_wavelet_destroy(self.h)
            self.h = None


class HotPath:
    """One context = one GPU.  Thread-compatible, not thread-safe."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.art_hp_create(ctypes.byref(h), device)
        if rc != 0:
            raise HotPathError(rc, "art_hp_create(device=%d) failed (2 = no CUDA device; there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.art_hp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise HotPathError(rc, (self.lib.art_hp_last_error(self.h) or b"").decode())

    # -- context ---------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.art_hp_set_stream(self.h, ctypes.c_void_p(cuda_stream or 0)))

    def get_stream(self):
        return self.lib.art_hp_get_stream(self.h) or 0

    def sync(self):
        self._check(self.lib.art_hp_sync(self.h))

    def launch_count(self):
        return int(self.lib.art_hp_launch_count(self.h))

    def profile_enable(self, on=True):
        self._check(self.lib.art_hp_profile_enable(self.h, 1 if on else 0))

    def profile_collect(self):
        """{kernel name: (total ms, calls)} since profile_enable(True)."""
        n = self.lib.art_hp_profile_collect(self.h)
        out = {}
        for k in range(max(n, 0)):
            name, ms, calls = ctypes.c_char_p(), ctypes.c_double(), ctypes.c_int()
            self._check(self.lib.art_hp_profile_entry(self.h, k, ctypes.byref(name), ctypes.byref(ms), ctypes.byref(calls)))
            out[name.value.decode()] = (ms.value, calls.value)
        return out

    def pinned(self, H, W, dtype=np.float32):
        return PinnedArray(self.lib, H, W, dtype)

    # -- demosaic ----------------------------------------------------------
    def demosaic_bayer(self, method, raw, filters, red=None, green=None, blue=None, initial_gain=1.0, border=4):
        """Host entry: numpy (H, W) float32 in, three (H, W) float32 planes out."""
        raw = np.asarray(raw)
        if raw.dtype != np.float32 or raw.ndim != 2 or raw.strides[1] != 4:
            raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        outs = []
        for o in (red, green, blue):
            outs.append(np.empty((H, W), np.float32) if o is None else o)
        tabs = [row_table(a) for a in [raw] + outs]
        self._check(self.lib.art_hp_demosaic_bayer(self.h, method, W, H, filters, tabs[0], tabs[1], tabs[2], tabs[3],
                                                   float(initial_gain), int(border)))
        return tuple(outs)

    def demosaic_bayer_dev(self, method, W, H, filters, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch,
                           initial_gain=1.0, border=4):
        """Device entry: raw device addresses (ints), pitches in floats; asynchronous."""
        self._check(self.lib.art_hp_demosaic_bayer_dev(self.h, method, W, H, filters, d_raw, raw_pitch,
                                                       d_r, d_g, d_b, out_pitch, float(initial_gain), int(border)))

    def demosaic_bayer_rows_dev(self, method, W, H, filters, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch,
                                row_begin, row_end, initial_gain=1.0, border=4):
        """Row-band form: only output rows [row_begin,row_end); pointers address row 0 of the frame."""
        self._check(self.lib.art_hp_demosaic_bayer_rows_dev(self.h, method, W, H, filters, d_raw, raw_pitch, d_r, d_g, d_b,
                                                            out_pitch, float(initial_gain), int(border), row_begin, row_end))

    def wavelet_decompose_dev(self, d_src, pitch, W, H, maxlvl, subsampling=1):
        """wavelet_decomposition on a device plane; returns a WaveletDev handle object."""
        h = ctypes.c_void_p()
        self._check(self.lib.art_hp_wavelet_decompose_dev(self.h, d_src, pitch, W, H, int(maxlvl), int(subsampling), ctypes.byref(h)))
        return WaveletDev(self, h, W, H)

    def wavelet_mad_dev(self, wv, d_madL):
        self._check(self.lib.art_hp_wavelet_mad_dev(self.h, wv.h, d_madL))

    def wavelet_denoise_L_dev(self, wL, d_noisevarlum, d_madL, scale=1.0):
        self._check(self.lib.art_hp_wavelet_denoise_L_dev(self.h, wL.h, d_noisevarlum, d_madL, float(scale)))

    def wavelet_denoise_AB_dev(self, wL, wab, d_noisevarchrom, d_madL, noisevar_ab, use_ccurve=False, autoch=False, scale=1.0):
        self._check(self.lib.art_hp_wavelet_denoise_AB_dev(self.h, wL.h, wab.h, d_noisevarchrom, d_madL, float(noisevar_ab),
                                                           int(use_ccurve), int(autoch), float(scale)))

    def boxblur(self, src, radius, dst=None):
        H, W = src.shape
        if dst is src:
            tab = row_table(src)
            self._check(self.lib.art_hp_boxblur(self.h, tab, tab, int(radius), W, H))
            return src
        if dst is None:
            dst = np.empty_like(src)
        self._check(self.lib.art_hp_boxblur(self.h, row_table(src), row_table(dst), int(radius), W, H))
        return dst

    def guided_filter(self, guide, src, r, epsilon, subsampling=0, dst=None):
        H, W = src.shape
        if dst is None:
            dst = np.empty_like(src)
        gt = row_table(guide)
        stab = gt if src is guide else row_table(src)
        self._check(self.lib.art_hp_guided_filter(self.h, W, H, gt, stab, row_table(dst), int(r), float(epsilon), int(subsampling)))
        return dst

    def rgb_denoise(self, r, g, b, params, wprof, calclum=None, want_residuals=False):
        """denoise::RGB_denoise in place on three host (H, W) float32 planes.  params: DenoiseParams; calclum: optional
        three half-resolution planes (needed with the chroma noise curve).  Returns (nresi, highresi) if asked."""
        H, W = r.shape
        wp = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(wprof, dtype=np.float64).reshape(9)])
        res = (ctypes.c_float * 2)() if want_residuals else None
        cl = [row_table(p) for p in calclum] if calclum is not None else [None, None, None]
        self._check(self.lib.art_hp_rgb_denoise(self.h, row_table(r), row_table(g), row_table(b), W, H, ctypes.byref(params.c_struct()),
                                                wp, cl[0], cl[1], cl[2], res))
        return (float(res[0]), float(res[1])) if want_residuals else None

    def rgb_denoise_dev(self, d_r, d_g, d_b, pitch, W, H, params, wprof, d_calclum=None, calclum_pitch=0, want_residuals=False):
        wp = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(wprof, dtype=np.float64).reshape(9)])
        res = (ctypes.c_float * 2)() if want_residuals else None
        cl = list(d_calclum) if d_calclum is not None else [None, None, None]
        self._check(self.lib.art_hp_rgb_denoise_dev(self.h, d_r, d_g, d_b, pitch, W, H, ctypes.byref(params.c_struct()), wp,
                                                    cl[0], cl[1], cl[2], calclum_pitch, res))
        return (float(res[0]), float(res[1])) if want_residuals else None

    def detail_mask(self, src, scaling, threshold, ceiling, factor, blur_type=2, blur=2.0):
        """denoise::detail_mask on a host (H, W) float32 array; returns the mask."""
        H, W = src.shape
        mask = np.empty_like(src)
        self._check(self.lib.art_hp_detail_mask(self.h, row_table(src), row_table(mask), W, H, scaling, threshold, ceiling, factor,
                                                int(blur_type), blur))
        return mask

    def detail_mask_dev(self, d_src, src_pitch, d_mask, mask_pitch, W, H, scaling, threshold, ceiling, factor, blur_type=2, blur=2.0):
        self._check(self.lib.art_hp_detail_mask_dev(self.h, d_src, src_pitch, d_mask, mask_pitch, W, H, scaling, threshold, ceiling,
                                                    factor, int(blur_type), blur))

    def nlmeans(self, img, normcoeff, strength, detail_thresh, scale=1.0):
        """denoise::NLMeans, in place on a host (H, W) float32 array."""
        H, W = img.shape
        self._check(self.lib.art_hp_nlmeans(self.h, row_table(img), W, H, normcoeff, int(strength), int(detail_thresh), scale))
        return img

    def nlmeans_dev(self, d_img, pitch, W, H, normcoeff, strength, detail_thresh, scale=1.0):
        self._check(self.lib.art_hp_nlmeans_dev(self.h, d_img, pitch, W, H, normcoeff, int(strength), int(detail_thresh), scale))

    def develop(self, raw, params, red=None, green=None, blue=None):
        """art_hp_develop on a host (H, W) float32 CFA plane; returns the three developed planes."""
        H, W = raw.shape
        out = [p if p is not None else np.empty(params.out_shape(H, W), np.float32) for p in (red, green, blue)]
        assert all(p.shape == params.out_shape(H, W) for p in out), "output planes must be %s" % (params.out_shape(H, W),)
        c = params.c_struct()
        self._check(self.lib.art_hp_develop(self.h, ctypes.byref(c), W, H, row_table(raw), row_table(out[0]), row_table(out[1]), row_table(out[2])))
        return out

    def develop_submit(self, raw, params, red, green, blue):
        """Queue one frame (pinned planes, see HotPath.pinned) and return; develop_wait() collects the oldest queued frame."""
        H, W = raw.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_develop_submit(self.h, ctypes.byref(c), W, H, row_table(raw), row_table(red), row_table(green), row_table(blue)))

    def develop_submit_packed(self, raw, params, out, bps=16, is_float=False):
        """Queue one frame whose result leaves as interleaved scanlines (Imagefloat::getScanline on the device): `out` is a pinned
        (H_out, 3 * W_out) array of uint16 / uint8 / float32 (see HotPath.pinned_bytes)."""
        H, W = raw.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_develop_submit_packed(self.h, ctypes.byref(c), W, H, row_table(raw), int(bps), int(bool(is_float)),
                                                          out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))

    def scanlines(self, r, g, b, bps=16, is_float=False):
        """Imagefloat::getScanline for every row: three host (H, W) float32 planes -> (H, 3 W) interleaved samples."""
        H, W = r.shape
        dt = {(8, False): np.uint8, (16, False): np.uint16, (16, True): np.uint16, (32, True): np.float32}[(int(bps), bool(is_float))]
        out = np.zeros((H, 3 * W), dt)
        self._check(self.lib.art_hp_scanlines(self.h, W, H, row_table(r), row_table(g), row_table(b), int(bps), int(bool(is_float)),
                                              out.ctypes.data_as(ctypes.c_void_p), out.strides[0]))
        return out

    # ---- dual demosaic (dual_demosaic_RT.cc) and its flat-region demosaicer VNG4 (vng4_demosaic_RT.cc) ----
    def demosaic_vng4(self, raw, prefilters):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        out = [np.zeros((H, W), np.float32) for _ in range(3)]
        self._check(self.lib.art_hp_demosaic_vng4(self.h, W, H, int(prefilters), row_table(raw), *[row_table(o) for o in out]))
        return out

    def dual_demosaic_bayer(self, method, second, raw, filters, prefilters=0, contrast=20.0, auto_contrast=True, initial_gain=1.0, border=4):
        """Returns ((r, g, b), contrast): contrast is the percent value dual_demosaic_RT hands back (the automatic threshold x 100)."""
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        out = [np.zeros((H, W), np.float32) for _ in range(3)]
        c = ctypes.c_double(float(contrast))
        self._check(self.lib.art_hp_dual_demosaic_bayer(self.h, int(method), int(second), W, H, int(filters), int(prefilters), row_table(raw),
                                                        *[row_table(o) for o in out], float(initial_gain), int(border), ctypes.byref(c), int(bool(auto_contrast))))
        return out, c.value

    def dual_demosaic_xtrans(self, raw, xtrans, rgb_cam, passes=3, use_cielab=True, contrast=20.0, auto_contrast=True):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        out = [np.zeros((H, W), np.float32) for _ in range(3)]
        xt = np.ascontiguousarray(xtrans, dtype=np.int32)
        cam = np.ascontiguousarray(rgb_cam, dtype=np.float32)
        c = ctypes.c_double(float(contrast))
        self._check(self.lib.art_hp_dual_demosaic_xtrans(self.h, int(passes), int(bool(use_cielab)), W, H, xt.ctypes.data_as(ctypes.c_void_p),
                                                         cam.ctypes.data_as(ctypes.c_void_p), row_table(raw), *[row_table(o) for o in out],
                                                         ctypes.byref(c), int(bool(auto_contrast))))
        return out, c.value

    def prophoto_blue(self, r, g, b):
        """proPhotoBlue (improcfun.cc L312-357) in place on three host (H, W) float32 planes."""
        H, W = r.shape
        self._check(self.lib.art_hp_prophoto_blue(self.h, W, H, row_table(r), row_table(g), row_table(b)))
        return r, g, b

    def black_and_white(self, r, g, b, params):
        """ImProcFunctions::blackAndWhite's pixel loops in place on three host (H, W) float32 working-space RGB planes."""
        H, W = r.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_black_and_white(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c)))
        return r, g, b

    def tone_equalizer(self, r, g, b, params):
        """ImProcFunctions::toneEqualizer in place on three host (H, W) float32 working-space RGB planes."""
        H, W = r.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_tone_equalizer(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c)))
        return r, g, b

    def hsl_equalizer(self, r, g, b, params):
        """ImProcFunctions::hslEqualizer in place on three host (H, W) float32 working-space RGB planes."""
        H, W = r.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_hsl_equalizer(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c)))
        return r, g, b

    def hsl_equalizer_dev(self, W, H, d_r, d_g, d_b, pitch, params):
        c = params.c_struct()
        self._check(self.lib.art_hp_hsl_equalizer_dev(self.h, W, H, d_r, d_g, d_b, pitch, ctypes.byref(c)))

    def channel_mixer(self, r, g, b, matrix):
        """ImProcFunctions::channelMixer's loop in place on three host (H, W) float32 planes; matrix = 9 floats (RGB_MATRIX: sliders / 1000.f)."""
        H, W = r.shape
        m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(9)
        self._check(self.lib.art_hp_channel_mixer(self.h, W, H, row_table(r), row_table(g), row_table(b), m.ctypes.data_as(ctypes.c_void_p)))
        return r, g, b

    # ---- preprocess: hot / dead pixel filter (badpixels.cc) ----
    def find_hot_dead_pixels(self, raw, thresh=100.0, hot=True, dead=True, xtrans=None, bad_map=None):
        """findHotDeadPixels on a host (H, W) float32 CFA plane; returns (map, count), map = (H, W) uint8 with the new marks OR-ed into bad_map."""
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        H, W = raw.shape
        m = np.zeros((H, W), np.uint8) if bad_map is None else np.ascontiguousarray(bad_map, dtype=np.uint8).copy()
        xt = None if xtrans is None else np.ascontiguousarray(xtrans, dtype=np.int32)
        n = ctypes.c_int(0)
        self._check(self.lib.art_hp_find_hot_dead_pixels(self.h, W, H, None if xt is None else xt.ctypes.data_as(ctypes.c_void_p), row_table(raw),
                                                         float(thresh), int(bool(hot)), int(bool(dead)), m.ctypes.data_as(ctypes.c_void_p), m.strides[0],
                                                         ctypes.byref(n)))
        return m, n.value

    def interpolate_bad_pixels_bayer(self, raw, filters, bad_map):
        """interpolateBadPixelsBayer in place on a host (H, W) float32 Bayer plane; returns the number of pixels interpolated."""
        H, W = raw.shape
        m = np.ascontiguousarray(bad_map, dtype=np.uint8)
        n = ctypes.c_int(0)
        self._check(self.lib.art_hp_interpolate_bad_pixels_bayer(self.h, W, H, int(filters), row_table(raw), m.ctypes.data_as(ctypes.c_void_p), m.strides[0],
                                                                 ctypes.byref(n)))
        return n.value

    def interpolate_bad_pixels_xtrans(self, raw, xtrans, bad_map):
        """interpolateBadPixelsXtrans (raster order) in place on a host (H, W) float32 X-Trans plane; returns the number of pixels interpolated."""
        H, W = raw.shape
        m = np.ascontiguousarray(bad_map, dtype=np.uint8)
        xt = np.ascontiguousarray(xtrans, dtype=np.int32)
        n = ctypes.c_int(0)
        self._check(self.lib.art_hp_interpolate_bad_pixels_xtrans(self.h, W, H, xt.ctypes.data_as(ctypes.c_void_p), row_table(raw), m.ctypes.data_as(ctypes.c_void_p),
                                                                  m.strides[0], ctypes.byref(n)))
        return n.value

    def resize_lanczos(self, planes, scale, size=None):
        """ImProcFunctions::Lanczos on three host (H, W) float32 planes; size = (dH, dW), by default resizeScale's int(n * scale + 0.5)."""
        sH, sW = planes[0].shape
        dH, dW = size if size is not None else (max(1, int(sH * scale + 0.5)), max(1, int(sW * scale + 0.5)))
        src = [np.ascontiguousarray(p, dtype=np.float32) for p in planes]
        dst = [np.zeros((dH, dW), np.float32) for _ in range(3)]
        self._check(self.lib.art_hp_resize_lanczos(self.h, sW, sH, *[row_table(p) for p in src], dW, dH, *[row_table(p) for p in dst], float(scale)))
        return dst

    def develop_wait(self):
        self._check(self.lib.art_hp_develop_wait(self.h))

    def develop_pending(self):
        return int(self.lib.art_hp_develop_pending(self.h))

    def develop_dev(self, params, W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch):
        c = params.c_struct()
        self._check(self.lib.art_hp_develop_dev(self.h, ctypes.byref(c), W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch))

    # ---- preprocess: green equilibration (green_equil_RT.cc), in place on a host (H, W) float32 Bayer plane ----
    def green_equilibrate_global(self, raw, filters, border=4):
        H, W = raw.shape
        self._check(self.lib.art_hp_green_equilibrate_global(self.h, W, H, int(filters), row_table(raw), int(border)))
        return raw

    def green_equilibrate(self, raw, filters, thresh, thresh_map=None):
        H, W = raw.shape
        self._check(self.lib.art_hp_green_equilibrate(self.h, W, H, int(filters), row_table(raw), float(thresh),
                                                      row_table(thresh_map) if thresh_map is not None else None))
        return raw

    # ---- one frame across GPUs (ABI version 3) ----
    def band_plan(self, params, W, H, own_begin, own_end, halo=200):
        return band_plan(params, W, H, own_begin, own_end, halo)

    def set_allreduce(self, fn):
        """fn(d_buf: int, count: int, stream: int) -> int sums `count` int32 at device address d_buf over the ranks, in place, on the stream."""
        if fn is None:
            self._allreduce_cb = None
            self._check(self.lib.art_hp_set_allreduce(self.h, None, None))
            return
        self._allreduce_cb = ALLREDUCE_FN(lambda user, buf, count, stream: int(fn(int(buf or 0), int(count), int(stream or 0)) or 0))
        self._check(self.lib.art_hp_set_allreduce(self.h, ctypes.cast(self._allreduce_cb, ctypes.c_void_p), None))

    def comm_unique_id(self):
        buf = (ctypes.c_ubyte * 128)()
        rc = self.lib.art_hp_comm_unique_id(buf)
        if rc:
            raise HotPathError(rc, "NCCL is not available (libnccl.so.2)")
        return bytes(buf)

    def comm_init(self, unique_id, rank, nranks):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.art_hp_comm_init(self.h, buf, int(rank), int(nranks)))

    def comm_destroy(self):
        self._check(self.lib.art_hp_comm_destroy(self.h))

    def develop_band_dev(self, params, W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch, plan):
        c = params.c_struct()
        self._check(self.lib.art_hp_develop_band_dev(self.h, ctypes.byref(c), W, H, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch, ctypes.byref(plan)))

    def demosaic_xtrans(self, raw, xtrans, rgb_cam, passes=3, use_cielab=True, out=None):
        """RawImageSource::xtrans_interpolate(passes, useCieLab) on a host (H, W) float32 X-Trans CFA plane.
        xtrans: 6x6 colours (0 R, 1 G, 2 B) as RawImage::getXtransMatrix; rgb_cam: 3x4 as RawImage::getRgbCam."""
        H, W = raw.shape
        out = out if out is not None else [np.empty((H, W), np.float32) for _ in range(3)]
        xt = (ctypes.c_int * 36)(*[int(v) for v in np.asarray(xtrans).reshape(36)])
        cam = (ctypes.c_float * 12)(*[float(v) for v in np.asarray(rgb_cam, dtype=np.float32).reshape(12)])
        self._check(self.lib.art_hp_demosaic_xtrans(self.h, int(passes), int(bool(use_cielab)), W, H, xt, cam, row_table(raw),
                                                    row_table(out[0]), row_table(out[1]), row_table(out[2])))
        return out

    def demosaic_xtrans_dev(self, W, H, xtrans, rgb_cam, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch, passes=3, use_cielab=True):
        xt = (ctypes.c_int * 36)(*[int(v) for v in np.asarray(xtrans).reshape(36)])
        cam = (ctypes.c_float * 12)(*[float(v) for v in np.asarray(rgb_cam, dtype=np.float32).reshape(12)])
        self._check(self.lib.art_hp_demosaic_xtrans_dev(self.h, int(passes), int(bool(use_cielab)), W, H, xt, cam, d_raw, raw_pitch,
                                                        d_r, d_g, d_b, out_pitch))

    def sharpen_usm(self, r, g, b, params, ws):
        """ImProcFunctions::sharpening with method "usm", in place on three host (H, W) float32 planes."""
        H, W = r.shape
        wsc = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(ws, dtype=np.float64).reshape(9)])
        c = params.c_struct()
        self._check(self.lib.art_hp_sharpen_usm(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c), wsc))

    def sharpen_usm_dev(self, W, H, d_r, d_g, d_b, pitch, params, ws):
        wsc = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(ws, dtype=np.float64).reshape(9)])
        c = params.c_struct()
        self._check(self.lib.art_hp_sharpen_usm_dev(self.h, W, H, d_r, d_g, d_b, pitch, ctypes.byref(c), wsc))

    def color_chain(self, r, g, b, params):
        """The fused per-pixel chain of ImProcFunctions::process, in place on three host (H, W) float32 planes."""
        H, W = r.shape
        c = params.c_struct()
        self._check(self.lib.art_hp_color_chain(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c)))

    def lab_histogram(self, r, g, b, params):
        """labAdjustments' hist16 (65536 uint32 counts of (int)L) of three host planes after the stages of `params` that precede the Lab stage."""
        H, W = r.shape
        c = params.c_struct()
        hist = np.zeros(65536, np.uint32)
        self._check(self.lib.art_hp_lab_histogram(self.h, W, H, row_table(r), row_table(g), row_table(b), ctypes.byref(c), hist.ctypes.data_as(ctypes.c_void_p)))
        return hist

    def color_chain_dev(self, W, H, d_r, d_g, d_b, pitch, params):
        c = params.c_struct()
        self._check(self.lib.art_hp_color_chain_dev(self.h, W, H, d_r, d_g, d_b, pitch, ctypes.byref(c)))

    def denoise_compute_params(self, r, g, b, mul, do_clip, cam2work, wprof, gamma=1.7, aggressive=0):
        """ImProcFunctions::denoiseComputeParams (AUTOMATIC chroma) on the demosaiced camera-space planes; returns
        (chrominance, chrominanceRedGreen, chrominanceBlueYellow) as float32 and the 9 x 15 per-crop statistics."""
        H, W = r.shape
        m = (ctypes.c_float * 3)(*[float(x) for x in mul])
        c2w = None if cam2work is None else (ctypes.c_double * 9)(*[float(x) for x in np.asarray(cam2work, dtype=np.float64).reshape(9)])
        wp = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(wprof, dtype=np.float64).reshape(9)])
        out3 = np.zeros(3, np.float32)
        stats = np.zeros((9, 15), np.float32)
        self._check(self.lib.art_hp_denoise_compute_params(self.h, W, H, row_table(r), row_table(g), row_table(b), m, int(bool(do_clip)), c2w, wp,
                                                           float(gamma), int(aggressive), out3.ctypes.data_as(ctypes.c_void_p),
                                                           stats.ctypes.data_as(ctypes.c_void_p)))
        return out3, stats

    def denoise_guided_smoothing(self, r, g, b, ws, guided_chroma_radius=3, scale=1.0):
        """denoise::denoiseGuidedSmoothing, in place on three host (H, W) float32 planes."""
        H, W = r.shape
        wsc = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(ws, dtype=np.float64).reshape(9)])
        self._check(self.lib.art_hp_denoise_guided_smoothing(self.h, W, H, row_table(r), row_table(g), row_table(b), wsc, int(guided_chroma_radius), float(scale)))

    def fattal(self, r, g, b, threshold, amount, satcontrol, ws):
        """ImProcFunctions::dynamicRangeCompression (ToneMapFattal02), in place on three host (H, W) float32 planes."""
        H, W = r.shape
        wsc = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(ws, dtype=np.float64).reshape(9)])
        self._check(self.lib.art_hp_fattal(self.h, W, H, row_table(r), row_table(g), row_table(b), int(threshold), int(amount),
                                           int(bool(satcontrol)), wsc))

    def fattal_dev(self, W, H, d_r, d_g, d_b, pitch, threshold, amount, satcontrol, ws):
        wsc = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(ws, dtype=np.float64).reshape(9)])
        self._check(self.lib.art_hp_fattal_dev(self.h, W, H, d_r, d_g, d_b, pitch, int(threshold), int(amount), int(bool(satcontrol)), wsc))

    def median_denoise(self, src, median_type, upper_bound=None, dst=None):
        """denoise::Median_Denoise, one iteration; upper_bound=None selects the overload without a bound."""
        H, W = src.shape
        if dst is None:
            dst = np.empty_like(src)
        self._check(self.lib.art_hp_median_denoise(self.h, row_table(src), row_table(dst), W, H, int(median_type),
                                                   int(upper_bound is not None), float(upper_bound or 0.0)))
        return dst

    def redft00_2d(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty_like(a)
        self._check(self.lib.art_hp_redft00_2d(self.h, a.shape[0], a.shape[1], a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def gauss(self, src, sigma, dst=None, gausstype=0):
        """Host entry.  dst=None -> out of place into a new array; dst is src -> the in-place variants."""
        H, W = src.shape
        if dst is src:
            tab = row_table(src)
            self._check(self.lib.art_hp_gauss(self.h, tab, tab, W, H, float(sigma), gausstype))
            return src
        if dst is None:
            dst = np.empty_like(src)
        self._check(self.lib.art_hp_gauss(self.h, row_table(src), row_table(dst), W, H, float(sigma), gausstype))
        return dst

    def gauss_dev(self, d_src, src_pitch, d_dst, dst_pitch, W, H, sigma, gausstype=0):
        self._check(self.lib.art_hp_gauss_dev(self.h, d_src, src_pitch, d_dst, dst_pitch, W, H, float(sigma), gausstype))

    def scale_colors_bayer(self, raw, filters, cblacksom, scale_mul):
        """Host entry, in place on a (H, W) float32 array; returns chmax[3]."""
        H, W = raw.shape
        tab = row_table(raw)
        bl = (ctypes.c_float * 4)(*[float(x) for x in cblacksom])
        mu = (ctypes.c_float * 4)(*[float(x) for x in scale_mul])
        ch = (ctypes.c_float * 3)()
        self._check(self.lib.art_hp_scale_colors_bayer(self.h, W, H, filters, tab, bl, mu, ch))
        return [float(ch[0]), float(ch[1]), float(ch[2])]

    def scale_colors_xtrans(self, raw, xtrans, cblacksom, scale_mul):
        """scaleColors' X-Trans branch, in place on a (H, W) float32 array; returns chmax[3]."""
        H, W = raw.shape
        xt = np.ascontiguousarray(xtrans, dtype=np.int32)
        bl = (ctypes.c_float * 3)(*[float(x) for x in cblacksom[:3]])
        mu = (ctypes.c_float * 3)(*[float(x) for x in scale_mul[:3]])
        ch = (ctypes.c_float * 3)()
        self._check(self.lib.art_hp_scale_colors_xtrans(self.h, W, H, xt.ctypes.data_as(ctypes.c_void_p), row_table(raw), bl, mu, ch))
        return [float(ch[0]), float(ch[1]), float(ch[2])]

    def scale_colors_bayer_dev(self, W, H, filters, d_raw, pitch, cblacksom, scale_mul):
        bl = (ctypes.c_float * 4)(*[float(x) for x in cblacksom])
        mu = (ctypes.c_float * 4)(*[float(x) for x in scale_mul])
        ch = (ctypes.c_float * 3)()
        self._check(self.lib.art_hp_scale_colors_bayer_dev(self.h, W, H, filters, d_raw, pitch, bl, mu, ch))
        return [float(ch[0]), float(ch[1]), float(ch[2])]

    @staticmethod
    def _mul_mat(mul, mat):
        m = (ctypes.c_float * 3)(*[float(x) for x in mul])
        mm = None
        if mat is not None:
            mm = (ctypes.c_double * 9)(*[float(x) for x in np.asarray(mat, dtype=np.float64).reshape(9)])
        return m, mm

    def scale_convert(self, red, green, blue, mul, do_clip, mat=None):
        """Host entry, in place on three (H, W) float32 arrays: getImage gain/clip + 3x3 camera->working."""
        H, W = red.shape
        tabs = [row_table(a) for a in (red, green, blue)]
        m, mm = self._mul_mat(mul, mat)
        self._check(self.lib.art_hp_scale_convert(self.h, W, H, tabs[0], tabs[1], tabs[2], m, int(bool(do_clip)), mm))

    def scale_convert_dev(self, W, H, d_r, d_g, d_b, pitch, mul, do_clip, mat=None):
        m, mm = self._mul_mat(mul, mat)
        self._check(self.lib.art_hp_scale_convert_dev(self.h, W, H, d_r, d_g, d_b, pitch, m, int(bool(do_clip)), mm))

    def border_interpolate2_dev(self, W, H, filters, lborders, d_raw, raw_pitch, d_r, d_g, d_b, out_pitch):
        self._check(self.lib.art_hp_border_interpolate2_dev(self.h, W, H, filters, lborders, d_raw, raw_pitch,
                                                            d_r, d_g, d_b, out_pitch))
