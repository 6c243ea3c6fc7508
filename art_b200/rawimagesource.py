"""Host-side mirror of the reference's RawImageSource surface for the hot path.

Reference: rtengine/rawimagesource.h L119-122, L258-294, the dispatcher RawImageSource::demosaic (rtengine/rawimagesource.cc
L1854-1962), dual_demosaic_RT's fall-through (rtengine/dual_demosaic_RT.cc L39-72) and the preprocess members the hot path covers
(scaleColors L2677-2859, findHotDeadPixels / interpolateBadPixelsBayer / interpolateBadPixelsXtrans in rtengine/badpixels.cc,
green_equilibrate / green_equilibrate_global).  Same member names, method strings (procparams.cc L3015-3034, L3097-3104) and
argument meaning; the bodies call the C-ABI through art_b200.HotPath.  Methods of the reference that are not on the hot path
raise HotPathError(ART_HP_ERR_UNSUPPORTED) instead of silently falling back -- the reference would run them on the CPU.
"""
import numpy as np

from . import api

ST_BAYER, ST_FUJI_XTRANS = 1, 2          # rtengine/rawimage.h getSensorType()
UNSUPPORTED = 5                          # ART_HP_ERR_UNSUPPORTED

# RAWParams::BayerSensor::Method -> (first demosaicer, second demosaicer of the dual methods or None)
_BAYER = {
    "amaze": (api.BAYER_AMAZE, None), "rcd": (api.BAYER_RCD, None), "vng4": ("vng4", None),
    "amazebilinear": (api.BAYER_AMAZE, 0), "amazevng4": (api.BAYER_AMAZE, 1),        # ART_HP_DUAL_BILINEAR = 0, ART_HP_DUAL_VNG4 = 1
    "rcdbilinear": (api.BAYER_RCD, 0), "rcdvng4": (api.BAYER_RCD, 1),
}
_BAYER_OFF_PATH = ("hphd", "ahd", "eahd", "dcb", "dcbbilinear", "dcbvng4", "igv", "lmmse", "fast", "mono", "pixelshift", "none")
# RAWParams::XTransSensor::Method -> (passes, useCieLab, dual)
_XTRANS = {
    "3-pass (best)": (3, True, False), "1-pass (medium)": (1, False, False),
    "4-pass": (3, True, True), "2-pass": (1, False, True),                           # dual_demosaic_RT L63-67: FOUR_PASS = 3 passes + CIELab, else 1 pass
}
_XTRANS_ALIASES = {"three_pass": "3-pass (best)", "one_pass": "1-pass (medium)", "four_pass": "4-pass", "two_pass": "2-pass"}
_XTRANS_OFF_PATH = ("fast", "mono", "none")


class RawImageSource:
    """Holds one CFA frame (rawData) and its demosaiced planes (red, green, blue).

    rawData is in the domain scaleColors leaves it in (rawimagesource.cc L2677-2859): float32, 0..65535.  A Bayer source is
    described by `filters` (the dcraw CFA descriptor, RawImage::FC, rtengine/rawimage.h L186-189) and, for VNG4, `prefilters`
    (RawImage::prefilters: the CFA with the second green as colour 3); an X-Trans source by `xtrans` (6x6 colours,
    RawImage::getXtransMatrix) and `rgb_cam` (3x4, RawImage::getRgbCam).
    """

    def __init__(self, rawData, filters=0, hot_path=None, initialGain=1.0, border=4, xtrans=None, rgb_cam=None, prefilters=None):
        self.rawData = np.ascontiguousarray(rawData, dtype=np.float32)
        self.H, self.W = self.rawData.shape
        self.filters = int(filters)
        self.prefilters = None if prefilters is None else int(prefilters)
        self.xtrans = None if xtrans is None else np.ascontiguousarray(xtrans, dtype=np.int32).reshape(6, 6)
        self.rgb_cam = None if rgb_cam is None else np.ascontiguousarray(rgb_cam, dtype=np.float32).reshape(3, 4)
        self.initialGain = float(initialGain)
        self.border = int(border)
        self.red = np.empty((self.H, self.W), np.float32)      # allocated at load, rawimagesource.cc L1458-1460
        self.green = np.empty((self.H, self.W), np.float32)
        self.blue = np.empty((self.H, self.W), np.float32)
        self._hp = hot_path or api.HotPath(0)

    # ---- RawImage
    def getSensorType(self):
        return ST_FUJI_XTRANS if self.xtrans is not None else ST_BAYER

    def FC(self, row, col):
        return (self.filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3

    def XTRANSFC(self, row, col):
        return int(self.xtrans[row % 6][col % 6])

    def _take(self, planes):
        for dst, src in zip((self.red, self.green, self.blue), planes):
            if src is not dst:
                dst[...] = src

    # ---- demosaicers
    def rcd_demosaic(self):
        """rtengine/rcd_demosaic.cc L51"""
        self._hp.demosaic_bayer(api.BAYER_RCD, self.rawData, self.filters, self.red, self.green, self.blue, self.initialGain, self.border)

    def amaze_demosaic_RT(self, winx=0, winy=0, winw=None, winh=None):
        """rtengine/amaze_demosaic_RT.cc L41 -- the reference only ever calls it on the full frame
        (rawimagesource.cc L1873: amaze_demosaic_RT(0, 0, W, H, rawData, red, green, blue))."""
        winw = self.W if winw is None else winw
        winh = self.H if winh is None else winh
        if (winx, winy, winw, winh) != (0, 0, self.W, self.H):
            raise api.HotPathError(UNSUPPORTED, "amaze_demosaic_RT: only the full-frame window is supported")
        self._hp.demosaic_bayer(api.BAYER_AMAZE, self.rawData, self.filters, self.red, self.green, self.blue, self.initialGain, self.border)

    def vng4_demosaic(self):
        """rtengine/vng4_demosaic_RT.cc L32"""
        if self.prefilters is None:
            raise api.HotPathError(1, "vng4_demosaic needs RawImage::prefilters")
        self._take(self._hp.demosaic_vng4(self.rawData, self.prefilters))

    def xtrans_interpolate(self, passes, useCieLab):
        """rtengine/xtrans_demosaic.cc L181"""
        self._take(self._hp.demosaic_xtrans(self.rawData, self.xtrans, self.rgb_cam, int(passes), bool(useCieLab)))

    def dual_demosaic_RT(self, isBayer, method, contrast, autoContrast):
        """rtengine/dual_demosaic_RT.cc L39-152; returns the contrast threshold it hands back (in / out in the reference).
        contrast == 0 without autoContrast runs the first demosaicer alone (L43-72)."""
        if isBayer:
            first, second = _BAYER[method]
            if contrast == 0.0 and not autoContrast:
                (self.amaze_demosaic_RT if first == api.BAYER_AMAZE else self.rcd_demosaic)()
                return contrast
            if second == 1 and self.prefilters is None:
                raise api.HotPathError(1, "the VNG4 dual methods need RawImage::prefilters")
            planes, c = self._hp.dual_demosaic_bayer(first, second, self.rawData, self.filters, self.prefilters or 0, contrast, autoContrast,
                                                      self.initialGain, self.border)
        else:
            passes, cielab, _ = _XTRANS[method]
            if contrast == 0.0 and not autoContrast:
                self.xtrans_interpolate(passes, cielab)
                return contrast
            planes, c = self._hp.dual_demosaic_xtrans(self.rawData, self.xtrans, self.rgb_cam, passes, cielab, contrast, autoContrast)
        self._take(planes)
        return c

    def demosaic(self, method="rcd", autoContrast=False, contrastThreshold=0.0, dualDemosaicContrast=20.0):
        """RawImageSource::demosaic(raw, autoContrast, contrastThreshold), rawimagesource.cc L1854-1949: `method` is
        raw.bayersensor.method / raw.xtranssensor.method as its procparams string, dualDemosaicContrast the sensor's
        dualDemosaicContrast.  Returns (red, green, blue); the contrast threshold a dual method reports is left in
        self.contrastThreshold (the reference's in / out argument)."""
        m = method.lower()
        self.contrastThreshold = contrastThreshold
        if self.getSensorType() == ST_BAYER:
            if m in _BAYER_OFF_PATH or m not in _BAYER:
                raise api.HotPathError(UNSUPPORTED, "Bayer demosaic method %r is not on the hot path" % method)
            first, second = _BAYER[m]
            if second is not None:      # L1875-1887: without autoContrast the sensor's threshold goes in by value and nothing comes back
                c = self.dual_demosaic_RT(True, m, contrastThreshold if autoContrast else dualDemosaicContrast, autoContrast)
                if autoContrast:
                    self.contrastThreshold = c
            elif first == "vng4":
                self.vng4_demosaic()
            elif first == api.BAYER_AMAZE:
                self.amaze_demosaic_RT(0, 0, self.W, self.H)
            else:
                self.rcd_demosaic()
        else:
            m = _XTRANS_ALIASES.get(m, m)
            if m in _XTRANS_OFF_PATH or m not in _XTRANS:
                raise api.HotPathError(UNSUPPORTED, "X-Trans demosaic method %r is not on the hot path" % method)
            passes, cielab, dual = _XTRANS[m]
            if dual:
                c = self.dual_demosaic_RT(False, m, contrastThreshold if autoContrast else dualDemosaicContrast, autoContrast)
                if autoContrast:
                    self.contrastThreshold = c
            else:
                self.xtrans_interpolate(passes, cielab)
        return self.red, self.green, self.blue

    # ---- preprocess members on the hot path (rawimagesource.cc preprocess(), L1380-1530)
    def scaleColors(self, cblacksom, scale_mul):
        """the per-pixel loop of scaleColors (L2731-2826); the caller keeps computing cblacksom / scale_mul (L2700-2719).  Returns chmax."""
        if self.getSensorType() == ST_BAYER:
            return self._hp.scale_colors_bayer(self.rawData, self.filters, cblacksom, scale_mul)
        return self._hp.scale_colors_xtrans(self.rawData, self.xtrans, cblacksom, scale_mul)

    def findHotDeadPixels(self, bpMap, thresh, findHotPixels, findDeadPixels):
        """badpixels.cc L477-627; bpMap: (H, W) uint8 standing in for PixelsMap, updated in place.  Returns the number of pixels found."""
        m, n = self._hp.find_hot_dead_pixels(self.rawData, thresh, findHotPixels, findDeadPixels, self.xtrans, bpMap)
        bpMap[...] = m
        return n

    def interpolateBadPixelsBayer(self, bitmapBads):
        """badpixels.cc L36-180"""
        return self._hp.interpolate_bad_pixels_bayer(self.rawData, self.filters, bitmapBads)

    def interpolateBadPixelsXtrans(self, bitmapBads):
        """badpixels.cc L288-475 (the one-thread order)"""
        return self._hp.interpolate_bad_pixels_xtrans(self.rawData, self.xtrans, bitmapBads)

    def green_equilibrate_global(self):
        return self._hp.green_equilibrate_global(self.rawData, self.filters, self.border)

    def green_equilibrate(self, thresh, thresh_map=None):
        return self._hp.green_equilibrate(self.rawData, self.filters, thresh, thresh_map)
