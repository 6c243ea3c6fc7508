"""Host-side mirror of the reference's RawImageSource surface for the hot path.

Reference: rtengine/rawimagesource.h L119-122, L258-294 and the dispatcher
RawImageSource::demosaic, rtengine/rawimagesource.cc L1854-1962.  Same member
names and argument meaning; the bodies call the C-ABI (include/art_hotpath.h).
"""
import numpy as np

from . import api


class RawImageSource:
    """Holds one CFA frame (rawData) and its demosaiced planes (red, green, blue).

    rawData is in the domain scaleColors leaves it in (rawimagesource.cc L2677-2859):
    float32, 0..65535.  `filters` is the dcraw CFA descriptor (RawImage::FC,
    rtengine/rawimage.h L186-189).
    """

    def __init__(self, rawData, filters, hot_path=None, initialGain=1.0, border=4):
        self.rawData = np.ascontiguousarray(rawData, dtype=np.float32)
        self.H, self.W = self.rawData.shape
        self.filters = int(filters)
        self.initialGain = float(initialGain)
        self.border = int(border)
        self.red = np.empty((self.H, self.W), np.float32)      # allocated at load, rawimagesource.cc L1458-1460
        self.green = np.empty((self.H, self.W), np.float32)
        self.blue = np.empty((self.H, self.W), np.float32)
        self._hp = hot_path or api.HotPath(0)

    def FC(self, row, col):
        return (self.filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3

    def rcd_demosaic(self):
        """rtengine/rcd_demosaic.cc L51"""
        self._hp.demosaic_bayer(api.BAYER_RCD, self.rawData, self.filters, self.red, self.green, self.blue,
                                self.initialGain, self.border)

    def amaze_demosaic_RT(self, winx=0, winy=0, winw=None, winh=None):
        """rtengine/amaze_demosaic_RT.cc L41 -- the reference only ever calls it on the full frame
        (rawimagesource.cc L1879: amaze_demosaic_RT(0, 0, W, H, rawData, red, green, blue))."""
        winw = self.W if winw is None else winw
        winh = self.H if winh is None else winh
        if (winx, winy, winw, winh) != (0, 0, self.W, self.H):
            raise api.HotPathError(5, "amaze_demosaic_RT: only the full-frame window is supported")
        self._hp.demosaic_bayer(api.BAYER_AMAZE, self.rawData, self.filters, self.red, self.green, self.blue,
                                self.initialGain, self.border)

    def demosaic(self, method="rcd"):
        """RawImageSource::demosaic dispatch, rawimagesource.cc L1872-1924 (Bayer subset)."""
        m = method.lower()
        if m == "amaze":
            self.amaze_demosaic_RT(0, 0, self.W, self.H)
        elif m == "rcd":
            self.rcd_demosaic()
        else:
            raise api.HotPathError(5, "demosaic method %r not on the hot path" % method)
        return self.red, self.green, self.blue
