"""The host-side mirror of RawImageSource::demosaic's dispatch (rawimagesource.cc L1854-1949, dual_demosaic_RT.cc L39-72) and of the preprocess
members, on the CPU: a recording stand-in for art_b200.HotPath shows which C-ABI entry each reference method string reaches and with which
arguments.  (The entries themselves are checked against the oracle in the -m gpu tests.)"""
import numpy as np
import pytest

import art_b200
from art_b200 import api
from art_b200.rawimagesource import RawImageSource, ST_BAYER, ST_FUJI_XTRANS

RGGB = 0x94949494


class Recorder:
    """stands in for HotPath: records (entry, args) and returns planes of the right shape"""

    def __init__(self):
        self.calls = []

    def _planes(self, raw):
        return [np.full(raw.shape, k + 1.0, np.float32) for k in range(3)]

    def demosaic_bayer(self, method, raw, filters, red, green, blue, initial_gain, border):
        self.calls.append(("demosaic_bayer", method, filters, initial_gain, border))
        red[...], green[...], blue[...] = 1.0, 2.0, 3.0

    def demosaic_vng4(self, raw, prefilters):
        self.calls.append(("demosaic_vng4", prefilters))
        return self._planes(raw)

    def demosaic_xtrans(self, raw, xtrans, rgb_cam, passes, use_cielab):
        self.calls.append(("demosaic_xtrans", passes, use_cielab))
        return self._planes(raw)

    def dual_demosaic_bayer(self, method, second, raw, filters, prefilters, contrast, auto_contrast, initial_gain, border):
        self.calls.append(("dual_demosaic_bayer", method, second, prefilters, contrast, auto_contrast))
        return self._planes(raw), 12.5

    def dual_demosaic_xtrans(self, raw, xtrans, rgb_cam, passes, use_cielab, contrast, auto_contrast):
        self.calls.append(("dual_demosaic_xtrans", passes, use_cielab, contrast, auto_contrast))
        return self._planes(raw), 7.0

    def scale_colors_bayer(self, raw, filters, cblacksom, scale_mul):
        self.calls.append(("scale_colors_bayer", filters))
        return [1.0, 2.0, 3.0]

    def scale_colors_xtrans(self, raw, xtrans, cblacksom, scale_mul):
        self.calls.append(("scale_colors_xtrans",))
        return [1.0, 2.0, 3.0]

    def find_hot_dead_pixels(self, raw, thresh, hot, dead, xtrans, bad_map):
        self.calls.append(("find_hot_dead_pixels", thresh, hot, dead, xtrans is not None))
        m = bad_map.copy()
        m[1, 1] = 1
        return m, 1

    def interpolate_bad_pixels_bayer(self, raw, filters, m):
        self.calls.append(("interpolate_bad_pixels_bayer", filters))
        return int(m.sum())

    def interpolate_bad_pixels_xtrans(self, raw, xtrans, m):
        self.calls.append(("interpolate_bad_pixels_xtrans",))
        return int(m.sum())


def bayer(rec, **kw):
    return RawImageSource(np.zeros((12, 16), np.float32), RGGB, hot_path=rec, prefilters=0xb4b4b4b4, **kw)


def xtrans(rec):
    xt = np.array([[1, 1, 0, 1, 1, 2], [1, 1, 2, 1, 1, 0], [2, 0, 1, 0, 2, 1], [1, 1, 2, 1, 1, 0], [1, 1, 0, 1, 1, 2], [0, 2, 1, 2, 0, 1]])
    return RawImageSource(np.zeros((12, 18), np.float32), hot_path=rec, xtrans=xt, rgb_cam=np.eye(3, 4))


@pytest.mark.parametrize("method,want", [
    ("amaze", ("demosaic_bayer", api.BAYER_AMAZE, RGGB, 1.0, 4)), ("RCD", ("demosaic_bayer", api.BAYER_RCD, RGGB, 1.0, 4)),
    ("vng4", ("demosaic_vng4", 0xb4b4b4b4)),
    ("amazebilinear", ("dual_demosaic_bayer", api.BAYER_AMAZE, 0, 0xb4b4b4b4, 20.0, False)),
    ("amazevng4", ("dual_demosaic_bayer", api.BAYER_AMAZE, 1, 0xb4b4b4b4, 20.0, False)),
    ("rcdbilinear", ("dual_demosaic_bayer", api.BAYER_RCD, 0, 0xb4b4b4b4, 20.0, False)),
    ("rcdvng4", ("dual_demosaic_bayer", api.BAYER_RCD, 1, 0xb4b4b4b4, 20.0, False)),
])
def test_bayer_dispatch(method, want):
    rec = Recorder()
    src = bayer(rec)
    assert src.getSensorType() == ST_BAYER
    r, g, b = src.demosaic(method)
    assert rec.calls == [want]
    assert (r == 1).all() and (g == 2).all() and (b == 3).all() and r is src.red


def test_dual_threshold_semantics():
    """rawimagesource.cc L1875-1887: without autoContrast the sensor's dualDemosaicContrast goes in by value and contrastThreshold is untouched;
    with it contrastThreshold goes in and comes back; dual_demosaic_RT.cc L43-72: contrast 0 without autoContrast = the first demosaicer alone"""
    rec = Recorder()
    src = bayer(rec)
    src.demosaic("amazevng4", autoContrast=False, contrastThreshold=3.0, dualDemosaicContrast=35.0)
    assert rec.calls[-1][4:] == (35.0, False) and src.contrastThreshold == 3.0
    src.demosaic("amazevng4", autoContrast=True, contrastThreshold=3.0, dualDemosaicContrast=35.0)
    assert rec.calls[-1][4:] == (3.0, True) and src.contrastThreshold == 12.5
    src.demosaic("rcdvng4", autoContrast=False, dualDemosaicContrast=0.0)
    assert rec.calls[-1] == ("demosaic_bayer", api.BAYER_RCD, RGGB, 1.0, 4)
    x = xtrans(rec)
    x.demosaic("4-pass", autoContrast=False, dualDemosaicContrast=0.0)
    assert rec.calls[-1] == ("demosaic_xtrans", 3, True)


@pytest.mark.parametrize("method,want", [
    ("3-pass (best)", ("demosaic_xtrans", 3, True)), ("1-pass (medium)", ("demosaic_xtrans", 1, False)), ("three_pass", ("demosaic_xtrans", 3, True)),
    ("4-pass", ("dual_demosaic_xtrans", 3, True, 20.0, False)), ("2-pass", ("dual_demosaic_xtrans", 1, False, 20.0, False)),
])
def test_xtrans_dispatch(method, want):
    rec = Recorder()
    src = xtrans(rec)
    assert src.getSensorType() == ST_FUJI_XTRANS and src.XTRANSFC(7, 8) == int(src.xtrans[1][2])
    src.demosaic(method)
    assert rec.calls == [want]


@pytest.mark.parametrize("method", ["lmmse", "igv", "dcb", "ahd", "hphd", "eahd", "fast", "mono", "pixelshift", "none", "no such method"])
def test_methods_off_the_hot_path_fail_loudly(method):
    rec = Recorder()
    with pytest.raises(art_b200.HotPathError):
        bayer(rec).demosaic(method)
    assert rec.calls == []
    if method in ("fast", "mono", "none", "no such method"):
        with pytest.raises(art_b200.HotPathError):
            xtrans(rec).demosaic(method)


def test_amaze_window_and_missing_prefilters():
    rec = Recorder()
    with pytest.raises(art_b200.HotPathError):
        bayer(rec).amaze_demosaic_RT(0, 0, 8, 8)
    src = RawImageSource(np.zeros((12, 16), np.float32), RGGB, hot_path=rec)
    with pytest.raises(art_b200.HotPathError):
        src.demosaic("vng4")
    with pytest.raises(art_b200.HotPathError):
        src.demosaic("amazevng4")
    src.demosaic("amazebilinear")            # the bilinear dual method needs no prefilters
    assert rec.calls[-1][:3] == ("dual_demosaic_bayer", api.BAYER_AMAZE, 0)


def test_fc_and_preprocess_members():
    rec = Recorder()
    src = bayer(rec)
    assert [src.FC(0, 0), src.FC(0, 1), src.FC(1, 0), src.FC(1, 1)] == [0, 1, 1, 2]
    assert src.scaleColors([0] * 4, [1] * 4) == [1.0, 2.0, 3.0] and rec.calls[-1] == ("scale_colors_bayer", RGGB)
    bp = np.zeros((12, 16), np.uint8)
    assert src.findHotDeadPixels(bp, 100.0, True, False) == 1 and bp[1, 1] == 1 and rec.calls[-1] == ("find_hot_dead_pixels", 100.0, True, False, False)
    assert src.interpolateBadPixelsBayer(bp) == 1
    x = xtrans(rec)
    x.scaleColors([0] * 4, [1] * 4)
    assert rec.calls[-1] == ("scale_colors_xtrans",)
    x.findHotDeadPixels(np.zeros((12, 18), np.uint8), 50.0, True, True)
    assert rec.calls[-1] == ("find_hot_dead_pixels", 50.0, True, True, True)
    assert x.interpolateBadPixelsXtrans(np.ones((12, 18), np.uint8)) == 216
