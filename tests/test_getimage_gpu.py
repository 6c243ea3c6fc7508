"""getImage's geometry and highlight stage on the device (art_hp_develop with `tran`, `hr_blend`; ABI version 3) against the pinned oracle
(tests/test_oracle_getimage.py): demosaic -> crop by the border -> gains / clip -> "Blend" highlight reconstruction -> coarse rotation / mirrors ->
camera->working matrix.  Bit-exact."""
import ctypes

import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth
from art_b200.api import DevelopParams
from test_oracle_getimage import HLMAX, getimage

pytestmark = pytest.mark.gpu
CAM2WORK = np.array([[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]], np.float64)
MUL = (1.9, 1.0, 1.6)


def oracle_develop(raw, f, tran, hr, do_clip, border=4):
    P = oracle.port()
    dm = P.rcd(raw, f)
    crop = [np.ascontiguousarray(p[border:-border, border:-border]) for p in dm]
    turned = getimage(P.lib, "artoracle_getimage", crop, MUL, do_clip, hr, tran)
    return P.scale_convert(turned, (1.0, 1.0, 1.0), 0, CAM2WORK)


@pytest.mark.parametrize("tran", list(range(16)))
@pytest.mark.parametrize("hr,do_clip", [(0, 1), (1, 0)])
def test_develop_geometry_and_highlights(hot_path, tran, hr, do_clip):
    W, H = 333, 250
    f = synth.RGGB
    raw = synth.bayer_frame(W, H, f, seed=40 + tran)
    raw[60:90, 100:160] = 65535.0          # a blown patch: with mul > 1 it lands above the clip point
    params = DevelopParams(method=art_b200.BAYER_RCD, filters=f, mul=MUL, do_clip=bool(do_clip), cam2work=CAM2WORK, tran=tran, hr_blend=bool(hr), hlmax=HLMAX)
    got = hot_path.develop(raw, params)
    want = oracle_develop(raw, f, tran, hr, do_clip)
    assert got[0].shape == want[0].shape == params.out_shape(H, W)
    for g, w, ch in zip(got, want, "RGB"):
        assert np.array_equal(g, w), "%s: %d of %d differ (tran %d, hr %d)" % (ch, int((g != w).sum()), g.size, tran, hr)


def test_turned_frame_through_the_later_stages(hot_path):
    """a quarter turn in front of denoise + Fattal: the later stages see the turned frame (the same result as developing a frame that was turned
    by hand after the getImage stage)"""
    from art_b200.api import DenoiseParams
    PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
    W, H = 420, 300
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=9)
    kw = dict(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=MUL, do_clip=True, cam2work=CAM2WORK, wprof=PROPHOTO)
    turned = hot_path.develop(raw, DevelopParams(tran=1, **kw))                                         # getImage alone, turned
    dn = DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15)
    got = hot_path.develop(raw, DevelopParams(tran=1, denoise=dn, fattal=(30, 20, 0), **kw))
    want = [t.copy() for t in turned]
    hot_path.rgb_denoise(want[0], want[1], want[2], dn, PROPHOTO)
    hot_path.fattal(want[0], want[1], want[2], 30, 20, 0, PROPHOTO)
    for g, w in zip(got, want):
        assert g.shape == (W - 8, H - 8) and np.array_equal(g, w)
