"""GPU parity for the automatic chroma estimator (art_hp_denoise_compute_params = ImProcFunctions::denoiseComputeParams for
DenoiseParams::ChrominanceMethod::AUTOMATIC, the reference's default) through the C-ABI against the oracle port, which
tests/test_oracle_denoise_auto.py pins bit-exact to the reference's RGB_denoise_info / calcautodn_info / nine-crop combination
compiled in place.  Bit-exact on all 9 x 15 per-crop statistics (wavelet MADs, the raster-order running sums) and on the three
resulting chrominance parameters; then the develop entry with the reference's default DenoiseParams."""
import ctypes

import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth
from art_b200.api import DenoiseParams, DevelopParams
from test_oracle_denoise import PROPHOTO, rgb_frame
from test_oracle_denoise_auto import EXPCOMP, NAMES, auto

pytestmark = pytest.mark.gpu
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
CAM2WORK = np.array([[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]], np.float64)
MUL = (1.9, 1.0, 1.6)


def oracle_estimate(planes, mul, do_clip, cam2work, gamma=1.7, aggressive=0):
    """denoiseComputeParams over the port: nine crops of half the frame (ipdenoise.cc L866-877, L905-914), getImage's gain / clip per crop,
    provicalc = the crop's even pixels through the camera->working matrix"""
    P = oracle.port()
    H, W = planes[0].shape
    crW, crH = W // 2, H // 2
    cw = [50, W // 2 - crW // 2, W - crW - 50]
    chh = [50, H // 2 - crH // 2, H - crH - 50]
    stats = np.zeros((9, 15), np.float32)
    for wcr in range(3):
        for hcr in range(3):
            crop = [np.ascontiguousarray(p[chh[hcr]:chh[hcr] + crH, cw[wcr]:cw[wcr] + crW]) for p in planes]
            crop = P.scale_convert(crop, mul, do_clip, None)
            calc = P.scale_convert([np.ascontiguousarray(c[::2, ::2]) for c in crop], (1.0, 1.0, 1.0), False, cam2work)
            out = np.zeros(15, np.float32)
            wp = PROPHOTO.copy()
            rc = P.lib.artoracle_denoise_info(crop[0].ctypes.data_as(fp), crop[1].ctypes.data_as(fp), crop[2].ctypes.data_as(fp), crW, crH,
                                              calc[0].ctypes.data_as(fp), calc[1].ctypes.data_as(fp), calc[2].ctypes.data_as(fp),
                                              ctypes.c_double(gamma), int(aggressive), ctypes.c_double(1.0), ctypes.c_double(EXPCOMP),
                                              wp.ctypes.data_as(dp), out.ctypes.data_as(fp))
            assert rc == 0
            stats[hcr * 3 + wcr] = out
    return auto(P.lib, "artoracle_denoise_auto_params", np.ascontiguousarray(stats), 1, aggressive), stats


@pytest.mark.parametrize("W,H", [(600, 400), (701, 467), (1203, 807), (2100, 1400)])
@pytest.mark.parametrize("gamma,aggressive", [(1.7, 0), (3.0, 1), (1.0, 0)])
def test_compute_params_matches_oracle(hot_path, W, H, gamma, aggressive):
    planes = rgb_frame(H, W, seed=W + H, noise=1200.0, hot=True)
    planes[0][H // 2:, : W // 2] *= 1.7           # saturated reds: the red_yel / skin counters
    planes[2][H // 2:, : W // 2] *= 0.3
    want3, wstats = oracle_estimate(planes, MUL, True, CAM2WORK, gamma, aggressive)
    got3, gstats = hot_path.denoise_compute_params(planes[0], planes[1], planes[2], MUL, True, CAM2WORK, PROPHOTO, gamma, aggressive)
    for k in range(9):
        for j, name in enumerate(NAMES):
            assert gstats[k, j] == wstats[k, j], "crop %d %s: %r vs %r" % (k, name, gstats[k, j], wstats[k, j])
    assert np.array_equal(got3, want3), (got3, want3)
    assert np.isfinite(got3).all() and got3[0] > 0


def test_develop_with_reference_default_denoise_params(hot_path):
    """DenoiseParams() as the reference constructs it (procparams.cc L1901-1918) but enabled: AUTOMATIC chroma, luminance 0 -- the estimate
    feeds the chroma-only RGB_denoise; bit-exact against the oracle chain fed with the oracle's estimate."""
    from test_develop_gpu import crop, run_chain_denoise
    W, H = 645, 404
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=31)
    P = oracle.port()
    dm = crop(P.amaze(raw, synth.RGGB, 1.0, 4), 4)
    est, _ = oracle_estimate(dm, MUL, True, CAM2WORK)
    planes = P.scale_convert(dm, MUL, True, CAM2WORK)
    dn = (0.0, 0.0, 0, float(est[0]) * 1.0, float(est[1]) * 1.0, float(est[2]) * 1.0, 1.7, 1.0)
    want = run_chain_denoise(P.lib, planes, dn, None)
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=MUL, do_clip=True, cam2work=CAM2WORK, wprof=PROPHOTO,
                           denoise=DenoiseParams(chrominanceMethod=1, chrominanceAutoFactor=1.0))
    got = hot_path.develop(raw, params)
    for x, y, ch in zip(got, want, "RGB"):
        assert np.array_equal(x, y), "%s: %d of %d differ" % (ch, int((x != y).sum()), x.size)
    assert any((x != p).any() for x, p in zip(got, planes)), "the denoise stage did nothing"
