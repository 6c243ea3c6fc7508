"""GPU parity for denoise::RGB_denoise (through the C-ABI) against the oracle port (pinned bit-exact to the reference
function in test_oracle_denoise.py).

* chroma-only settings never reach the block DCT: bit-exact.
* with luminance denoising the two 64x64 block DCTs run as fp32 matrix products on the GPU while the oracle (like
  the reference) delegates them to an FFTW-style routine (here a double-precision stand-in, fftw3f being absent):
  tolerance 1e-4 relative (BASELINE.json north_star) plus 0.02 absolute on the 0..65535 scale for values near zero.
"""
import numpy as np
import pytest

import oracle
from test_oracle_denoise import CASES, PROPHOTO, PROPHOTO_INV, calclum_of, noise_ccurve, rgb_frame, run

pytestmark = pytest.mark.gpu


def gpu_run(hot_path, planes, params, curve, aggressive=0, lab=0):
    from art_b200.api import DenoiseParams
    lum, det, thr, chroma, rg, by, gamma, scale = params
    p = DenoiseParams(luminance=lum, luminanceDetail=det, luminanceDetailThreshold=thr, chrominance=chroma, chrominanceRedGreen=rg,
                      chrominanceBlueYellow=by, gamma=gamma, scale=scale, noiseCCurve=curve[0] if curve else None, aggressive=aggressive,
                      colorSpace=lab, wprof_inverse=PROPHOTO_INV if lab else None)
    out = [q.copy() for q in planes]
    res = hot_path.rgb_denoise(out[0], out[1], out[2], p, PROPHOTO, calclum=calclum_of(planes) if curve else None, want_residuals=True)
    return out, res


@pytest.mark.parametrize("W,H,params,curve,hot", CASES + [(1203, 807, (30, 50, 0, 15, 0, 0, 1.7, 1.0), True, True)])
def test_rgb_denoise_matches_oracle(hot_path, W, H, params, curve, hot):
    planes = rgb_frame(H, W, seed=W * 7 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    want, wres = run(oracle.port().lib, "artoracle_rgb_denoise", planes, params, cc)
    got, gres = gpu_run(hot_path, planes, params, cc)
    exact = params[0] == 0            # no luminance denoising -> no DCT
    for name, x, y in zip("rgb", got, want):
        if exact:
            assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
        else:
            err = np.abs(x - y)
            lim = 1e-4 * np.abs(y) + 0.02
            assert (err <= lim).all(), "%s: max |err| %g at value %g, %d of %d over tolerance" % (
                name, float(err.max()), float(y.flat[err.argmax()]), int((err > lim).sum()), x.size)
    assert np.array_equal(np.float32(gres), wres), (gres, wres)          # chroma residual statistics never see the DCT


@pytest.mark.parametrize("W,H,params,curve,hot", CASES)
def test_rgb_denoise_aggressive_matches_oracle(hot_path, W, H, params, curve, hot):
    """DenoiseParams::aggressive: two more wavelet levels, BiShrink + standard shrinkage of the chroma channels, luminance shrunk twice"""
    planes = rgb_frame(H, W, seed=W * 5 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    want, wres = run(oracle.port().lib, "artoracle_rgb_denoise_ex", planes, params, cc, aggressive=1)
    got, gres = gpu_run(hot_path, planes, params, cc, aggressive=1)
    exact = params[0] == 0
    for name, x, y in zip("rgb", got, want):
        if exact:
            assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
        else:
            err = np.abs(x - y)
            lim = 1e-4 * np.abs(y) + 0.05        # 1e-4 relative; 0.05 on the 0..65535 scale (8e-7 of full scale) for samples near zero: the
            #                                      luminance is shrunk twice in this mode, so the DCT stage sees a larger residual
            assert (err <= lim).all(), "%s: max |err| %g, %d of %d over tolerance" % (name, float(err.max()), int((err > lim).sum()), x.size)
    assert np.array_equal(np.float32(gres), wres), (gres, wres)


@pytest.mark.parametrize("W,H,params,curve,hot", CASES)
def test_rgb_denoise_lab_colour_space_matches_oracle(hot_path, W, H, params, curve, hot):
    """DenoiseParams::colorSpace == LAB"""
    planes = rgb_frame(H, W, seed=W * 3 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    want, wres = run(oracle.port().lib, "artoracle_rgb_denoise_ex2", planes, params, cc, with_inverse=True, aggressive=0, lab=1)
    got, gres = gpu_run(hot_path, planes, params, cc, lab=1)
    exact = params[0] == 0
    for name, x, y in zip("rgb", got, want):
        if exact:
            assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
        else:
            err = np.abs(x - y)
            lim = 1e-4 * np.abs(y) + 0.05
            assert (err <= lim).all(), "%s: max |err| %g at %g, %d of %d over tolerance" % (name, float(err.max()), float(y.flat[err.argmax()]), int((err > lim).sum()), x.size)
    assert np.array_equal(np.float32(gres), wres), (gres, wres)


def test_rgb_denoise_error_budget(hot_path, capsys):
    """Reports how far the fp32 GPU DCT is from the double-precision stand-in (the number quoted in DESIGN.md)."""
    W, H = 640, 480
    planes = rgb_frame(H, W, seed=77)
    params = (30, 50, 0, 15, 0, 0, 1.7, 1.0)
    want, _ = run(oracle.port().lib, "artoracle_rgb_denoise", planes, params, noise_ccurve())
    got, _ = gpu_run(hot_path, planes, params, noise_ccurve())
    worst = max(float(np.abs(x - y).max()) for x, y in zip(got, want))
    rel = max(float((np.abs(x - y) / np.maximum(np.abs(y), 1.0)).max()) for x, y in zip(got, want))
    with capsys.disabled():
        print("\n[rgb_denoise] max |GPU - oracle| = %.4g (0..65535 scale), max relative = %.3g" % (worst, rel))
    assert rel < 1e-4


def test_rgb_denoise_rejects_unsupported(hot_path):
    from art_b200.api import DenoiseParams, HotPathError
    planes = rgb_frame(64, 64, seed=1)
    with pytest.raises(HotPathError):
        hot_path.rgb_denoise(planes[0], planes[1], planes[2], DenoiseParams(luminance=10, colorSpace=1), PROPHOTO)      # LAB without the inverse matrix
    with pytest.raises(HotPathError):        # curve without calclum
        hot_path.rgb_denoise(planes[0], planes[1], planes[2], DenoiseParams(luminance=10, noiseCCurve=noise_ccurve()[0]), PROPHOTO)
