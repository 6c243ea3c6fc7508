"""Pin the chroma-smoothing oracle (oracle/smoothing_port.c: denoiseGuidedSmoothing) against the reference's own guided_smoothing /
guidedFilterLog / guidedFilter compiled in place (oracle/_ref).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle
from test_oracle_denoise import PROPHOTO, rgb_frame

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)


def run(lib, name, planes, radius, scale):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    ws = PROPHOTO.copy()
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, ws.ctypes.data_as(dp), int(radius), ctypes.c_double(scale))
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (264, 200), (301, 203)])
@pytest.mark.parametrize("radius,scale", [(3, 1.0), (1, 1.0), (10, 1.0), (3, 2.0), (3, 8.0), (0, 1.0)])
@pytest.mark.parametrize("hot", [False, True])
def test_port_matches_reference(W, H, radius, scale, hot):
    planes = rgb_frame(H, W, seed=W + radius, noise=2500.0, hot=hot)
    planes[1][H // 2, W // 3] = 0.0
    planes[0][H // 2, W // 3] = 0.0
    planes[2][H // 2, W // 3] = 0.0                      # a black pixel: the bump guard
    got = run(oracle.port().lib, "artoracle_denoise_guided_smoothing", planes, radius, scale)
    want = run(oracle.ref().lib, "artref_denoise_guided_smoothing", planes, radius, scale)
    for g, w, ch in zip(got, want, "RGB"):
        eq = (g == w) | (np.isnan(g) & np.isnan(w))
        assert eq.all(), "%s: %d of %d differ, max abs %g" % (ch, int((~eq).sum()), g.size, float(np.nanmax(np.abs(g - w))))
    if radius == 0:
        assert all(np.array_equal(g, p) for g, p in zip(got, planes))


@needs_ref
def test_luminance_is_kept_and_chroma_is_smoothed():
    planes = rgb_frame(200, 264, seed=4, noise=2500.0)
    out = run(oracle.port().lib, "artoracle_denoise_guided_smoothing", planes, 3, 1.0)
    w = PROPHOTO[1].astype(np.float32)
    yin = planes[0] * w[0] + planes[1] * w[1] + planes[2] * w[2]
    yout = out[0] * w[0] + out[1] * w[1] + out[2] * w[2]
    assert np.allclose(yin, yout, rtol=2e-4, atol=2.0)
    assert np.std(out[0] - yout) < np.std(planes[0] - yin)
