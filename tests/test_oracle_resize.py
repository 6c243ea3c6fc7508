"""Pin the Lanczos resampler's oracle (oracle/resize_port.c) against the reference's own ImProcFunctions::Lanczos compiled in place
(oracle/_ref).  Bit-exact: down- and upscaling, ragged source widths (SSE2 4-column groups + scalar tail), one-pixel outputs."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)


def planes3(H, W, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    base = 20000 + 15000 * np.sin(0.11 * x) * np.cos(0.07 * y) + 9000 * ((x // 13 + y // 9) % 2)
    return [np.ascontiguousarray(base * k + rng.normal(0, 700, (H, W)), dtype=np.float32) for k in (1.0, -0.3, 0.45)]


def lanczos(lib, name, src, dW, dH, scale):
    sH, sW = src[0].shape
    dst = [np.zeros((dH, dW), np.float32) for _ in range(3)]
    rc = getattr(lib, name)(src[0].ctypes.data_as(fp), src[1].ctypes.data_as(fp), src[2].ctypes.data_as(fp), sW, sH,
                            dst[0].ctypes.data_as(fp), dst[1].ctypes.data_as(fp), dst[2].ctypes.data_as(fp), dW, dH, ctypes.c_float(scale))
    assert rc == 0
    return dst


# (sW, sH, scale): the destination size follows resizeScale's rounding (ipresize.cc L296-297: int(w * scale + 0.5))
CASES = [(640, 427, 0.5), (641, 430, 0.3), (203, 131, 0.77), (97, 64, 0.11), (120, 80, 1.0), (90, 61, 1.6), (33, 17, 3.0), (1000, 7, 0.25),
         (5, 300, 0.4), (16, 16, 0.0625)]


@needs_ref
@pytest.mark.parametrize("sW,sH,scale", CASES)
def test_port_matches_reference(sW, sH, scale):
    src = planes3(sH, sW, sW * 7 + sH)
    dW, dH = max(1, int(sW * scale + 0.5)), max(1, int(sH * scale + 0.5))
    got = lanczos(oracle.port().lib, "artoracle_lanczos", src, dW, dH, scale)
    want = lanczos(oracle.ref().lib, "artref_lanczos", src, dW, dH, scale)
    for g, w, ch in zip(got, want, "012"):
        assert np.array_equal(g, w), "plane %s: %d of %d differ, max abs %g" % (ch, int((g != w).sum()), g.size, float(np.abs(g - w).max()))


def test_lanczos_properties():
    """Size-independent: weights are normalised, so a constant image stays constant (to rounding) and scale 1 is the identity."""
    c = [np.full((50, 70), v, np.float32) for v in (1234.5, -77.25, 40000.0)]
    out = lanczos(oracle.port().lib, "artoracle_lanczos", c, 35, 25, 0.5)
    for o, p in zip(out, c):
        assert np.allclose(o, p[0, 0], rtol=3e-6)
    src = planes3(40, 60, 3)
    out = lanczos(oracle.port().lib, "artoracle_lanczos", src, 60, 40, 1.0)
    for o, p in zip(out, src):
        assert np.allclose(o, p, rtol=1e-5, atol=1e-2)
