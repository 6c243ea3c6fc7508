"""Pin the tone equalizer oracle (oracle/toneeq_port.c) against the reference's own tone_eq() compiled in place (oracle/_ref, shim_tone.cc:
iptoneequalizer.cc L68-338 over the reference's guidedFilter / guidedFilterLog, sleef and LUT.h).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)


def image(H, W, seed, bright=1.0):
    """a scene spanning ~14 EV with smooth regions, texture, a blown patch (Y > 1 after the gain) and a few negative samples"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    ev = -12.0 + 13.0 * (xx / max(W - 1, 1)) + 1.5 * np.sin(yy / 9.0)
    base = 65535.0 * bright * 2.0 ** ev
    planes = [(base * rng.uniform(0.6, 1.3) * (1 + 0.2 * rng.normal(0, 1, (H, W)))).astype(np.float32) for _ in range(3)]
    for p in planes:
        p[H // 3: H // 3 + 5, W // 2: W // 2 + 9] = 65535.0 * 3.0       # Y > 1: the direct evaluation, whole SSE2 groups
        p[-1, -1] = 65535.0 * 40.0                                      # row tail above the LIM(.., 32)
    planes[0][0, ::7] *= -0.5
    return [np.ascontiguousarray(p, np.float32) for p in planes]


def run(lib, name, planes, bands, regularization, pivot, scale):
    out = [p.copy() for p in planes]
    H, W = out[0].shape
    b = np.array(bands, np.int32)
    rc = getattr(lib, name)(*[p.ctypes.data_as(fp) for p in out], W, H, PROPHOTO.ctypes.data_as(dp), b.ctypes.data_as(ip), int(regularization),
                            ctypes.c_double(pivot), ctypes.c_double(scale))
    assert rc == 0
    return out


def port_teq(planes, bands, regularization, pivot, scale):
    return run(oracle.port().lib, "artoracle_tone_equalizer", planes, bands, regularization, pivot, scale)


def same(a, b):
    for x, y, ch in zip(a, b, "RGB"):
        eq = (x == y) | (np.isnan(x) & np.isnan(y))
        assert eq.all(), "%s: %d of %d differ, first at %s: %r vs %r" % (ch, int((~eq).sum()), x.size, np.argwhere(~eq)[0], x[~eq][0], y[~eq][0])


BANDS = [(0, 0, 0, 0, 0), (40, 25, 0, -20, -35), (-100, 100, -50, 50, 100), (15, 0, 0, 0, -15)]


# regularization > 1 runs guidedFilter with radius 350 / scale (and twice that for regularization 2); f_mean clamps the window to the frame
CASES = [(96, 64, 2, 0.0, 1.0), (131, 77, 4, 0.5, 1.0), (403, 301, 3, 0.0, 1.0),
         (96, 64, 0, 0.0, 1.0), (131, 77, 0, 1.5, 1.0), (131, 77, 1, 0.0, 1.0), (403, 301, 1, -2.0, 1.0), (403, 301, 1, 0.0, 2.0),
         (403, 301, 3, 0.0, 4.0), (403, 301, 4, 1.0, 4.0), (803, 602, 2, 0.0, 4.0), (1203, 900, 3, -1.0, 1.0)]


@needs_ref
@pytest.mark.parametrize("W,H,regularization,pivot,scale", CASES)
@pytest.mark.parametrize("bands", BANDS)
def test_toneeq_port_matches_reference(W, H, bands, regularization, pivot, scale):
    planes = image(H, W, W + H + regularization)
    same(port_teq(planes, bands, regularization, pivot, scale), run(oracle.ref().lib, "artref_tone_equalizer", planes, bands, regularization, pivot, scale))


@needs_ref
def test_toneeq_lifts_shadows():
    planes = image(120, 200, 3)
    out = port_teq(planes, (80, 60, 0, 0, 0), 1, 0.0, 1.0)
    dark = planes[1] < 65535.0 * 2.0 ** -7
    assert np.median(out[1][dark] / np.maximum(planes[1][dark], 1e-3)) > 1.5
