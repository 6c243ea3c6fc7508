"""Pin the dual-demosaic oracle (oracle/dual_port.c: "<first demosaicer> + bilinear", manual contrast) against the reference's own
Color::RGB2L, buildBlendMask and bayer_bilinear_demosaic(blend, ...) compiled in place (oracle/_ref) and chained as dual_demosaic_RT
chains them.  Bit-exact, step by step and end to end behind the AMaZE / RCD ports."""
import ctypes

import numpy as np
import pytest

import oracle
from art_b200 import synth

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
F = ctypes.c_float
XYZ_RGB = np.array([0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227], np.float32)
FILTERS = [0x94949494, 0x16161616, 0x61616161, 0x49494949]


def P(a):
    return a.ctypes.data_as(fp)


def ref_chain(raw, filters, planes, contrast):
    lib = oracle.ref().lib
    H, W = raw.shape
    r, g, b = [p.copy() for p in planes]
    L = np.zeros((H, W), np.float32)
    assert lib.artref_rgb2l(P(r), P(g), P(b), P(L), W, H, P(XYZ_RGB)) == 0
    blend = np.zeros((H, W), np.float32)
    contrastf = np.float32(contrast / 100.0)
    assert lib.artref_blend_mask(P(L), P(blend), W, H, F(contrastf), F(1.0), F(2.0)) == 0
    assert lib.artref_bilinear_blend(P(raw), P(blend), W, H, ctypes.c_uint(filters), P(r), P(g), P(b)) == 0
    return [r, g, b], blend, L


def port_chain(raw, filters, planes, contrast):
    lib = oracle.port().lib
    H, W = raw.shape
    r, g, b = [p.copy() for p in planes]
    blend = np.zeros((H, W), np.float32)
    assert lib.artoracle_dual_bilinear(P(raw), W, H, ctypes.c_uint(filters), P(r), P(g), P(b), ctypes.c_double(contrast), P(blend)) == 0
    return [r, g, b], blend


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 77), (33, 95)])
def test_rgb2l(W, H):
    rng = np.random.default_rng(W)
    planes = [np.ascontiguousarray(rng.uniform(-2000, 90000, (H, W)), dtype=np.float32) for _ in range(3)]     # inside and outside [0, 65535]
    planes[1][:, : W // 2] = np.abs(planes[1][:, : W // 2]) % 60000                                              # whole groups on the LUT path
    planes[0][:, : W // 2] = np.abs(planes[0][:, : W // 2]) % 60000
    planes[2][:, : W // 2] = np.abs(planes[2][:, : W // 2]) % 60000
    a, b = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
    assert oracle.port().lib.artoracle_rgb2l(P(planes[0]), P(planes[1]), P(planes[2]), P(a), W, H, P(XYZ_RGB)) == 0
    assert oracle.ref().lib.artref_rgb2l(P(planes[0]), P(planes[1]), P(planes[2]), P(b), W, H, P(XYZ_RGB)) == 0
    assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 77), (301, 203)])
@pytest.mark.parametrize("contrast", [20.0, 3.0, 75.0])
@pytest.mark.parametrize("first", ["amaze", "rcd"])
def test_dual_bilinear(filters, W, H, contrast, first):
    raw = synth.bayer_frame(W, H, filters, seed=W + H)
    planes = list(getattr(oracle.port(), first)(raw, filters))
    got, gb = port_chain(raw, filters, planes, contrast)
    want, wb, _ = ref_chain(raw, filters, planes, contrast)
    assert np.array_equal(gb, wb)
    for x, y, ch in zip(got, want, "RGB"):
        assert np.array_equal(x, y), "%s: %d differ" % (ch, int((x != y).sum()))
    assert 0.0 < gb.min() <= gb.max() <= 1.001          # the blurred sigmoid
    assert any((x != p).any() for x, p in zip(got, planes))
