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


def lum_scene(W, H, seed, kind):
    """L-like plane: texture + flat patches in the accepted luminance range.  kind 0: a very flat patch (pass 0 finds it); 1: only
    moderately flat patches (pass 1 + the fine scan); 2: everything busy or out of range (threshold 0)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 9000 + 5000 * np.sin(0.21 * x) * np.cos(0.17 * y) + rng.normal(0, 900, (H, W))
    if kind == 0:
        ph, pw = min(170, H - H // 4), min(170, W - W // 3)
        img[H // 4: H // 4 + ph, W // 3: W // 3 + pw] = 8000 + rng.normal(0, 80, (ph, pw))
    elif kind == 1:
        img[H // 2: H // 2 + 47, W // 5: W // 5 + 52] = 6000 + rng.normal(0, 170, (47, 52))
    else:
        img[: H // 2] = 30000 + rng.normal(0, 10, (H // 2, W))
    return np.ascontiguousarray(img, dtype=np.float32)


@needs_ref
@pytest.mark.parametrize("W,H", [(400, 300), (333, 251), (163, 170)])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_auto_contrast_threshold(W, H, kind):
    lum = lum_scene(W, H, W + kind, kind)
    f = oracle.port().lib.artoracle_auto_contrast_threshold
    f.restype = ctypes.c_float
    got = f(P(lum), W, H, F(0.2), F(1.0))
    thr = ctypes.c_float(0.2)
    blend = np.zeros((H, W), np.float32)
    assert oracle.ref().lib.artref_blend_mask_ex(P(lum), P(blend), W, H, ctypes.byref(thr), F(1.0), 1, F(2.0)) == 0
    assert got == thr.value, (got, thr.value)
    if kind == 2:
        assert got == 0.0
    else:
        assert 0.0 < got < 1.0
    # and the mask built with that threshold is the one the reference built
    mine = np.zeros((H, W), np.float32)
    assert oracle.port().lib.artoracle_blend_mask(P(lum), P(mine), W, H, F(got), F(1.0), F(2.0)) == 0
    assert np.array_equal(mine, blend)


@needs_ref
@pytest.mark.parametrize("filters", [0x94949494, 0x49494949])
@pytest.mark.parametrize("W,H", [(301, 203), (640, 427)])
def test_dual_bilinear_auto_contrast(filters, W, H):
    """autoContrast (the reference's default for the dual methods): the threshold comes from the flattest tile of the frame's L."""
    raw = synth.bayer_frame(W, H, filters, seed=W)
    planes = list(oracle.port().amaze(raw, filters))
    lib = oracle.port().lib
    r, g, b = [p.copy() for p in planes]
    c = ctypes.c_double(20.0)
    blend = np.zeros((H, W), np.float32)
    assert lib.artoracle_dual_bilinear_ex(P(raw), W, H, ctypes.c_uint(filters), P(r), P(g), P(b), ctypes.byref(c), 1, P(blend)) == 0
    # the reference's pieces, chained as dual_demosaic_RT.cc L101-127 chains them
    rl = oracle.ref().lib
    rr, rg, rb = [p.copy() for p in planes]
    L = np.zeros((H, W), np.float32)
    rl.artref_rgb2l(P(rr), P(rg), P(rb), P(L), W, H, P(XYZ_RGB))
    thr = ctypes.c_float(np.float32(20.0 / 100.0))
    rblend = np.zeros((H, W), np.float32)
    rl.artref_blend_mask_ex(P(L), P(rblend), W, H, ctypes.byref(thr), F(1.0), 1, F(2.0))
    rl.artref_bilinear_blend(P(raw), P(rblend), W, H, ctypes.c_uint(filters), P(rr), P(rg), P(rb))
    assert c.value == float(np.float32(thr.value) * np.float32(100.0))
    assert np.array_equal(blend, rblend)
    for x, y in zip((r, g, b), (rr, rg, rb)):
        assert np.array_equal(x, y)


@needs_ref
@pytest.mark.parametrize("dy,dx", [(0, 0), (2, 5), (4, 1)])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (131, 140), (17, 40)])
def test_xtrans_fast_blend(dy, dx, W, H):
    """The X-Trans half of the dual demosaic: fast_xtrans_interpolate_blend over arbitrary first-demosaicer planes and blend mask."""
    xt = np.ascontiguousarray(synth.xtrans_matrix(dy, dx), np.int32)
    raw = synth.xtrans_frame(W, H, xt, seed=W + dy)
    rng = np.random.default_rng(H + dx)
    planes = [np.ascontiguousarray(rng.uniform(0, 65535, (H, W)), dtype=np.float32) for _ in range(3)]
    blend = np.ascontiguousarray(rng.uniform(0, 1, (H, W)), dtype=np.float32)
    blend[::5, ::3] = 1.0
    blend[1::7, ::2] = 0.0
    ip = ctypes.POINTER(ctypes.c_int)
    a = [p.copy() for p in planes]
    b = [p.copy() for p in planes]
    assert oracle.port().lib.artoracle_xtrans_fast_blend(W, H, xt.ctypes.data_as(ip), P(raw), P(blend), P(a[0]), P(a[1]), P(a[2])) == 0
    assert oracle.ref().lib.artref_xtrans_fast_blend(W, H, xt.ctypes.data_as(ip), P(raw), P(blend), P(b[0]), P(b[1]), P(b[2])) == 0
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    if W > 16 and H > 16:
        assert any((x != p).any() for x, p in zip(a, planes))
        assert all(np.array_equal(x[:8], p[:8]) and np.array_equal(x[:, :8], p[:, :8]) for x, p in zip(a, planes))


@needs_ref
@pytest.mark.parametrize("filters,pf", [(0x94949494, 0xb4b4b4b4), (0x61616161, 0xe1e1e1e1)])
@pytest.mark.parametrize("auto", [0, 1])
def test_dual_vng4(filters, pf, auto):
    """AMAZEVNG4: the flat-region planes come from VNG4; chained from the reference's own pieces as dual_demosaic_RT.cc L101-147 chains them."""
    W, H = 301, 203
    raw = synth.bayer_frame(W, H, filters, seed=77)
    planes = list(oracle.port().amaze(raw, filters))
    r, g, b = [p.copy() for p in planes]
    c = ctypes.c_double(20.0)
    blend = np.zeros((H, W), np.float32)
    assert oracle.port().lib.artoracle_dual_vng4(P(raw), W, H, ctypes.c_uint(pf), P(r), P(g), P(b), ctypes.byref(c), auto, P(blend)) == 0
    rl = oracle.ref().lib
    L = np.zeros((H, W), np.float32)
    rl.artref_rgb2l(P(planes[0]), P(planes[1]), P(planes[2]), P(L), W, H, P(XYZ_RGB))
    thr = ctypes.c_float(np.float32(0.2))
    rblend = np.zeros((H, W), np.float32)
    rl.artref_blend_mask_ex(P(L), P(rblend), W, H, ctypes.byref(thr), F(1.0), auto, F(2.0))
    tmp = [np.zeros((H, W), np.float32) for _ in range(3)]
    rl.artref_vng4(W, H, ctypes.c_uint(pf), P(raw), P(tmp[0]), P(tmp[1]), P(tmp[2]), 1)   # one thread: the stock VNG4 races with its own border pass (test_oracle_vng4.py)
    assert np.array_equal(blend, rblend)
    one = np.float32(1.0)
    for mine, first, flat in zip((r, g, b), planes, tmp):
        want = rblend * first + (one - rblend) * flat           # intp(blend, first, flat), float32 throughout
        assert np.array_equal(mine, want)
