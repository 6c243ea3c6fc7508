"""Pin the colour / curve chain oracle (oracle/chain_port.c) against the reference's own pixel loops compiled in place
(oracle/_ref, shim_chain.cc): expcomp, saturationVibrance, filmlike_clip + Standard / Adobe tone curves, rgbCurves,
labAdjustments between Imagefloat::setMode(LAB) and setMode(RGB).  Bit-exact, including the 4-wide SSE2 groups versus the
scalar row tails, out-of-range pixels that push a whole group onto the scalar Lab route, and LUT extrapolation."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
F = ctypes.c_float
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
PROPHOTO_INV = np.array([[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0, 0.0, 1.2118128]], np.float64)
SIZES = [(64, 8), (67, 5), (5, 9), (3, 3), (130, 17)]


def image(H, W, seed, wild=True):
    """working-space RGB with in-range values, some negatives / overrange samples and exact ties between channels"""
    rng = np.random.default_rng(seed)
    planes = [rng.uniform(0, 65535, (H, W)).astype(np.float32) for _ in range(3)]
    if wild:
        m = rng.random((H, W))
        for p in planes:
            p[m < 0.03] *= -0.2
            p[(m > 0.95)] *= 1.7
        planes[1][m > 0.9] = planes[0][m > 0.9]            # r == g
        planes[2][(m > 0.8) & (m < 0.85)] = planes[1][(m > 0.8) & (m < 0.85)]   # g == b
        planes[0][:, : W // 5] *= 1e-3                       # deep shadows
    return planes


def curve_lut(n=65536, gamma=0.8, top=65535.0, seed=0):
    x = np.arange(n, dtype=np.float64) / (n - 1)
    y = x ** gamma * (1 + 0.05 * np.sin(7 * x + seed))
    return (np.clip(y, 0, None) * top).astype(np.float32)


def call(lib, name, planes, *args):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, *args)
    assert rc == 0
    return out


def same(a, b):
    for x, y, ch in zip(a, b, "RGB"):
        eq = (x == y) | (np.isnan(x) & np.isnan(y))
        assert eq.all(), "%s: %d of %d differ, first at %s: %r vs %r" % (ch, int((~eq).sum()), x.size, np.argwhere(~eq)[0], x[~eq][0], y[~eq][0])


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("ev,black", [(0.0, 0.0), (1.3, 0.01), (-0.7, -0.02)])
def test_expcomp(W, H, ev, black):
    planes = image(H, W, W + H)
    args = (F(np.float32(2.0 ** ev)), F(np.float32(black * 2000.0)))
    same(call(oracle.port().lib, "artoracle_chain_expcomp", planes, *args), call(oracle.ref().lib, "artref_chain_expcomp", planes, *args))


MIXERS = [(1000, 0, 0, 0, 1000, 0, 0, 0, 1000), (800, 300, -100, -50, 1100, -50, 20, -400, 1380), (-200, 600, 600, 333, 333, 334, 0, 0, -1000)]


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("mixer", MIXERS)
def test_channel_mixer(W, H, mixer):
    """ImProcFunctions::channelMixer's loop (ipchmixer.cc L200-230), RGB_MATRIX coefficients float(v) / 1000.f; NaN samples show the two clamps."""
    planes = image(H, W, W * 5 + H)
    planes[0][H // 2, ::3] = np.nan
    m = (np.array(mixer, np.float32) / np.float32(1000.0)).astype(np.float32)
    same(call(oracle.port().lib, "artoracle_chmixer", planes, m.ctypes.data_as(fp)), call(oracle.ref().lib, "artref_chmixer", planes, m.ctypes.data_as(fp)))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("sat,vib", [(30, 0), (0, 40), (-50, -30), (100, 100)])
def test_saturation(W, H, sat, vib):
    planes = image(H, W, W * 3 + H)
    args = (sat, vib, PROPHOTO.ctypes.data_as(dp))
    same(call(oracle.port().lib, "artoracle_chain_saturation", planes, *args), call(oracle.ref().lib, "artref_chain_saturation", planes, *args))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("whitept", [1.0, 2.0])
def test_tonecurve(W, H, mode, whitept):
    planes = image(H, W, W * 5 + H)
    lut = curve_lut(gamma=0.6, seed=mode)
    args = (mode, lut.ctypes.data_as(fp), F(whitept))
    same(call(oracle.port().lib, "artoracle_chain_tonecurve", planes, *args), call(oracle.ref().lib, "artref_chain_tonecurve", planes, *args))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES + [(301, 211)])
@pytest.mark.parametrize("mode", [3, 4, 5])
@pytest.mark.parametrize("whitept", [1.0, 2.0])
@pytest.mark.parametrize("gamma", [0.6, 1.4])
def test_tonecurve_other_modes(W, H, mode, whitept, gamma):
    """WEIGHTEDSTD (3), SATANDVALBLENDING (4), LUMINANCE (5) against the reference's own Apply members (curves.h L474-668)"""
    planes = image(H, W, W * 7 + H + mode)
    planes[0][0, :2] = planes[1][0, :2] = planes[2][0, :2] = 0.0          # black: LuminanceToneCurve's 0.00001 guard, rgb2hsvtc's grey branch
    lut = curve_lut(gamma=gamma, seed=mode)
    if mode == 4:
        lut[20000:20100] = np.arange(20000, 20100, dtype=np.float32)      # newLum == lum for some pixels: the early return
        planes[0][-1, :] = planes[1][-1, :] = planes[2][-1, :] = 20050.0
    args = (mode, lut.ctypes.data_as(fp), F(whitept), PROPHOTO.ctypes.data_as(dp))
    same(call(oracle.port().lib, "artoracle_chain_tonecurve_ex", planes, *args), call(oracle.ref().lib, "artref_chain_tonecurve_ex", planes, *args))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("which", [(1, 1, 1), (1, 0, 0), (0, 1, 1)])
def test_rgbcurves(W, H, which):
    planes = image(H, W, W * 7 + H)
    luts = [curve_lut(gamma=0.7 + 0.2 * i, seed=i) if w else None for i, w in enumerate(which)]
    args = tuple(l.ctypes.data_as(fp) if l is not None else None for l in luts)
    same(call(oracle.port().lib, "artoracle_chain_rgbcurves", planes, *args), call(oracle.ref().lib, "artref_chain_rgbcurves", planes, *args))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("wild", [False, True])
def test_rgb2lab_and_back(W, H, wild):
    planes = image(H, W, W * 11 + H, wild)
    args = (PROPHOTO.ctypes.data_as(dp), PROPHOTO_INV.ctypes.data_as(dp))
    a = call(oracle.port().lib, "artoracle_chain_rgb2lab", planes, *args, 0)
    b = call(oracle.ref().lib, "artref_chain_rgb2lab", planes, *args, 0)
    same(a, b)
    same(call(oracle.port().lib, "artoracle_chain_rgb2lab", a, *args, 1), call(oracle.ref().lib, "artref_chain_rgb2lab", b, *args, 1))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("chroma", [1.0, 1.35, 0.4])
def test_lab_adjustments(W, H, chroma):
    planes = image(H, W, W * 13 + H)
    lc = np.concatenate([curve_lut(32768, 0.85, 32767.0), np.array([32768.0, 32769.0], np.float32)]).astype(np.float32)
    ac = curve_lut(65536, 1.1, 65535.0, 1)
    bc = curve_lut(65536, 0.9, 65535.0, 2)
    args = (lc.ctypes.data_as(fp), ac.ctypes.data_as(fp), bc.ctypes.data_as(fp), F(chroma), PROPHOTO.ctypes.data_as(dp), PROPHOTO_INV.ctypes.data_as(dp))
    same(call(oracle.port().lib, "artoracle_chain_lab", planes, *args), call(oracle.ref().lib, "artref_chain_lab", planes, *args))


def softlight_lut(strength):
    """ImProcFunctions::softLight's table f[i] = sl(strength / 100, i), built by the reference's own code"""
    if not oracle.have_ref():      # no oracle/_ref on this machine: the same Pegtop blend in numpy -- any table serves the parity of the apply loop
        x = np.arange(65536, dtype=np.float64) / 65535.0
        v = np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.maximum(x, 1e-12) ** (1 / 2.4) - 0.055)
        v = v * v + 2 * v * v - 2 * v * v * v
        lin = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4) * 65535.0
        b = strength / 100.0
        return (b * lin + (1 - b) * x * 65535.0).astype(np.float32)
    lut = np.zeros(65536, np.float32)
    z = [np.zeros((1, 1), np.float32) for _ in range(3)]
    assert oracle.ref().lib.artref_softlight(z[0].ctypes.data_as(fp), z[1].ctypes.data_as(fp), z[2].ctypes.data_as(fp), 1, 1, int(strength), lut.ctypes.data_as(fp)) == 0
    return lut


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("strength", [1, 30, 100])
def test_softlight(W, H, strength):
    """the port's apply loop against ImProcFunctions::softLight's own (ipsoftlight.cc L44-81), over the reference-built table"""
    planes = image(H, W, W * 3 + H + strength)
    planes[0][0, 0] = np.nan
    lut = softlight_lut(strength)
    want = call(oracle.ref().lib, "artref_softlight", planes, int(strength), None)
    same(call(oracle.port().lib, "artoracle_chain_softlight", planes, lut.ctypes.data_as(fp)), want)
    assert max(float(np.nanmax(np.abs(w - p))) for w, p in zip(want, planes)) > 1.0


def bw_tables(seed, gamma=True, cast=True):
    """tables shaped like ImProcFunctions::blackAndWhite builds them (ipbw.cc L264-272, L321-341): three gamma tables, the colour cast's u / v tables"""
    x = np.arange(65536, dtype=np.float64) / 65535.0
    g = [(x ** e * 65535.0).astype(np.float32) for e in (0.8, 1.0, 1.25)] if gamma else [None, None, None]
    if cast:
        y = x ** 0.9 * 65535.0
        c = np.exp(-((x - 0.5) / 0.25) ** 2)
        ul, vl = (y * 0.08 * c * np.cos(0.7 + seed)).astype(np.float32), (y * 0.08 * c * np.sin(0.7 + seed)).astype(np.float32)
    else:
        ul = vl = None
    return g + [ul, vl]


def bw_args(mix, kcorec, tabs):
    return (PROPHOTO.ctypes.data_as(dp), F(mix[0]), F(mix[1]), F(mix[2]), F(kcorec), *[t.ctypes.data_as(fp) if t is not None else None for t in tabs])


@needs_ref
@pytest.mark.parametrize("W,H", SIZES + [(301, 211)])
@pytest.mark.parametrize("gamma,cast", [(False, False), (True, False), (False, True), (True, True)])
def test_black_and_white(W, H, gamma, cast):
    """the port against ImProcFunctions::blackAndWhite's own pixel loops (ipbw.cc L283-312, L343-362) and Imagefloat's YUV round trip"""
    planes = image(H, W, W * 5 + H + gamma)
    tabs = bw_tables(W, gamma, cast)
    args = bw_args((0.43, 0.33, 0.30), 1.06, tabs)
    same(call(oracle.port().lib, "artoracle_bw", planes, *args), call(oracle.ref().lib, "artref_bw", planes, *args))


def blue_image(H, W, seed):
    """saturated blues and other pixels with r == 0 or g == 0 (what proPhotoBlue touches), negatives (left alone), ordinary pixels"""
    planes = image(H, W, seed)
    rng = np.random.default_rng(seed + 1)
    m = rng.random((H, W))
    planes[0][m < 0.25] = 0.0
    planes[1][(m > 0.2) & (m < 0.45)] = 0.0
    planes[2][(m > 0.4) & (m < 0.5)] = 0.0
    planes[2][m > 0.97] = 65535.0
    return planes


@needs_ref
@pytest.mark.parametrize("W,H", SIZES + [(301, 211)])
def test_prophoto_blue(W, H):
    """the port against the reference's proPhotoBlue (improcfun.cc L312-357) with Color::rgb2hsv / hsv2rgb"""
    planes = blue_image(H, W, W * 3 + H)
    want = call(oracle.ref().lib, "artref_prophoto_blue", planes)
    same(call(oracle.port().lib, "artoracle_prophoto_blue", planes), want)
    assert sum(int((w != p).sum()) for w, p in zip(want, planes)) > 0
