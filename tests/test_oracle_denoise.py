"""Pin the RGB_denoise oracle (oracle/denoise_port.c) against the reference's own RGB_denoise compiled in place
(oracle/_ref; both over the same DCT stand-in, since fftw3f is absent: parity is unpinned at the FFTW boundary).
Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
F = ctypes.c_float

PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)


def rgb_frame(H, W, seed, noise=1500.0, hot=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    base = 20000 + 15000 * np.sin(0.05 * x) * np.cos(0.04 * y) + 8000 * ((x // 31 + y // 19) % 2)
    planes = [np.clip(base * k + rng.normal(0, noise, (H, W)), 0, 65535).astype(np.float32) for k in (0.8, 1.0, 0.6)]
    if hot:                                   # saturated colour patch: exercises c_h > 3000 and values above the LUT range
        planes[0][: H // 4, : W // 4] = 70000.0
        planes[2][: H // 4, : W // 4] = 300.0
    return planes


def noise_ccurve():
    """A plausible NoiseCurve LUT (501 entries, floor 0.01 as NoiseCurve::Set enforces, ipdenoise.cc L691-701)."""
    x = np.arange(501, dtype=np.float64) / 500.0
    v = 0.05 + 0.45 * np.clip((0.35 - x) / 0.30, 0, 1) ** 2
    v = np.maximum(v, 0.01).astype(np.float32)
    return v, float(np.float32(v.sum(dtype=np.float32)))


def calclum_of(planes):
    return [np.ascontiguousarray(p[::2, ::2]) for p in planes]


PROPHOTO_INV = np.array([[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0, 0.0, 1.2118128]], np.float64)   # iccmatrices.h prophoto_xyz


def run(lib, fname, planes, params, ccurve=None, with_inverse=False, aggressive=None, lab=None):
    H, W = planes[0].shape
    out = [p.copy() for p in planes]
    p = np.array(params, np.float64)
    wp = PROPHOTO.copy()
    wpi = PROPHOTO_INV.copy() if lab is not None else np.linalg.inv(wp)
    res = np.zeros(2, np.float32)
    args = [out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, p.ctypes.data_as(dp), wp.ctypes.data_as(dp)]
    if with_inverse:
        args.append(wpi.ctypes.data_as(dp))
    if ccurve is not None:
        lut, s = ccurve
        cl = calclum_of(planes)
        args += [lut.ctypes.data_as(fp), F(s), cl[0].ctypes.data_as(fp), cl[1].ctypes.data_as(fp), cl[2].ctypes.data_as(fp)]
    else:
        args += [None, F(0), None, None, None]
    args.append(res.ctypes.data_as(fp))
    if aggressive is not None:          # the *_ex entries: DenoiseParams::aggressive
        args.append(int(aggressive))
    if lab is not None:                 # the *_ex2 entries: DenoiseParams::colorSpace == LAB
        args.append(int(lab))
    assert getattr(lib, fname)(*args) == 0
    return out, res


CASES = [
    # W, H, (luminance, detail, detail_thresh, chroma, chromaRG, chromaBY, gamma, scale), curve, hot
    (160, 120, (30, 50, 0, 15, 0, 0, 1.7, 1.0), False, False),
    (203, 131, (30, 50, 0, 15, 0, 0, 1.7, 1.0), True, True),
    (131, 203, (60, 20, 40, 40, 10, -20, 1.3, 1.0), True, False),      # detail mask, uneven chroma sliders
    (300, 260, (0, 50, 0, 25, 0, 0, 1.7, 1.0), True, False),           # chroma only: no luminance shrink, no DCT
    (300, 260, (45, 80, 0, 0, 0, 0, 1.0, 1.0), False, False),          # gamma 1: LUTs bypassed; chroma 0 -> 0.001
    (260, 300, (30, 50, 0, 90, 60, 60, 1.7, 2.0), True, True),         # 8 wavelet levels asked, scale 2
]


@needs_ref
@pytest.mark.parametrize("W,H,params,curve,hot", CASES)
def test_rgb_denoise(W, H, params, curve, hot):
    planes = rgb_frame(H, W, seed=W * 7 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    a, ra = run(oracle.port().lib, "artoracle_rgb_denoise", planes, params, cc)
    b, rb = run(oracle.ref().lib, "artref_rgb_denoise", planes, params, cc, with_inverse=True)
    for name, x, y in zip("rgb", a, b):
        assert np.isfinite(y).all()
        assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
    assert np.array_equal(ra, rb), (ra, rb)
    assert not np.array_equal(b[1], planes[1])


@needs_ref
@pytest.mark.parametrize("W,H,params,curve,hot", CASES)
def test_rgb_denoise_aggressive(W, H, params, curve, hot):
    """DenoiseParams::aggressive (QUALITY_HIGH): two more wavelet levels, BiShrink for the chroma channels, qhighFactor 1 / 0.9"""
    planes = rgb_frame(H, W, seed=W * 5 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    a, ra = run(oracle.port().lib, "artoracle_rgb_denoise_ex", planes, params, cc, aggressive=1)
    b, rb = run(oracle.ref().lib, "artref_rgb_denoise_ex", planes, params, cc, with_inverse=True, aggressive=1)
    for name, x, y in zip("rgb", a, b):
        assert np.isfinite(y).all()
        assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
    assert np.array_equal(ra, rb), (ra, rb)
    std, _ = run(oracle.ref().lib, "artref_rgb_denoise", planes, params, cc, with_inverse=True)
    assert any(not np.array_equal(x, y) for x, y in zip(b, std)), "aggressive mode changed nothing"


@needs_ref
@pytest.mark.parametrize("W,H,params,curve,hot", CASES)
@pytest.mark.parametrize("aggressive", [0, 1])
def test_rgb_denoise_lab_colour_space(W, H, params, curve, hot, aggressive):
    """DenoiseParams::colorSpace == LAB: denoiseIGammaTab + Color::rgb2lab on the way in, Color::lab2rgb + denoiseGammaTab on the way out"""
    planes = rgb_frame(H, W, seed=W * 3 + H, hot=hot)
    cc = noise_ccurve() if curve else None
    a, ra = run(oracle.port().lib, "artoracle_rgb_denoise_ex2", planes, params, cc, with_inverse=True, aggressive=aggressive, lab=1)
    b, rb = run(oracle.ref().lib, "artref_rgb_denoise_ex2", planes, params, cc, with_inverse=True, aggressive=aggressive, lab=1)
    for name, x, y in zip("rgb", a, b):
        assert np.isfinite(y).all()
        assert np.array_equal(x, y), "%s: %d of %d differ, max %g" % (name, int((x != y).sum()), x.size, float(np.abs(x - y).max()))
    assert np.array_equal(ra, rb), (ra, rb)
    std, _ = run(oracle.ref().lib, "artref_rgb_denoise_ex", planes, params, cc, with_inverse=True, aggressive=aggressive)
    assert any(not np.array_equal(x, y) for x, y in zip(b, std)), "the colour space changed nothing"


@needs_ref
def test_gamma_lut_matches_reference_use():
    """The gamma LUT is only observable through RGB_denoise; pin a value-by-value property instead: monotone, and
    continuous across the three construction ranges."""
    lut = np.zeros(65536, np.float32)
    lib = oracle.port().lib
    gam, th = np.float32(1.7), np.float32(0.001)
    slope = np.float32(np.exp(np.log(float(th)) / float(gam)) / float(th))
    lib.artoracle_gammaf2lut(lut.ctypes.data_as(fp), F(gam), F(th), F(slope), F(65535.0), F(65535.0))
    assert np.all(np.diff(lut) > 0)
    assert abs(lut[65535] - 65535.0) < 1.0


def test_block_dct_standin_matches_scipy_pocketfft():
    """A second opinion on the FFTW boundary of detail_recovery (FTblockDN.cc L1604, L1614: 64 x 64 REDFT10 / REDFT01, fftw3f absent): the fp64
    cosine-sum stand-in against scipy's pocketfft DCT-II / DCT-III, in double (definition) and in float32 (what a float FFT library returns)."""
    import ctypes
    import scipy.fft
    fp_ = ctypes.POINTER(ctypes.c_float)
    lib = oracle.port().lib
    rng = np.random.default_rng(12)
    for n in (64, 8):
        for kind, tp in ((0, 2), (1, 3)):
            a = (rng.standard_normal((n, n)) * 300).astype(np.float32)
            out = np.zeros_like(a)
            assert lib.artoracle_block_dct(n, kind, a.ctypes.data_as(fp_), out.ctypes.data_as(fp_)) == 0
            want64 = scipy.fft.dctn(a.astype(np.float64), type=tp)
            want32 = scipy.fft.dctn(a, type=tp)
            assert want32.dtype == np.float32
            scale = np.abs(want64).max()
            assert np.abs(out - want64).max() <= 1e-7 * scale          # the stand-in is the definition rounded once to float
            assert np.abs(out - want32).max() <= 4e-6 * scale          # a float32 FFT sits a few ulp of the largest coefficient away
