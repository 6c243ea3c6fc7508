"""Pin the Fattal oracle (oracle/fattal_port.c) against the reference's own tmo_fattal02.cc compiled in place
(oracle/_ref, shim_fattal.cc).  Bit-exact: both sides run the same double-precision REDFT00 stand-in
(oracle/dct_standin.h; fftw3f is absent, parity unpinned at that boundary) and the same libm powf."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
F = ctypes.c_float
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)


def scene(H, W, seed, dark_frac=0.25):
    """working-space RGB 0..65535 with a wide dynamic range, deep shadows (below the 65.535 median floor) and noise"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    base = 2500.0 * np.exp(2.6 * np.sin(x / 37.0) * np.cos(y / 23.0)) + 40.0
    base[: int(H * dark_frac)] *= 0.004
    base[:, : W // 9] = 0.2
    out = []
    for gain in (1.0, 0.9, 0.55):
        out.append(np.clip(base * gain * rng.uniform(0.7, 1.3, (H, W)), 0, 65535).astype(np.float32))
    return out


def fattal(lib, fname, planes, threshold, amount, sat):
    r, g, b = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = r.shape
    rc = getattr(lib, fname)(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, threshold, amount, sat,
                             PROPHOTO.ctypes.data_as(dp))
    assert rc == 0
    return r, g, b


def test_standin_matches_cosine_sum():
    lib = oracle.port().lib
    rng = np.random.default_rng(7)
    import scipy.fft
    for n0, n1 in ((5, 9), (34, 21), (66, 131), (3, 3)):
        a = rng.standard_normal((n0, n1)).astype(np.float32)
        out = np.zeros_like(a)
        lib.artoracle_redft00_2d(n0, n1, a.ctypes.data_as(fp), out.ctypes.data_as(fp))
        want = scipy.fft.dctn(a.astype(np.float64), type=1)
        assert np.abs(out - want).max() <= 2e-7 * np.abs(want).max()


@needs_ref
def test_standin_1d_matches_literal_definition():
    lib = oracle.ref().lib
    rng = np.random.default_rng(8)
    for n in (3, 4, 5, 17, 34, 66, 131, 209):
        x = rng.standard_normal(n)
        a, b = np.zeros(n), np.zeros(n)
        lib.artref_redft00_1d(n, x.ctypes.data_as(dp), a.ctypes.data_as(dp))
        lib.artref_redft00_1d_naive(n, x.ctypes.data_as(dp), b.ctypes.data_as(dp))
        assert np.abs(a - b).max() < 1e-11 * max(1.0, np.abs(b).max())


@needs_ref
def test_find_fast_dim():
    p, r = oracle.port().lib, oracle.ref().lib
    for d in list(range(1, 700)) + [1919, 1920, 1921, 4000, 5464, 6240, 8192, 8193, 12288, 20000, 32767]:
        assert p.artoracle_find_fast_dim(d) == r.artref_find_fast_dim(d)
    assert p.artoracle_find_fast_dim(5464) == 5632 and p.artoracle_find_fast_dim(8192) == 8192


@needs_ref
@pytest.mark.parametrize("mtype", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("use_upper", [0, 1])
@pytest.mark.parametrize("W,H", [(37, 29), (8, 9), (131, 64)])
def test_median_denoise(mtype, use_upper, W, H):
    rng = np.random.default_rng(W + 10 * mtype)
    src = rng.uniform(1, 200, (H, W)).astype(np.float32)
    a = np.zeros_like(src)
    b = np.zeros_like(src)
    assert oracle.port().lib.artoracle_median_denoise(src.ctypes.data_as(fp), a.ctypes.data_as(fp), F(65.535), use_upper, W, H, mtype) == 0
    s2 = src.copy()
    assert oracle.ref().lib.artref_median_denoise(s2.ctypes.data_as(fp), b.ctypes.data_as(fp), F(65.535), use_upper, W, H, mtype, 1) == 0
    assert np.array_equal(a, b)
    # in place, the way tone mapping calls it
    s3 = src.copy()
    oracle.ref().lib.artref_median_denoise(s3.ctypes.data_as(fp), s3.ctypes.data_as(fp), F(65.535), use_upper, W, H, mtype, 1)
    assert np.array_equal(a, s3)


@needs_ref
@pytest.mark.parametrize("W,H,detail", [(65, 49, 3), (97, 81, 0), (161, 129, 3), (41, 321, 2), (17, 9, 3)])
def test_tmo_fattal02(W, H, detail):
    rng = np.random.default_rng(W)
    Y = (np.exp(rng.uniform(0, 10, (H, W))) + 1).astype(np.float32)
    a, b = Y.copy(), Y.copy()
    oracle.port().lib.artoracle_tmo_fattal02(W, H, a.ctypes.data_as(fp), a.ctypes.data_as(fp), F(1.3), F(0.94), F(0.013), detail)
    oracle.ref().lib.artref_tmo_fattal02(W, H, b.ctypes.data_as(fp), F(1.3), F(0.94), F(0.013), detail)
    assert np.isfinite(a).all()
    assert np.array_equal(a, b), "%d of %d differ, max rel %g" % (int((a != b).sum()), a.size, float(np.abs(a / b - 1).max()))


@needs_ref
@pytest.mark.parametrize("W,H,threshold,amount,sat", [
    (300, 200, 30, 20, 0),
    (301, 203, 30, 20, 1),
    (203, 301, -50, 80, 1),        # portrait, negative threshold
    (640, 480, 0, 100, 0),
    (97, 64, 100, 1, 1),
    (2100, 1400, 30, 20, 1),       # beyond RT_dimension_cap: bilinear down/up of H and FI, 5x5 soft median
])
def test_fattal(W, H, threshold, amount, sat):
    planes = scene(H, W, seed=W + H)
    a = fattal(oracle.port().lib, "artoracle_fattal", planes, threshold, amount, sat)
    b = fattal(oracle.ref().lib, "artref_fattal", planes, threshold, amount, sat)
    for p, q, ch in zip(a, b, "RGB"):
        assert np.isfinite(p).all()
        assert np.array_equal(p, q), "%s: %d of %d differ" % (ch, int((p != q).sum()), p.size)
    assert not np.array_equal(a[0], planes[0])


@needs_ref
def test_fattal_large_medians():
    """r >= 2 and r >= 3 select the strong 5x5 and the 7x7 medians (tmo_fattal02.cc L1104-1114); thin strips keep it cheap"""
    for W, H in ((3900, 40), (60, 5800)):
        planes = scene(H, W, seed=W)
        a = fattal(oracle.port().lib, "artoracle_fattal", planes, 30, 20, 0)
        b = fattal(oracle.ref().lib, "artref_fattal", planes, 30, 20, 0)
        for p, q in zip(a, b):
            assert np.array_equal(p, q)
