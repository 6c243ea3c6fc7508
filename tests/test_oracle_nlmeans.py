"""Pin the NL-means oracle (oracle/nlmeans_port.c: detail_mask, laplacian, rescaleBilinear, NLMeans) against the
reference's own functions compiled in place (oracle/_ref).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
F = ctypes.c_float


def luminance(H, W, seed, noise=900.0, dark=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 18000 + 14000 * np.sin(0.045 * x) * np.cos(0.06 * y) + 9000 * ((x // 23 + y // 17) % 2) + rng.normal(0, noise, size=(H, W))
    if dark:
        img[: H // 3] *= 1e-4          # near-black region: exercises the flush-to-zero arithmetic
        img[: H // 5, : W // 2] = 0
    return np.clip(img, 0, 65535).astype(np.float32)


def detail_mask(lib, fname, src, scaling, threshold, ceiling, factor, blur_type, blur):
    H, W = src.shape
    src = np.ascontiguousarray(src)
    out = np.zeros_like(src)
    rc = getattr(lib, fname)(src.ctypes.data_as(fp), out.ctypes.data_as(fp), W, H, F(scaling), F(threshold), F(ceiling), F(factor), blur_type, F(blur))
    assert rc == 0
    return out


def nlmeans(lib, fname, img, normcoeff, strength, detail_thresh, scale):
    H, W = img.shape
    out = np.ascontiguousarray(img).copy()
    rc = getattr(lib, fname)(out.ctypes.data_as(fp), W, H, F(normcoeff), strength, detail_thresh, F(scale))
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (131, 97), (7, 40), (203, 77), (322, 251)])
@pytest.mark.parametrize("blur_type,blur", [(2, 2.0), (2, 0.5), (1, 2.0), (0, 0.0)])
def test_detail_mask(W, H, blur_type, blur):
    img = luminance(H, W, seed=W * 3 + H)
    a = detail_mask(oracle.port().lib, "artoracle_detail_mask", img, 65535.0, 65.535, 65535.0, 0.3, blur_type, blur)
    b = detail_mask(oracle.ref().lib, "artref_detail_mask", img, 65535.0, 65.535, 65535.0, 0.3, blur_type, blur)
    assert np.array_equal(a, b), "%d of %d differ" % (int((a != b).sum()), a.size)


@needs_ref
@pytest.mark.parametrize("W,H,strength,detail,scale,dark", [
    (64, 48, 50, 50, 1.0, False),
    (131, 97, 100, 0, 1.0, False),
    (150, 150, 20, 80, 1.0, True),
    (283, 161, 70, 30, 1.0, True),      # several tiles, ragged last tile, vector and scalar columns
    (283, 161, 70, 30, 2.0, False),     # scale 2: search radius 3, patch radius 1
    (300, 290, 35, 100, 1.5, False),
])
def test_nlmeans(W, H, strength, detail, scale, dark):
    img = luminance(H, W, seed=W + 5 * H + strength, dark=dark)
    a = nlmeans(oracle.port().lib, "artoracle_nlmeans", img, 65535.0, strength, detail, scale)
    b = nlmeans(oracle.ref().lib, "artref_nlmeans", img, 65535.0, strength, detail, scale)
    assert np.isfinite(b).all()
    assert np.array_equal(a, b), "%d of %d differ, max %g" % (int((a != b).sum()), a.size, float(np.abs(a - b).max()))
    assert not np.array_equal(b, img)


@needs_ref
def test_nlmeans_strength_zero_is_identity():
    img = luminance(40, 40, 1)
    a = nlmeans(oracle.port().lib, "artoracle_nlmeans", img, 65535.0, 0, 50, 1.0)
    assert np.array_equal(a, img)
