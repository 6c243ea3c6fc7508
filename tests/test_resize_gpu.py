"""GPU parity for the Lanczos resampler (art_hp_resize_lanczos = ImProcFunctions::Lanczos, rtengine/ipresize.cc L38-207) through the C-ABI
against the oracle port, which tests/test_oracle_resize.py pins bit-exact to the reference's own function compiled in place.  The weights
(sleef sine, IEEE division, serial normalisation) and both tap sums are reproduced in the reference's order: bit-exact."""
import time

import numpy as np
import pytest

import art_b200
import oracle
from test_oracle_resize import CASES, lanczos, planes3

pytestmark = pytest.mark.gpu

MORE = [(1283, 857, 0.5), (1920, 1280, 0.333), (801, 603, 1.25), (2048, 64, 0.0625), (64, 2048, 0.031), (350, 233, 2.0), (4, 3, 0.5), (1, 1, 1.0),
        (700, 500, 0.02)]


@pytest.mark.parametrize("sW,sH,scale", CASES + MORE)
def test_lanczos_matches_oracle(hot_path, sW, sH, scale):
    src = planes3(sH, sW, sW * 7 + sH)
    dW, dH = max(1, int(sW * scale + 0.5)), max(1, int(sH * scale + 0.5))
    want = lanczos(oracle.port().lib, "artoracle_lanczos", src, dW, dH, scale)
    got = hot_path.resize_lanczos(src, scale)
    for g, w, ch in zip(got, want, "012"):
        assert g.shape == w.shape
        assert np.array_equal(g, w), "plane %s: %d of %d differ, max abs %g" % (ch, int((g != w).sum()), g.size, float(np.abs(g - w).max()))


def test_destination_size_is_the_callers(hot_path):
    """The reference passes any destination size it likes (resizeScale rounds, crops change it): taps are clipped to the source."""
    src = planes3(120, 180, 5)
    want = lanczos(oracle.port().lib, "artoracle_lanczos", src, 61, 47, 0.4)
    got = hot_path.resize_lanczos(src, 0.4, size=(47, 61))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_full_frame_downscale_properties(hot_path):
    """configs[1]'s frame (8192 x 5464) to half size: constant planes stay constant to rounding (normalised weights), a checksum of the
    device result on a 512-row slab equals the oracle's on the same rows (the slab carries its taps' reach), and the timing is printed."""
    sW, sH, scale = 8192, 5464, 0.5
    rng = np.random.default_rng(3)
    y, x = np.mgrid[0:sH, 0:sW].astype(np.float32)
    base = (20000 + 15000 * np.sin(0.011 * x) * np.cos(0.007 * y)).astype(np.float32)
    src = [np.ascontiguousarray(base * k + rng.normal(0, 500, (sH, sW)).astype(np.float32)) for k in (1.0, -0.3, 0.45)]
    hot_path.resize_lanczos(src, scale)
    t0 = time.perf_counter()
    got = hot_path.resize_lanczos(src, scale)
    dt = time.perf_counter() - t0
    print("\n[lanczos] %dx%d -> %dx%d through the host entry (pageable memory, copies included): %.1f ms" % (sW, sH, got[0].shape[1], got[0].shape[0], dt * 1e3))
    # rows [0, 200) of the output depend on source rows [0, 2 * 200 + 6): compare that corner exactly against the oracle on a cropped source
    crop = [np.ascontiguousarray(p[:1024]) for p in src]
    want = lanczos(oracle.port().lib, "artoracle_lanczos", crop, sW // 2, 512, scale)
    for g, w in zip(got, want):
        assert np.array_equal(g[:500], w[:500])
    c = [np.full((600, 800), v, np.float32) for v in (1234.5, -77.25, 40000.0)]
    for o, p in zip(hot_path.resize_lanczos(c, 0.37), c):
        assert np.allclose(o, p[0, 0], rtol=3e-6)


def test_rejects_bad_arguments(hot_path):
    src = planes3(20, 30, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.resize_lanczos(src, 0.0)
    with pytest.raises(art_b200.HotPathError):
        hot_path.resize_lanczos(src, 0.0004, size=(1, 1))
