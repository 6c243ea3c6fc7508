"""Pin the green-equilibration oracle (oracle/greeneq_port.c) against the reference's own RawImageSource::green_equilibrate(_global)
compiled in place (oracle/_ref).  Bit-exact over the four Bayer phases, ragged widths (SSE2 groups + scalar tail), constant and
per-pixel thresholds."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
FILTERS = [0x94949494, 0x16161616, 0x61616161, 0x49494949]       # RGGB, BGGR, GRBG, GBRG


def mosaic(H, W, seed, imbalance=1.04):
    """A Bayer frame whose two green phases differ by a few percent in flat areas and by a lot in a textured strip."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 6000 + 4000 * np.sin(0.045 * x) * np.cos(0.03 * y) + 2500 * ((x // 23 + y // 17) % 2) + rng.normal(0, 40, (H, W))
    img[y % 2 == 1] *= imbalance                                   # rows of the second green phase
    img[:, W // 2: W // 2 + 24] += 1500 * ((x[:, W // 2: W // 2 + 24] + y[:, W // 2: W // 2 + 24]) % 2)   # Nyquist texture: left alone
    return np.ascontiguousarray(np.clip(img, 0, 65535), dtype=np.float32)


@needs_ref
@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95), (24, 12), (301, 203)])
@pytest.mark.parametrize("border", [4, 0])
def test_global(filters, W, H, border):
    raw = mosaic(H, W, W + H)
    a, b = raw.copy(), raw.copy()
    assert oracle.port().lib.artoracle_green_equilibrate_global(a.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), border) == 0
    assert oracle.ref().lib.artref_green_equilibrate_global(b.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), border) == 0
    assert np.array_equal(a, b)
    assert (a != raw).any()


@needs_ref
@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95), (24, 12), (13, 9), (301, 203)])
@pytest.mark.parametrize("thresh", [0.01, 0.05, 0.5, "map"])
def test_local(filters, W, H, thresh):
    raw = mosaic(H, W, 3 * W + H)
    a, b = raw.copy(), raw.copy()
    tmap = None
    if thresh == "map":
        tmap = np.ascontiguousarray(np.random.default_rng(W).uniform(0.0, 0.2, (H, W)), dtype=np.float32)
    t = ctypes.c_float(0.0 if tmap is not None else thresh)
    tp = tmap.ctypes.data_as(fp) if tmap is not None else None
    assert oracle.port().lib.artoracle_green_equilibrate(a.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), t, tp) == 0
    assert oracle.ref().lib.artref_green_equilibrate(b.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), t, tp) == 0
    n = int((a != b).sum())
    assert n == 0, "%d of %d differ, first at %s" % (n, a.size, np.argwhere(a != b)[0])
    if W >= 64 and thresh != 0.01:
        assert (a != raw).any(), "nothing was equilibrated"
    # only green sites change
    yy, xx = np.nonzero(a != raw)
    fc = (filters >> ((((yy << 1) & 14) + (xx & 1)) << 1)) & 3
    assert (fc == 1).all()
