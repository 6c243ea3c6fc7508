"""GPU parity for the float box blur and the guided filter (through the C-ABI) against the oracle.  Bit-exact."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def image(H, W, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 0.45 + 0.3 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, 0.03, size=(H, W))
    img[H // 3:H // 2, W // 4:W // 2] += 0.2
    return (np.clip(img, 0, 1) * scale).astype(np.float32)


@pytest.mark.parametrize("radius", [0, 1, 2, 5, 11])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (1003, 517)])
@pytest.mark.parametrize("inplace", [False, True])
def test_boxblur(hot_path, radius, W, H, inplace):
    img = image(H, W, seed=W * H, scale=65535.0)
    want = oracle.port().boxblur(img, radius, inplace)
    if inplace:
        got = img.copy()
        hot_path.boxblur(got, radius, dst=got)
    else:
        got = hot_path.boxblur(img, radius)
    assert np.array_equal(got, want), "%d samples differ" % int((got != want).sum())


@pytest.mark.parametrize("W,H,r,sub", [(200, 150, 4, 0), (200, 150, 1, 0), (700, 640, 10, 0), (700, 640, 7, 0),
                                       (333, 257, 6, 2), (640, 480, 9, 3), (801, 603, 8, 4), (2048, 1536, 15, 0)])
@pytest.mark.parametrize("eps", [1e-4, 0.01])
def test_guided_filter(hot_path, W, H, r, sub, eps):
    guide = image(H, W, seed=W + H)
    src = image(H, W, seed=W * 3 + H) * 0.8 + 0.1 * guide
    want = oracle.port().guided_filter(guide, src, r, eps, sub)
    got = hot_path.guided_filter(guide, src, r, eps, sub)
    assert np.array_equal(got, want), "%d samples differ, max abs %g" % (int((got != want).sum()), float(np.abs(got - want).max()))
