"""Pin the box-blur / guided-filter oracle (oracle/guided_port.c) against the reference's own boxblur.h and
guidedfilter.cc compiled in place (oracle/_ref).  Bit-exact."""
import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")


def image(H, W, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 0.45 + 0.3 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, 0.03, size=(H, W))
    img[H // 3:H // 2, W // 4:W // 2] += 0.2
    return (np.clip(img, 0, 1) * scale).astype(np.float32)


@needs_ref
@pytest.mark.parametrize("radius", [0, 1, 2, 5, 11])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95)])
@pytest.mark.parametrize("inplace", [False, True])
def test_boxblur_port_matches_reference(radius, W, H, inplace):
    img = image(H, W, seed=W * H, scale=65535.0)
    got = oracle.port().boxblur(img, radius, inplace)
    want = oracle.ref().boxblur(img, radius, inplace)
    n = int((got != want).sum())
    assert n == 0, "%d samples differ, max abs %g" % (n, float(np.abs(got - want).max()))


@needs_ref
@pytest.mark.parametrize("W,H,r,sub", [(200, 150, 4, 0), (200, 150, 1, 0), (700, 640, 10, 0), (700, 640, 7, 0),
                                       (333, 257, 6, 2), (640, 480, 9, 3), (801, 603, 8, 4)])
@pytest.mark.parametrize("eps", [1e-4, 0.01])
def test_guided_port_matches_reference(W, H, r, sub, eps):
    guide = image(H, W, seed=W + H)
    src = image(H, W, seed=W * 3 + H) * 0.8 + 0.1 * guide
    got = oracle.port().guided_filter(guide, src, r, eps, sub)
    want = oracle.ref().guided_filter(guide, src, r, eps, sub)
    n = int((got != want).sum())
    assert n == 0, "%d samples differ, max abs %g" % (n, float(np.abs(got - want).max()))
