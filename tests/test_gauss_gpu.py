"""GPU parity for gaussianBlur (through the C-ABI) against the oracle.  Bit-exact on every dispatch branch."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def image(H, W, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 20000 + 15000 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, 800, size=(H, W))
    img[H // 3:H // 2, W // 4:W // 2] += 20000
    return np.clip(img, 0, 65535).astype(np.float32)


@pytest.mark.parametrize("sigma", [0.2, 0.3, 0.5, 0.6, 0.9, 2.5, 7.7, 24.9, 25.0, 40.0])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95), (1003, 517)])
@pytest.mark.parametrize("inplace", [False, True])
def test_gauss_matches_oracle(hot_path, sigma, W, H, inplace):
    img = image(H, W, seed=W * H)
    want = oracle.port().gauss(img, sigma, inplace)
    if inplace:
        got = img.copy()
        hot_path.gauss(got, sigma, dst=got)
    else:
        got = hot_path.gauss(img, sigma)
    n = int((got != want).sum())
    assert n == 0, "%d samples differ, max abs %g" % (n, float(np.abs(got - want).max()))


def test_gauss_unsupported_types(hot_path):
    import art_b200
    img = image(32, 32, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.gauss(img, 2.0, gausstype=1)
