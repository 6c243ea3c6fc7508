"""Bayer green equilibration on the device (greeneq.cu, through the C-ABI) against the pinned oracle (tests/test_oracle_greeneq.py): bit-exact over
the four Bayer phases, ragged widths (SSE2 groups + scalar tail), constant and per-pixel thresholds, and a 45 MP frame."""
import ctypes

import numpy as np
import pytest

import oracle
from test_oracle_greeneq import FILTERS, fp, mosaic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95), (24, 12), (301, 203), (1203, 807)])
@pytest.mark.parametrize("border", [4, 0])
def test_global(hot_path, filters, W, H, border):
    raw = mosaic(H, W, W + H)
    want = raw.copy()
    assert oracle.port().lib.artoracle_green_equilibrate_global(want.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), border) == 0
    got = hot_path.green_equilibrate_global(raw.copy(), filters, border)
    assert np.array_equal(got, want), "%d of %d differ" % (int((got != want).sum()), got.size)
    assert (got != raw).any()


@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 37), (33, 95), (24, 12), (13, 9), (301, 203), (1203, 807)])
@pytest.mark.parametrize("thresh", [0.01, 0.05, 0.5, "map"])
def test_local(hot_path, filters, W, H, thresh):
    raw = mosaic(H, W, 3 * W + H)
    want = raw.copy()
    tmap = None
    if thresh == "map":
        tmap = np.ascontiguousarray(np.random.default_rng(W).uniform(0.0, 0.2, (H, W)), dtype=np.float32)
    t = 0.0 if tmap is not None else thresh
    assert oracle.port().lib.artoracle_green_equilibrate(want.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), ctypes.c_float(t),
                                                         tmap.ctypes.data_as(fp) if tmap is not None else None) == 0
    got = hot_path.green_equilibrate(raw.copy(), filters, t, tmap)
    n = int((got != want).sum())
    assert n == 0, "%d of %d differ, first at %s" % (n, got.size, np.argwhere(got != want)[0])


def test_config1_frame(hot_path):
    """both passes on the 45 MP frame of BASELINE.json configs[1]"""
    import time
    W, H, f = 8192, 5464, FILTERS[0]
    raw = mosaic(H, W, 5)
    want = raw.copy()
    assert oracle.port().lib.artoracle_green_equilibrate_global(want.ctypes.data_as(fp), W, H, ctypes.c_uint(f), 4) == 0
    assert oracle.port().lib.artoracle_green_equilibrate(want.ctypes.data_as(fp), W, H, ctypes.c_uint(f), ctypes.c_float(0.05), None) == 0
    got = raw.copy()
    t0 = time.perf_counter()
    hot_path.green_equilibrate_global(got, f, 4)
    hot_path.green_equilibrate(got, f, 0.05)
    dt = time.perf_counter() - t0
    assert np.array_equal(got, want)
    print("\n[green equilibration] %dx%d both passes through the host entries (pageable memory, copies included): %.1f ms" % (W, H, dt * 1e3))
