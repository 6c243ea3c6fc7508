"""Pin the hot / dead pixel filter's oracle (oracle/badpixels_port.c) against the reference's own RawImageSource::findHotDeadPixels and
::interpolateBadPixelsBayer compiled in place (oracle/_ref).  Bit-exact: Bayer and X-Trans detection, one and several reference threads (its
per-thread row chunks with their five-row ring), hot only / dead only / both, clustered bad pixels that force the fallback mean."""
import ctypes

import numpy as np
import pytest

import oracle
from art_b200 import synth

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
ip = ctypes.POINTER(ctypes.c_int)
bp = ctypes.POINTER(ctypes.c_ubyte)


def spiky(raw, seed, n):
    """Plant hot and dead samples (single ones, a 2x2 clump and a same-colour cluster) into a frame."""
    rng = np.random.default_rng(seed)
    out = raw.copy()
    H, W = out.shape
    ys, xs = rng.integers(0, H, n), rng.integers(0, W, n)
    out[ys[: n // 2], xs[: n // 2]] = 65535.0
    out[ys[n // 2:], xs[n // 2:]] = 0.0
    if H > 20 and W > 20:
        out[10:12, 10:12] = 60000.0
        out[14:19:2, 14:19:2] = 64000.0
    return out


def find(lib, name, raw, xt, thresh, hot, dead, *extra):
    H, W = raw.shape
    m = np.zeros((H, W), np.uint8)
    x = None if xt is None else np.ascontiguousarray(xt, np.int32)
    n = getattr(lib, name)(raw.ctypes.data_as(fp), W, H, None if x is None else x.ctypes.data_as(ip), ctypes.c_float(thresh), int(hot), int(dead),
                           m.ctypes.data_as(bp), *extra)
    return m, n


def interpolate(lib, name, raw, filters, m):
    out = raw.copy()
    H, W = out.shape
    n = getattr(lib, name)(out.ctypes.data_as(fp), W, H, ctypes.c_uint(filters), np.ascontiguousarray(m).ctypes.data_as(bp))
    return out, n


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (131, 97), (301, 203), (17, 12)])
@pytest.mark.parametrize("thresh,hot,dead", [(100.0, 1, 1), (40.0, 1, 0), (250.0, 0, 1)])
@pytest.mark.parametrize("threads", [1, 5])
def test_find_bayer(W, H, thresh, hot, dead, threads):
    if H - 4 < 2 * threads:
        # a reference thread whose static chunk has fewer than two rows re-evaluates its neighbour's last row against a ring that lacks
        # that row's upper context: frames under 2 x threads + 4 rows depend on the thread count.  The one-thread order is the oracle.
        pytest.skip("reference chunks under two rows")
    raw = spiky(synth.bayer_frame(W, H, synth.RGGB, seed=W + H), W, max(4, W * H // 300))
    got, gn = find(oracle.port().lib, "artoracle_find_hot_dead", raw, None, thresh, hot, dead)
    want, wn = find(oracle.ref().lib, "artref_find_hot_dead", raw, None, thresh, hot, dead, threads)
    assert gn == wn and np.array_equal(got, want)
    if W > 60:
        assert gn > 0


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (131, 97), (301, 203)])
@pytest.mark.parametrize("dy,dx", [(0, 0), (2, 5)])
@pytest.mark.parametrize("threads", [1, 4])
def test_find_xtrans(W, H, dy, dx, threads):
    xt = synth.xtrans_matrix(dy, dx)
    raw = spiky(synth.xtrans_frame(W, H, xt, seed=W + dy), H, max(4, W * H // 300))
    got, gn = find(oracle.port().lib, "artoracle_find_hot_dead", raw, xt, 100.0, 1, 1)
    want, wn = find(oracle.ref().lib, "artref_find_hot_dead", raw, xt, 100.0, 1, 1, threads)
    assert gn == wn and np.array_equal(got, want)
    assert gn > 0


@needs_ref
@pytest.mark.parametrize("filters", sorted(synth.BAYER_FILTERS.values()))
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (301, 203)])
def test_interpolate_bayer(filters, W, H):
    raw = spiky(synth.bayer_frame(W, H, filters, seed=W * 2 + H), W + 1, max(4, W * H // 200))
    m, n = find(oracle.port().lib, "artoracle_find_hot_dead", raw, None, 100.0, 1, 1)
    m[20:25, 20:25] = 1          # a block of bad pixels: no good pair for the centre, the fallback mean (and a pixel with nothing to use)
    m[30:35:2, 30:35:2] = 1
    got, gn = interpolate(oracle.port().lib, "artoracle_interpolate_bad_bayer", raw, filters, m)
    want, wn = interpolate(oracle.ref().lib, "artref_interpolate_bad_bayer", raw, filters, m)
    assert gn == wn and np.array_equal(got, want)
    assert (got != raw).sum() > 0 and np.array_equal(got[m == 0], raw[m == 0])


def interpolate_xtrans(lib, name, raw, xt, m, *extra):
    out = raw.copy()
    H, W = out.shape
    x = np.ascontiguousarray(xt, np.int32)
    n = getattr(lib, name)(out.ctypes.data_as(fp), W, H, x.ctypes.data_as(ip), np.ascontiguousarray(m).ctypes.data_as(bp), *extra)
    return out, n


def xtrans_bad_map(raw, xt, seed):
    """detected spikes + clumps: adjacent bad pixels of different colours (the unchecked virtual-pixel neighbours), runs along a row and a column
    (chains of pixels that each read the one rewritten before it), a solid block (pixels without a valid pair)"""
    m, _ = find(oracle.port().lib, "artoracle_find_hot_dead", raw, xt, 100.0, 1, 1)
    H, W = raw.shape
    rng = np.random.default_rng(seed)
    ys, xs = rng.integers(2, H - 3, 40), rng.integers(2, W - 3, 40)
    for y, x in zip(ys, xs):
        m[y, x] = 1
        m[y + rng.integers(-1, 2), x + rng.integers(-1, 2)] = 1
    m[20, 10:30] = 1
    m[12:34, 40] = 1
    m[30:36, 20:26] = 1
    return m


@needs_ref
@pytest.mark.parametrize("dy,dx", [(0, 0), (1, 2), (4, 5), (3, 1)])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (301, 203)])
def test_interpolate_xtrans(dy, dx, W, H):
    """interpolateBadPixelsXtrans in raster order = the reference on one thread (its parallel schedule is not a function of the input)"""
    xt = synth.xtrans_matrix(dy, dx)
    raw = spiky(synth.xtrans_frame(W, H, xt, seed=W + H + dy), W + 3, max(4, W * H // 150))
    m = xtrans_bad_map(raw, xt, W + dx)
    got, gn = interpolate_xtrans(oracle.port().lib, "artoracle_interpolate_bad_xtrans", raw, xt, m)
    want, wn = interpolate_xtrans(oracle.ref().lib, "artref_interpolate_bad_xtrans", raw, xt, m, 1)
    assert gn == wn and gn > 0 and np.array_equal(got, want)
    assert np.array_equal(got[m == 0], raw[m == 0])
