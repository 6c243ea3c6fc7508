"""GPU parity for the hot / dead pixel filter (art_hp_find_hot_dead_pixels, art_hp_interpolate_bad_pixels_bayer = RawImageSource::
findHotDeadPixels / interpolateBadPixelsBayer, rtengine/badpixels.cc) through the C-ABI against the oracle port, which
tests/test_oracle_badpixels.py pins bit-exact to the reference's own functions compiled in place.  Maps, counts and samples: bit-exact."""
import time

import numpy as np
import pytest

import oracle
from art_b200 import synth
from test_oracle_badpixels import find, interpolate, spiky

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (131, 97), (301, 203), (17, 12), (5, 5), (4, 9), (1029, 515)])
@pytest.mark.parametrize("thresh,hot,dead", [(100.0, 1, 1), (40.0, 1, 0), (250.0, 0, 1)])
def test_find_bayer(hot_path, W, H, thresh, hot, dead):
    raw = spiky(synth.bayer_frame(W, H, synth.RGGB, seed=W + H), W, max(4, W * H // 300))
    want, wn = find(oracle.port().lib, "artoracle_find_hot_dead", raw, None, thresh, hot, dead)
    got, gn = hot_path.find_hot_dead_pixels(raw, thresh, hot, dead)
    assert gn == wn and np.array_equal(got, want)


@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (131, 97), (301, 203), (1029, 515)])
@pytest.mark.parametrize("dy,dx", [(0, 0), (2, 5), (4, 1)])
def test_find_xtrans(hot_path, W, H, dy, dx):
    xt = synth.xtrans_matrix(dy, dx)
    raw = spiky(synth.xtrans_frame(W, H, xt, seed=W + dy), H, max(4, W * H // 300))
    want, wn = find(oracle.port().lib, "artoracle_find_hot_dead", raw, xt, 100.0, 1, 1)
    got, gn = hot_path.find_hot_dead_pixels(raw, 100.0, True, True, xtrans=xt)
    assert gn == wn and gn > 0 and np.array_equal(got, want)


def test_marks_are_ored_into_the_callers_map(hot_path):
    raw = spiky(synth.bayer_frame(200, 150, synth.RGGB, seed=3), 5, 80)
    known = np.zeros((150, 200), np.uint8)
    known[7, 9] = known[100, 33] = 1
    got, gn = hot_path.find_hot_dead_pixels(raw, 100.0, True, True, bad_map=known)
    fresh, fn = hot_path.find_hot_dead_pixels(raw, 100.0, True, True)
    assert gn == fn and np.array_equal(got, fresh | known)


@pytest.mark.parametrize("filters", sorted(synth.BAYER_FILTERS.values()))
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (301, 203), (1029, 515)])
def test_interpolate_bayer(hot_path, filters, W, H):
    raw = spiky(synth.bayer_frame(W, H, filters, seed=W * 2 + H), W + 1, max(4, W * H // 200))
    m, _ = find(oracle.port().lib, "artoracle_find_hot_dead", raw, None, 100.0, 1, 1)
    m[20:25, 20:25] = 1
    m[30:35:2, 30:35:2] = 1
    want, wn = interpolate(oracle.port().lib, "artoracle_interpolate_bad_bayer", raw, filters, m)
    got = raw.copy()
    gn = hot_path.interpolate_bad_pixels_bayer(got, filters, m)
    assert gn == wn and np.array_equal(got, want)


def test_full_frame(hot_path):
    """configs[1]'s frame: detection and repair against the oracle at full size (the oracle does 45 MP in a few seconds), timing printed."""
    W, H, f = 8192, 5464, synth.RGGB
    raw = spiky(synth.bayer_frame(W, H, f, seed=11), 17, 4000)
    want, wn = find(oracle.port().lib, "artoracle_find_hot_dead", raw, None, 100.0, 1, 1)
    hot_path.find_hot_dead_pixels(raw, 100.0, True, True)
    t0 = time.perf_counter()
    got, gn = hot_path.find_hot_dead_pixels(raw, 100.0, True, True)
    t1 = time.perf_counter()
    assert gn == wn and gn >= 3000 and np.array_equal(got, want)
    fixed_want, n1 = interpolate(oracle.port().lib, "artoracle_interpolate_bad_bayer", raw, f, want)
    fixed = raw.copy()
    t2 = time.perf_counter()
    n2 = hot_path.interpolate_bad_pixels_bayer(fixed, f, got)
    t3 = time.perf_counter()
    assert n1 == n2 and np.array_equal(fixed, fixed_want)
    print("\n[hot / dead pixels] 8192x5464 through the host entries (pageable memory, copies included): find %.1f ms (%d marked), interpolate %.1f ms"
          % ((t1 - t0) * 1e3, gn, (t3 - t2) * 1e3))


@pytest.mark.parametrize("dy,dx", [(0, 0), (1, 2), (4, 5)])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (301, 203), (1029, 515)])
def test_interpolate_xtrans(hot_path, dy, dx, W, H):
    """art_hp_interpolate_bad_pixels_xtrans = interpolateBadPixelsXtrans in raster order (the oracle is pinned to the reference on one thread),
    incl. runs of bad pixels along a row and a column whose members read the ones rewritten before them"""
    from test_oracle_badpixels import interpolate_xtrans, xtrans_bad_map
    xt = synth.xtrans_matrix(dy, dx)
    raw = spiky(synth.xtrans_frame(W, H, xt, seed=W + H + dy), W + 3, max(4, W * H // 150))
    m = xtrans_bad_map(raw, xt, W + dx)
    want, wn = interpolate_xtrans(oracle.port().lib, "artoracle_interpolate_bad_xtrans", raw, xt, m)
    got = raw.copy()
    gn = hot_path.interpolate_bad_pixels_xtrans(got, xt, m)
    assert gn == wn and np.array_equal(got, want)


def test_interpolate_xtrans_full_frame_and_bad_layout(hot_path):
    """configs[3]'s 26 MP frame; a 6x6 table that is not an X-Trans layout is refused (the reference's scan runs off the frame there)"""
    import art_b200
    from test_oracle_badpixels import interpolate_xtrans
    W, H = 6240, 4160
    xt = synth.xtrans_matrix()
    raw = spiky(synth.xtrans_frame(W, H, xt, seed=3), 5, 6000)
    m, _ = find(oracle.port().lib, "artoracle_find_hot_dead", raw, xt, 100.0, 1, 1)
    m[1000, 500:560] = 1
    want, wn = interpolate_xtrans(oracle.port().lib, "artoracle_interpolate_bad_xtrans", raw, xt, m)
    got = raw.copy()
    gn = hot_path.interpolate_bad_pixels_xtrans(got, xt, m)
    assert gn == wn and gn > 1000 and np.array_equal(got, want)
    bad = np.zeros((6, 6), np.int32)
    bad[0, 0] = 2        # a lone blue site: nothing of its colour at distance 2
    with pytest.raises(art_b200.HotPathError):
        hot_path.interpolate_bad_pixels_xtrans(raw[:64, :64].copy(), bad, m[:64, :64])
