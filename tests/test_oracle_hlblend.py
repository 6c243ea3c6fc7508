"""Pin the "Blend" highlight-reconstruction oracle (oracle/pointwise_port.c: artoracle_hl_blend) against the reference's own
RawImageSource::HLRecovery_blend compiled in place (oracle/_ref).  Bit-exact; unclipped pixels are left alone."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
F = ctypes.c_float


def line(n, seed, hlmax):
    rng = np.random.default_rng(seed)
    rgb = [rng.uniform(100.0, 1.25 * m, n).astype(np.float32) for m in hlmax]       # up to 25 % above each clip point
    rgb[0][::7] = hlmax[0]
    rgb[1][3::11] = 65535.0
    rgb[2][5::13] = 0.95 * 65535.0                                                    # exactly at the clip threshold
    for k in range(3):
        rgb[k][: n // 5] = rng.uniform(100.0, 20000.0, n // 5).astype(np.float32)     # a stretch of unclipped pixels
    return rgb


def run(lib, name, rgb, hlmax, maxval=65535.0):
    out = [p.copy() for p in rgb]
    h = np.array(hlmax, np.float32)
    assert getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), out[0].size, F(maxval), h.ctypes.data_as(fp)) == 0
    return out


@needs_ref
@pytest.mark.parametrize("hlmax", [(65535.0, 65535.0, 65535.0), (124000.0, 65535.0, 98000.0), (40000.0, 70000.0, 52000.0)])
@pytest.mark.parametrize("n", [1, 37, 4001])
def test_port_matches_reference(hlmax, n):
    rgb = line(max(n, 16), n + int(hlmax[0]), hlmax)
    rgb = [p[:n].copy() for p in rgb]
    got = run(oracle.port().lib, "artoracle_hl_blend", rgb, hlmax)
    want = run(oracle.ref().lib, "artref_hl_blend", rgb, hlmax)
    for g, w, ch in zip(got, want, "RGB"):
        eq = (g == w) | (np.isnan(g) & np.isnan(w))
        assert eq.all(), "%s: %d differ" % (ch, int((~eq).sum()))


def test_unclipped_pixels_are_untouched_and_clipped_ones_change():
    hlmax = (124000.0, 65535.0, 98000.0)
    rgb = line(4001, 5, hlmax)
    out = run(oracle.port().lib, "artoracle_hl_blend", rgb, hlmax)
    clipped = (rgb[0] > 0.95 * 65535) | (rgb[1] > 0.95 * 65535) | (rgb[2] > 0.95 * 65535)
    for o, p in zip(out, rgb):
        assert np.array_equal(o[~clipped], p[~clipped])
    assert any((o[clipped] != p[clipped]).any() for o, p in zip(out, rgb))
