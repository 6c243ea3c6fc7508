"""Pin the per-pixel oracle (oracle/pointwise_port.c) against the reference's own CLIP + matrix loop
(oracle/_ref, cut from rawimagesource.cc L3197-3211).  Bit-exact."""
import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")

# ProPhoto working space inverse x a plausible camera matrix (values only need to be generic doubles)
MAT = np.array([[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0000000, 0.0000000, 1.2118128]]) @ \
      np.array([[0.6594, 0.2521, 0.0528], [0.2661, 0.9712, -0.2373], [0.0292, -0.2046, 1.0003]])


def planes(H, W, seed, lo=-500.0, hi=80000.0):
    rng = np.random.default_rng(seed)
    return [rng.uniform(lo, hi, size=(H, W)).astype(np.float32) for _ in range(3)]


@needs_ref
@pytest.mark.parametrize("do_clip", [0, 1])
@pytest.mark.parametrize("use_mat", [False, True])
def test_port_matches_reference(do_clip, use_mat):
    p = planes(97, 131, seed=3)
    mul = (1.9371, 1.0, 1.4182)
    got = oracle.port().scale_convert(p, mul, do_clip, MAT if use_mat else None)
    want = oracle.ref().scale_convert(p, mul, do_clip, MAT if use_mat else None)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_port_numpy_restatement():
    """Independent numpy restatement: double accumulation left to right, one rounding."""
    p = planes(40, 50, seed=4)
    mul = np.float32([2.1, 1.0, 1.6])
    got = oracle.port().scale_convert(p, mul, 1, MAT)
    x = [np.clip(q * m, np.float32(0), np.float32(65535)) for q, m in zip(p, mul)]
    for k in range(3):
        want = (MAT[k, 0] * x[0].astype(np.float64) + MAT[k, 1] * x[1].astype(np.float64) + MAT[k, 2] * x[2].astype(np.float64)).astype(np.float32)
        assert np.array_equal(got[k], want)


@needs_ref
@pytest.mark.parametrize("pattern", ["RGGB", "BGGR", "GRBG", "GBRG"])
def test_scale_colors_port_matches_reference(pattern):
    from art_b200 import synth
    f = synth.BAYER_FILTERS[pattern]
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 16384, size=(75, 102)).astype(np.float32)
    black = (511.0, 512.5, 510.0, 513.25)
    mul = (8.1234, 4.0625, 6.3321, 4.0631)
    got, gmax = oracle.port().scale_colors_bayer(raw, f, black, mul)
    want, wmax = oracle.ref().scale_colors_bayer(raw, f, black, mul)
    assert np.array_equal(got, want) and gmax == wmax


def scale_colors_xtrans(lib, name, raw, xt, black, mul):
    import ctypes
    out = np.array(raw, dtype=np.float32, order="C", copy=True)
    H, W = out.shape
    x = np.ascontiguousarray(xt, dtype=np.int32)
    bl = (ctypes.c_float * 4)(*[float(v) for v in black])
    mu = (ctypes.c_float * 4)(*[float(v) for v in mul])
    ch = (ctypes.c_float * 3)()
    assert getattr(lib, name)(W, H, x.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                              ctypes.c_long(W), bl, mu, ch) == 0
    return out, [float(ch[0]), float(ch[1]), float(ch[2])]


@needs_ref
@pytest.mark.parametrize("dy,dx", [(0, 0), (2, 5), (4, 1)])
def test_scale_colors_xtrans_port_matches_reference(dy, dx):
    """scaleColors' X-Trans branch (rawimagesource.cc L2795-2826) against the reference's own loop."""
    from art_b200 import synth
    xt = synth.xtrans_matrix(dy, dx)
    rng = np.random.default_rng(12 + dy)
    raw = rng.integers(0, 16384, size=(75, 103)).astype(np.float32)
    black = (1023.0, 1024.5, 1022.0, 0.0)
    mul = (7.9, 4.0625, 6.3321, 0.0)
    got, gmax = scale_colors_xtrans(oracle.port().lib, "artoracle_scale_colors_xtrans", raw, xt, black, mul)
    want, wmax = scale_colors_xtrans(oracle.ref().lib, "artref_scale_colors_xtrans", raw, xt, black, mul)
    assert np.array_equal(got, want) and gmax == wmax
