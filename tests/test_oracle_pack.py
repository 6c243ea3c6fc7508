"""Pin the output-packing oracle (oracle/pack_port.c) against the reference's own Imagefloat::getScanline and DNG_FloatToHalf compiled
in place (oracle/_ref).  Bit-exact: 16-bit (clamp + truncate), 8-bit (rounded), float32, half; the half conversion over EVERY float."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)


def frame(H, W, seed):
    rng = np.random.default_rng(seed)
    planes = [rng.uniform(-500.0, 70000.0, (H, W)).astype(np.float32) for _ in range(3)]
    edge = np.array([0.0, -0.0, 0.49, 0.5, 0.999, 1.0, 127.5, 128.0, 255.5, 256.0, 65534.5, 65534.999, 65535.0, 65535.5, 1e9, -1e9,
                     np.inf, -np.inf, np.nan, 1e-30, 65535.0 * 2.0 ** -15, 65535.0 * 2.0 ** -25, 65535.0 * 6.1e-5], np.float32)
    planes[0].flat[: edge.size] = edge
    planes[1].flat[: edge.size] = edge[::-1]
    planes[2].flat[: edge.size] = np.roll(edge, 5)
    return planes


def scan(lib, name, planes, bps, is_float):
    H, W = planes[0].shape
    dt = {(8, 0): np.uint8, (16, 0): np.uint16, (16, 1): np.uint16, (32, 1): np.float32}[(bps, is_float)]
    out = np.zeros((H, W * 3), dt)
    rc = getattr(lib, name)(planes[0].ctypes.data_as(fp), planes[1].ctypes.data_as(fp), planes[2].ctypes.data_as(fp), W, H, bps, is_float,
                            out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("bps,is_float", [(8, 0), (16, 0), (16, 1), (32, 1)])
@pytest.mark.parametrize("W,H", [(64, 48), (67, 35), (1, 1), (301, 203)])
def test_scanlines(bps, is_float, W, H):
    planes = frame(max(H, 1), max(W, 1), W + H) if W * H >= 23 else [np.full((H, W), v, np.float32) for v in (1.5, 65534.9, -3.0)]
    got = scan(oracle.port().lib, "artoracle_scanlines", planes, bps, is_float)
    want = scan(oracle.ref().lib, "artref_scanlines", planes, bps, is_float)
    assert got.tobytes() == want.tobytes()


def test_integer_packing_rules():
    """Known answers: 16-bit truncates after the clamp, 8-bit divides by 257 with rounding, NaN packs as 0."""
    p = [np.array([[0.0, 0.99, 1.0, 65534.99, 65535.0, 70000.0, -5.0, np.nan]], np.float32)] * 3
    o16 = scan(oracle.port().lib, "artoracle_scanlines", p, 16, 0)[0, ::3]
    assert o16.tolist() == [0, 0, 1, 65534, 65535, 65535, 0, 0]
    p = [np.array([[0.0, 128.0, 129.0, 257.0, 32896.0, 65535.0, 65407.0]], np.float32)] * 3
    o8 = scan(oracle.port().lib, "artoracle_scanlines", p, 8, 0)[0, ::3]
    assert o8.tolist() == [0, 0, 1, 1, 128, 255, 255]


@needs_ref
def test_float_to_half_every_float():
    """All 2^32 bit patterns, compared inside the reference library (a few seconds with OpenMP)."""
    fn = ctypes.cast(oracle.port().lib.artoracle_float_to_half, ctypes.c_void_p)
    mm = oracle.ref().lib.artref_float_to_half_mismatches
    mm.restype = ctypes.c_longlong
    mm.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
    bad = mm(0, 0x80000000, fn) + mm(0x80000000, 0xFFFFFFFF, fn)
    assert bad == 0
