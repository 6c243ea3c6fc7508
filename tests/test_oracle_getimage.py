"""Pin the getImage oracle (oracle/pointwise_port.c: artoracle_getimage -- gains / clip, "Blend" highlight reconstruction, coarse rotation and the
mirrors) against the reference's own rotateLine, CLIP and HLRecovery_blend compiled in place (oracle/_ref: artref_getimage).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
HLMAX = np.float32([61000.0, 65535.0, 58000.0])


def getimage(lib, name, planes, mul, do_clip, hr, tran):
    H, W = planes[0].shape
    swap = (tran & 3) in (1, 3)
    out = [np.full((W, H) if swap else (H, W), np.nan, np.float32) for _ in range(3)]
    m = np.float32(mul)
    rc = getattr(lib, name)(W, H, *[p.ctypes.data_as(fp) for p in planes], ctypes.c_long(W), m.ctypes.data_as(fp), int(do_clip), int(hr),
                            HLMAX.ctypes.data_as(fp), int(tran), *[o.ctypes.data_as(fp) for o in out], ctypes.c_long(out[0].shape[1]))
    assert rc == 0
    return out


def planes(H, W, seed):
    rng = np.random.default_rng(seed)
    p = [rng.uniform(0, 40000.0, size=(H, W)).astype(np.float32) for _ in range(3)]
    for q in p:                                   # blown patches so that the highlight branch runs
        q[rng.integers(0, H, 40), rng.integers(0, W, 40)] = 70000.0
    return p


@needs_ref
@pytest.mark.parametrize("tran", list(range(16)))
@pytest.mark.parametrize("hr,do_clip", [(0, 1), (1, 0), (0, 0)])
def test_port_matches_reference(tran, hr, do_clip):
    p = planes(53, 71, seed=tran)
    mul = (1.9371, 1.0, 1.4182)
    got = getimage(oracle.port().lib, "artoracle_getimage", p, mul, do_clip, hr, tran)
    want = getimage(oracle.ref().lib, "artref_getimage", p, mul, do_clip, hr, tran)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("tran", list(range(16)))
def test_port_against_numpy(tran):
    """the geometry restated with numpy: rot90 by the quarter turns, then the mirrors"""
    p = planes(20, 33, seed=100 + tran)
    got = getimage(oracle.port().lib, "artoracle_getimage", p, (1.0, 1.0, 1.0), 0, 0, tran)
    for g, q in zip(got, p):
        w = {0: q, 1: np.rot90(q, -1), 2: np.rot90(q, 2), 3: np.rot90(q, 1)}[tran & 3]
        if tran & 8:
            w = w[:, ::-1]
        if tran & 4:
            w = w[::-1]
        assert np.array_equal(g, w)


def transform_rect(lib, name, W, H, border, x, y, w, h, skip, tran):
    out = (ctypes.c_int * 4)()
    assert getattr(lib, name)(W, H, border, x, y, w, h, skip, tran, out) == 0
    return tuple(out)


def getimage_pp(lib, name, planes, mul, do_clip, hr, tran, rect, skip):
    H, W = planes[0].shape
    sx1, sy1, iw, ih = rect
    swap = (tran & 3) in (1, 3)
    out = [np.full((iw, ih) if swap else (ih, iw), np.nan, np.float32) for _ in range(3)]
    m = np.float32(mul)
    rc = getattr(lib, name)(W, H, *[p.ctypes.data_as(fp) for p in planes], ctypes.c_long(W), m.ctypes.data_as(fp), int(do_clip), int(hr),
                            HLMAX.ctypes.data_as(fp), int(tran), sx1, sy1, iw, ih, skip, *[o.ctypes.data_as(fp) for o in out], ctypes.c_long(out[0].shape[1]))
    assert rc == 0
    return out


# PreviewProps windows: the whole frame, an interior crop, crops touching the far edges, an oversized request; skips 1 .. 5
WINDOWS = [(0, 0, 10000, 10000), (5, 7, 31, 22), (30, 20, 200, 200), (0, 11, 63, 40), (12, 0, 17, 45)]


@needs_ref
@pytest.mark.parametrize("tran", list(range(16)))
@pytest.mark.parametrize("skip", [1, 2, 3, 5])
@pytest.mark.parametrize("border", [4, 7, 0])
def test_transform_rect_matches_reference(tran, skip, border):
    for W, H in ((71, 53), (64, 48), (130, 33)):
        for (x, y, w, h) in WINDOWS:
            a = transform_rect(oracle.port().lib, "artoracle_transform_rect", W, H, border, x, y, w, h, skip, tran)
            b = transform_rect(oracle.ref().lib, "artref_transform_rect", W, H, border, x, y, w, h, skip, tran)
            assert a == b, (W, H, x, y, w, h, a, b)


@needs_ref
@pytest.mark.parametrize("tran", [0, 1, 2, 3, 4, 8, 13])
@pytest.mark.parametrize("skip", [1, 2, 3, 5])
@pytest.mark.parametrize("hr,do_clip", [(0, 1), (1, 0)])
def test_preview_form_matches_reference(tran, skip, hr, do_clip):
    """getImage for a PreviewProps window at skip >= 1: transformRect + the reference's own skip x skip box sum, CLIP, HLRecovery_blend, rotateLine."""
    p = planes(53, 71, seed=tran + skip)
    mul = tuple(np.float32(v) / np.float32(skip * skip) for v in (1.9371, 1.0, 1.4182))
    for (x, y, w, h) in WINDOWS:
        rect = transform_rect(oracle.ref().lib, "artref_transform_rect", 71, 53, 4, x, y, w, h, skip, tran)
        if rect[2] < 1 or rect[3] < 1:
            continue
        got = getimage_pp(oracle.port().lib, "artoracle_getimage_pp", p, mul, do_clip, hr, tran, rect, skip)
        want = getimage_pp(oracle.ref().lib, "artref_getimage_pp", p, mul, do_clip, hr, tran, rect, skip)
        for g, w_ in zip(got, want):
            assert np.array_equal(g, w_)
            assert not np.isnan(g).any()
