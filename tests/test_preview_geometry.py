"""Host logic of the preview path without a GPU: art_hp_develop_size (getSize + the window check of art_develop_geometry2) against the reference's
own arithmetic -- RawImageSource::getSize is ceil(w / skip) x ceil(h / skip) (rawimagesource.cc L1199-1203) -- and DevelopParams.out_shape;
the source origin art_hp_develop uses is checked on the GPU against the oracle's transformRect (tests/test_preview_gpu.py)."""
import ctypes

import numpy as np
import pytest

import art_b200
from art_b200.api import DevelopParams


def develop_size(params, W, H):
    lib = art_b200.load_library()
    c = params.c_struct()
    w, h, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.art_hp_develop_size(ctypes.byref(c), W, H, ctypes.byref(w), ctypes.byref(h), ctypes.byref(b))
    return rc, w.value, h.value, b.value


@pytest.mark.parametrize("tran", list(range(16)))
@pytest.mark.parametrize("method,border", [(art_b200.BAYER_AMAZE, 4), (art_b200.XTRANS_3PASS, 7)])
def test_whole_frame_size(tran, method, border):
    W, H = 645, 404
    p = DevelopParams(method=method, tran=tran)
    rc, w, h, b = develop_size(p, W, H)
    assert rc == 0 and b == border
    turned = (tran & 3) in (1, 3)
    assert (w, h) == ((H - 2 * border, W - 2 * border) if turned else (W - 2 * border, H - 2 * border))
    assert p.out_shape(H, W) == (h, w)


@pytest.mark.parametrize("tran", [0, 1, 2, 3, 4, 8, 13])
@pytest.mark.parametrize("skip", [1, 2, 3, 5, 8])
def test_window_size_and_bounds(tran, skip):
    W, H, border = 645, 404, 4
    turned = (tran & 3) in (1, 3)
    fw, fh = ((H, W) if turned else (W, H))
    fw, fh = fw - 2 * border, fh - 2 * border           # getFullSize in the turned orientation
    rng = np.random.default_rng(tran * 10 + skip)
    for _ in range(40):
        x, y = int(rng.integers(0, fw - 30)), int(rng.integers(0, fh - 30))
        w, h = int(rng.integers(24, fw - x + 1)), int(rng.integers(24, fh - y + 1))
        p = DevelopParams(tran=tran, pp=(x, y, w, h, skip))
        rc, ow, oh, b = develop_size(p, W, H)
        assert rc == 0 and b == border
        assert (ow, oh) == (w // skip + (w % skip > 0), h // skip + (h % skip > 0))      # RawImageSource::getSize
        assert p.out_shape(H, W) == (oh, ow)
    for bad in [(0, 0, fw + 1, fh, skip), (1, 0, fw, fh, skip), (0, 1, fw, fh, skip), (-1, 0, 10, 10, skip), (0, 0, 0, 10, skip)]:
        rc, *_ = develop_size(DevelopParams(tran=tran, pp=bad), W, H)
        assert rc != 0
