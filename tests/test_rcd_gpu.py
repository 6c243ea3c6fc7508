"""GPU parity for RCD: the CUDA path (through the C-ABI) against the oracle.  Bit-exact."""
import os

import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cmp(got, want, what):
    for g, w, ch in zip(got, want, "RGB"):
        n = int((g != w).sum())
        assert n == 0, "%s plane %s: %d of %d samples differ (max abs %.6g)" % (
            what, ch, n, g.size, float(np.abs(g - w).max()))


@pytest.mark.parametrize("name", ["rcd_rggb_scene", "rcd_gbrg_noise"])
def test_cuda_matches_golden(hot_path, name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    raw = z["raw"].astype(np.float32)
    got = hot_path.demosaic_bayer(art_b200.BAYER_RCD, raw, int(z["filters"]))
    _cmp(got, (z["red"], z["green"], z["blue"]), name)


@pytest.mark.parametrize("pattern", ["RGGB", "BGGR", "GRBG", "GBRG"])
@pytest.mark.parametrize("W,H,kind", [(640, 500, "scene"), (401, 367, "noise"), (177, 195, "scene"),
                                      (352, 40, "noise"), (1000, 700, "scene")])
def test_cuda_matches_oracle(hot_path, pattern, W, H, kind):
    f = synth.BAYER_FILTERS[pattern]
    raw = synth.bayer_frame(W, H, f, seed=W + H) if kind == "scene" else synth.random_frame(W, H, seed=W * H)
    got = hot_path.demosaic_bayer(art_b200.BAYER_RCD, raw, f)
    _cmp(got, oracle.port().rcd(raw, f), "%s %dx%d %s" % (pattern, W, H, kind))


def test_cuda_matches_reference_bodies_config1(hot_path):
    """BASELINE config 1: 4000x3000 RGGB, against the reference's own code when oracle/_ref travelled."""
    f = synth.RGGB
    raw = synth.bayer_frame(4000, 3000, f, seed=1001)
    got = hot_path.demosaic_bayer(art_b200.BAYER_RCD, raw, f)
    want = oracle.ref(det=True).rcd(raw, f) if oracle.have_ref() else oracle.port().rcd(raw, f)
    _cmp(got, want, "config1")


def test_rawimagesource_mirror(hot_path):
    f = synth.GRBG
    raw = synth.bayer_frame(300, 260, f, seed=9)
    src = art_b200.RawImageSource(raw, f, hot_path=hot_path)
    src.demosaic("rcd")
    _cmp((src.red, src.green, src.blue), oracle.port().rcd(raw, f), "mirror")


def test_pinned_and_strided_host_buffers(hot_path):
    """Row tables over pinned memory (direct DMA) and over a padded, non-contiguous parent (staged)."""
    f = synth.RGGB
    W, H = 500, 333
    raw = synth.bayer_frame(W, H, f, seed=21)
    want = oracle.port().rcd(raw, f)
    pin = [hot_path.pinned(H, W) for _ in range(4)]
    pin[0].array[:] = raw
    got = hot_path.demosaic_bayer(art_b200.BAYER_RCD, pin[0].array, f, pin[1].array, pin[2].array, pin[3].array)
    _cmp(got, want, "pinned")
    big = np.zeros((H, W + 24), np.float32)
    big[:, 8:8 + W] = raw
    outs = [np.zeros((H, W + 24), np.float32)[:, 8:8 + W] for _ in range(3)]
    got = hot_path.demosaic_bayer(art_b200.BAYER_RCD, big[:, 8:8 + W], f, *outs)
    _cmp(got, want, "strided")


def test_invalid_arguments(hot_path):
    raw = np.zeros((64, 64), np.float32)
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_bayer(art_b200.BAYER_RCD, raw, 0xFFFFFFFF)      # colour 3 in the CFA
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_bayer(7, raw, synth.RGGB)                        # unknown method
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_bayer(art_b200.BAYER_RCD, raw[:8, :8], synth.RGGB)


def test_rawimagesource_mirror_other_sensors_and_methods(hot_path):
    """the dispatcher mirror through the real library: X-Trans 3-pass against the oracle, VNG4 and a dual method against the entries called directly"""
    import ctypes
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int)
    xt = synth.xtrans_matrix()
    cam = np.array(synth.XTRANS_RGB_CAM, np.float32)
    xraw = synth.xtrans_frame(210, 150, xt, seed=4)
    src = art_b200.RawImageSource(xraw, hot_path=hot_path, xtrans=xt, rgb_cam=cam)
    got = src.demosaic("3-pass (best)")
    want = [np.empty_like(xraw) for _ in range(3)]
    assert oracle.port().lib.artoracle_xtrans(210, 150, np.ascontiguousarray(xt, np.int32).ctypes.data_as(ip), cam.ctypes.data_as(fp), 3, 1,
                                              xraw.ctypes.data_as(fp), *[w.ctypes.data_as(fp) for w in want]) == 0
    _cmp(got, want, "mirror x-trans")
    f, pre = synth.RGGB, 0xb4b4b4b4
    raw = synth.bayer_frame(300, 260, f, seed=10)
    b = art_b200.RawImageSource(raw, f, hot_path=hot_path, prefilters=pre)
    _cmp([p.copy() for p in b.demosaic("vng4")], hot_path.demosaic_vng4(raw, pre), "mirror vng4")
    direct, c = hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 0, raw, f, pre, 0.0, True, 1.0, 4)
    _cmp([p.copy() for p in b.demosaic("amazebilinear", autoContrast=True)], direct, "mirror dual")
    assert b.contrastThreshold == c
