"""GPU parity for the dual demosaic and VNG4 (art_hp_demosaic_vng4, art_hp_dual_demosaic_bayer / _xtrans = RawImageSource::vng4_demosaic,
dual_demosaic_RT) through the C-ABI against the oracle ports, which tests/test_oracle_vng4.py / test_oracle_dual.py pin bit-exact to the
reference's own functions compiled in place.  Bit-exact: demosaicers, Color::RGB2L, the automatic contrast threshold, the blend mask and
the mixes."""
import ctypes
import time

import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth
from test_oracle_dual import FILTERS, P, XYZ_RGB, lum_scene
from test_oracle_vng4 import PREFILTERS, vng4

pytestmark = pytest.mark.gpu
F = ctypes.c_float


def same(got, want, what):
    for g, w, ch in zip(got, want, "RGB"):
        assert np.array_equal(g, w), "%s plane %s: %d of %d differ, first at %s" % (what, ch, int((g != w).sum()), g.size, np.argwhere(g != w)[0])


@pytest.mark.parametrize("filters", sorted(PREFILTERS))
@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (130, 77), (33, 95), (8, 8), (301, 203), (1023, 517)])
def test_vng4_matches_oracle(hot_path, filters, W, H):
    raw = synth.bayer_frame(W, H, filters, seed=W * 3 + H)
    raw[H // 3, W // 2] = 0.0
    raw[H // 2: H // 2 + 3, 2: 9] = 65535.0
    pf = PREFILTERS[filters]
    same(hot_path.demosaic_vng4(raw, pf), vng4(oracle.port().lib, "artoracle_vng4", raw, pf), "VNG4")


def oracle_dual_bayer(raw, filters, first, second, contrast, auto):
    lib = oracle.port().lib
    H, W = raw.shape
    r, g, b = [p.copy() for p in getattr(oracle.port(), first)(raw, filters)]
    c = ctypes.c_double(contrast)
    if contrast == 0.0 and not auto:
        return [r, g, b], 0.0
    if second == "bilinear":
        assert lib.artoracle_dual_bilinear_ex(P(raw), W, H, ctypes.c_uint(filters), P(r), P(g), P(b), ctypes.byref(c), int(auto), None) == 0
    else:
        assert lib.artoracle_dual_vng4(P(raw), W, H, ctypes.c_uint(PREFILTERS[filters]), P(r), P(g), P(b), ctypes.byref(c), int(auto), None) == 0
    return [r, g, b], c.value


@pytest.mark.parametrize("filters", FILTERS)
@pytest.mark.parametrize("first", ["amaze", "rcd"])
@pytest.mark.parametrize("second", ["bilinear", "vng4"])
@pytest.mark.parametrize("W,H,contrast,auto", [(130, 97, 20.0, False), (301, 203, 3.0, False), (301, 203, 20.0, True), (640, 427, 20.0, True),
                                               (97, 130, 75.0, False), (203, 301, 0.0, False)])
def test_dual_bayer_matches_oracle(hot_path, filters, first, second, W, H, contrast, auto):
    raw = synth.bayer_frame(W, H, filters, seed=W + H)
    want, wc = oracle_dual_bayer(raw, filters, first, second, contrast, auto)
    got, gc = hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE if first == "amaze" else art_b200.BAYER_RCD,
                                           0 if second == "bilinear" else 1, raw, filters, PREFILTERS[filters], contrast, auto)
    assert gc == wc, (gc, wc)
    same(got, want, "%s + %s" % (first, second))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("W,H", [(400, 300), (333, 251), (163, 170)])
def test_auto_threshold_on_crafted_luminance(hot_path, W, H, kind):
    """The three outcomes of buildBlendMask's search (a very flat tile in pass 0; the pixel-by-pixel scan of pass 1; nothing flat:
    threshold 0 and no blur) on frames whose first-demosaicer luminance is crafted: a grey frame (R = G = B on every site) makes
    AMaZE / RCD return nearly that grey, so the oracle chain and the device see the same L and must pick the same tile."""
    lum = lum_scene(W, H, W + kind, kind)
    raw = np.ascontiguousarray(np.clip(lum, 0, 65535), dtype=np.float32)
    f = 0x94949494
    for second in ("bilinear", "vng4"):
        want, wc = oracle_dual_bayer(raw, f, "rcd", second, 20.0, True)
        got, gc = hot_path.dual_demosaic_bayer(art_b200.BAYER_RCD, 0 if second == "bilinear" else 1, raw, f, PREFILTERS[f], 20.0, True)
        assert gc == wc, (gc, wc)
        same(got, want, "auto %d %s" % (kind, second))


@pytest.mark.parametrize("passes,lab", [(3, True), (1, False)])
@pytest.mark.parametrize("W,H,contrast,auto", [(131, 140, 20.0, False), (330, 270, 20.0, True), (200, 97, 0.0, False)])
def test_dual_xtrans_matches_oracle(hot_path, passes, lab, W, H, contrast, auto):
    lib = oracle.port().lib
    ip = ctypes.POINTER(ctypes.c_int)
    xt = np.ascontiguousarray(synth.xtrans_matrix(), np.int32)
    cam = np.array(synth.XTRANS_RGB_CAM, np.float32)
    raw = synth.xtrans_frame(W, H, xt, seed=W + passes)
    first = [np.empty_like(raw) for _ in range(3)]
    assert lib.artoracle_xtrans(W, H, xt.ctypes.data_as(ip), P(cam), passes, int(lab), P(raw), *[P(p) for p in first]) == 0
    want, wc = [p.copy() for p in first], 0.0
    if contrast != 0.0 or auto:
        L = np.zeros((H, W), np.float32)
        assert lib.artoracle_rgb2l(P(first[0]), P(first[1]), P(first[2]), P(L), W, H, P(XYZ_RGB)) == 0
        thr = np.float32(contrast / 100.0)
        if auto:
            lib.artoracle_auto_contrast_threshold.restype = ctypes.c_float
            thr = np.float32(lib.artoracle_auto_contrast_threshold(P(L), W, H, F(thr), F(1.0)))
        blend = np.zeros((H, W), np.float32)
        assert lib.artoracle_blend_mask(P(L), P(blend), W, H, F(thr), F(1.0), F(2.0)) == 0
        assert lib.artoracle_xtrans_fast_blend(W, H, xt.ctypes.data_as(ip), P(raw), P(blend), *[P(p) for p in want]) == 0
        wc = float(thr * np.float32(100.0))
    got, gc = hot_path.dual_demosaic_xtrans(raw, xt, cam, passes, lab, contrast, auto)
    assert gc == wc, (gc, wc)
    same(got, want, "X-Trans dual")


def test_full_frame_dual(hot_path):
    """configs[1]'s frame through AMAZEVNG4, with the automatic threshold and with a manual 20 %: the whole frame against the oracle chain
    (the Gaussian blur of the blend mask is a whole-column recurrence, so a cut of the frame is no witness), timing printed."""
    W, H, f = 8192, 5464, 0x94949494
    raw = synth.bayer_frame(W, H, f, seed=9)
    hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 1, raw, f, PREFILTERS[f], 20.0, True)
    t0 = time.perf_counter()
    got, gc = hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 1, raw, f, PREFILTERS[f], 20.0, True)
    dt = time.perf_counter() - t0
    print("\n[dual demosaic] AMaZE + VNG4, automatic threshold %.0f %%, 8192x5464 through the host entry (pageable memory, copies included): %.1f ms" % (gc, dt * 1e3))
    want, wc = oracle_dual_bayer(raw, f, "amaze", "vng4", 20.0, True)
    assert gc == wc, (gc, wc)
    same(got, want, "45 MP, automatic threshold")
    got, gc = hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 1, raw, f, PREFILTERS[f], 20.0, False)
    want, wc = oracle_dual_bayer(raw, f, "amaze", "vng4", 20.0, False)
    assert gc == wc, (gc, wc)
    same(got, want, "45 MP, threshold 20 %")
    assert any((g != a).any() for g, a in zip(got, hot_path.demosaic_bayer(art_b200.BAYER_AMAZE, raw, f, initial_gain=1.0, border=4)))      # the blend did something


def test_rejects_bad_arguments(hot_path):
    raw = synth.bayer_frame(64, 64, 0x94949494, seed=1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 0, raw, 0x94949494, 0xb4b4b4b4, 20.0, True)      # automatic threshold under 80x80
    with pytest.raises(art_b200.HotPathError):
        hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 1, raw, 0x94949494, 0x1e1e1e1e, 20.0, False)     # prefilters of another phase
    with pytest.raises(art_b200.HotPathError):
        hot_path.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 7, raw, 0x94949494, 0xb4b4b4b4, 20.0, False)
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_vng4(raw[:6], 0xb4b4b4b4)
