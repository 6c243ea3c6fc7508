"""GPU parity for the preview path of art_hp_develop (ABI 4): a PreviewProps window (crop + skip) of the demosaiced frame as
ImProcCoordinator / Crop ask RawImageSource::getImage for it (improccoordinator.cc L377, dcrop.cc L204), then the stages at scale = skip.
The oracle chain: AMaZE -> artoracle_transform_rect + artoracle_getimage_pp (both pinned against the reference's own transformRect, box sum,
CLIP, rotateLine in tests/test_oracle_getimage.py) -> matrix -> denoise at that scale.  Bit-exact through the chroma-only denoise."""
import ctypes

import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth
from art_b200.api import DenoiseParams, DevelopParams
from test_develop_gpu import CAM2WORK, MUL, adjust_params, guided_smoothing, run_chain_denoise
from test_oracle_getimage import getimage_pp, transform_rect

pytestmark = pytest.mark.gpu


def oracle_preview(raw, pp, tran, dn, guided, border=4):
    P = oracle.port()
    H, W = raw.shape
    x, y, w, h, skip = pp
    planes = [np.ascontiguousarray(p) for p in P.amaze(raw, synth.RGGB, 1.0, 4)]
    rect = list(transform_rect(P.lib, "artoracle_transform_rect", W, H, border, x, y, w, h, skip, tran))
    ow, oh = (w + skip - 1) // skip, (h + skip - 1) // skip            # getSize: the image getImage fills
    rect[2], rect[3] = (oh, ow) if tran & 1 else (ow, oh)               # imwidth / imheight clamped to the image, in the source orientation
    mul = tuple(np.float32(m) / np.float32(skip * skip) for m in MUL)   # rm /= area (L928-931)
    out = getimage_pp(P.lib, "artoracle_getimage_pp", planes, mul, 1, 0, tran, tuple(rect), skip)
    out = list(P.scale_convert(out, (1.0, 1.0, 1.0), False, CAM2WORK))
    if dn is not None:
        out = run_chain_denoise(P.lib, out, adjust_params(dn, dn[7]), None)
        if guided:
            out = guided_smoothing(out, guided, dn[7])
    return out, mul


@pytest.mark.parametrize("W,H,pp,tran", [
    (322, 260, (0, 0, 314, 252, 1), 0),            # the whole frame through the window form
    (322, 260, (0, 0, 314, 252, 2), 0),            # preview at scale 2
    (322, 260, (0, 0, 314, 252, 3), 0),            # ... 3: ragged last row / column, clamped box origins
    (645, 404, (0, 0, 637, 396, 4), 0),
    (645, 404, (100, 60, 300, 200, 1), 0),         # a detail crop (dcrop) at 1:1
    (645, 404, (101, 61, 333, 207, 2), 0),
    (645, 404, (337, 196, 300, 200, 1), 0),        # touching the far corner
    (322, 260, (0, 0, 252, 314, 2), 1),            # TR_R90: the window is in the turned frame's coordinates
    (322, 260, (10, 20, 200, 150, 2), 2),
    (322, 260, (5, 7, 120, 200, 3), 3),
    (322, 260, (30, 10, 250, 200, 2), 8),          # TR_HFLIP
    (322, 260, (30, 10, 200, 250, 2), 13),         # TR_R90 | TR_VFLIP | TR_HFLIP
])
def test_preview_window_matches_oracle(hot_path, W, H, pp, tran):
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=W + H + tran)
    want, mul = oracle_preview(raw, pp, tran, None, 0)
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=mul, do_clip=True, cam2work=CAM2WORK, tran=tran, pp=pp)
    got = hot_path.develop(raw, params)
    for g, w, ch in zip(got, want, "RGB"):
        assert g.shape == w.shape, (g.shape, w.shape)
        assert np.array_equal(g, w), "%s: %d of %d differ" % (ch, int((g != w).sum()), g.size)


@pytest.mark.parametrize("skip,guided", [(2, 0), (2, 5), (3, 3)])
def test_preview_with_denoise_at_scale(hot_path, skip, guided):
    """ImProcFunctions at scale = skip: adjust_params for RGB_denoise, the guided radius divided by the scale (chroma only: bit-exact)."""
    W, H = 1290, 808
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=skip)
    pp = (0, 0, W - 8, H - 8, skip)
    dn = (0, 0, 0, 15, 0, 0, 1.7, float(skip))
    want, mul = oracle_preview(raw, pp, 0, dn, guided)
    dnp = DenoiseParams(luminance=0, luminanceDetail=0, luminanceDetailThreshold=0, chrominance=15, chrominanceRedGreen=0, chrominanceBlueYellow=0,
                        gamma=1.7, scale=float(skip))
    from test_oracle_denoise import PROPHOTO
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=mul, do_clip=True, cam2work=CAM2WORK, denoise=dnp, wprof=PROPHOTO,
                           guided_chroma_radius=guided, pp=pp)
    got = hot_path.develop(raw, params)
    for g, w, ch in zip(got, want, "RGB"):
        assert np.array_equal(g, w), "%s: %d of %d differ" % (ch, int((g != w).sum()), g.size)


def test_window_outside_the_frame_is_refused(hot_path):
    raw = synth.bayer_frame(322, 260, synth.RGGB, seed=1)
    for pp in [(0, 0, 315, 252, 1), (10, 0, 314, 252, 2), (0, -1, 100, 100, 1)]:
        with pytest.raises(art_b200.HotPathError):
            hot_path.develop(raw, DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=MUL, do_clip=True, cam2work=CAM2WORK, pp=pp))


def test_full_size_preview(hot_path):
    """configs[1]'s frame at preview scale 4 and a 1:1 detail crop of it: same bits as the oracle on the window (the box sums only read the
    window's pixels, so the oracle runs on a cut of the demosaiced frame), timing printed."""
    import time
    W, H = 8192, 5464
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=21)
    pp = (0, 0, W - 8, H - 8, 4)
    mul = tuple(np.float32(m) / np.float32(16) for m in MUL)
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=mul, do_clip=True, cam2work=CAM2WORK, pp=pp)
    hot_path.develop(raw, params)
    t0 = time.perf_counter()
    got = hot_path.develop(raw, params)
    dt = time.perf_counter() - t0
    print("\n[preview] 8192x5464 -> %dx%d (skip 4) through art_hp_develop (pageable memory, copies included): %.1f ms" % (got[0].shape[1], got[0].shape[0], dt * 1e3))
    full = hot_path.demosaic_bayer(art_b200.BAYER_AMAZE, raw, synth.RGGB, initial_gain=1.0, border=4)     # bit-exact to the oracle (test_amaze_gpu.py)
    P = oracle.port()
    want = getimage_pp(P.lib, "artoracle_getimage_pp", [np.ascontiguousarray(p) for p in full], mul, 1, 0, 0, (4, 4, (W - 8 + 3) // 4, (H - 8 + 3) // 4), 4)
    want = P.scale_convert(want, (1.0, 1.0, 1.0), False, CAM2WORK)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
