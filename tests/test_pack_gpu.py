"""GPU parity for the output packing (art_hp_scanlines = Imagefloat::getScanline for every row, and the packed batch-queue entry
art_hp_develop_submit_packed) through the C-ABI against the oracle port, which tests/test_oracle_pack.py pins bit-exact to the reference's
own function compiled in place (the half conversion over every float).  Integer work: bit-exact."""
import numpy as np
import pytest

import art_b200
import oracle
from art_b200 import synth
from art_b200.api import DenoiseParams, DevelopParams
from test_oracle_pack import frame, scan

pytestmark = pytest.mark.gpu
FORMATS = [(8, 0), (16, 0), (16, 1), (32, 1)]


@pytest.mark.parametrize("bps,is_float", FORMATS)
@pytest.mark.parametrize("W,H", [(64, 48), (67, 35), (301, 203), (1021, 300)])
def test_scanlines_match_oracle(hot_path, bps, is_float, W, H):
    planes = frame(H, W, W + H)
    want = scan(oracle.port().lib, "artoracle_scanlines", planes, bps, is_float)
    got = hot_path.scanlines(planes[0], planes[1], planes[2], bps, bool(is_float))
    assert got.dtype == want.dtype and got.tobytes() == want.tobytes()


@pytest.mark.parametrize("bps,is_float", [(16, 0), (8, 0), (16, 1)])
def test_packed_batch_queue_equals_packing_the_float_planes(hot_path, bps, is_float):
    from test_develop_gpu import CAM2WORK, MUL
    from test_oracle_denoise import PROPHOTO
    W, H = 322, 260
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=MUL, do_clip=True, cam2work=CAM2WORK, wprof=PROPHOTO,
                           denoise=DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15, gamma=1.7), fattal=(30, 20, 0))
    Ho, Wo = params.out_shape(H, W)
    dt = np.uint8 if bps == 8 else np.uint16
    raws = [hot_path.pinned(H, W) for _ in range(2)]
    outs = [hot_path.pinned(Ho, 3 * Wo, dt) for _ in range(2)]
    frames = [synth.bayer_frame(W, H, synth.RGGB, seed=400 + k) for k in range(3)]
    got = []
    for k, f in enumerate(frames):
        if k >= 2:
            hot_path.develop_wait()
            got.append(outs[k & 1].array.copy())
        raws[k & 1].array[:] = f
        hot_path.develop_submit_packed(raws[k & 1].array, params, outs[k & 1].array, bps, bool(is_float))
    for k in (1, 2):
        hot_path.develop_wait()
        got.append(outs[k & 1].array.copy())
    for f, g in zip(frames, got):
        planes = hot_path.develop(f, params)
        want = scan(oracle.port().lib, "artoracle_scanlines", [np.ascontiguousarray(p) for p in planes], bps, is_float)
        assert g.tobytes() == want.tobytes()
    with pytest.raises(art_b200.HotPathError):       # a pageable output is refused
        hot_path.develop_submit_packed(raws[0].array, params, np.zeros((Ho, 3 * Wo), dt), bps, bool(is_float))
