"""GPU parity for the X-Trans demosaic (art_hp_demosaic_xtrans = RawImageSource::xtrans_interpolate + border) through the C-ABI
against the oracle port, which test_oracle_xtrans.py pins bit-exact to the reference compiled in place.  Bit-exact."""
import numpy as np
import pytest

import art_b200
from art_b200 import synth
from test_oracle_xtrans import CAM, MODES, port_xtrans, same

pytestmark = pytest.mark.gpu

SIZES = [(300, 260), (131, 140), (24, 24), (23, 40), (120, 40), (40, 121), (233, 119), (215, 217), (500, 333)]


@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("passes,lab", MODES)
@pytest.mark.parametrize("dy,dx", [(0, 0), (1, 2), (4, 5)])
def test_xtrans_matches_oracle(hot_path, W, H, passes, lab, dy, dx):
    xt = synth.xtrans_matrix(dy, dx)
    raw = synth.xtrans_frame(W, H, xt, seed=W + dy)
    same(hot_path.demosaic_xtrans(raw, xt, CAM, passes, lab), port_xtrans(raw, xt, passes, lab))


@pytest.mark.parametrize("passes,lab", [(1, 0), (3, 1)])
def test_xtrans_noise_frame(hot_path, passes, lab):
    xt = synth.xtrans_matrix(2, 1)
    raw = synth.random_frame(333, 251, seed=3)
    raw[40:60, 50:90] = 65535.0
    raw[100:120, 10:70] = 0.0
    same(hot_path.demosaic_xtrans(raw, xt, CAM, passes, lab), port_xtrans(raw, xt, passes, lab))


@pytest.mark.parametrize("passes,lab", [(1, 0), (3, 1)])
def test_xtrans_more_tiles_than_resident_ctas(hot_path, passes, lab):
    """2.4 MP: 16 x 16 = 256+ tiles... every CTA of the persistent grid walks several tiles"""
    W, H = 2000, 1700
    xt = synth.xtrans_matrix(5, 1)
    raw = synth.xtrans_frame(W, H, xt, seed=11)
    same(hot_path.demosaic_xtrans(raw, xt, CAM, passes, lab), port_xtrans(raw, xt, passes, lab))


def test_xtrans_is_repeatable(hot_path):
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(700, 500, xt, seed=5)
    a = hot_path.demosaic_xtrans(raw, xt, CAM, 3, 1)
    b = hot_path.demosaic_xtrans(raw, xt, CAM, 3, 1)
    same(a, b)


def test_xtrans_full_size_properties(hot_path):
    """BASELINE configs[3] size (6240 x 4160, 26 MP): finite, non-negative, native samples pass through unchanged"""
    W, H = 6240, 4160
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(W, H, xt, seed=1004)
    out = hot_path.demosaic_xtrans(raw, xt, CAM, 3, 1)
    cmap = xt[np.arange(H)[:, None] % 6, np.arange(W)[None, :] % 6]
    for ch, p in enumerate(out):
        assert np.isfinite(p).all() and (p >= 0).all()
    # outside the 8-pixel border every direction plane keeps the native sample, so the average does too (up to the sum / count rounding)
    inner = np.zeros((H, W), bool)
    inner[16:-16, 16:-16] = True
    for ch, p in enumerate(out):
        m = inner & (cmap == ch)
        assert np.abs(p[m] - raw[m]).max() <= 1e-3 * 65535 / 64


def test_xtrans_device_form_with_pitch(hot_path):
    torch = pytest.importorskip("torch")
    W, H, rp, op = 203, 141, 224, 256
    xt = synth.xtrans_matrix(3, 2)
    raw = synth.xtrans_frame(W, H, xt, seed=7)
    d_raw = torch.zeros((H, rp), dtype=torch.float32, device="cuda")
    d_raw[:, :W] = torch.from_numpy(raw).cuda()
    outs = [torch.zeros((H, op), dtype=torch.float32, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()
    hot_path.demosaic_xtrans_dev(W, H, xt, CAM, d_raw.data_ptr(), rp, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), op, 3, 1)
    hot_path.sync()
    same([o[:, :W].cpu().numpy() for o in outs], port_xtrans(raw, xt, 3, 1))


def test_xtrans_rejects_bad_arguments(hot_path):
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(64, 64, xt, seed=1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_xtrans(raw, xt, CAM, passes=2)
    with pytest.raises(art_b200.HotPathError):
        hot_path.demosaic_xtrans(raw, np.ones((6, 6), np.int32), CAM)
