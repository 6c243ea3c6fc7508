"""GPU parity for the wavelet decomposition / reconstruction (through the C-ABI) against the oracle.  Bit-exact
on every subband of every level and on the reconstruction (with coefficients modified on the device)."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def image(H, W, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 20000 + 15000 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, 800, size=(H, W))
    return np.clip(img, 0, 65535).astype(np.float32)


@pytest.mark.parametrize("W,H,maxlvl", [(128, 96, 5), (131, 97, 3), (200, 77, 1), (701, 523, 8), (1003, 517, 6)])
def test_wavelet_matches_oracle(hot_path, W, H, maxlvl):
    import torch
    img = image(H, W, seed=W * H + maxlvl)
    ref = oracle.port().wavelet(img, maxlvl, 1)
    pitch = W + 5
    d_img = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
    d_img[:, :W] = torch.from_numpy(img).cuda()
    wv = hot_path.wavelet_decompose_dev(d_img.data_ptr(), pitch, W, H, maxlvl, 1)
    hot_path.sync()
    assert wv.maxlevel() == ref.maxlevel()
    for lvl in range(maxlvl):
        assert wv.dims(lvl) == ref.dims(lvl)
        for d in (1, 2, 3):
            got = wv.band(lvl, d)
            assert np.array_equal(got, ref.band(lvl, d)), "level %d band %d: %d differ" % (lvl, d, int((got != ref.band(lvl, d)).sum()))
    assert np.array_equal(wv.band(maxlvl - 1, 0), ref.band(maxlvl - 1, 0)), "lowpass"
    # modify coefficients identically on both sides, then reconstruct
    ref.band(0, 1)[...] *= 0.5
    ref.band(maxlvl - 1, 3)[...] *= 0.25
    wv.set_band(0, 1, ref.band(0, 1))
    wv.set_band(maxlvl - 1, 3, ref.band(maxlvl - 1, 3))
    for blend in (1.0,):
        want = ref.reconstruct(img.copy(), blend=blend)
        wv.reconstruct_dev(d_img.data_ptr(), pitch, blend)
        hot_path.sync()
        got = d_img[:, :W].cpu().numpy()
        assert np.array_equal(got, want), "%d samples differ" % int((got != want).sum())
    wv.close(); ref.close()


def test_wavelet_blend(hot_path):
    import torch
    W, H = 300, 200
    img = image(H, W, 9)
    ref = oracle.port().wavelet(img, 4, 1)
    d_img = torch.from_numpy(img).cuda()
    wv = hot_path.wavelet_decompose_dev(d_img.data_ptr(), W, W, H, 4, 1)
    want = ref.reconstruct(img.copy(), blend=0.3)
    wv.reconstruct_dev(d_img.data_ptr(), W, 0.3)
    hot_path.sync()
    assert np.array_equal(d_img.cpu().numpy(), want)
    wv.close(); ref.close()
