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


def band_to_numpy(torch, wv, lvl, d):
    h, w, _ = wv.dims(lvl if d else wv.maxlevel() - 1)
    buf = torch.empty((h, w), dtype=torch.float32, device="cuda")
    ptr = wv.band_ptr(lvl, d)
    assert ptr
    import art_b200  # noqa: F401
    # device-to-device copy through torch's raw pointer view
    src = torch.from_dlpack(_dl(torch, ptr, h * w)).view(h, w)
    buf.copy_(src)
    return buf.cpu().numpy(), src


def _dl(torch, ptr, n):
    import numpy as _np  # noqa: F401

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Holder(), device="cuda")


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
    views = {}
    for lvl in range(maxlvl):
        assert wv.dims(lvl) == ref.dims(lvl)
        for d in (1, 2, 3):
            got, view = band_to_numpy(torch, wv, lvl, d)
            views[(lvl, d)] = view
            assert np.array_equal(got, ref.band(lvl, d)), "level %d band %d: %d differ" % (lvl, d, int((got != ref.band(lvl, d)).sum()))
    got, _ = band_to_numpy(torch, wv, maxlvl - 1, 0)
    assert np.array_equal(got, ref.band(maxlvl - 1, 0)), "lowpass"
    # modify coefficients identically on both sides, then reconstruct with a blend
    views[(0, 1)].mul_(0.5)
    views[(maxlvl - 1, 3)].mul_(0.25)
    ref.band(0, 1)[...] *= 0.5
    ref.band(maxlvl - 1, 3)[...] *= 0.25
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
