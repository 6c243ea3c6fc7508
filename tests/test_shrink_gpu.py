"""GPU parity for the FTblockDN wavelet shrinkage (MadRgb table, WaveletDenoiseAllAB, WaveletDenoiseAllL) through
the C-ABI against the oracle port (itself pinned bit-exact to the reference functions in test_oracle_shrink.py).
Bit-exact on the MAD table and on every subband of both decompositions."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
fp = ctypes.POINTER(ctypes.c_float)


def image(H, W, seed, amp=800.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 20000 + 15000 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, amp, size=(H, W))
    return np.clip(img, 0, 65535).astype(np.float32)


def port_mad_table(w):
    f = oracle.port().lib.artoracle_madrgb
    f.restype = ctypes.c_float
    m = np.zeros((8, 3), np.float32)
    for lvl in range(w.maxlevel()):
        for d in (1, 2, 3):
            b = np.ascontiguousarray(w.band(lvl, d))
            v = f(b.ctypes.data_as(fp), b.size)
            m[lvl, d - 1] = np.float32(v) * np.float32(v)
    return m


@pytest.mark.parametrize("W,H,levels,scale,nab,ccurve,autoch", [
    (128, 96, 5, 1.0, 2.25, 0, 0),
    (131, 97, 5, 2.0, 2.25, 0, 0),
    (203, 77, 4, 1.0, 0.0005, 0, 1),       # autoch raises a tiny noisevar_ab to 0.02
    (322, 251, 6, 1.0, 3.0, 1, 0),         # noise C-curve: madab is not multiplied by noisevar_ab
    (1003, 517, 7, 1.0, 1.5, 0, 0),
    (2051, 1027, 5, 1.0, 2.25, 0, 0),
])
def test_shrink_matches_oracle(hot_path, W, H, levels, scale, nab, ccurve, autoch):
    import torch
    lum = image(H, W, seed=W + H)
    chroma = ((image(H, W, seed=W * H, amp=1500.0) - 20000) * 0.3).astype(np.float32)
    port = oracle.port()
    pL, pab = port.wavelet(lum, levels, 1), port.wavelet(chroma, levels, 1)
    d_lum, d_chroma = torch.from_numpy(lum).cuda(), torch.from_numpy(chroma).cuda()
    gL = hot_path.wavelet_decompose_dev(d_lum.data_ptr(), W, W, H, levels, 1)
    gab = hot_path.wavelet_decompose_dev(d_chroma.data_ptr(), W, W, H, levels, 1)

    want_mad = port_mad_table(pL)
    d_mad = torch.zeros((8, 3), dtype=torch.float32, device="cuda")
    hot_path.wavelet_mad_dev(gL, d_mad.data_ptr())
    hot_path.sync()
    got_mad = d_mad.cpu().numpy()
    assert np.array_equal(got_mad[:levels], want_mad[:levels]), (got_mad, want_mad)

    h2, w2, _ = pL.dims(0)
    rng = np.random.default_rng(7)
    nvl = rng.uniform(0.5, 4.0, size=h2 * w2).astype(np.float32)
    nvc = rng.uniform(0.5, 2.0, size=h2 * w2).astype(np.float32)
    d_nvl, d_nvc = torch.from_numpy(nvl).cuda(), torch.from_numpy(nvc).cuda()

    assert port.lib.artoracle_wavelet_denoise_AB(pL.h, pab.h, nvc.ctypes.data_as(fp), want_mad.ctypes.data_as(fp),
                                                 ctypes.c_float(nab), ccurve, autoch, ctypes.c_double(scale)) == 0
    assert port.lib.artoracle_wavelet_denoise_L(pL.h, nvl.ctypes.data_as(fp), want_mad.ctypes.data_as(fp), ctypes.c_double(scale)) == 0
    hot_path.wavelet_denoise_AB_dev(gL, gab, d_nvc.data_ptr(), d_mad.data_ptr(), nab, ccurve, autoch, scale)
    hot_path.wavelet_denoise_L_dev(gL, d_nvl.data_ptr(), d_mad.data_ptr(), scale)
    hot_path.sync()
    for name, g, p in (("ab", gab, pab), ("L", gL, pL)):
        for lvl in range(levels):
            for d in (1, 2, 3):
                a, b = g.band(lvl, d), p.band(lvl, d)
                assert np.array_equal(a, b), "%s level %d band %d: %d of %d differ, max |d| %g" % (
                    name, lvl, d, int((a != b).sum()), a.size, float(np.abs(a - b).max()))
    # and the reconstruction of the shrunk coefficients
    want = pL.reconstruct(lum.copy())
    gL.reconstruct_dev(d_lum.data_ptr(), W, 1.0)
    hot_path.sync()
    assert np.array_equal(d_lum.cpu().numpy(), want)
    for o in (pL, pab, gL, gab):
        o.close()


def test_shrink_rejects_mismatched_decompositions(hot_path):
    import torch
    from art_b200.api import HotPathError
    a = torch.zeros((64, 64), dtype=torch.float32, device="cuda")
    b = torch.zeros((64, 80), dtype=torch.float32, device="cuda")
    wa = hot_path.wavelet_decompose_dev(a.data_ptr(), 64, 64, 64, 3, 1)
    wb = hot_path.wavelet_decompose_dev(b.data_ptr(), 80, 80, 64, 3, 1)
    d = torch.zeros(4096, dtype=torch.float32, device="cuda")
    with pytest.raises(HotPathError):
        hot_path.wavelet_denoise_AB_dev(wa, wb, d.data_ptr(), d.data_ptr(), 2.0)
    wa.close(); wb.close()
