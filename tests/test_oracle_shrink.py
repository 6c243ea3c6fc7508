"""Pin the wavelet-shrinkage oracle (oracle/shrink_port.c) against the reference's own MadRgb / ShrinkAllL /
ShrinkAllAB / WaveletDenoiseAll* compiled in place (oracle/_ref).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)


def image(H, W, seed, amp=800.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 20000 + 15000 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, amp, size=(H, W))
    return np.clip(img, 0, 65535).astype(np.float32)


def mad_table(w, lib, fname):
    f = getattr(lib, fname)
    f.restype = ctypes.c_float
    m = np.zeros((8, 3), np.float32)
    for lvl in range(w.maxlevel()):
        for d in (1, 2, 3):
            b = np.ascontiguousarray(w.band(lvl, d))
            v = f(b.ctypes.data_as(fp), b.size)
            m[lvl, d - 1] = np.float32(v) * np.float32(v)          # madL = SQR(MadRgb(..)), FTblockDN.cc L2311-2320
    return m


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 4097])
def test_madrgb(n):
    rng = np.random.default_rng(n)
    data = (rng.normal(0, 300, size=n)).astype(np.float32)
    p, r = oracle.port().lib, oracle.ref().lib
    p.artoracle_madrgb.restype = ctypes.c_float
    r.artref_madrgb.restype = ctypes.c_float
    assert p.artoracle_madrgb(data.ctypes.data_as(fp), n) == r.artref_madrgb(data.ctypes.data_as(fp), n)


@needs_ref
@pytest.mark.parametrize("W,H,levels", [(128, 96, 5), (131, 97, 5), (203, 77, 4), (322, 251, 6)])
@pytest.mark.parametrize("scale", [1.0, 2.0])
def test_denoise_L_and_AB(W, H, levels, scale):
    lum = image(H, W, seed=W + H)
    chroma = (image(H, W, seed=W * H, amp=1500.0) - 20000) * 0.3
    port, ref = oracle.port(), oracle.ref()
    objs = {}
    for name, o in (("port", port), ("ref", ref)):
        objs[name] = (o.wavelet(lum, levels, 1), o.wavelet(chroma, levels, 1))
    madp = mad_table(objs["port"][0], port.lib, "artoracle_madrgb")
    madr = mad_table(objs["ref"][0], ref.lib, "artref_madrgb")
    assert np.array_equal(madp, madr)
    h2, w2, _ = objs["port"][0].dims(0)
    rng = np.random.default_rng(7)
    nvl = (rng.uniform(0.5, 4.0, size=h2 * w2)).astype(np.float32)       # noisevarlum
    nvc = (rng.uniform(0.5, 2.0, size=h2 * w2)).astype(np.float32)       # noisevarchrom
    # chroma first (uses the unshrunk L coefficients, FTblockDN.cc L2328-2402), then luma
    rc = port.lib.artoracle_wavelet_denoise_AB(objs["port"][0].h, objs["port"][1].h, nvc.ctypes.data_as(fp), madp.ctypes.data_as(fp),
                                               ctypes.c_float(2.25), 0, 0, ctypes.c_double(scale))
    assert rc == 0
    rc = ref.lib.artref_wavelet_denoise_AB(objs["ref"][0].h, objs["ref"][1].h, nvc.ctypes.data_as(fp), madr.ctypes.data_as(fp),
                                           ctypes.c_float(2.25), 0, 0, ctypes.c_double(scale))
    assert rc == 0
    rc = port.lib.artoracle_wavelet_denoise_L(objs["port"][0].h, nvl.ctypes.data_as(fp), madp.ctypes.data_as(fp), ctypes.c_double(scale))
    assert rc == 0
    rc = ref.lib.artref_wavelet_denoise_L(objs["ref"][0].h, nvl.ctypes.data_as(fp), madr.ctypes.data_as(fp), ctypes.c_double(scale))
    assert rc == 0
    for k in (0, 1):
        for lvl in range(levels):
            for d in (1, 2, 3):
                a, b = objs["port"][k].band(lvl, d), objs["ref"][k].band(lvl, d)
                assert np.array_equal(a, b), "%s level %d band %d: %d differ" % ("L" if k == 0 else "ab", lvl, d, int((a != b).sum()))
    for v in objs.values():
        v[0].close(); v[1].close()
