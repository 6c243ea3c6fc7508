"""GPU parity for denoise::detail_mask and denoise::NLMeans (through the C-ABI) against the oracle port (pinned
bit-exact to the reference functions in test_oracle_nlmeans.py).  Bit-exact, including the flush-to-zero cases."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
fp = ctypes.POINTER(ctypes.c_float)
F = ctypes.c_float


def luminance(H, W, seed, noise=900.0, dark=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 18000 + 14000 * np.sin(0.045 * x) * np.cos(0.06 * y) + 9000 * ((x // 23 + y // 17) % 2) + rng.normal(0, noise, size=(H, W))
    if dark:
        img[: H // 3] *= 1e-4
        img[: H // 5, : W // 2] = 0
    return np.clip(img, 0, 65535).astype(np.float32)


def port_detail_mask(src, scaling, threshold, ceiling, factor, blur_type, blur):
    H, W = src.shape
    out = np.zeros_like(src)
    assert oracle.port().lib.artoracle_detail_mask(src.ctypes.data_as(fp), out.ctypes.data_as(fp), W, H, F(scaling), F(threshold), F(ceiling),
                                                   F(factor), blur_type, F(blur)) == 0
    return out


def port_nlmeans(img, normcoeff, strength, detail, scale):
    H, W = img.shape
    out = img.copy()
    assert oracle.port().lib.artoracle_nlmeans(out.ctypes.data_as(fp), W, H, F(normcoeff), strength, detail, F(scale)) == 0
    return out


@pytest.mark.parametrize("W,H", [(64, 48), (131, 97), (7, 40), (803, 517)])
@pytest.mark.parametrize("blur_type,blur", [(2, 2.0), (1, 2.0), (0, 0.0)])
def test_detail_mask(hot_path, W, H, blur_type, blur):
    img = luminance(H, W, seed=W * 3 + H)
    want = port_detail_mask(img, 65535.0, 65.535, 65535.0, 0.3, blur_type, blur)
    got = hot_path.detail_mask(img, 65535.0, 65.535, 65535.0, 0.3, blur_type, blur)
    assert np.array_equal(got, want), "%d of %d differ" % (int((got != want).sum()), got.size)


@pytest.mark.parametrize("W,H,strength,detail,scale,dark", [
    (64, 48, 50, 50, 1.0, False),
    (131, 97, 100, 0, 1.0, False),
    (150, 150, 20, 80, 1.0, True),
    (283, 161, 70, 30, 1.0, True),
    (283, 161, 70, 30, 2.0, False),
    (300, 290, 35, 100, 1.5, False),
    (1203, 807, 60, 50, 1.0, True),
])
def test_nlmeans_matches_oracle(hot_path, W, H, strength, detail, scale, dark):
    img = luminance(H, W, seed=W + 5 * H + strength, dark=dark)
    want = port_nlmeans(img, 65535.0, strength, detail, scale)
    got = hot_path.nlmeans(img.copy(), 65535.0, strength, detail, scale)
    assert np.array_equal(got, want), "%d of %d differ, max %g" % (int((got != want).sum()), got.size, float(np.abs(got - want).max()))


def test_nlmeans_device_plane_with_pitch(hot_path):
    import torch
    W, H, pitch = 333, 222, 352
    img = luminance(H, W, seed=5)
    want = port_nlmeans(img, 65535.0, 40, 60, 1.0)
    d = torch.full((H, pitch), 7.0, dtype=torch.float32, device="cuda")
    d[:, :W] = torch.from_numpy(img).cuda()
    hot_path.nlmeans_dev(d.data_ptr(), pitch, W, H, 65535.0, 40, 60, 1.0)
    hot_path.sync()
    out = d.cpu().numpy()
    assert np.array_equal(out[:, :W], want)
    assert (out[:, W:] == 7.0).all()


def test_nlmeans_strength_zero_and_bad_scale(hot_path):
    from art_b200.api import HotPathError
    img = luminance(40, 40, 1)
    assert np.array_equal(hot_path.nlmeans(img.copy(), 65535.0, 0, 50, 1.0), img)
    with pytest.raises(HotPathError):
        hot_path.nlmeans(img.copy(), 65535.0, 50, 50, 0.05)


def test_nlmeans_throughput_report(hot_path, capsys):
    """Not a parity test: prints the device time of one 2048x1536 plane so the round log carries a number."""
    import torch
    W, H = 2048, 1536
    img = luminance(H, W, seed=11)
    d = torch.from_numpy(img).cuda()
    hot_path.nlmeans_dev(d.data_ptr(), W, W, H, 65535.0, 50, 50, 1.0)
    hot_path.sync()
    d.copy_(torch.from_numpy(img).cuda())
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    old = hot_path.get_stream() if hasattr(hot_path, "get_stream") else None
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    hot_path.nlmeans_dev(d.data_ptr(), W, W, H, 65535.0, 50, 50, 1.0)
    hot_path.sync()
    dt = time.perf_counter() - t0
    with capsys.disabled():
        print("\n[nlmeans] %dx%d: %.2f ms (%.1f Mpixel/s, host clock around launch+sync)" % (W, H, dt * 1e3, W * H / dt / 1e6))
