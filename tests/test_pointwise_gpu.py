"""GPU parity for the per-pixel stages (through the C-ABI) against the oracle.  Bit-exact."""
import numpy as np
import pytest

import art_b200
import oracle

pytestmark = pytest.mark.gpu

MAT = np.array([[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0000000, 0.0000000, 1.2118128]]) @ \
      np.array([[0.6594, 0.2521, 0.0528], [0.2661, 0.9712, -0.2373], [0.0292, -0.2046, 1.0003]])


def planes(H, W, seed, lo=-500.0, hi=80000.0):
    rng = np.random.default_rng(seed)
    return [rng.uniform(lo, hi, size=(H, W)).astype(np.float32) for _ in range(3)]


@pytest.mark.parametrize("W,H", [(640, 480), (401, 33), (7, 5), (1003, 257)])
@pytest.mark.parametrize("do_clip", [0, 1])
@pytest.mark.parametrize("use_mat", [False, True])
def test_scale_convert_host(hot_path, W, H, do_clip, use_mat):
    p = planes(H, W, seed=W + H)
    mul = (1.9371, 1.0, 1.4182)
    want = oracle.port().scale_convert(p, mul, do_clip, MAT if use_mat else None)
    got = [q.copy() for q in p]
    hot_path.scale_convert(got[0], got[1], got[2], mul, do_clip, MAT if use_mat else None)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_scale_convert_dev_unaligned(hot_path):
    """Device entry on planes whose pitch/base are not 16-byte aligned (scalar kernel path)."""
    import torch
    H, W, pitch = 50, 333, 335
    p = planes(H, W, seed=9)
    mul = (0.7, 1.0, 2.3)
    want = oracle.port().scale_convert(p, mul, 1, MAT)
    dev = []
    for q in p:
        t = torch.zeros(H * pitch + 1, dtype=torch.float32, device="cuda")
        v = t[1:].view(H, pitch)
        v[:, :W] = torch.from_numpy(q).cuda()
        dev.append((t, v))
    hot_path.scale_convert_dev(W, H, dev[0][1].data_ptr(), dev[1][1].data_ptr(), dev[2][1].data_ptr(), pitch, mul, 1, MAT)
    hot_path.sync()
    for (t, v), w in zip(dev, want):
        assert np.array_equal(v[:, :W].cpu().numpy(), w)


def test_demosaic_then_convert_resident(hot_path):
    """The resident chain the pipeline uses: demosaic -> gain/clip/3x3 without leaving the GPU."""
    import torch
    from art_b200 import synth
    f = synth.RGGB
    W, H = 512, 384
    raw = synth.bayer_frame(W, H, f, seed=5)
    d_raw = torch.from_numpy(raw).cuda()
    outs = [torch.empty((H, W), dtype=torch.float32, device="cuda") for _ in range(3)]
    hot_path.demosaic_bayer_dev(art_b200.BAYER_AMAZE, W, H, f, d_raw.data_ptr(), W, outs[0].data_ptr(), outs[1].data_ptr(),
                                outs[2].data_ptr(), W, 1.0, 4)
    mul = (2.0, 1.0, 1.5)
    hot_path.scale_convert_dev(W, H, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), W, mul, 1, MAT)
    hot_path.sync()
    want = oracle.port().scale_convert(oracle.port().amaze(raw, f), mul, 1, MAT)
    for o, w in zip(outs, want):
        assert np.array_equal(o.cpu().numpy(), w)


@pytest.mark.parametrize("pattern", ["RGGB", "GBRG"])
@pytest.mark.parametrize("W,H", [(640, 480), (333, 77)])
def test_scale_colors_bayer(hot_path, pattern, W, H):
    from art_b200 import synth
    f = synth.BAYER_FILTERS[pattern]
    rng = np.random.default_rng(W)
    raw = rng.integers(0, 16384, size=(H, W)).astype(np.float32)
    black = (511.0, 512.5, 510.0, 513.25)
    mul = (8.1234, 4.0625, 6.3321, 4.0631)
    want, wmax = oracle.port().scale_colors_bayer(raw, f, black, mul)
    got = raw.copy()
    gmax = hot_path.scale_colors_bayer(got, f, black, mul)
    assert np.array_equal(got, want) and gmax == wmax


@pytest.mark.parametrize("dy,dx", [(0, 0), (2, 5), (4, 1)])
@pytest.mark.parametrize("W,H", [(640, 480), (333, 77), (6240, 417)])
def test_scale_colors_xtrans(hot_path, dy, dx, W, H):
    """scaleColors' X-Trans branch (rawimagesource.cc L2795-2826) through the C-ABI against the oracle (pinned in test_oracle_pointwise.py)."""
    from art_b200 import synth
    from test_oracle_pointwise import scale_colors_xtrans
    xt = synth.xtrans_matrix(dy, dx)
    rng = np.random.default_rng(W + dy)
    raw = rng.integers(0, 16384, size=(H, W)).astype(np.float32)
    black = (1023.0, 1024.5, 1022.0)
    mul = (7.9, 4.0625, 6.3321)
    want, wmax = scale_colors_xtrans(oracle.port().lib, "artoracle_scale_colors_xtrans", raw, xt, black + (0.0,), mul + (0.0,))
    got = raw.copy()
    gmax = hot_path.scale_colors_xtrans(got, xt, black, mul)
    assert np.array_equal(got, want) and gmax == wmax
