"""GPU parity for unsharp-mask sharpening (art_hp_sharpen_usm = ImProcFunctions::sharpening, method "usm") through the C-ABI
against the oracle port, which test_oracle_usm.py pins bit-exact to the reference functions compiled in place.  Bit-exact."""
import numpy as np
import pytest

import oracle
from art_b200.api import SharpenParams
from test_oracle_usm import CASES, PROPHOTO, SIZES, run, same, scene

pytestmark = pytest.mark.gpu


def gpu(hp, planes, **kw):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    kw = dict(kw)
    if "thr" in kw:
        kw["threshold"] = kw.pop("thr")
    if "halo" in kw:
        kw["halocontrol"] = bool(kw.pop("halo"))
    if "halo_amount" in kw:
        kw["halocontrol_amount"] = kw.pop("halo_amount")
    hp.sharpen_usm(out[0], out[1], out[2], SharpenParams(**kw), PROPHOTO)
    return out


@pytest.mark.parametrize("W,H", SIZES + [(1023, 517)])
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("wild", [False, True])
def test_usm_matches_oracle(hot_path, W, H, case, wild):
    planes = scene(W, H, W * 3 + H + case, wild)
    want, _ = run(oracle.port().lib, "artoracle_usm", planes, **CASES[case])
    same(gpu(hot_path, planes, **CASES[case]), want)


def test_usm_too_small_or_disabled_is_identity(hot_path):
    planes = scene(7, 40, 1)
    same(gpu(hot_path, planes), planes)
    planes = scene(64, 40, 2)
    same(gpu(hot_path, planes, amount=0), planes)


def test_usm_rejects_bad_edges_tolerance(hot_path):
    import art_b200
    planes = scene(64, 40, 3)
    with pytest.raises(art_b200.HotPathError) as e:
        gpu(hot_path, planes, edgesonly=True, edges_tolerance=0)
    assert e.value.code == 1


from test_oracle_usm import EDGES_CASES, run_edges  # noqa: E402


@pytest.mark.parametrize("W,H", [(64, 48), (9, 8), (12, 11), (130, 77), (1023, 517)])
@pytest.mark.parametrize("case", range(len(EDGES_CASES)))
def test_usm_edgesonly_matches_oracle(hot_path, W, H, case):
    """edgesonly: every one of bilateral2.h's 21 kernels, the copy below sigma 0.45, both sides of the dispatch thresholds, range
    sigma extremes, with and without halo control; the range LUT is rebuilt when (kernel scale, tolerance) changes between calls."""
    planes = scene(W, H, W * 5 + H + case, wild=bool(case % 3 == 1))
    want = run_edges(oracle.port().lib, "artoracle_usm_ex", planes, **EDGES_CASES[case])
    same(gpu(hot_path, planes, edgesonly=True, **EDGES_CASES[case]), want)


def test_usm_device_form_with_pitch(hot_path):
    torch = pytest.importorskip("torch")
    W, H, pitch = 203, 141, 224
    planes = scene(W, H, 9)
    dev = [torch.zeros((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    for d, p in zip(dev, planes):
        d[:, :W] = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    hot_path.sharpen_usm_dev(W, H, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), pitch, SharpenParams(radius=0.9, amount=300), PROPHOTO)
    hot_path.sync()
    want, _ = run(oracle.port().lib, "artoracle_usm", planes, radius=0.9, amount=300)
    same([d[:, :W].cpu().numpy() for d in dev], want)


# ---- "rld" route
from test_oracle_usm import RLD_CASES, run_rld  # noqa: E402


def gpu_rld(hp, planes, **kw):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    p = SharpenParams(method="rld", contrast=kw.get("contrast", 20.0), deconvradius=kw.get("radius", 0.75), deconvamount=kw.get("amount", 100),
                      scale=kw.get("scale", 1.0), deconvCornerBoost=kw.get("boost", 0.0), deconvCornerLatitude=kw.get("latitude", 25),
                      offset_x=kw.get("ox", 0), offset_y=kw.get("oy", 0), full_width=kw.get("fw", 0), full_height=kw.get("fh", 0))
    hp.sharpen_usm(out[0], out[1], out[2], p, PROPHOTO)
    return out


@pytest.mark.parametrize("W,H", SIZES + [(1023, 517)])
@pytest.mark.parametrize("case", [c for c in range(len(RLD_CASES)) if RLD_CASES[c].get("radius", 0.75) != 0.22])
@pytest.mark.parametrize("wild", [False, True])
def test_rld_matches_oracle(hot_path, W, H, case, wild):
    planes = scene(W, H, W * 7 + H + case, wild)
    want, _ = run_rld(oracle.port().lib, "artoracle_rld", planes, **RLD_CASES[case])
    same(gpu_rld(hot_path, planes, **RLD_CASES[case]), want)


def test_rld_rejects_large_sigma(hot_path):
    import art_b200
    planes = scene(64, 40, 3)
    with pytest.raises(art_b200.HotPathError) as e:
        gpu_rld(hot_path, planes, radius=30.0)       # sigma >= 25 would take gaussianBlur's all-double branch
    assert e.value.code == 5


from test_oracle_usm import BOOST_CASES, run_rld_ex  # noqa: E402


@pytest.mark.parametrize("W,H", [(64, 48), (301, 203), (130, 77), (1023, 517)])
@pytest.mark.parametrize("case", range(len(BOOST_CASES)))
def test_rld_corner_boost_matches_oracle(hot_path, W, H, case):
    planes = scene(W, H, W * 11 + H + case, wild=bool(case & 1))
    same(gpu_rld(hot_path, planes, **BOOST_CASES[case]), run_rld_ex(oracle.port().lib, "artoracle_rld_ex", planes, **BOOST_CASES[case]))
