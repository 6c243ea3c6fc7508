"""GPU parity for art_hp_hsl_equalizer (ImProcFunctions::hslEqualizer, iphsl.cc L29-221) through the C-ABI against oracle/hsl_port.c,
which tests/test_oracle_hsl.py pins bit-exact to the reference's own function body.  Bit-exact."""
import numpy as np
import pytest

from art_b200.api import HslParams
from test_oracle_hsl import CASES, COEFF, PROPHOTO, image, polyline, port_hsl, same

pytestmark = pytest.mark.gpu


def gpu_hsl(hp, planes, hc, sc, lc, smoothing, scale):
    pn = int(1000 / scale)
    cv = [polyline(c, True, pn) for c in (hc, sc, lc)] + [polyline(COEFF, True, 1000)]
    as_arg = lambda c: None if c[0] == 0 else (c[1], c[2], c[3])
    out = [p.copy() for p in planes]
    hp.hsl_equalizer(out[0], out[1], out[2], HslParams(hcurve=as_arg(cv[0]), scurve=as_arg(cv[1]), lcurve=as_arg(cv[2]), coeff=as_arg(cv[3]),
                                                        smoothing=smoothing, scale=scale, ws=PROPHOTO))
    return out


@pytest.mark.parametrize("W,H", [(96, 64), (131, 77), (300, 201), (1021, 403)])
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("smoothing,scale", [(0, 1.0), (5, 1.0), (10, 1.0), (7, 2.0)])
def test_hsl_matches_oracle(hot_path, W, H, case, smoothing, scale):
    planes = image(H, W, W + H + smoothing)
    hc, sc, lc = CASES[case]
    same(gpu_hsl(hot_path, planes, hc, sc, lc, smoothing, scale), port_hsl(planes, hc, sc, lc, smoothing, scale))


def test_hsl_large_frame(hot_path):
    planes = image(2000, 3008, 5)
    hc, sc, lc = CASES["all"]
    same(gpu_hsl(hot_path, planes, hc, sc, lc, 5, 1.0), port_hsl(planes, hc, sc, lc, 5, 1.0))


def test_hsl_device_form_with_pitch(hot_path):
    torch = pytest.importorskip("torch")
    W, H, pitch = 203, 141, 224
    planes = image(H, W, 99)
    dev = [torch.zeros((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    for d, p in zip(dev, planes):
        d[:, :W] = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    hc, sc, lc = CASES["all"]
    cv = [polyline(c, True, 1000) for c in (hc, sc, lc, COEFF)]
    hot_path.hsl_equalizer_dev(W, H, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), pitch,
                               HslParams(*[(c[1], c[2], c[3]) for c in cv], smoothing=4, scale=1.0, ws=PROPHOTO))
    hot_path.sync()
    same([d[:, :W].cpu().numpy() for d in dev], port_hsl(planes, hc, sc, lc, 4, 1.0))


def test_hsl_rejects_bad_parameters(hot_path):
    import art_b200
    planes = image(16, 16, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.hsl_equalizer(planes[0], planes[1], planes[2], HslParams(smoothing=1, scale=1.0))                 # no ws
    n, px, py, dy = polyline(CASES["all"][1])
    with pytest.raises(art_b200.HotPathError):
        hot_path.hsl_equalizer(planes[0], planes[1], planes[2], HslParams(scurve=(px, py, dy), ws=PROPHOTO))       # S curve without coeff
