"""GPU parity for art_hp_tone_equalizer (ImProcFunctions::toneEqualizer, iptoneequalizer.cc) through the C-ABI against oracle/toneeq_port.c,
which tests/test_oracle_toneeq.py pins bit-exact to the reference's own tone_eq().  Bit-exact."""
import numpy as np
import pytest

from art_b200.api import ToneEqParams
from test_oracle_toneeq import BANDS, CASES, PROPHOTO, image, port_teq, same

pytestmark = pytest.mark.gpu


def gpu_teq(hp, planes, bands, regularization, pivot, scale):
    out = [p.copy() for p in planes]
    hp.tone_equalizer(out[0], out[1], out[2], ToneEqParams(bands, regularization, pivot, scale, PROPHOTO))
    return out


@pytest.mark.parametrize("W,H,regularization,pivot,scale", CASES + [(2003, 1501, 2, 0.5, 1.0)])
@pytest.mark.parametrize("bands", BANDS)
def test_toneeq_matches_oracle(hot_path, W, H, bands, regularization, pivot, scale):
    planes = image(H, W, W + H + regularization)
    same(gpu_teq(hot_path, planes, bands, regularization, pivot, scale), port_teq(planes, bands, regularization, pivot, scale))


def test_toneeq_rejects_bad_parameters(hot_path):
    import art_b200
    planes = image(64, 96, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.tone_equalizer(planes[0], planes[1], planes[2], ToneEqParams((1, 2, 3, 4, 5), 0, 0.0, 1.0, None))            # no ws
    with pytest.raises(art_b200.HotPathError):
        hot_path.tone_equalizer(planes[0], planes[1], planes[2], ToneEqParams((1, 2, 3, 4, 5), 1, 0.0, 0.0, PROPHOTO))        # scale 0


def test_toneeq_device_form_with_unaligned_pitch(hot_path):
    """the device entry on planes whose rows are not 16-byte aligned (pitch W + 1): the scalar-access path of k_teq_apply"""
    import ctypes
    torch = pytest.importorskip("torch")
    W, H, pitch = 203, 141, 205
    planes = image(H, W, 77)
    dev = [torch.zeros((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    for d, p in zip(dev, planes):
        d[:, :W] = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    c = ToneEqParams((40, 25, 0, -20, -35), 1, 0.5, 1.0, PROPHOTO).c_struct()
    hot_path._check(hot_path.lib.art_hp_tone_equalizer_dev(hot_path.h, W, H, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), pitch, ctypes.byref(c)))
    hot_path.sync()
    same([d[:, :W].cpu().numpy() for d in dev], port_teq(planes, (40, 25, 0, -20, -35), 1, 0.5, 1.0))
