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
