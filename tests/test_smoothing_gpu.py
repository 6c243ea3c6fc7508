"""GPU parity for denoise::denoiseGuidedSmoothing (art_hp_denoise_guided_smoothing) through the C-ABI against the oracle port, which
tests/test_oracle_smoothing.py pins bit-exact to the reference's own guided_smoothing / guidedFilterLog / guidedFilter compiled in
place.  Bit-exact: the log encoding and the chroma transfer follow the reference's scalar sleef forms, the guided filter is
guided.cu's (bit-exact on its own in test_guided_gpu.py)."""
import numpy as np
import pytest

from test_develop_gpu import guided_smoothing
from test_oracle_denoise import PROPHOTO, rgb_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(64, 48), (67, 53), (264, 200), (301, 203), (1203, 807)])
@pytest.mark.parametrize("radius,scale", [(3, 1.0), (1, 1.0), (10, 1.0), (3, 2.0), (3, 8.0), (0, 1.0)])
def test_guided_smoothing_matches_oracle(hot_path, W, H, radius, scale):
    planes = rgb_frame(H, W, seed=W + radius, noise=2500.0, hot=(W % 2 == 1))
    for p in planes:
        p[H // 2, W // 3] = 0.0                             # a black pixel: the bump guard
    want = guided_smoothing(planes, radius, scale)
    got = [p.copy() for p in planes]
    hot_path.denoise_guided_smoothing(got[0], got[1], got[2], PROPHOTO, radius, scale)
    for g, w, ch in zip(got, want, "RGB"):
        eq = (g == w) | (np.isnan(g) & np.isnan(w))
        assert eq.all(), "%s: %d of %d differ, max abs %g" % (ch, int((~eq).sum()), g.size, float(np.nanmax(np.abs(g - w))))
    if radius == 0:
        assert all(np.array_equal(g, p) for g, p in zip(got, planes))
