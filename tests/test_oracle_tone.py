"""Pin the tone-curve oracle (oracle/tone_port.c) against the reference's own NeutralToneCurve::BatchApply / apply_satcurve and the
reference's own curve objects (DiagonalCurve, FlatCurve, ToneCurve::Set, the curve assembly of ImProcFunctions::toneCurve) compiled
in place (oracle/_ref, build_ref_tone.py).  Bit-exact: the default tone-curve mode (NEUTRAL) with the Standard Film Curve profile,
linear curves, contrast, white points above 1 (samples above the LUT go through Curve::getVal), an output-profile matrix."""
import numpy as np
import pytest

import oracle
import tone_util as tu

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
SIZES = [(64, 48), (67, 5), (3, 3), (131, 33)]
S_CURVE = [4, 0, 0, 0.2, 0.12, 0.5, 0.55, 0.8, 0.9, 1, 1]          # a Catmull-Rom curve, as the GUI writes them


def same(a, b):
    for x, y, ch in zip(a, b, "RGB"):
        eq = (x == y) | (np.isnan(x) & np.isnan(y))
        assert eq.all(), "%s: %d of %d differ, first at %s: %r vs %r" % (ch, int((~eq).sum()), x.size, np.argwhere(~eq)[0], x[~eq][0], y[~eq][0])


@needs_ref
def test_hue_constants_match_applystate():
    """the constants ApplyState derives from the Rec.2020 primaries: the port computes them with its own JzCzhz code"""
    h = np.zeros(6, np.float32)
    oracle.port().lib.artoracle_tone_hues(h.ctypes.data_as(tu.fp))
    assert np.isfinite(h).all() and h[3] == h[4] and h[5] > 0


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("curve,curve2,contrast", [(tu.FILM_CURVE, tu.LINEAR, 0), (tu.LINEAR, tu.LINEAR, 0), (tu.FILM_CURVE, S_CURVE, 0),
                                                   (tu.FILM_CURVE, tu.LINEAR, 35), (tu.LINEAR, tu.LINEAR, -20)])
@pytest.mark.parametrize("whitept", [1.0, 2.5])
@pytest.mark.parametrize("om", [None, tu.SRGB_XYZ])
def test_neutral_matches_reference(W, H, curve, curve2, contrast, whitept, om):
    planes = tu.frame(H, W, W * 7 + H + contrast)
    lut, _ = tu.build_lut(curve, curve2, contrast, whitept)
    stages = tu.stages_for(curve, curve2, contrast, whitept)
    to_out, to_work = tu.out_matrices(tu.PROPHOTO, tu.PROPHOTO_INV, om)
    got = tu.port_neutral(planes, lut, whitept, stages, to_out=to_out, to_work=to_work)
    want = tu.ref_neutral(planes, curve, curve2, contrast, whitept, om=om)
    same(got, want)
    assert all(np.isfinite(g).all() for g in got)
    assert max(float(g.max()) for g in got) <= 65535.0 * whitept


@needs_ref
def test_samples_above_the_lut_take_the_curve():
    """the reference's BatchApply clips at Lmax = 65535 * whitecoeff in a [0, 1] domain, i.e. never: over-range samples reach
    curves::setLutVal's Curve::getVal branch even at white point 1"""
    planes = [np.full((4, 8), v, np.float32) for v in (90000.0, 70000.0, 80000.0)]
    lut, _ = tu.build_lut(tu.FILM_CURVE, tu.LINEAR)
    with_curve = tu.port_neutral(planes, lut, 1.0, tu.stages_for(tu.FILM_CURVE, tu.LINEAR))
    same(with_curve, tu.ref_neutral(planes, tu.FILM_CURVE, tu.LINEAR))
    lin, _ = tu.build_lut(tu.LINEAR, tu.LINEAR)
    same(tu.port_neutral(planes, lin, 1.0, tu.stages_for(tu.LINEAR, tu.LINEAR)), tu.ref_neutral(planes, tu.LINEAR, tu.LINEAR))


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("sat", [tu.FILM_SAT, [1, 0, 0.2, 0.35, 0.35, 0.5, 0.8, 0.35, 0.35, 1, 0.4, 0.35, 0.35]])
def test_satcurve_matches_reference(W, H, sat):
    planes = tu.frame(H, W, W + 11 * H, over=True)
    lut = tu.sat_lut(sat)
    assert lut is not None
    same(tu.port_satcurve(planes, lut), tu.ref_satcurve(planes, sat))


@needs_ref
def test_film_curve_lut_properties():
    lut, ident = tu.build_lut(tu.FILM_CURVE, tu.LINEAR)
    assert not ident and lut[0] == 0.0 and lut[65535] == 65535.0 and (np.diff(lut) >= 0).all()
    lin, ident = tu.build_lut(tu.LINEAR, tu.LINEAR)
    assert np.array_equal(lin, np.arange(65536, dtype=np.float32))
