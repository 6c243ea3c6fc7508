"""GPU parity for the reference's DEFAULT tone-curve path through the C-ABI (art_hp_color_chain, tonecurve_mode 2 = NEUTRAL, and
the saturation curve): against the oracle port (pinned bit-exact to the reference's NeutralToneCurve::BatchApply / apply_satcurve in
tests/test_oracle_tone.py) on seeded frames, and against the committed golden vectors generated from the reference itself
(tests/golden/tone_film.npz: the Standard Film Curve profile).  Everything is restated operation by operation, so the bar is
bit-exact; the one exception is libm's powf on samples whose PQ argument leaves [0, 1] -- see close()."""
import os

import numpy as np
import pytest

import oracle
import tone_util as tu
from art_b200.api import ChainParams

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tone_film.npz")
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
S_CURVE = [4, 0, 0, 0.2, 0.12, 0.5, 0.55, 0.8, 0.9, 1, 1]


def gpu(hp, planes, **kw):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    hp.color_chain(out[0], out[1], out[2], ChainParams(ws=tu.PROPHOTO, iws=tu.PROPHOTO_INV, **kw))
    return out


def close(got, want, what):
    """Bit-identical but for the samples whose PQ argument leaves [0, 1]: there the reference calls libm's powf (color.cc L67-85) -- an
    external-library boundary like FFTW's: glibc's powf is within 0.8 ulp, not correctly rounded, and picks an FMA variant by CPU, so the
    reference's own bits differ between machines.  The device evaluates fp64 pow rounded once.  One ulp there is amplified by PQ's exponents
    (134 forward, 1 / 0.0075 in the inverse), and JzCzhz -> RGB spreads it over the pixel's channels: those samples are held to 1e-4 of the
    pixel's largest channel (north_star's float tolerance); at least 99 % of all samples must be bit-identical."""
    scale = np.maximum.reduce([np.abs(w.astype(np.float64)) for w in want])
    worst, differ, total = 0.0, 0, 0
    for g, w, ch in zip(got, want, "RGB"):
        ne = (g != w) & ~(np.isnan(g) & np.isnan(w))
        differ += int(ne.sum()); total += g.size
        if ne.any():
            err = np.abs(g[ne].astype(np.float64) - w[ne])
            lim = 1e-4 * scale[ne] + 0.02
            worst = max(worst, float((err / (scale[ne] + 0.02)).max()))
            bad = err > lim
            assert not bad.any(), "%s %s: %d samples beyond 1e-4 of the pixel scale, worst %g; (got, want, scale): %s" % (
                what, ch, int(bad.sum()), worst, [(float(a), float(b), float(c)) for a, b, c in zip(g[ne][bad][:4], w[ne][bad][:4], scale[ne][bad][:4])])
    assert differ <= 0.01 * total, "%s: %d of %d samples differ (more than the libm boundary explains)" % (what, differ, total)
    return worst


def test_golden_standard_film_curve(hot_path):
    z = np.load(GOLD)
    planes = list(z["inp"])
    stages = [(1, z["poly_last"][:1], z["poly_last"][1:], 0, 0, 0), (0, None, None, 0, 0, 0)]
    got = gpu(hot_path, planes, tonecurve=(2, z["lut"]), stages=stages)
    close(got, list(z["neutral"]), "neutral")
    got = gpu(hot_path, planes, tonecurve=(2, z["lut"]), stages=stages, satcurve=z["satlut"])
    close(got, list(z["sat"]), "neutral + satcurve")
    stages_c = [(2, None, None, z["contrast_ab"][0], z["contrast_ab"][1], 1.0)] + stages
    got = gpu(hot_path, planes, tonecurve=(2, z["lut_contrast"]), stages=stages_c, to_out=z["to_out"], to_work=z["to_work"])
    close(got, list(z["neutral_contrast"]), "neutral, contrast 30, sRGB output matrix")


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 5), (3, 3), (1021, 300)])
@pytest.mark.parametrize("curve,curve2,contrast", [(tu.FILM_CURVE, tu.LINEAR, 0), (tu.LINEAR, tu.LINEAR, 0), (tu.FILM_CURVE, S_CURVE, 0),
                                                   (tu.FILM_CURVE, tu.LINEAR, 35), (tu.LINEAR, tu.LINEAR, -20)])
@pytest.mark.parametrize("whitept", [1.0, 2.5])
@pytest.mark.parametrize("om", [None, tu.SRGB_XYZ])
def test_neutral_matches_oracle(hot_path, W, H, curve, curve2, contrast, whitept, om):
    planes = tu.frame(H, W, W * 7 + H + contrast)
    lut, _ = tu.build_lut(curve, curve2, contrast, whitept)
    stages = tu.stages_for(curve, curve2, contrast, whitept)
    to_out, to_work = tu.out_matrices(tu.PROPHOTO, tu.PROPHOTO_INV, om)
    want = tu.port_neutral(planes, lut, whitept, stages, to_out=to_out, to_work=to_work)
    got = gpu(hot_path, planes, tonecurve=(2, lut), whitept=whitept, stages=stages, to_out=to_out, to_work=to_work)
    close(got, want, "neutral")


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 5), (1021, 300)])
@pytest.mark.parametrize("mode", [0, 1])
def test_std_and_filmlike_with_white_point(hot_path, W, H, mode):
    """STD / FILMLIKE with a white point above 1: filmlike_clip at 65535 * whitept, the LUT above it clipped (no curve stages)"""
    import ctypes
    planes = tu.frame(H, W, W + H + mode)
    lut, _ = tu.build_lut(tu.FILM_CURVE, tu.LINEAR, 0, 2.0)
    want = [p.copy() for p in planes]
    assert oracle.port().lib.artoracle_chain_tonecurve(*[p.ctypes.data_as(tu.fp) for p in want], W, H, mode, lut.ctypes.data_as(tu.fp), ctypes.c_float(2.0)) == 0
    got = gpu(hot_path, planes, tonecurve=(mode, lut), whitept=2.0)
    for g, w in zip(got, want):
        assert np.array_equal(g, w, equal_nan=True)


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (67, 5), (1021, 300)])
@pytest.mark.parametrize("sat", [tu.FILM_SAT, [1, 0, 0.2, 0.35, 0.35, 0.5, 0.8, 0.35, 0.35, 1, 0.4, 0.35, 0.35]])
def test_satcurve_matches_oracle(hot_path, W, H, sat):
    planes = tu.frame(H, W, W + 11 * H)
    lut = tu.sat_lut(sat)
    close(gpu(hot_path, planes, satcurve=lut), tu.port_satcurve(planes, lut), "satcurve")


@needs_ref
def test_default_chain_order(hot_path):
    """exposure -> saturation -> NEUTRAL curve -> satcurve -> rgb curves -> Lab, fused, against the stages applied one by one"""
    from test_chain_gpu import lab_luts, oracle_chain
    from test_oracle_chain import curve_lut
    planes = tu.frame(130, 203, 5)
    lut, _ = tu.build_lut(tu.FILM_CURVE, tu.LINEAR)
    stages = tu.stages_for(tu.FILM_CURVE, tu.LINEAR)
    satl = tu.sat_lut(tu.FILM_SAT)
    rgbc = [curve_lut(gamma=0.9, seed=i) for i in range(3)]
    lab = lab_luts() + (1.2,)
    want = oracle_chain(planes, exposure=(0.4, 0.0), saturation=(20, 10))
    want = tu.port_satcurve(tu.port_neutral(want, lut, 1.0, stages), satl)
    want = oracle_chain(want, rgbcurves=rgbc, lab=lab)
    got = gpu(hot_path, planes, exposure=(0.4, 0.0), saturation=(20, 10), tonecurve=(2, lut), stages=stages, satcurve=satl, rgbcurves=rgbc, lab=lab)
    close(got, want, "whole chain")
