"""The host-side mirror of ImProcFunctions::process (improcfun.cc L567-641) on the CPU: a recording stand-in for art_b200.HotPath shows the order
in which the steps of a stage reach the C-ABI, that disabled steps are skipped like the reference's `enabled` early-outs, that STAGE_3's per-pixel
steps go down as ONE fused chain call, and that an enabled step which is not on the hot path fails loudly."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import art_b200
from art_b200.improcfun import ImProcFunctions, OUTPUT, PREVIEW, STAGE_0, STAGE_1, STAGE_2, STAGE_3, THUMBNAIL


class Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def entry(*args, **kw):
            self.calls.append((name, args[3:]))
        return entry


def planes():
    return [np.zeros((4, 8), np.float32) for _ in range(3)]


LUT = np.linspace(0, 65535, 65536).astype(np.float32)
WS = np.eye(3)


def full_params():
    return NS(
        fattal=NS(enabled=True, threshold=30, amount=20, satcontrol=False),
        chmixer=NS(enabled=True, matrix=[1, 0, 0, 0, 1, 0, 0, 0, 1]),
        exposure=NS(enabled=True, expcomp=0.5, black=0.01),
        hsl=NS(enabled=True, params="HSL"), toneEqualizer=NS(enabled=True, params="TEQ"), workingProfile="ProPhoto",
        sharpening=NS(enabled=True, params="USM"),
        saturation=NS(enabled=True, saturation=20, vibrance=10),
        toneCurve=NS(enabled=True, mode=2, lut=LUT, whitept=1.0, stages=None, satcurve=LUT, to_out=None, to_work=None),
        rgbCurves=NS(enabled=True, luts=[LUT, None, LUT]), labCurve=NS(enabled=True, lcurve=LUT[:32770], acurve=LUT, bcurve=LUT, chroma=1.1),
        softlight=NS(enabled=True, lut=LUT), blackwhite=NS(enabled=True, params="BW"))


def test_stage_order_matches_the_reference():
    rec = Recorder()
    ipf = ImProcFunctions(full_params(), rec, 1.0, WS, WS)
    r, g, b = planes()
    assert ipf.process(OUTPUT, STAGE_0, r, g, b) is False
    assert [c[0] for c in rec.calls] == ["fattal"] and rec.calls[0][1][:3] == (30, 20, False)
    rec.calls.clear()
    ipf.process(OUTPUT, STAGE_1, r, g, b)
    assert [c[0] for c in rec.calls] == ["channel_mixer", "color_chain", "hsl_equalizer", "tone_equalizer", "prophoto_blue"]     # improcfun.cc L581-587
    exp = rec.calls[1][1][0]
    assert exp.exposure == (0.5, 0.01) and exp.tonecurve is None and exp.saturation is None       # the exposure step alone
    assert rec.calls[2][1] == ("HSL",) and rec.calls[3][1] == ("TEQ",)
    rec.calls.clear()
    ipf.process(OUTPUT, STAGE_2, r, g, b)
    assert rec.calls == [("sharpen_usm", ("USM", WS))]
    rec.calls.clear()
    ipf.process(OUTPUT, STAGE_3, r, g, b)
    assert [c[0] for c in rec.calls] == ["color_chain", "black_and_white"]                        # five per-pixel steps fused, then blackAndWhite (L611-628)
    ch = rec.calls[0][1][0]
    assert ch.saturation == (20, 10) and ch.tonecurve[0] == 2 and ch.satcurve is LUT and ch.rgbcurves[1] is None and ch.lab[3] == 1.1 and ch.softlight is LUT
    assert ch.exposure is None and ch.ws is WS and ch.iws is WS


def test_disabled_steps_are_skipped_and_thumbnails_are_not_sharpened():
    rec = Recorder()
    p = full_params()
    p.chmixer.enabled = p.hsl.enabled = p.blackwhite.enabled = p.saturation.enabled = p.softlight.enabled = False
    p.workingProfile = "Rec2020"
    ipf = ImProcFunctions(p, rec, 1.0, WS, WS)
    r, g, b = planes()
    ipf.process(OUTPUT, STAGE_1, r, g, b)
    assert [c[0] for c in rec.calls] == ["color_chain", "tone_equalizer"]
    rec.calls.clear()
    ipf.process(THUMBNAIL, STAGE_2, r, g, b)                  # L594: sharpening only in the OUTPUT and PREVIEW pipelines
    assert rec.calls == []
    ipf.process(PREVIEW, STAGE_3, r, g, b)
    ch = rec.calls[0][1][0]
    assert [c[0] for c in rec.calls] == ["color_chain"] and ch.saturation is None and ch.softlight is None and ch.tonecurve is not None
    rec.calls.clear()
    for s in ("toneCurve", "rgbCurves", "labCurve"):
        getattr(p, s).enabled = False
    ipf.process(OUTPUT, STAGE_3, r, g, b)
    assert rec.calls == []                                     # nothing enabled: no launch at all


@pytest.mark.parametrize("stage,step", [(STAGE_0, "dehaze"), (STAGE_2, "impulseDenoise"), (STAGE_2, "defringe"), (STAGE_2, "colorcorrection"),
                                        (STAGE_2, "smoothing"), (STAGE_3, "gradient"), (STAGE_3, "textureBoost"), (STAGE_3, "grain"), (STAGE_3, "logenc"),
                                        (STAGE_3, "filmSimulation"), (STAGE_3, "localContrast")])
def test_enabled_steps_off_the_hot_path_fail_loudly(stage, step):
    rec = Recorder()
    p = full_params()
    setattr(p, step, NS(enabled=True))
    with pytest.raises(art_b200.HotPathError):
        ImProcFunctions(p, rec, 1.0, WS, WS).process(OUTPUT, stage, *planes())
    assert rec.calls == []                                     # refused before anything ran
    setattr(p, step, NS(enabled=False))
    ImProcFunctions(p, rec, 1.0, WS, WS).process(OUTPUT, stage, *planes())
    assert rec.calls
