"""Host-side mirror of ImProcFunctions::process for the steps the hot path covers.

Reference: rtengine/improcfun.cc L567-641 -- process(pipeline, stage, img) runs a fixed list of steps per stage, each a member that
returns at once when its `enabled` flag is off.  This mirror keeps the stage names, the ORDER of the steps and the enabled tests, and
sends each step to its C-ABI entry through art_b200.HotPath; contiguous per-pixel steps of STAGE_3 (saturationVibrance, toneCurve,
rgbCurves, labAdjustments, softLight) go down as ONE fused art_hp_color_chain call, which is the point of that entry.  A step of the
reference that is enabled but not on the hot path raises HotPathError(ART_HP_ERR_UNSUPPORTED): nothing is silently skipped and
nothing falls back to a CPU path here (the reference's caller keeps those steps on its side).

`params` is any object with the attributes used below (a types.SimpleNamespace in the tests), named after procparams.h:
  dehaze, fattal (enabled, threshold, amount, satcontrol), chmixer (enabled, matrix: 9 floats as ipchmixer.cc L156-164 computes them),
  exposure (enabled, expcomp, black), hsl (enabled, params: api.HslParams), toneEqualizer (enabled, params: api.ToneEqParams),
  workingProfile, sharpening (enabled, params: api.SharpenParams), impulseDenoise, defringe, colorcorrection, smoothing, gradient,
  pcvignette, textureBoost, grain, logenc, saturation (enabled, saturation, vibrance), dcp_look, filmSimulation,
  toneCurve (enabled, mode, lut, whitept, stages, satcurve, to_out, to_work), rgbCurves (enabled, luts), labCurve (enabled, lcurve,
  acurve, bcurve, chroma), softlight (enabled, lut), localContrast, blackwhite (enabled, params: api.BwParams).
Host-built inputs (curves, LUTs, matrices) arrive ready-made, as everywhere at this boundary.
"""
from . import api

UNSUPPORTED = 5      # ART_HP_ERR_UNSUPPORTED
STAGE_0, STAGE_1, STAGE_2, STAGE_3 = range(4)                       # ImProcFunctions::Stage
OUTPUT, PREVIEW, THUMBNAIL = "OUTPUT", "PREVIEW", "THUMBNAIL"       # ImProcFunctions::Pipeline


def _on(params, name):
    step = getattr(params, name, None)
    return bool(step is not None and getattr(step, "enabled", False))


class ImProcFunctions:
    def __init__(self, params, hot_path, scale=1.0, ws=None, iws=None):
        self.params, self._hp, self.scale, self.ws, self.iws = params, hot_path, float(scale), ws, iws
        self.cur_pipeline = OUTPUT

    def _off_path(self, *names):
        for n in names:
            if _on(self.params, n):
                raise api.HotPathError(UNSUPPORTED, "ImProcFunctions::%s is enabled and not on the hot path" % n)

    # ---- the steps, named as in improcfun.h
    def dynamicRangeCompression(self, r, g, b):
        p = self.params.fattal
        self._hp.fattal(r, g, b, p.threshold, p.amount, getattr(p, "satcontrol", False), self.ws)

    def channelMixer(self, r, g, b):
        self._hp.channel_mixer(r, g, b, self.params.chmixer.matrix)

    def exposure(self, r, g, b):
        p = self.params.exposure
        self._hp.color_chain(r, g, b, api.ChainParams(exposure=(p.expcomp, p.black)))

    def hslEqualizer(self, r, g, b):
        self._hp.hsl_equalizer(r, g, b, self.params.hsl.params)

    def toneEqualizer(self, r, g, b):
        self._hp.tone_equalizer(r, g, b, self.params.toneEqualizer.params)

    def sharpening(self, r, g, b):
        self._hp.sharpen_usm(r, g, b, self.params.sharpening.params, self.ws)

    def blackAndWhite(self, r, g, b):
        self._hp.black_and_white(r, g, b, self.params.blackwhite.params)

    def _stage3_chain(self, r, g, b):
        """saturationVibrance, toneCurve, rgbCurves, labAdjustments, softLight (improcfun.cc L611-625) in one fused pass"""
        P, kw = self.params, {}
        if _on(P, "saturation"):
            kw["saturation"] = (P.saturation.saturation, P.saturation.vibrance)
        if _on(P, "toneCurve"):
            t = P.toneCurve
            kw["tonecurve"] = (t.mode, t.lut)
            for name in ("whitept", "stages", "to_out", "to_work", "satcurve"):
                if getattr(t, name, None) is not None:
                    kw[name] = getattr(t, name)
        if _on(P, "rgbCurves"):
            kw["rgbcurves"] = P.rgbCurves.luts
        if _on(P, "labCurve"):
            c = P.labCurve
            kw["lab"] = (c.lcurve, c.acurve, c.bcurve, c.chroma)
        if _on(P, "softlight"):
            kw["softlight"] = P.softlight.lut
        if kw:
            self._hp.color_chain(r, g, b, api.ChainParams(ws=self.ws, iws=self.iws, **kw))

    def process(self, pipeline, stage, r, g, b):
        """ImProcFunctions::process(pipeline, stage, img), improcfun.cc L567-641, in place on three host planes; returns `stop` (always False: the
        steps that can stop the pipeline -- the preview colour maps -- are not on the hot path)."""
        self.cur_pipeline = pipeline
        P = self.params
        if stage == STAGE_0:                    # L576-579
            self._off_path("dehaze")
            if _on(P, "fattal"):
                self.dynamicRangeCompression(r, g, b)
        elif stage == STAGE_1:                  # L580-588
            if _on(P, "chmixer"):
                self.channelMixer(r, g, b)
            if _on(P, "exposure"):
                self.exposure(r, g, b)
            if _on(P, "hsl"):
                self.hslEqualizer(r, g, b)
            if _on(P, "toneEqualizer"):
                self.toneEqualizer(r, g, b)
            if getattr(P, "workingProfile", "") == "ProPhoto":
                self._hp.prophoto_blue(r, g, b)
        elif stage == STAGE_2:                  # L589-603; a stage is refused as a whole, before any of its steps has run
            self._off_path("dcp_look_early", "colorcorrection", "smoothing")
            if pipeline in (OUTPUT, PREVIEW):
                self._off_path("impulseDenoise", "defringe")
                if _on(P, "sharpening"):
                    self.sharpening(r, g, b)
        elif stage == STAGE_3:                  # L604-640
            self._off_path("gradient", "pcvignette", "textureBoost", "grain", "logenc", "localContrast")
            self._off_path("dcp_look", "filmSimulation")          # they sit between saturationVibrance / toneCurve / rgbCurves in the reference's order
            if pipeline == PREVIEW:
                self._off_path("prsharpening")
            self._stage3_chain(r, g, b)
            if _on(P, "blackwhite"):
                self.blackAndWhite(r, g, b)
        else:
            raise api.HotPathError(1, "unknown stage %r" % (stage,))
        return False
