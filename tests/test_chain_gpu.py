"""GPU parity for the fused per-pixel colour / curve chain (art_hp_color_chain = the per-pixel stages of
ImProcFunctions::process) through the C-ABI against the oracle port, which test_oracle_chain.py pins bit-exact to the
reference's own loops compiled in place.  The oracle applies the stages one after the other as the reference does (one
full pass each); the GPU applies them in one pass.  Bit-exact, SSE2 groups and scalar row tails included."""
import ctypes

import numpy as np
import pytest

import oracle
from art_b200.api import ChainParams
from test_oracle_chain import F, PROPHOTO, PROPHOTO_INV, call, curve_lut, dp, fp, image, same, softlight_lut

pytestmark = pytest.mark.gpu

SIZES = [(64, 8), (67, 5), (5, 9), (3, 3), (130, 17), (1021, 300)]


def lab_luts():
    lc = np.concatenate([curve_lut(32768, 0.85, 32767.0), np.array([32768.0, 32769.0], np.float32)]).astype(np.float32)
    return lc, curve_lut(65536, 1.1, 65535.0, 1), curve_lut(65536, 0.9, 65535.0, 2)


def oracle_chain(planes, exposure=None, saturation=None, tonecurve=None, rgbcurves=None, lab=None, whitept=1.0, softlight=None):
    lib = oracle.port().lib
    out = planes
    if exposure is not None:
        ev, black = exposure
        out = call(lib, "artoracle_chain_expcomp", out, F(np.float32(2.0) ** np.float32(ev)), F(np.float32(black) * np.float32(2000.0)))
    if saturation is not None and (saturation[0] or saturation[1]):
        out = call(lib, "artoracle_chain_saturation", out, saturation[0], saturation[1], PROPHOTO.ctypes.data_as(dp))
    if tonecurve is not None and tonecurve[0] >= 3:     # WEIGHTEDSTD / SATANDVALBLENDING / LUMINANCE
        out = call(lib, "artoracle_chain_tonecurve_ex", out, tonecurve[0], tonecurve[1].ctypes.data_as(fp), F(whitept), PROPHOTO.ctypes.data_as(dp))
    elif tonecurve is not None:
        out = call(lib, "artoracle_chain_tonecurve", out, tonecurve[0], tonecurve[1].ctypes.data_as(fp), F(whitept))
    if rgbcurves is not None:
        out = call(lib, "artoracle_chain_rgbcurves", out, *[c.ctypes.data_as(fp) if c is not None else None for c in rgbcurves])
    if lab is not None:
        out = call(lib, "artoracle_chain_lab", out, lab[0].ctypes.data_as(fp), lab[1].ctypes.data_as(fp), lab[2].ctypes.data_as(fp), F(lab[3]),
                   PROPHOTO.ctypes.data_as(dp), PROPHOTO_INV.ctypes.data_as(dp))
    if softlight is not None:
        out = call(lib, "artoracle_chain_softlight", out, softlight.ctypes.data_as(fp))
    return out


def gpu_chain(hp, planes, **kw):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    hp.color_chain(out[0], out[1], out[2], ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, **kw))
    return out


STAGES = {
    "exposure": dict(exposure=(1.3, 0.01)),
    "exposure_neg": dict(exposure=(-0.7, -0.02)),
    "saturation": dict(saturation=(30, 0)),
    "vibrance": dict(saturation=(-50, -30)),
    "satvib": dict(saturation=(100, 100)),
    "tone_std": dict(tonecurve=(0, curve_lut(gamma=0.6, seed=0))),
    "tone_film": dict(tonecurve=(1, curve_lut(gamma=0.6, seed=1))),
    "rgbcurves": dict(rgbcurves=[curve_lut(gamma=0.7, seed=0), None, curve_lut(gamma=1.1, seed=2)]),
    "lab": dict(lab=lab_luts() + (1.35,)),
    "lab_lowchroma": dict(lab=lab_luts() + (0.4,)),
}
STAGES["tone_weighted"] = dict(tonecurve=(3, curve_lut(gamma=0.6, seed=3)))
STAGES["tone_satval"] = dict(tonecurve=(4, curve_lut(gamma=0.6, seed=4)))
STAGES["tone_satval_dark"] = dict(tonecurve=(4, curve_lut(gamma=1.5, seed=4)))
STAGES["tone_luminance"] = dict(tonecurve=(5, curve_lut(gamma=0.6, seed=5)))
STAGES["tone_weighted_white2"] = dict(tonecurve=(3, curve_lut(gamma=0.7, seed=3)), whitept=2.0)
STAGES["tone_luminance_white2"] = dict(tonecurve=(5, curve_lut(gamma=1.3, seed=5)), whitept=2.0)
STAGES["all_weighted"] = dict(STAGES["exposure"], **STAGES["satvib"], **STAGES["tone_weighted"], **STAGES["rgbcurves"], **STAGES["lab"])
STAGES["all_satval"] = dict(STAGES["exposure_neg"], **STAGES["saturation"], **STAGES["tone_satval"], **STAGES["lab_lowchroma"])
STAGES["all_luminance"] = dict(STAGES["exposure"], **STAGES["vibrance"], **STAGES["tone_luminance"], **STAGES["rgbcurves"], **STAGES["lab"])
STAGES["softlight"] = dict(softlight=softlight_lut(30))
STAGES["all_softlight"] = dict(STAGES["exposure"], **STAGES["saturation"], **STAGES["tone_film"], **STAGES["lab"], softlight=softlight_lut(100))
STAGES["all_std"] = dict(STAGES["exposure"], **STAGES["satvib"], **STAGES["tone_std"], **STAGES["rgbcurves"], **STAGES["lab"])
STAGES["all_film"] = dict(STAGES["exposure_neg"], **STAGES["saturation"], **STAGES["tone_film"],
                          rgbcurves=[curve_lut(gamma=0.9, seed=i) for i in range(3)], **STAGES["lab_lowchroma"])


@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("name", sorted(STAGES))
def test_chain_matches_oracle(hot_path, W, H, name):
    planes = image(H, W, W * 17 + H)
    same(gpu_chain(hot_path, planes, **STAGES[name]), oracle_chain(planes, **STAGES[name]))


def test_chain_rejects_perceptual_and_luminance_without_ws(hot_path):
    import art_b200
    planes = image(8, 8, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.color_chain(planes[0], planes[1], planes[2], ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, tonecurve=(6, curve_lut())))
    with pytest.raises(art_b200.HotPathError):
        hot_path.color_chain(planes[0], planes[1], planes[2], ChainParams(tonecurve=(5, curve_lut())))


def test_chain_in_range_image(hot_path):
    planes = image(240, 516, 7, wild=False)
    same(gpu_chain(hot_path, planes, **STAGES["all_std"]), oracle_chain(planes, **STAGES["all_std"]))


def test_chain_nothing_enabled_is_identity(hot_path):
    planes = image(33, 70, 3)
    same(gpu_chain(hot_path, planes), planes)


def test_chain_device_form_with_pitch(hot_path):
    torch = pytest.importorskip("torch")
    W, H, pitch = 203, 41, 224
    planes = image(H, W, 99)
    dev = [torch.zeros((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    for d, p in zip(dev, planes):
        d[:, :W] = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    hot_path.color_chain_dev(W, H, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), pitch,
                             ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, **STAGES["all_film"]))
    hot_path.sync()
    same([d[:, :W].cpu().numpy() for d in dev], oracle_chain(planes, **STAGES["all_film"]))


def test_chain_rejects_missing_matrix(hot_path):
    import art_b200
    planes = image(8, 8, 1)
    with pytest.raises(art_b200.HotPathError):
        hot_path.color_chain(planes[0], planes[1], planes[2], ChainParams(saturation=(10, 0)))


@pytest.mark.parametrize("W,H", [(64, 48), (67, 35), (301, 203), (1021, 77)])
@pytest.mark.parametrize("mixer", [(1000, 0, 0, 0, 1000, 0, 0, 0, 1000), (800, 300, -100, -50, 1100, -50, 20, -400, 1380), (-200, 600, 600, 333, 333, 334, 0, 0, -1000)])
def test_channel_mixer_matches_oracle(hot_path, W, H, mixer):
    """art_hp_channel_mixer = ImProcFunctions::channelMixer's loop (ipchmixer.cc L200-230); the oracle is pinned to it in test_oracle_chain.py.
    Bit-exact, NaN samples included (the vector groups clamp them to 0, the row tail keeps them)."""
    import ctypes
    import oracle
    rng = np.random.default_rng(W + H)
    planes = [np.ascontiguousarray(rng.uniform(-3000, 70000, (H, W)), dtype=np.float32) for _ in range(3)]
    planes[0][H // 2, ::3] = np.nan
    m = (np.array(mixer, np.float32) / np.float32(1000.0)).astype(np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    want = [p.copy() for p in planes]
    assert oracle.port().lib.artoracle_chmixer(*[p.ctypes.data_as(fp) for p in want], W, H, m.ctypes.data_as(fp)) == 0
    got = hot_path.channel_mixer(*[p.copy() for p in planes], m)
    for g, w in zip(got, want):
        assert ((g == w) | (np.isnan(g) & np.isnan(w))).all()


@pytest.mark.parametrize("W,H", SIZES + [(2001, 1333)])
@pytest.mark.parametrize("name", ["none", "exposure", "all_before_lab"])
def test_lab_histogram_matches_oracle(hot_path, W, H, name):
    """art_hp_lab_histogram = labAdjustments' hist16 (iplabadjustments.cc L307-334): the oracle's stages, its rgb -> Lab (pinned to
    Imagefloat::setMode(LAB) in test_oracle_chain.py), then hist[LIM((int)L, 0, 65535)]++ (LUT<T>::operator[](int)).  Exact; planes untouched."""
    kw = {"none": {}, "exposure": STAGES["exposure"],
          "all_before_lab": dict(STAGES["exposure"], **STAGES["satvib"], **STAGES["tone_weighted"], **STAGES["rgbcurves"])}[name]
    planes = image(H, W, W * 13 + H, wild=(name != "none"))
    pre = oracle_chain(planes, **kw)
    lab = call(oracle.port().lib, "artoracle_chain_rgb2lab", pre, PROPHOTO.ctypes.data_as(dp), PROPHOTO_INV.ctypes.data_as(dp), 0)
    L = lab[1]
    idx = np.clip(np.where(np.isnan(L), -2147483648.0, L), -2147483648.0, 2147483520.0).astype(np.int64)      # cvttss2si: NaN / overflow -> INT_MIN
    want = np.bincount(np.clip(idx, 0, 65535).ravel(), minlength=65536).astype(np.uint32)
    keep = [p.copy() for p in planes]
    got = hot_path.lab_histogram(planes[0], planes[1], planes[2], ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, **kw))
    assert int(got.sum()) == W * H
    assert np.array_equal(got, want)
    same(planes, keep)


@pytest.mark.parametrize("W,H", SIZES + [(2001, 1333)])
@pytest.mark.parametrize("gamma,cast", [(False, False), (True, False), (False, True), (True, True)])
def test_black_and_white_matches_oracle(hot_path, W, H, gamma, cast):
    """art_hp_black_and_white = ImProcFunctions::blackAndWhite's pixel loops (ipbw.cc L283-312, L343-362); the oracle is pinned to them in test_oracle_chain.py"""
    from art_b200.api import BwParams
    from test_oracle_chain import bw_args, bw_tables
    planes = image(H, W, W * 5 + H + gamma)
    tabs = bw_tables(W, gamma, cast)
    want = call(oracle.port().lib, "artoracle_bw", planes, *bw_args((0.43, 0.33, 0.30), 1.06, tabs))
    got = [p.copy() for p in planes]
    hot_path.black_and_white(got[0], got[1], got[2], BwParams((0.43, 0.33, 0.30), 1.06, tabs[:3] if gamma else None, tabs[3:] if cast else None, PROPHOTO))
    same(got, want)


@pytest.mark.parametrize("W,H", SIZES + [(2001, 1333)])
def test_prophoto_blue_matches_oracle(hot_path, W, H):
    """art_hp_prophoto_blue = proPhotoBlue (improcfun.cc L312-357); the oracle is pinned to it in test_oracle_chain.py"""
    from test_oracle_chain import blue_image
    planes = blue_image(H, W, W * 3 + H)
    want = call(oracle.port().lib, "artoracle_prophoto_blue", planes)
    got = [p.copy() for p in planes]
    hot_path.prophoto_blue(got[0], got[1], got[2])
    same(got, want)


def test_black_and_white_device_form_with_unaligned_pitch(hot_path):
    """the device entry on planes whose rows are not 16-byte aligned: the scalar-access path of k_bw"""
    import ctypes
    torch = pytest.importorskip("torch")
    from art_b200.api import BwParams
    from test_oracle_chain import bw_args, bw_tables
    W, H, pitch = 203, 41, 205
    planes = image(H, W, 31)
    tabs = bw_tables(5, True, True)
    want = call(oracle.port().lib, "artoracle_bw", planes, *bw_args((0.3, 0.5, 0.2), 0.97, tabs))
    dev = [torch.zeros((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    for d, p in zip(dev, planes):
        d[:, :W] = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    c = BwParams((0.3, 0.5, 0.2), 0.97, tabs[:3], tabs[3:], PROPHOTO).c_struct()
    hot_path._check(hot_path.lib.art_hp_black_and_white_dev(hot_path.h, W, H, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), pitch, ctypes.byref(c)))
    hot_path.sync()
    same([d[:, :W].cpu().numpy() for d in dev], want)


def test_improcfunctions_process_through_the_library(hot_path):
    """art_b200.ImProcFunctions.process (the mirror of improcfun.cc L567-641) through the real entries: STAGE_1 = channelMixer, exposure, toneEqualizer
    and STAGE_3 = one fused chain call + blackAndWhite, against the oracle ports applied in the reference's order"""
    import ctypes
    from types import SimpleNamespace as NS
    from art_b200.api import BwParams, ToneEqParams
    from art_b200.improcfun import ImProcFunctions, OUTPUT, STAGE_1, STAGE_3
    from test_oracle_chain import bw_args, bw_tables
    ip = ctypes.POINTER(ctypes.c_int)
    W, H = 203, 97
    planes = image(H, W, 123, wild=False)
    mix = (np.array([800, 300, -100, -50, 1100, -50, 20, -400, 1380], np.float32) / np.float32(1000.0)).astype(np.float32)
    bands = np.array([30, 10, 0, -10, -25], np.int32)
    tabs = bw_tables(3, True, False)
    lut = curve_lut(gamma=0.7, seed=9)
    P = NS(chmixer=NS(enabled=True, matrix=mix), exposure=NS(enabled=True, expcomp=0.4, black=0.0), hsl=NS(enabled=False),
           toneEqualizer=NS(enabled=True, params=ToneEqParams(bands, 0, 0.0, 1.0, PROPHOTO)), workingProfile="Rec2020",
           saturation=NS(enabled=True, saturation=15, vibrance=5), toneCurve=NS(enabled=True, mode=0, lut=lut), rgbCurves=NS(enabled=False),
           labCurve=NS(enabled=False), softlight=NS(enabled=True, lut=softlight_lut(30)),
           blackwhite=NS(enabled=True, params=BwParams((0.43, 0.33, 0.30), 1.06, tabs[:3], None, PROPHOTO)))
    got = [p.copy() for p in planes]
    ipf = ImProcFunctions(P, hot_path, 1.0, PROPHOTO, PROPHOTO_INV)
    ipf.process(OUTPUT, STAGE_1, *got)
    lib = oracle.port().lib
    want = call(lib, "artoracle_chmixer", planes, mix.ctypes.data_as(fp))
    want = oracle_chain(want, exposure=(0.4, 0.0))
    want = call(lib, "artoracle_tone_equalizer", want, PROPHOTO.ctypes.data_as(dp), bands.ctypes.data_as(ip), 0, ctypes.c_double(0.0), ctypes.c_double(1.0))
    same(got, want)
    ipf.process(OUTPUT, STAGE_3, *got)
    want = oracle_chain(want, saturation=(15, 5), tonecurve=(0, lut), softlight=softlight_lut(30))
    want = call(lib, "artoracle_bw", want, *bw_args((0.43, 0.33, 0.30), 1.06, tabs))
    same(got, want)
