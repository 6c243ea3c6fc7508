"""Pin the automatic chroma estimator's oracle (oracle/denoise_port.c: artoracle_denoise_info, artoracle_denoise_auto_params) against
the reference's own RGB_denoise_info / calcautodn_info / the nine-crop combination of denoiseComputeParams (ipdenoise.cc) compiled
in place (oracle/_ref).  Bit-exact on all fifteen statistics and the three resulting chrominance parameters."""
import ctypes

import numpy as np
import pytest

import oracle
from test_oracle_denoise import PROPHOTO, rgb_frame

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
D = ctypes.c_double
EXPCOMP = float(np.log(np.float32(5.0)) / np.log(np.float32(2.0)))      # std::log(5.f) / std::log(2.f), ipdenoise.cc L939
NAMES = ["chaut", "Nb", "redaut", "blueaut", "maxredaut", "maxblueaut", "minredaut", "minblueaut", "chromina", "sigma", "lumema", "sigma_L",
         "redyel", "skinc", "nsknc"]


def crop(H, W, seed, kind):
    """kind 0: moderately coloured noisy scene; 1: saturated reds / yellows and skin-like tones (the red_yel / skin counters); 2: grey with
    exact neutral pixels (a = b = 0: xatan2f's zero branches) and values past the gamma LUT."""
    planes = rgb_frame(H, W, seed, noise=900.0 + 400.0 * kind, hot=(kind == 1))
    if kind == 1:
        planes[0][H // 2:, : W // 2] *= 1.9
        planes[2][H // 2:, : W // 2] *= 0.2
        planes[1][: H // 3, W // 2:] *= 0.75
    if kind == 2:
        g = planes[1].copy()
        planes = [g.copy(), g.copy(), g.copy()]
        planes[0][::7, ::5] += 900.0
        planes[2][3::11, 1::3] = 14000.0
    return [np.ascontiguousarray(np.clip(p, 0, 65535), dtype=np.float32) for p in planes]


def info(lib, name, planes, gamma=1.7, aggressive=0, scale=1.0, init=None):
    H, W = planes[0].shape
    calc = [np.ascontiguousarray(p[::2, ::2]) for p in planes]           # provicalc: the crop's even pixels (convertColorSpace is the host's)
    out = np.zeros(15, np.float32) if init is None else np.array(init, np.float32)
    wp = PROPHOTO.copy()
    rc = getattr(lib, name)(planes[0].ctypes.data_as(fp), planes[1].ctypes.data_as(fp), planes[2].ctypes.data_as(fp), W, H,
                            calc[0].ctypes.data_as(fp), calc[1].ctypes.data_as(fp), calc[2].ctypes.data_as(fp),
                            D(gamma), int(aggressive), D(scale), D(EXPCOMP), wp.ctypes.data_as(dp), out.ctypes.data_as(fp))
    assert rc == 0
    return out


def same(a, b):
    for k, (x, y) in enumerate(zip(a, b)):
        assert x == y or (np.isnan(x) and np.isnan(y)), "%s: %r vs %r" % (NAMES[k], x, y)


CASES = [dict(), dict(gamma=1.0), dict(gamma=3.0, aggressive=1), dict(scale=2.0), dict(scale=8.0, aggressive=1)]


@needs_ref
@pytest.mark.parametrize("W,H", [(450, 300), (257, 263), (128, 130), (701, 467)])
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_denoise_info(W, H, kind, case):
    planes = crop(H, W, W + 3 * H + kind, kind)
    got = info(oracle.port().lib, "artoracle_denoise_info", planes, **CASES[case])
    want = info(oracle.ref().lib, "artref_denoise_info", planes, **CASES[case])
    same(got, want)
    assert got[1] == 3 * max(2, int(5 - np.ceil(np.log(CASES[case].get("scale", 1.0)))))      # three subbands per level
    assert got[0] > 0 and got[8] >= 100.0 and 2.0 <= got[10] <= 32768.0


@needs_ref
def test_denoise_info_counters_see_red_and_skin():
    st = info(oracle.ref().lib, "artref_denoise_info", crop(300, 450, 5, 1))
    assert st[12] > 10000.0 and st[13] > 0.0 and 0.0 < st[14] < 1.0


def nine_stats(seed, kinds=(0, 1, 2)):
    rows = []
    for k in range(9):
        planes = crop(200 + 8 * k, 260 - 6 * k, seed * 31 + k, kinds[k % len(kinds)])
        rows.append(info(oracle.port().lib, "artoracle_denoise_info", planes, aggressive=seed & 1))
    return np.ascontiguousarray(np.stack(rows), dtype=np.float32)


def auto(lib, name, stats, raw=1, aggressive=0):
    out = np.zeros(3, np.float32)
    rc = getattr(lib, name)(stats.ctypes.data_as(fp), int(raw), int(aggressive), out.ctypes.data_as(fp))
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("raw", [1, 0])
def test_denoise_auto_params(seed, raw):
    stats = nine_stats(seed, kinds=[(0, 1, 2), (0,), (1,), (2,), (1, 0), (2, 1)][seed])
    got = auto(oracle.port().lib, "artoracle_denoise_auto_params", stats, raw, seed & 1)
    want = auto(oracle.ref().lib, "artref_denoise_auto_params", stats, raw, seed & 1)
    assert np.array_equal(got, want), (got, want)
    assert np.isfinite(got).all() and got[0] > 0


@needs_ref
def test_denoise_auto_params_branches():
    """calcautodn_info's threshold ladders, driven by synthetic statistics on both sides of every limit."""
    rng = np.random.default_rng(11)
    for trial in range(300):
        st = np.zeros((9, 15), np.float32)
        st[:, 0] = rng.choice([50, 150, 250, 350, 450, 600, 700, 1200], 9) * rng.uniform(0.9, 1.1, 9)        # chaut
        st[:, 1] = 15
        st[:, 4] = st[:, 0] * rng.uniform(0.8, 3.0, 9)                                                          # maxredaut
        st[:, 5] = st[:, 0] * rng.uniform(0.8, 3.0, 9)                                                          # maxblueaut
        st[:, 6] = st[:, 0] * rng.uniform(0.1, 0.9, 9)
        st[:, 7] = st[:, 0] * rng.uniform(0.1, 0.9, 9)
        st[:, 8] = rng.choice([1500, 2500, 4000, 7000, 12000], 9)                                               # chromina
        st[:, 10] = rng.choice([1000, 3000, 9000, 25000], 9)                                                    # lumema
        st[:, 12] = rng.choice([0, 6000, 13000], 9)                                                             # redyel
        st[:, 13] = rng.choice([0, 1100, 1300], 9)                                                              # skinc
        st[:, 14] = rng.choice([0.1, 0.35, 0.6], 9)                                                             # nsknc
        st = np.ascontiguousarray(st)
        for raw in (0, 1):
            got = auto(oracle.port().lib, "artoracle_denoise_auto_params", st, raw, trial & 1)
            want = auto(oracle.ref().lib, "artref_denoise_auto_params", st, raw, trial & 1)
            assert np.array_equal(got, want), (trial, raw, got, want)
