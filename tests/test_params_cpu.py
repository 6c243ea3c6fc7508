"""Host-side marshalling of the late round-2 parameter blocks (no GPU): the ctypes structures the Python mirror hands to the C-ABI carry the
values, array lengths and NULLs the header documents (include/art_hotpath.h: art_hp_hsl_params, art_hp_toneeq_params, art_hp_bw_params,
art_hp_chain_params::softlight_lut)."""
import numpy as np

from art_b200.api import BwParams, ChainParams, HslParams, ToneEqParams

WS = np.arange(9, dtype=np.float64).reshape(3, 3) + 0.5


def test_hsl_params_marshalling():
    px, py, dy = np.linspace(0, 1, 7), np.linspace(0.4, 0.6, 7), np.full(7, 0.25)
    p = HslParams(hcurve=(px, py, dy), scurve=None, lcurve=(px[:3], py[:3], dy[:3]), coeff=(px, py, dy), smoothing=4, scale=2.0, ws=WS)
    c = p.c_struct()
    assert (c.hcurve.n, c.scurve.n, c.lcurve.n, c.coeff.n) == (7, 0, 3, 7)
    assert not c.scurve.poly_x and not c.scurve.poly_y and not c.scurve.dy_by_dx          # identity curve: n = 0, NULL arrays
    assert [c.hcurve.poly_x[i] for i in range(7)] == list(px) and [c.lcurve.poly_y[i] for i in range(3)] == list(py[:3])
    assert c.coeff.dy_by_dx[6] == 0.25 and c.smoothing == 4 and c.scale == 2.0
    assert [c.ws[i] for i in range(9)] == list(WS.ravel())
    assert not HslParams().c_struct().ws                                                 # no matrix: NULL (the entry refuses it)


def test_toneeq_params_marshalling():
    c = ToneEqParams((10, -20, 30, -40, 50), regularization=3, pivot=-1.5, scale=4.0, ws=WS).c_struct()
    assert list(c.bands) == [10, -20, 30, -40, 50] and c.regularization == 3 and c.pivot == -1.5 and c.scale == 4.0
    assert [c.ws[i] for i in range(9)] == list(WS.ravel())


def test_bw_params_marshalling():
    t = [np.full(65536, float(k), np.float32) for k in range(5)]
    c = BwParams((0.4, 0.35, 0.25), 1.1, gamma=t[:3], cast=t[3:], ws=WS).c_struct()
    assert abs(c.bwr - 0.4) < 1e-7 and abs(c.bwb - 0.25) < 1e-7 and abs(c.kcorec - 1.1) < 1e-7
    assert c.gamma_g[65535] == 1.0 and c.ulut[0] == 3.0 and c.vlut[17] == 4.0
    c0 = BwParams((1, 0, 0)).c_struct()
    assert not c0.gamma_r and not c0.ulut and not c0.ws and c0.kcorec == 1.0


def test_chain_params_new_fields():
    lut = np.linspace(0, 65535, 65536).astype(np.float32)
    c = ChainParams(tonecurve=(5, lut), softlight=lut, ws=WS, iws=WS).c_struct()
    assert c.tonecurve_mode == 5 and c.softlight_lut[65535] == lut[65535] and c.tonecurve_lut[1] == lut[1]
    assert not ChainParams().c_struct().softlight_lut
