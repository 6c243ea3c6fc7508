"""Pin the HSL equalizer oracle (oracle/hsl_port.c) against the reference's own ImProcFunctions::hslEqualizer body compiled in place
(oracle/_ref, shim_tone.cc: iphsl.cc L29-221 over the reference's FlatCurve, guidedFilter, Color::yuv2hsl / hsl2yuv and the
Imagefloat rgb_to_yuv / yuv_to_rgb / multiply loops).  Bit-exact."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
FCT_MinMaxCPoints = 1.0
# ImProcFunctions::hslEqualizer's local `coeff` curve, iphsl.cc L119-123
COEFF = [FCT_MinMaxCPoints, 0.25, 0.0, 0.5, 0.18, 1, 1, 0, 0.35]
IDENT = [0.0]        # FCT_Linear: "considered as identity" (flatcurves.cc)


def flat_points(seed, npts=6, amp=0.35):
    """FCT_MinMaxCPoints control points (x, y, left tangent, right tangent) around the neutral 0.5, like the GUI's equalizer curves"""
    rng = np.random.default_rng(seed)
    xs = np.sort(rng.uniform(0.0, 1.0, npts))
    pts = [FCT_MinMaxCPoints]
    for x in xs:
        pts += [float(x), float(np.clip(0.5 + rng.uniform(-amp, amp), 0, 1)), 0.35, 0.35]
    return pts


def curve_key(pts, periodic, poly_pn):
    import hashlib
    return hashlib.sha1(np.array(list(pts) + [float(periodic), float(poly_pn)], np.float64).tobytes()).hexdigest()[:16]


_GOLDEN = None


def polyline(pts, periodic=True, poly_pn=1000):
    """(n, px, py, dy) of the FlatCurve the reference builds from the control points; n = 0 for an identity curve.  Read from the committed
    fixture tests/golden/hsl_polylines.npz (made by tests/golden/make_hsl_golden.py from the reference's constructor) when it holds the curve,
    else built by oracle/_ref."""
    global _GOLDEN
    if _GOLDEN is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hsl_polylines.npz")
        _GOLDEN = dict(np.load(path)) if os.path.exists(path) else {}
    key = curve_key(pts, periodic, poly_pn)
    if key + "_n" in _GOLDEN:
        n = int(_GOLDEN[key + "_n"][0])
        return n, _GOLDEN[key + "_x"].copy(), _GOLDEN[key + "_y"].copy(), _GOLDEN[key + "_d"].copy()
    return polyline_from_reference(pts, periodic, poly_pn)


def polyline_from_reference(pts, periodic=True, poly_pn=1000):
    lib = oracle.ref().lib
    cap = 65536
    px, py, dy = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    a = np.array(pts, np.float64)
    lib.artref_flat_polyline.restype = ctypes.c_int
    n = lib.artref_flat_polyline(a.ctypes.data_as(dp), len(pts), int(periodic), int(poly_pn), px.ctypes.data_as(dp), py.ctypes.data_as(dp), dy.ctypes.data_as(dp), cap)
    assert n <= cap
    return n, px[:n].copy(), py[:n].copy(), dy[:n].copy()


def image(H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 20000 + 15000 * np.sin(xx / 17.0) * np.cos(yy / 23.0)
    planes = [(base * rng.uniform(0.5, 1.4) + rng.normal(0, 2500, (H, W)) + 9000 * np.sin(xx / (5.0 + 3 * c) + c)).astype(np.float32) for c in range(3)]
    planes[0][: H // 6] *= 0.02            # near-grey shadows: small chroma, every hue
    planes[1][: H // 6] *= 0.02
    planes[2][: H // 6] *= 0.02
    planes[2][:, : W // 7] = planes[0][:, : W // 7] = planes[1][:, : W // 7]      # u = v = 0: xatan2f(0, 0)
    planes[0][H // 2, ::5] *= -0.3         # negative samples
    return [np.ascontiguousarray(np.clip(p, -8000, 90000), np.float32) for p in planes]


def port_hsl(planes, hc, sc, lc, smoothing, scale, poly_pn=None):
    pn = poly_pn if poly_pn is not None else int(1000 / scale)
    curves = [polyline(c, True, pn) for c in (hc, sc, lc)] + [polyline(COEFF, True, 1000)]
    out = [p.copy() for p in planes]
    H, W = out[0].shape
    args = []
    for n, px, py, dy in curves:
        args += [n, px.ctypes.data_as(dp), py.ctypes.data_as(dp), dy.ctypes.data_as(dp)]
    rc = oracle.port().lib.artoracle_hsl_equalizer(*[p.ctypes.data_as(fp) for p in out], W, H, PROPHOTO.ctypes.data_as(dp), *args, int(smoothing), ctypes.c_double(scale))
    assert rc == 0
    return out


def ref_hsl(planes, hc, sc, lc, smoothing, scale):
    out = [p.copy() for p in planes]
    H, W = out[0].shape
    a = [np.array(c, np.float64) for c in (hc, sc, lc)]
    rc = oracle.ref().lib.artref_hsl_equalizer(*[p.ctypes.data_as(fp) for p in out], W, H, PROPHOTO.ctypes.data_as(dp),
                                               a[0].ctypes.data_as(dp), len(hc), a[1].ctypes.data_as(dp), len(sc), a[2].ctypes.data_as(dp), len(lc),
                                               int(smoothing), ctypes.c_double(scale))
    assert rc == 0
    return out


def same(a, b):
    for x, y, ch in zip(a, b, "RGB"):
        eq = (x == y) | (np.isnan(x) & np.isnan(y))
        assert eq.all(), "%s: %d of %d differ, first at %s: %r vs %r" % (ch, int((~eq).sum()), x.size, np.argwhere(~eq)[0], x[~eq][0], y[~eq][0])


CASES = {
    "all": (flat_points(1), flat_points(2), flat_points(3)),
    "h_only": (flat_points(4), IDENT, IDENT),
    "s_only": (IDENT, flat_points(5, amp=0.5), IDENT),
    "l_only": (IDENT, IDENT, flat_points(6)),
    "none": (IDENT, IDENT, IDENT),
}


@needs_ref
@pytest.mark.parametrize("W,H", [(96, 64), (131, 77), (300, 201)])
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("smoothing,scale", [(0, 1.0), (5, 1.0), (10, 1.0), (7, 2.0)])
def test_hsl_port_matches_reference(W, H, case, smoothing, scale):
    planes = image(H, W, W + H + smoothing)
    hc, sc, lc = CASES[case]
    same(port_hsl(planes, hc, sc, lc, smoothing, scale), ref_hsl(planes, hc, sc, lc, smoothing, scale))


@needs_ref
def test_hsl_changes_the_image_and_identity_round_trips_closely():
    planes = image(120, 160, 9)
    out = port_hsl(planes, *CASES["all"], 5, 1.0)
    assert max(float(np.abs(o - p).max()) for o, p in zip(out, planes)) > 100.0
    ident = port_hsl(planes, *CASES["none"], 5, 1.0)         # only rgb -> yuv -> hsl -> yuv -> rgb: float round-off
    for o, p in zip(ident, planes):
        assert np.allclose(o, p, rtol=0, atol=0.05)


@needs_ref
def test_flat_getval_of_the_polyline_matches_the_reference_curve():
    pts = flat_points(11)
    n, px, py, dy = polyline(pts)
    lib = oracle.ref().lib
    lib.artref_flat_getval.restype = ctypes.c_double
    a = np.array(pts, np.float64)
    for t in np.linspace(-0.2, 1.3, 301):
        tt = t + 1.0 if t < px[0] else t
        lo, hi = 0, n - 1
        while hi > 1 + lo:
            k = (hi + lo) // 2
            if px[k] > tt:
                hi = k
            else:
                lo = k
        want = lib.artref_flat_getval(a.ctypes.data_as(dp), len(pts), 1, 1000, ctypes.c_double(t))
        assert py[lo] + (tt - px[lo]) * dy[lo] == want


@needs_ref
def test_golden_polylines_are_the_reference_constructor_output():
    for pts, pn in ((COEFF, 1000), (CASES["all"][0], 1000), (CASES["all"][2], 500), (CASES["s_only"][1], 1000)):
        a, b = polyline(pts, True, pn), polyline_from_reference(pts, True, pn)
        assert a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
