"""Helpers shared by the tone-curve tests: the reference's curve objects (DiagonalCurve / FlatCurve / ToneCurve::Set compiled in
place in oracle/_ref) turned into what the hot path takes -- LUTs and Curve::getVal stage lists -- plus the port's entry points.
Test infrastructure only."""
import ctypes

import numpy as np

import oracle

fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
F = ctypes.c_float
D = ctypes.c_double

# rtdata/profiles/Standard Film Curve.arp: CurveMode=Neutral, Curve (spline), Curve2 linear, Saturation (flat curve), WhitePoint=1
FILM_CURVE = [1, 0, 0, 0.11, 0.09, 0.32, 0.47, 0.66, 0.87, 1, 1]
FILM_SAT = [1, 0, 0.48, 0.34, 0.35, 1, 0.48, 0.35, 0.35]
LINEAR = [0]
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
PROPHOTO_INV = np.array([[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0, 0.0, 1.2118128]], np.float64)
SRGB_XYZ = np.array([[0.4360747, 0.3850649, 0.1430804], [0.2225045, 0.7168786, 0.0606169], [0.0139322, 0.0971045, 0.7141733]], np.float32)


class Stage(ctypes.Structure):      # artoracle_curve_stage == art_hp_curve_stage
    _fields_ = [("kind", ctypes.c_int), ("poly_x", dp), ("poly_y", dp), ("n", ctypes.c_int), ("a", D), ("b", D), ("w", D)]


def arr(v):
    return np.ascontiguousarray(v, np.float64)


def build_lut(curve, curve2, contrast=0, whitept=1.0, scale=1.0, which=0):
    """ToneCurve::Set over the curve ImProcFunctions::toneCurve assembles; returns (lut, is_identity)"""
    c1, c2 = arr(curve), arr(curve2)
    lut = np.zeros(65536, np.float32)
    rc = oracle.ref().lib.artref_tone_build_lut(c1.ctypes.data_as(dp), len(c1), c2.ctypes.data_as(dp), len(c2), int(contrast), F(whitept), D(scale), which,
                                               lut.ctypes.data_as(fp))
    return lut, bool(rc)


def polyline(curve, curve2, which, whitept=1.0, scale=1.0):
    c1, c2 = arr(curve), arr(curve2)
    f = oracle.ref().lib.artref_tone_polyline
    cap = f(c1.ctypes.data_as(dp), len(c1), c2.ctypes.data_as(dp), len(c2), F(whitept), D(scale), which, None, None, 0)
    px, py = np.zeros(max(cap, 1)), np.zeros(max(cap, 1))
    n = f(c1.ctypes.data_as(dp), len(c1), c2.ctypes.data_as(dp), len(c2), F(whitept), D(scale), which, px.ctypes.data_as(dp), py.ctypes.data_as(dp), cap)
    assert n == cap
    return px[:n].copy(), py[:n].copy()


def stages_for(curve, curve2, contrast=0, whitept=1.0, scale=1.0):
    """the DoubleCurve chain of iptonecurve.cc L652-658 as (kind, poly_x, poly_y, a, b, w) tuples: [contrast], curve 1, curve 2"""
    out = []
    if contrast:
        ab = np.zeros(2)
        oracle.ref().lib.artref_tone_contrast_ab(int(contrast), F(whitept), ab.ctypes.data_as(dp))
        out.append((2, None, None, ab[0], ab[1], float(np.float32(whitept))))
    for which in (1, 2):
        px, py = polyline(curve, curve2, which, whitept, scale)
        out.append((1, px, py, 0.0, 0.0, 0.0) if len(px) else (0, None, None, 0.0, 0.0, 0.0))
    return out


def stage_array(stages):
    """ctypes array of Stage + the numpy arrays that must stay alive"""
    keep = []
    a = (Stage * max(1, len(stages)))()
    for i, (kind, px, py, ca, cb, cw) in enumerate(stages):
        a[i].kind = kind
        if kind == 1:
            keep += [px, py]
            a[i].poly_x = px.ctypes.data_as(dp); a[i].poly_y = py.ctypes.data_as(dp); a[i].n = len(px)
        a[i].a, a[i].b, a[i].w = ca, cb, cw
    return a, keep


def sat_lut(satcurve, scale=1.0):
    s = arr(satcurve)
    lut = np.zeros(65536, np.float32)
    rc = oracle.ref().lib.artref_tone_satlut(s.ctypes.data_as(dp), len(s), D(scale), lut.ctypes.data_as(fp))
    return None if rc else lut


def out_matrices(ws, iws, om):
    """ApplyState's to_out / to_work (curves.cc L868-872) in float, as linalgebra.h's inverse / dot_product compute them"""
    if om is None:
        return None, None
    f = np.float32
    m = om.astype(f)
    r00 = f(m[1, 1] * m[2, 2]) - f(m[2, 1] * m[1, 2]); r10 = f(m[2, 0] * m[1, 2]) - f(m[1, 0] * m[2, 2]); r20 = f(m[1, 0] * m[2, 1]) - f(m[2, 0] * m[1, 1])
    det = f(f(f(m[0, 0] * r00) + f(m[0, 1] * r10)) + f(m[0, 2] * r20))
    inv = np.array([[r00 / det, (f(m[2, 1] * m[0, 2]) - f(m[0, 1] * m[2, 2])) / det, (f(m[0, 1] * m[1, 2]) - f(m[1, 1] * m[0, 2])) / det],
                    [r10 / det, (f(m[0, 0] * m[2, 2]) - f(m[2, 0] * m[0, 2])) / det, (f(m[1, 0] * m[0, 2]) - f(m[0, 0] * m[1, 2])) / det],
                    [r20 / det, (f(m[2, 0] * m[0, 1]) - f(m[0, 0] * m[2, 1])) / det, (f(m[0, 0] * m[1, 1]) - f(m[1, 0] * m[0, 1])) / det]], f)

    def dot(a, b):
        res = np.zeros((3, 3), f)
        for i in range(3):
            for j in range(3):
                acc = f(0)
                for k in range(3):
                    acc = f(acc + f(a[i, k] * b[k, j]))
                res[i, j] = acc
        return res
    return dot(inv, ws.astype(f)), dot(iws.astype(f), m)


def ref_neutral(planes, curve, curve2, contrast=0, whitept=1.0, scale=1.0, ws=PROPHOTO, iws=PROPHOTO_INV, om=None):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    c1, c2 = arr(curve), arr(curve2)
    wsf, iwsf = np.ascontiguousarray(ws, np.float32), np.ascontiguousarray(iws, np.float32)
    omf = None if om is None else np.ascontiguousarray(om, np.float32)
    rc = oracle.ref().lib.artref_tone_neutral(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H,
                                              c1.ctypes.data_as(dp), len(c1), c2.ctypes.data_as(dp), len(c2), int(contrast), F(whitept), D(scale),
                                              wsf.ctypes.data_as(fp), iwsf.ctypes.data_as(fp), None if omf is None else omf.ctypes.data_as(fp))
    assert rc == 0
    return out


def port_neutral(planes, lut, whitept, stages, ws=PROPHOTO, iws=PROPHOTO_INV, to_out=None, to_work=None):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    wsf, iwsf = np.ascontiguousarray(ws, np.float32), np.ascontiguousarray(iws, np.float32)
    sa, keep = stage_array(stages or [])
    to = None if to_out is None else np.ascontiguousarray(to_out, np.float32)
    tw = None if to_work is None else np.ascontiguousarray(to_work, np.float32)
    rc = oracle.port().lib.artoracle_tone_neutral(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, lut.ctypes.data_as(fp),
                                                  F(whitept), sa, len(stages or []), wsf.ctypes.data_as(fp), iwsf.ctypes.data_as(fp),
                                                  None if to is None else to.ctypes.data_as(fp), None if tw is None else tw.ctypes.data_as(fp))
    assert rc == 0
    return out


def ref_satcurve(planes, satcurve, satcurve2=LINEAR, whitept=1.0, scale=1.0, ws=PROPHOTO, iws=PROPHOTO_INV):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    s1, s2 = arr(satcurve), arr(satcurve2)
    wsf, iwsf = np.ascontiguousarray(ws, np.float32), np.ascontiguousarray(iws, np.float32)
    oracle.ref().lib.artref_tone_satcurve(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, s1.ctypes.data_as(dp), len(s1),
                                          s2.ctypes.data_as(dp), len(s2), F(whitept), D(scale), wsf.ctypes.data_as(fp), iwsf.ctypes.data_as(fp))
    return out


def port_satcurve(planes, satlut, ws=PROPHOTO, iws=PROPHOTO_INV):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    wsf, iwsf = np.ascontiguousarray(ws, np.float32), np.ascontiguousarray(iws, np.float32)
    rc = oracle.port().lib.artoracle_tone_satcurve(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, satlut.ctypes.data_as(fp),
                                                   wsf.ctypes.data_as(fp), iwsf.ctypes.data_as(fp))
    assert rc == 0
    return out


def frame(H, W, seed, over=True):
    """working-space RGB: in-range values, deep shadows, exact zeros and ties, negatives and samples above 65535"""
    rng = np.random.default_rng(seed)
    planes = [rng.uniform(0, 65535, (H, W)).astype(np.float32) for _ in range(3)]
    m = rng.random((H, W))
    planes[0][:, : W // 5] *= 1e-3
    planes[1][:, : W // 7] *= 1e-3
    for p in planes:
        p[m < 0.03] *= -0.2
        if over:
            p[m > 0.93] *= 1.9
    planes[1][(m > 0.5) & (m < 0.55)] = planes[0][(m > 0.5) & (m < 0.55)]
    planes[2][(m > 0.6) & (m < 0.63)] = 0.0
    z = (m > 0.7) & (m < 0.71)
    for p in planes:
        p[z] = 0.0
    g = (m > 0.75) & (m < 0.8)
    planes[1][g] = planes[0][g]; planes[2][g] = planes[0][g]          # neutral greys
    return planes
