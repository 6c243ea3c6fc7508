"""Pin the unsharp-mask oracle (oracle/usm_port.c) against the reference's own functions compiled in place (oracle/_ref,
shim_usm.cc: apply_gamma, unsharp_mask, buildBlendMask, get_luminance, multiply, Threshold<int>::multiply, gaussianBlur).
Bit-exact on the blend mask and on the sharpened planes, SSE2 groups versus scalar row tails included."""
import ctypes

import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
D = ctypes.c_double
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
DEFAULT_THR = (20, 80, 2000, 1200)      # procparams.cc L1761
SIZES = [(64, 48), (67, 35), (9, 8), (8, 11), (130, 77), (301, 203)]


def scene(W, H, seed, wild=False):
    """working-space RGB: smooth gradients + edges + texture + noise, so that the blend mask spans ]0, 1]"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    base = 8000 + 20000 * (0.5 + 0.5 * np.sin(xx / 17.0 + yy / 29.0)) + 15000 * (xx > W * 0.6) + 9000 * (yy > H * 0.35)
    tex = 2500 * np.sin(xx * 1.3) * np.sin(yy * 0.9) * (xx < W * 0.4)
    planes = []
    for c, gain in enumerate((0.9, 1.0, 0.7)):
        p = (base * gain + tex + rng.normal(0, 150, (H, W))).astype(np.float32)
        planes.append(np.clip(p, 1.0, 65535.0).astype(np.float32))
    if wild:
        m = rng.random((H, W))
        for p in planes:
            p[m < 0.02] = 0.0                       # luminance 0: multiply() leaves the pixel alone
            p[m > 0.985] *= 1.6                     # above 65535: apply_gamma's pow_F branch
        planes[0][(m > 0.5) & (m < 0.51)] = -40.0   # negative luminance contributions
    return planes


def run(lib, name, planes, scale=1.0, contrast=20.0, radius=0.5, amount=200, thr=DEFAULT_THR, halo=0, halo_amount=85):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    blend = np.zeros((H, W), np.float32)
    t = (ctypes.c_int * 4)(*thr)
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, PROPHOTO.ctypes.data_as(dp),
                            D(scale), D(contrast), D(radius), int(amount), t, int(halo), int(halo_amount), blend.ctypes.data_as(fp))
    assert rc == 0
    return out, blend


def same(a, b, what="RGB"):
    for x, y, ch in zip(a, b, what):
        eq = (x == y) | (np.isnan(x) & np.isnan(y))
        assert eq.all(), "%s: %d of %d differ, first at %s: %r vs %r" % (ch, int((~eq).sum()), x.size, np.argwhere(~eq)[0], x[~eq][0], y[~eq][0])


CASES = [dict(), dict(radius=0.9, amount=350), dict(contrast=0.0), dict(contrast=55.0, radius=2.4, amount=80, thr=(10, 40, 1500, 600)),
         dict(radius=0.2), dict(scale=2.0, radius=1.5), dict(amount=0), dict(halo=1), dict(halo=1, halo_amount=30, radius=1.2, amount=400),
         dict(halo=1, halo_amount=100, contrast=0.0)]


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("wild", [False, True])
def test_usm(W, H, case, wild):
    planes = scene(W, H, W * 3 + H + case, wild)
    a, ba = run(oracle.port().lib, "artoracle_usm", planes, **CASES[case])
    b, bb = run(oracle.ref().lib, "artref_usm", planes, **CASES[case])
    same([ba], [bb], ["blend"])
    same(a, b)
    if CASES[case].get("amount", 200) >= 1:
        assert any((x != y).any() for x, y in zip(a, planes)), "sharpening changed nothing"


@needs_ref
def test_blend_mask_spans_range():
    planes = scene(301, 203, 5)
    _, blend = run(oracle.ref().lib, "artref_usm", planes)
    assert blend.min() < 0.05 and blend.max() > 0.95


# ---- "rld" (RL deconvolution) route: markImpulse + deconvsharpening over gaussianBlur's GAUSS_DIV / GAUSS_MULT forms
def run_rld(lib, name, planes, scale=1.0, contrast=20.0, radius=0.75, amount=100):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    imp = np.zeros((H, W), np.float32)
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, PROPHOTO.ctypes.data_as(dp),
                            D(scale), D(contrast), D(radius), int(amount), imp.ctypes.data_as(fp))
    assert rc == 0
    return out, imp


RLD_CASES = [dict(), dict(radius=0.5, amount=60), dict(radius=1.0, amount=150), dict(radius=0.84), dict(radius=1.15, contrast=0.0),
             dict(radius=0.3), dict(radius=0.22), dict(amount=0), dict(scale=2.0, radius=1.5),
             dict(radius=1.2), dict(radius=1.8, amount=70), dict(radius=2.5, contrast=5.0)]        # recursive GAUSS_DIV / GAUSS_MULT


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("case", range(len(RLD_CASES)))
@pytest.mark.parametrize("wild", [False, True])
def test_rld(W, H, case, wild):
    planes = scene(W, H, W * 7 + H + case, wild)
    a, ia = run_rld(oracle.port().lib, "artoracle_rld", planes, **RLD_CASES[case])
    b, ib = run_rld(oracle.ref().lib, "artref_rld", planes, **RLD_CASES[case])
    same([ia], [ib], ["impulse"])
    same(a, b)


@needs_ref
def test_rld_changes_the_image_and_marks_impulses():
    planes = scene(301, 203, 5)
    planes[1][50, 60] += 20000.0            # an isolated spike
    out, imp = run_rld(oracle.ref().lib, "artref_rld", planes)
    assert imp[50, 60] == 1 and 0 < imp.mean() < 0.2
    assert any((x != y).any() for x, y in zip(out, planes))


def run_rld_ex(lib, name, planes, boost, latitude=25, ox=0, oy=0, fw=0, fh=0, scale=1.0, contrast=20.0, radius=0.75, amount=100):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, PROPHOTO.ctypes.data_as(dp),
                            D(scale), D(contrast), D(radius), int(amount), None, D(boost), int(latitude), int(ox), int(oy), int(fw), int(fh))
    assert rc == 0
    return out


BOOST_CASES = [dict(boost=0.2), dict(boost=0.35, latitude=60, radius=0.6), dict(boost=0.3, latitude=0, ox=40, oy=25, fw=600, fh=420),
               dict(boost=0.25, latitude=200), dict(boost=0.005), dict(boost=0.4, radius=1.0), dict(boost=0.5, radius=1.4)]


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (301, 203), (130, 77)])
@pytest.mark.parametrize("case", range(len(BOOST_CASES)))
def test_rld_corner_boost(W, H, case):
    planes = scene(W, H, W * 11 + H + case, wild=bool(case & 1))
    same(run_rld_ex(oracle.port().lib, "artoracle_rld_ex", planes, **BOOST_CASES[case]), run_rld_ex(oracle.ref().lib, "artref_rld_ex", planes, **BOOST_CASES[case]))


# ---- edgesonly: the difference image is taken on a bilateral-filtered copy (bilateral2.h)
def run_edges(lib, name, planes, edges_radius=1.9, edges_tolerance=1800, scale=1.0, contrast=20.0, radius=0.5, amount=200, thr=DEFAULT_THR, halo=0, halo_amount=85):
    out = [np.ascontiguousarray(p).copy() for p in planes]
    H, W = out[0].shape
    t = (ctypes.c_int * 4)(*thr)
    rc = getattr(lib, name)(out[0].ctypes.data_as(fp), out[1].ctypes.data_as(fp), out[2].ctypes.data_as(fp), W, H, PROPHOTO.ctypes.data_as(dp),
                            D(scale), D(contrast), D(radius), int(amount), t, int(halo), int(halo_amount), None, 1, D(edges_radius), int(edges_tolerance))
    assert rc == 0
    return out


# one case per bilateral kernel (sigma 0.5 .. 2.5 in 0.1 steps; below 0.45 is a copy), thresholds hit on both sides, plus the other switches
EDGES_CASES = [dict(edges_radius=0.4 + 0.1 * k) for k in range(0, 23)] + [
    dict(edges_radius=0.45), dict(edges_radius=0.55), dict(edges_radius=2.45), dict(edges_radius=3.7),
    dict(edges_tolerance=10), dict(edges_tolerance=10000, radius=1.4), dict(scale=2.0, edges_radius=2.5),
    dict(halo=1), dict(halo=1, halo_amount=20, edges_radius=1.0, edges_tolerance=400, amount=500)]


@needs_ref
@pytest.mark.parametrize("W,H", [(64, 48), (9, 8), (12, 11), (130, 77)])
@pytest.mark.parametrize("case", range(len(EDGES_CASES)))
def test_usm_edgesonly(W, H, case):
    planes = scene(W, H, W * 5 + H + case, wild=bool(case % 3 == 1))
    same(run_edges(oracle.port().lib, "artoracle_usm_ex", planes, **EDGES_CASES[case]), run_edges(oracle.ref().lib, "artref_usm_ex", planes, **EDGES_CASES[case]))


@needs_ref
def test_usm_edgesonly_differs_from_plain():
    planes = scene(130, 77, 9)
    a = run_edges(oracle.port().lib, "artoracle_usm_ex", planes)
    b, _ = run(oracle.port().lib, "artoracle_usm", planes)
    assert any((x != y).any() for x, y in zip(a, b))
