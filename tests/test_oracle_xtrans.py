"""Pin the X-Trans oracle (oracle/xtrans_port.c) against the reference's own xtrans_interpolate / cielab /
xtransborder_interpolate compiled in place (oracle/_ref, shim_xtrans.cc; the "det" build clears the per-thread tile buffer
at the start of every tile).  Bit-exact for 1-pass and 3-pass, CIELab and YPbPr, every origin of the 6x6 matrix, ragged
and tiny frames.  Also measures the stock reference's schedule dependence (stale tile buffer) as its self-noise."""
import ctypes

import numpy as np
import pytest

import oracle
from art_b200 import synth

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")
fp = ctypes.POINTER(ctypes.c_float)
ip = ctypes.POINTER(ctypes.c_int)
CAM = np.array(synth.XTRANS_RGB_CAM, np.float32)


def ref_xtrans(raw, xt, passes, lab, det=True, threads=0, border_only=0):
    lib = oracle.ref(det).lib
    H, W = raw.shape
    o = [np.full((H, W), -1, np.float32) for _ in range(3)]
    xt = np.ascontiguousarray(xt, np.int32)
    assert lib.artref_xtrans(W, H, xt.ctypes.data_as(ip), CAM.ctypes.data_as(fp), passes, lab, raw.ctypes.data_as(fp),
                             *[x.ctypes.data_as(fp) for x in o], border_only, threads) == 0
    return o


def port_xtrans(raw, xt, passes, lab):
    lib = oracle.port().lib
    H, W = raw.shape
    o = [np.full((H, W), -1, np.float32) for _ in range(3)]
    xt = np.ascontiguousarray(xt, np.int32)
    assert lib.artoracle_xtrans(W, H, xt.ctypes.data_as(ip), CAM.ctypes.data_as(fp), passes, lab, raw.ctypes.data_as(fp),
                                *[x.ctypes.data_as(fp) for x in o]) == 0
    return o


def same(a, b):
    for x, y, ch in zip(a, b, "RGB"):
        assert np.array_equal(x, y), "%s: %d of %d differ, first at %s" % (ch, int((x != y).sum()), x.size, np.argwhere(x != y)[0])


SIZES = [(300, 260), (131, 140), (24, 24), (23, 40), (120, 40), (40, 121), (233, 119), (215, 217), (500, 333)]
MODES = [(1, 0), (3, 1), (1, 1), (3, 0)]


@needs_ref
@pytest.mark.parametrize("W,H", SIZES)
@pytest.mark.parametrize("passes,lab", MODES)
@pytest.mark.parametrize("dy,dx", [(0, 0), (1, 2), (4, 5), (3, 3)])
def test_port_matches_reference(W, H, passes, lab, dy, dx):
    xt = synth.xtrans_matrix(dy, dx)
    raw = synth.xtrans_frame(W, H, xt, seed=W + dy)
    same(port_xtrans(raw, xt, passes, lab), ref_xtrans(raw, xt, passes, lab))


@needs_ref
@pytest.mark.parametrize("passes,lab", [(1, 0), (3, 1)])
def test_port_matches_reference_on_noise(passes, lab):
    """uniform integer noise: the harshest input for the direction selection; clipped and zero samples included"""
    xt = synth.xtrans_matrix(2, 1)
    raw = synth.random_frame(333, 251, seed=3)
    raw[40:60, 50:90] = 65535.0
    raw[100:120, 10:70] = 0.0
    same(port_xtrans(raw, xt, passes, lab), ref_xtrans(raw, xt, passes, lab))


@needs_ref
def test_reference_is_thread_count_independent_when_buffer_is_cleared():
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(500, 333, xt, seed=9)
    same(ref_xtrans(raw, xt, 3, 1, threads=1), ref_xtrans(raw, xt, 3, 1, threads=8))


@needs_ref
def test_stock_reference_self_noise_is_confined_to_tile_borders_of_the_image_edge():
    """The stock build (buffer not cleared) differs from the canonical one only where the 5x5 homogeneity sums read bytes the
    tile never wrote: within 16 pixels of the image edge.  1-pass output does not depend on it at all."""
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(500, 333, xt, seed=9)
    same(ref_xtrans(raw, xt, 1, 0, det=False, threads=1), ref_xtrans(raw, xt, 1, 0))
    a, b = ref_xtrans(raw, xt, 3, 1, det=False, threads=1), ref_xtrans(raw, xt, 3, 1)
    H, W = raw.shape
    for x, y in zip(a, b):
        bad = np.argwhere(x != y)
        if len(bad):
            edge = np.minimum.reduce([bad[:, 0], bad[:, 1], H - 1 - bad[:, 0], W - 1 - bad[:, 1]])
            assert edge.max() < 16, edge.max()
            assert len(bad) < 0.01 * x.size


@needs_ref
@pytest.mark.parametrize("border", [8, 11])
def test_border(border):
    xt = synth.xtrans_matrix(1, 4)
    raw = synth.xtrans_frame(77, 65, xt, seed=2)
    want = ref_xtrans(raw, xt, 3, 1, border_only=border)
    got = [np.full(raw.shape, -1, np.float32) for _ in range(3)]
    xtc = np.ascontiguousarray(xt, np.int32)
    assert oracle.port().lib.artoracle_xtrans_border(77, 65, xtc.ctypes.data_as(ip), border, raw.ctypes.data_as(fp), *[x.ctypes.data_as(fp) for x in got]) == 0
    same(got, want)


def test_port_covers_every_pixel_once():
    """tiles and border together write every output sample (no -1 left)"""
    xt = synth.xtrans_matrix()
    raw = synth.xtrans_frame(215, 217, xt, seed=4)
    for passes, lab in MODES:
        out = port_xtrans(raw, xt, passes, lab)
        assert all((p >= 0).all() for p in out)
