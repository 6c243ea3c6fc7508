"""Pin the AMaZE oracle: our C restatement (oracle/amaze_port.c) against the committed golden vectors
(generated from the reference itself) and, when present, against the reference's own function body
compiled in place (oracle/_ref).  Bit-exact, including the tile sizes that trigger the reference's
border-overflow quirk.  Also measures the reference's own schedule-dependence."""
import os

import numpy as np
import pytest

import oracle
from art_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("name", ["amaze_rggb_scene", "amaze_grbg_noise"])
def test_port_matches_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    got = oracle.port().amaze(z["raw"].astype(np.float32), int(z["filters"]))
    for g, w, ch in zip(got, (z["red"], z["green"], z["blue"]), "RGB"):
        assert np.array_equal(g, w), "%s plane %s: %d samples differ" % (name, ch, int((g != w).sum()))


CASES = [
    # W, H, kind, gain, border  -- sizes chosen to hit: full tiles, ragged right/bottom tiles of every
    # residue class mod 4, odd widths, and the (W+16) mod 128 in (32,48) overflow quirk (409, 260)
    (640, 500, "scene", 1.0, 4), (401, 367, "noise", 1.0, 4), (300, 260, "scene", 1.0, 4),
    (409, 389, "noise", 1.0, 4), (170, 154, "scene", 1.0, 4), (518, 275, "scene", 1.7, 4),
    (333, 301, "noise", 2.5, 3), (262, 258, "scene", 0.8, 0),
]


@needs_ref
@pytest.mark.parametrize("pattern", ["RGGB", "BGGR", "GRBG", "GBRG"])
@pytest.mark.parametrize("W,H,kind,gain,border", CASES)
def test_port_matches_reference_body(pattern, W, H, kind, gain, border):
    f = synth.BAYER_FILTERS[pattern]
    raw = synth.bayer_frame(W, H, f, seed=W + H) if kind == "scene" else synth.random_frame(W, H, seed=W * H)
    got = oracle.port().amaze(raw, f, gain, border)
    want = oracle.ref(det=True).amaze(raw, f, gain, border)
    for g, w, ch in zip(got, want, "RGB"):
        assert np.array_equal(g, w), "plane %s: %d samples differ" % (ch, int((g != w).sum()))


@needs_ref
def test_reference_self_noise_floor():
    """Stock reference (stale per-thread scratch) vs the zero-scratch variant: the reference's own
    schedule-dependent noise.  It must be small -- it is the floor under any parity claim."""
    f = synth.RGGB
    raw = synth.bayer_frame(1400, 1100, f, seed=77, noise_a=40.0)
    det = oracle.ref(det=True).amaze(raw, f)
    stock = oracle.ref(det=False).amaze(raw, f)
    frac = max(float((d != s).mean()) for d, s in zip(det, stock))
    assert frac < 1e-3


def test_cfa_sites_pass_through_config2_size():
    """At the BASELINE configs[1] frame size (8192x5464): native CFA samples come back as
    G - (G - cfa), i.e. the input up to fp32 rounding of values <= 1 (<< 0.05 on the 0..65535 scale)."""
    f = synth.RGGB
    W, H = 8192, 5464
    raw = synth.bayer_frame(W, H, f, seed=1002)
    r, g, b = oracle.port().amaze(raw, f)
    planes = [r, g, b]
    for pr in range(2):
        for pc in range(2):
            k = int(synth.fc(f, pr, pc))
            got = planes[k][pr::2, pc::2]
            want = raw[pr::2, pc::2]
            assert np.max(np.abs(got - want)) < 0.05
    assert min(p.min() for p in planes) >= 0.0
