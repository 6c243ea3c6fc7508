"""Pin the wavelet oracle (oracle/wavelet_port.c) against the reference's own wavelet headers compiled
unmodified (oracle/_ref): every subband of every level, then the reconstruction.  Bit-exact."""
import numpy as np
import pytest

import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built and /root/reference absent")


def image(H, W, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 20000 + 15000 * np.sin(0.07 * x) * np.cos(0.05 * y) + rng.normal(0, 800, size=(H, W))
    return np.clip(img, 0, 65535).astype(np.float32)


@needs_ref
@pytest.mark.parametrize("W,H", [(128, 96), (131, 97), (200, 77), (64, 150)])
@pytest.mark.parametrize("maxlvl,subsamp", [(5, 1), (3, 1), (1, 1), (6, 1)])   # subsampling == 0 overflows the reference's own half-size ping-pong buffers (dec.h L159-175): not a valid use
def test_port_matches_reference(W, H, maxlvl, subsamp):
    img = image(H, W, seed=W * H + maxlvl)
    a = oracle.port().wavelet(img, maxlvl, subsamp)
    b = oracle.ref().wavelet(img, maxlvl, subsamp)
    assert a.maxlevel() == b.maxlevel() == maxlvl
    for lvl in range(maxlvl):
        assert a.dims(lvl) == b.dims(lvl)
        for d in (1, 2, 3):
            assert np.array_equal(a.band(lvl, d), b.band(lvl, d)), "level %d band %d" % (lvl, d)
    assert np.array_equal(a.band(maxlvl - 1, 0), b.band(maxlvl - 1, 0)), "lowpass"
    # shrink something so that the reconstruction is not the identity, the same way on both sides
    for w in (a, b):
        w.band(0, 1)[...] *= 0.5
        w.band(maxlvl - 1, 3)[...] *= 0.25
    ra = a.reconstruct(img.copy(), blend=1.0)
    rb = b.reconstruct(img.copy(), blend=1.0)
    assert np.array_equal(ra, rb)
    a.close(); b.close()


@needs_ref
def test_port_matches_reference_8_levels():
    """FTblockDN's maximum depth (levwav <= 8, FTblockDN.cc L2246-2293) on a frame large enough for skip = 64."""
    test_port_matches_reference(701, 523, 8, 1)


def test_perfect_reconstruction_property():
    """Size-independent property: decompose + reconstruct returns the input (to float rounding) away from the borders."""
    img = image(240, 320, 3)
    w = oracle.port().wavelet(img, 5, 1)
    out = w.reconstruct(img.copy())
    w.close()
    inner = (slice(70, -70), slice(70, -70))
    assert np.max(np.abs(out[inner] - img[inner])) < 0.05      # on a 0..65535 scale
