"""Deterministic synthetic sensor frames.

The reference has no test images; every parity test and the benchmark use frames
made here: a float32 CFA plane already in the value domain that
RawImageSource::scaleColors produces (reference rtengine/rawimagesource.cc
L2677-2859): integer-valued floats in 0..65535.

scene = 3 oriented sinusoids + 2 smooth-step edges + a zone-plate patch (Nyquist
content) + ~1 % of the area clipped (> 0.8*65535, exercises AMaZE's clip_pt8
branches); per-channel gains R 0.6 / G 1.0 / B 0.7 before mosaicking;
Poisson-Gaussian noise var = a*x + b.
"""
import numpy as np

RGGB = 0x94949494
BGGR = 0x16161616
GRBG = 0x61616161
GBRG = 0x49494949
BAYER_FILTERS = {"RGGB": RGGB, "BGGR": BGGR, "GRBG": GRBG, "GBRG": GBRG}


def fc(filters, row, col):
    """Colour of the CFA site -- reference rtengine/rawimage.h L186-189."""
    row = np.asarray(row)
    col = np.asarray(col)
    return (filters >> ((((row << 1) & 14) + (col & 1)) << 1)) & 3


def bayer_frame(width, height, filters=RGGB, seed=1001, noise_a=4.0, noise_b=100.0, block_rows=512):
    """Return a (height, width) float32 CFA plane, integer-valued, 0..65535."""
    rng = np.random.default_rng(seed)
    out = np.empty((height, width), dtype=np.float32)
    gains = np.array([0.6, 1.0, 0.7], dtype=np.float32)
    x = np.arange(width, dtype=np.float32)[None, :]
    cx, cy = 0.31 * width, 0.64 * height          # zone plate centre
    zr = 0.12 * min(width, height)                # zone plate radius
    hx, hy = 0.77 * width, 0.22 * height          # clipped-highlight blob centre
    hr = np.sqrt(0.01 * width * height / np.pi)   # ~1 % of the area
    cpat = np.array([[fc(filters, r, c) for c in range(2)] for r in range(2)])
    for r0 in range(0, height, block_rows):
        r1 = min(height, r0 + block_rows)
        y = np.arange(r0, r1, dtype=np.float32)[:, None]
        lum = (0.45
               + 0.18 * np.sin(0.0113 * x + 0.0071 * y)
               + 0.12 * np.sin(0.031 * x - 0.047 * y + 1.3)
               + 0.07 * np.sin(0.23 * x + 0.19 * y + 0.4))
        lum = lum + 0.15 * np.tanh((x - 0.55 * width + 0.3 * y) / 3.0)
        lum = lum + 0.10 * np.tanh((y - 0.4 * height - 0.2 * x) / 1.5)
        d2 = (x - cx) ** 2 + (y - cy) ** 2
        zone = 0.2 * np.cos(np.pi * d2 / (2.0 * zr)) * (d2 < zr * zr)
        lum = lum + zone
        hl = ((x - hx) ** 2 + (y - hy) ** 2) < hr * hr
        lum = np.clip(lum, 0.004, 0.92)
        rad = 200.0 + lum * (60000.0 - 200.0) / 0.92
        rad = np.where(hl, 65535.0 * 1.3, rad).astype(np.float32)
        blk = np.empty((r1 - r0, width), dtype=np.float32)
        for pr in range(2):
            for pc in range(2):
                rs = (pr - r0) % 2
                blk[rs::2, pc::2] = rad[rs::2, pc::2] * gains[cpat[pr, pc]]
        sigma = np.sqrt(noise_a * blk + noise_b, dtype=np.float32)
        blk = blk + sigma * rng.standard_normal(blk.shape, dtype=np.float32)
        out[r0:r1] = np.clip(np.rint(blk), 0.0, 65535.0)
    return out


# X-Trans 6x6 colour matrix (0 R, 1 G, 2 B), the layout dcraw reports for Fuji X-Trans sensors (passed explicitly);
# `xtrans_matrix(dy, dx)` gives the same pattern read from another origin (different camera crops start elsewhere in the 6x6 cell)
XTRANS = ((1, 1, 0, 1, 1, 2), (1, 1, 2, 1, 1, 0), (2, 0, 1, 0, 2, 1), (1, 1, 2, 1, 1, 0), (1, 1, 0, 1, 1, 2), (0, 2, 1, 2, 0, 1))
# camera RGB -> sRGB, 3x4 as RawImage::getRgbCam returns it: a fixed identity-ish matrix (rows sum to 1)
XTRANS_RGB_CAM = ((1.62, -0.48, -0.14, 0.0), (-0.21, 1.45, -0.24, 0.0), (0.02, -0.52, 1.50, 0.0))


def xtrans_matrix(dy=0, dx=0):
    return np.array([[XTRANS[(r + dy) % 6][(c + dx) % 6] for c in range(6)] for r in range(6)], dtype=np.int32)


def xtrans_frame(width, height, xtrans=None, seed=1004, noise_a=4.0, noise_b=100.0):
    """X-Trans CFA plane from the same scene as bayer_frame (the Bayer frame supplies the radiance, re-mosaicked 6x6)."""
    xt = np.asarray(xtrans if xtrans is not None else XTRANS, dtype=np.int32)
    rng = np.random.default_rng(seed)
    gains = np.array([0.6, 1.0, 0.7], dtype=np.float32)
    # bayer_frame with an all-green "CFA" (filters 0x55555555 -> colour 1 everywhere, gain 1) and no noise is the radiance itself
    rad = bayer_frame(width, height, 0x55555555, seed=seed, noise_a=0.0, noise_b=0.0)
    cmap = xt[np.arange(height)[:, None] % 6, np.arange(width)[None, :] % 6]
    blk = rad * gains[cmap]
    sigma = np.sqrt(noise_a * blk + noise_b, dtype=np.float32)
    blk = blk + sigma * rng.standard_normal(blk.shape, dtype=np.float32)
    return np.clip(np.rint(blk), 0.0, 65535.0).astype(np.float32)


def random_frame(width, height, seed=7, lo=0.0, hi=65535.0):
    """Uniform integer noise -- the harshest input for direction-select parity."""
    rng = np.random.default_rng(seed)
    return rng.integers(int(lo), int(hi) + 1, size=(height, width)).astype(np.float32)
