#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference itself (oracle/_ref/libartref_det.so, i.e. the
reference's own function bodies compiled where they lie under /root/reference).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Inputs are stored as uint16 (they are integer-valued), outputs as float32.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from art_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, method, W, H, pattern, kind, seed
    ("rcd_rggb_scene", "rcd", 230, 214, "RGGB", "scene", 11),
    ("rcd_gbrg_noise", "rcd", 203, 197, "GBRG", "noise", 12),
    ("amaze_rggb_scene", "amaze", 230, 214, "RGGB", "scene", 13),
    ("amaze_grbg_noise", "amaze", 203, 197, "GRBG", "noise", 14),
]


def make_input(W, H, pattern, kind, seed):
    f = synth.BAYER_FILTERS[pattern]
    if kind == "scene":
        return synth.bayer_frame(W, H, f, seed), f
    return synth.random_frame(W, H, seed), f


def main():
    ref = oracle.ref(det=True)
    for name, method, W, H, pattern, kind, seed in CASES:
        raw, f = make_input(W, H, pattern, kind, seed)
        if method == "rcd":
            r, g, b = ref.rcd(raw, f)
        else:
            if not hasattr(ref.lib, "artref_amaze"):
                continue
            r, g, b = ref.amaze(raw, f, 1.0, 4)
        assert np.all(raw == np.rint(raw)) and raw.min() >= 0 and raw.max() <= 65535
        np.savez_compressed(os.path.join(OUT, name + ".npz"), raw=raw.astype(np.uint16), filters=np.uint32(f),
                            red=r, green=g, blue=b, method=method)
        print(name, os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    main()
