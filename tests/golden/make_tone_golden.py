"""Generate tests/golden/tone_film.npz from the reference's own tone-curve code compiled in place (oracle/_ref; needs
/root/reference): the Standard Film Curve profile (rtdata/profiles/Standard Film Curve.arp) as the LUTs / curve stages the hot
path takes, one small input frame and the reference's output after NeutralToneCurve::BatchApply and after apply_satcurve."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tone_util as tu  # noqa: E402


def main():
    lut, _ = tu.build_lut(tu.FILM_CURVE, tu.LINEAR)
    px, py = tu.polyline(tu.FILM_CURVE, tu.LINEAR, 1)
    satlut = tu.sat_lut(tu.FILM_SAT)
    lut_c, _ = tu.build_lut(tu.FILM_CURVE, tu.LINEAR, contrast=30)
    ab = np.zeros(2)
    import oracle
    oracle.ref().lib.artref_tone_contrast_ab(30, tu.F(1.0), ab.ctypes.data_as(tu.dp))
    planes = tu.frame(64, 96, 2024)
    neutral = tu.ref_neutral(planes, tu.FILM_CURVE, tu.LINEAR)
    sat = tu.ref_satcurve(neutral, tu.FILM_SAT)
    neutral_c = tu.ref_neutral(planes, tu.FILM_CURVE, tu.LINEAR, contrast=30, om=tu.SRGB_XYZ)
    to_out, to_work = tu.out_matrices(tu.PROPHOTO, tu.PROPHOTO_INV, tu.SRGB_XYZ)
    np.savez_compressed(os.path.join(HERE, "tone_film.npz"), lut=lut, poly_last=np.array([px[-1], py[-1]]), satlut=satlut, lut_contrast=lut_c, contrast_ab=ab,
                        to_out=to_out, to_work=to_work, inp=np.stack(planes), neutral=np.stack(neutral), sat=np.stack(sat), neutral_contrast=np.stack(neutral_c))
    print("wrote tone_film.npz", os.path.getsize(os.path.join(HERE, "tone_film.npz")), "bytes")


if __name__ == "__main__":
    main()
