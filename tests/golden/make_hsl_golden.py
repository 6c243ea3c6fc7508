"""Golden FlatCurve polylines for the HSL equalizer tests: (poly_x, poly_y, dyByDx) as the reference's own FlatCurve constructor (flatcurves.cc,
compiled in place in oracle/_ref) builds them from the control points tests/test_oracle_hsl.py uses, at the poly_pn values the tests ask for.
Run in the build container (needs /root/reference for oracle/_ref):  python tests/golden/make_hsl_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import test_oracle_hsl as t    # noqa: E402

out = {}
curves = {"coeff": t.COEFF}
for name, (hc, sc, lc) in t.CASES.items():
    curves.update({name + "_h": hc, name + "_s": sc, name + "_l": lc})
curves["eleven"] = t.flat_points(11)
for name, pts in curves.items():
    for pn in (1000, 500):
        n, px, py, dy = t.polyline_from_reference(pts, True, pn)
        key = t.curve_key(pts, True, pn)
        out[key + "_n"] = np.array([n], np.int64)
        out[key + "_x"], out[key + "_y"], out[key + "_d"] = px, py, dy
np.savez_compressed(os.path.join(HERE, "hsl_polylines.npz"), **out)
print("wrote %d polylines" % (len(out) // 4))
