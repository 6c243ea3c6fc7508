"""diagnostic: where does the dual demosaic (manual threshold) leave the oracle as the frame grows?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import art_b200, oracle
from art_b200 import synth
from test_dual_gpu import oracle_dual_bayer, PREFILTERS
hp = art_b200.HotPath(0)
f = 0x94949494
for (W, H) in [(1100, 700), (2100, 1300), (4100, 1000), (8192, 600)]:
    raw = synth.bayer_frame(W, H, f, seed=9)
    for second in ("bilinear", "vng4"):
        want, _ = oracle_dual_bayer(raw, f, "amaze", second, 20.0, False)
        got, _ = hp.dual_demosaic_bayer(art_b200.BAYER_AMAZE, 0 if second == "bilinear" else 1, raw, f, PREFILTERS[f], 20.0, False)
        for g, w, ch in zip(got, want, "RGB"):
            d = g != w
            if d.any():
                ys, xs = np.nonzero(d)
                print(W, H, second, ch, int(d.sum()), "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "max abs", float(np.abs(g - w).max()), flush=True)
            else:
                print(W, H, second, ch, "equal", flush=True)
    v = hp.demosaic_vng4(raw, PREFILTERS[f])
    from test_oracle_vng4 import vng4
    vw = vng4(oracle.port().lib, "artoracle_vng4", raw, PREFILTERS[f])
    print(W, H, "vng4 alone:", [int((a != b).sum()) for a, b in zip(v, vw)], flush=True)
