"""Device-resident timing of the STAGE_1 / STAGE_3 additions at configs[1]'s frame (8192x5464): HSL equalizer, tone equalizer, the per-pixel
chain with each tone-curve mode, the Lab histogram.  CUDA events on torch's stream (handed to the library), 3 warm-up + 5 timed calls each;
also the per-kernel split from the library's own profiling spans.  Run on a GPU box: python tools/time_stage1.py > gpurun_out/stage1.txt"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import art_b200                                            # noqa: E402
from art_b200.api import ChainParams, HslParams, ToneEqParams   # noqa: E402
from test_oracle_chain import PROPHOTO, PROPHOTO_INV, curve_lut  # noqa: E402
from test_oracle_hsl import CASES, COEFF, polyline               # noqa: E402

W, H = 8192, 5464
hp = art_b200.HotPath(0)
rng = np.random.default_rng(1)
base = torch.from_numpy(rng.uniform(500, 60000, (H, W)).astype(np.float32)).cuda()
planes = [(base * s).contiguous() for s in (1.0, 0.8, 0.6)]
torch.cuda.synchronize()


def timed(name, fn, px=W * H, reps=5):
    for _ in range(3):
        fn()
    hp.sync()
    hp.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hp.sync()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    hp.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    prof = hp.profile_collect()
    hp.profile_enable(False)
    split = ", ".join("%s %.3f" % (k, v[0] / reps) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6])
    print("%-34s %8.3f ms  %8.0f Mpixel/s   [%s]" % (name, ms, px / ms / 1e3, split), flush=True)


def fresh():
    return [p.clone() for p in planes]


ptr = lambda ps: [p.data_ptr() for p in ps]
cv = [polyline(c, True, 1000) for c in CASES["all"]] + [polyline(COEFF, True, 1000)]
hsl = HslParams(*[(c[1], c[2], c[3]) for c in cv], smoothing=5, scale=1.0, ws=PROPHOTO)
buf = fresh()
timed("hsl_equalizer (S, L, H; smoothing 5)", lambda: hp.hsl_equalizer_dev(W, H, *ptr(buf), W, hsl))
hsl0 = HslParams(*[(c[1], c[2], c[3]) for c in cv], smoothing=0, scale=1.0, ws=PROPHOTO)
timed("hsl_equalizer (no smoothing)", lambda: hp.hsl_equalizer_dev(W, H, *ptr(buf), W, hsl0))
for reg in (0, 1, 3):
    te = ToneEqParams((40, 25, 0, -20, -35), reg, 0.0, 1.0, PROPHOTO)
    c = te.c_struct()
    import ctypes
    timed("tone_equalizer (regularization %d)" % reg,
          lambda: hp._check(hp.lib.art_hp_tone_equalizer_dev(hp.h, W, H, *ptr(buf), W, ctypes.byref(c))))
for mode, name in ((0, "STD"), (1, "FILMLIKE"), (3, "WEIGHTEDSTD"), (4, "SATANDVALBLENDING"), (5, "LUMINANCE")):
    cp = ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, exposure=(0.3, 0.0), saturation=(20, 10), tonecurve=(mode, curve_lut(gamma=0.7, seed=mode)))
    b2 = fresh()
    timed("chain: exposure + sat + %s" % name, lambda: hp.color_chain_dev(W, H, *ptr(b2), W, cp))
