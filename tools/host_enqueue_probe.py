#!/usr/bin/env python3
"""How long does the HOST take to queue one develop step (all launches, events, small uploads)?  If this approaches the device
time of a step, the batch queue becomes host-bound."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import art_b200  # noqa: E402
from art_b200 import synth  # noqa: E402
from art_b200.api import DenoiseParams, DevelopParams  # noqa: E402

W, H = 8192, 5464
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
CAM2WORK = np.array([[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]], np.float64)
params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=CAM2WORK,
                       denoise=DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15), fattal=(30, 20, 0), wprof=PROPHOTO)
hp = art_b200.HotPath(0)
raw = synth.bayer_frame(W, H, synth.RGGB, seed=1002)
pitch = (W + 31) // 32 * 32
d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
d_raw[:, :W] = torch.from_numpy(raw).cuda()
outs = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
for _ in range(3):
    hp.develop_dev(params, W, H, d_raw.data_ptr(), pitch, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), pitch)
hp.sync()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    hp.develop_dev(params, W, H, d_raw.data_ptr(), pitch, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), pitch)
    t1 = time.perf_counter()
    hp.sync()
    t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("host enqueue ms:", [round(a * 1e3, 2) for a, _ in ts], " enqueue+device ms:", [round(b * 1e3, 2) for _, b in ts])
