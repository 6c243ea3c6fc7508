#!/usr/bin/env python3
"""How much does a second context (its own streams and scratch) developing its own frame on the same GPU add?
Kernels whose parallelism is bounded by line count (box blurs, IIR sweeps, the block-DCT's serial phases) leave SMs idle that
another frame's kernels can use."""
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
raw = synth.bayer_frame(W, H, synth.RGGB, seed=1002)
pitch = (W + 31) // 32 * 32
for nctx in (1, 2, 3):
    ctxs = []
    for i in range(nctx):
        hp = art_b200.HotPath(0)
        d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
        d_raw[:, :W] = torch.from_numpy(raw).cuda()
        outs = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
        ctxs.append((hp, d_raw, outs))

    def step():
        for hp, d_raw, outs in ctxs:
            hp.develop_dev(params, W, H, d_raw.data_ptr(), pitch, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), pitch)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    n = 6
    t = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / n
    print("%d context(s): %.2f ms per step of %d frame(s) = %.1f Mpixel/s" % (nctx, dt * 1e3, nctx, nctx * W * H / dt / 1e6), flush=True)
    for hp, _, _ in ctxs:
        hp.close()
    del ctxs
    torch.cuda.empty_cache()
