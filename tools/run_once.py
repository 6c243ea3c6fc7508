#!/usr/bin/env python3
"""Run the device-resident hot path a few times on one synthetic frame (for ncu captures)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import art_b200  # noqa: E402
from art_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--method", default="amaze")
ap.add_argument("--width", type=int, default=8192)
ap.add_argument("--height", type=int, default=5464)
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
W, H = args.width, args.height
hp = art_b200.HotPath(0)
raw = synth.bayer_frame(W, H, synth.RGGB, seed=1002)
pitch = (W + 31) // 32 * 32
d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
d_raw[:, :W] = torch.from_numpy(raw).cuda()
outs = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
m = art_b200.BAYER_RCD if args.method == "rcd" else art_b200.BAYER_AMAZE
for _ in range(args.iters):
    hp.demosaic_bayer_dev(m, W, H, synth.RGGB, d_raw.data_ptr(), pitch, outs[0].data_ptr(), outs[1].data_ptr(),
                          outs[2].data_ptr(), pitch, 1.0, 4)
hp.sync()
print("done", float(outs[1][H // 2, W // 2]))
