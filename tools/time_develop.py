#!/usr/bin/env python3
"""Per-kernel timing of the device-resident develop pipeline on one synthetic frame (CUDA events around every launch)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import art_b200  # noqa: E402
from art_b200 import synth  # noqa: E402
from art_b200.api import DenoiseParams, DevelopParams, SharpenParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=8192)
ap.add_argument("--height", type=int, default=5464)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--no-denoise", action="store_true")
ap.add_argument("--no-fattal", action="store_true")
ap.add_argument("--nl", type=int, default=0)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--usm", action="store_true", help="add the unsharp mask of configs[2]")
args = ap.parse_args()
W, H = args.width, args.height
PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
CAM2WORK = np.array([[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]], np.float64)
hp = art_b200.HotPath(0)
raw = synth.bayer_frame(W, H, synth.RGGB, seed=1002)
pitch = (W + 31) // 32 * 32
d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
d_raw[:, :W] = torch.from_numpy(raw).cuda()
outs = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]
dn = None if args.no_denoise else DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15)
params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=CAM2WORK, denoise=dn,
                       nl_strength=args.nl, fattal=None if args.no_fattal else (30, 20, 0), wprof=PROPHOTO,
                       sharpen=SharpenParams(radius=0.5, amount=200) if args.usm else None)


def step():
    hp.develop_dev(params, W, H, d_raw.data_ptr(), pitch, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), pitch)


step(); hp.sync()
t = time.time()
for _ in range(args.iters):
    step()
hp.sync()
dt = (time.time() - t) / args.iters
print("develop %dx%d: %.2f ms/frame = %.1f Mpixel/s (host clock, device resident), launches/frame %d" % (W, H, dt * 1e3, W * H / dt / 1e6, hp.launch_count() // (args.iters + 1)))
if args.profile:
    hp.profile_enable(True)
    for _ in range(args.iters):
        step()
    hp.sync()
    stats = hp.profile_collect()
    tot = sum(ms for ms, _ in stats.values())
    print("profiled total %.2f ms/frame" % (tot / args.iters))
    for name, (ms, calls) in sorted(stats.items(), key=lambda s: -s[1][0])[:45]:
        print("  %-28s %8.3f ms/frame  %5.1f%%  %d launches/frame" % (name, ms / args.iters, 100 * ms / tot, calls // args.iters))
print("checksum", float(outs[1][H // 2, W // 2]), float(outs[0].float().mean()))
