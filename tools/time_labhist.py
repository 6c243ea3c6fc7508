"""Device time of art_hp_lab_histogram_dev at 8192x5464 (CUDA-event spans of the library).  python tools/time_labhist.py"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import art_b200                                   # noqa: E402
from art_b200.api import ChainParams              # noqa: E402
from test_oracle_chain import PROPHOTO, PROPHOTO_INV   # noqa: E402

W, H = 8192, 5464
hp = art_b200.HotPath(0)
rng = np.random.default_rng(1)
yy = torch.linspace(0.02, 1.0, H, device="cuda").reshape(H, 1)
base = (torch.from_numpy(rng.uniform(0.5, 1.0, (H, W)).astype(np.float32)).cuda() * yy * 60000).contiguous()
planes = [(base * s).contiguous() for s in (1.0, 0.8, 0.6)]
hist = np.zeros(65536, np.uint32)
c = ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, exposure=(0.3, 0.0)).c_struct()
call = lambda: hp._check(hp.lib.art_hp_lab_histogram_dev(hp.h, W, H, *[p.data_ptr() for p in planes], W, ctypes.byref(c), hist.ctypes.data_as(ctypes.c_void_p)))
for _ in range(2):
    call()
hp.profile_enable(True)
for _ in range(3):
    call()
prof = hp.profile_collect()
print({k: round(v[0] / 3, 3) for k, v in prof.items()}, int(hist.sum()) == W * H, int((hist > 0).sum()), "bins used")
