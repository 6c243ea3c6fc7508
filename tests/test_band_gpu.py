"""One frame across GPUs (ABI version 3, art_hp_develop_band_dev): the bands of several ranks, developed one after the other on this
GPU, reproduce the single-GPU frame on the rows they own.

The only exchange between ranks is the int32 sum of the wavelet subbands' MAD histograms.  On one GPU the ranks cannot run at the
same time, so the test plays the collective in two passes through art_hp_set_allreduce: pass 1 records every rank's histograms
(call by call), pass 2 replaces each buffer with the sum over the ranks -- what ncclAllReduce delivers when the ranks run side
by side (tools/band_check.py does that on 2+ GPUs with NCCL).  The histograms depend only on the input frame, so the replay is exact.

Tolerance: the box blurs of the wavelet shrinkage and of the DCT stage are running sums that restart at the band's first row, so
the bands are not bit-identical to the frame; north_star allows 1e-4 relative, the test holds 1e-5 of the pixel's largest channel (+ 0.02 on the
0..65535 scale), the measure tests/test_fullsize_gpu.py uses.
"""
import numpy as np
import pytest

import art_b200
from art_b200 import dist as adist
from art_b200 import synth
from art_b200.api import ChainParams, DenoiseParams, DevelopParams, SharpenParams

pytestmark = pytest.mark.gpu

PROPHOTO = np.array([[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]], np.float64)
CAM2WORK = np.array([[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]], np.float64)


def develop_in_bands(hp, params, raw, world, halo=200):
    import torch
    H, W = raw.shape
    Ho, Wo = params.out_shape(H, W)
    pitch = (Wo + 31) // 32 * 32
    bands = [b for b in adist.frame_bands(Ho, world) if b[1] > b[0]]
    plans = [hp.band_plan(params, W, H, b0, b1, halo) for b0, b1 in bands]
    recorded = [[] for _ in plans]
    result = [np.zeros((Ho, Wo), np.float32) for _ in range(3)]

    def run(k, hook):
        plan = plans[k]
        d_raw = torch.full((H, W), float("nan"), dtype=torch.float32, device="cuda")      # the rank holds only the rows its plan names
        d_raw[plan.raw_begin:plan.raw_end] = torch.from_numpy(raw[plan.raw_begin:plan.raw_end]).cuda()
        outs = [torch.full((Ho, pitch), float("nan"), dtype=torch.float32, device="cuda") for _ in range(3)]
        hp.set_allreduce(hook)
        try:
            hp.develop_band_dev(params, W, H, d_raw.data_ptr(), W, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), pitch, plan)
            hp.sync()
        finally:
            hp.set_allreduce(None)
        return [o[plan.own_begin:plan.own_end, :Wo].cpu().numpy() for o in outs]

    def as_tensor(ptr, count):
        class _Mem:          # a view of the library's device buffer
            __cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Mem(), device="cuda")

    for k in range(len(plans)):            # pass 1: every rank's own histograms, call by call
        def record(ptr, count, stream, k=k):
            torch.cuda.synchronize()
            recorded[k].append(as_tensor(ptr, count).clone())
            return 0
        run(k, record)
    ncalls = len(recorded[0])
    assert ncalls > 0 and all(len(r) == ncalls for r in recorded)
    sums = [sum(recorded[k][c] for k in range(len(plans))) for c in range(ncalls)]
    for k, plan in enumerate(plans):       # pass 2: the all-reduce's result
        calls = iter(sums)

        def replay(ptr, count, stream):
            torch.cuda.synchronize()
            as_tensor(ptr, count).copy_(next(calls))
            torch.cuda.synchronize()
            return 0
        rows = run(k, replay)
        for c in range(3):
            result[c][plan.own_begin:plan.own_end] = rows[c]
    return result, plans


@pytest.mark.parametrize("world,sharpen", [(2, False), (3, True)])
def test_bands_reproduce_the_frame(hot_path, world, sharpen):
    W, H = 1096, 1608 if world == 2 else 2008
    raw = synth.bayer_frame(W, H, synth.RGGB, seed=31 + world)
    x = np.arange(65536, dtype=np.float64) / 65535
    curve = (x ** 0.8 * 65535).astype(np.float32)
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=CAM2WORK,
                           denoise=DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15), fattal=None, wprof=PROPHOTO,
                           sharpen=SharpenParams(radius=0.5, amount=200) if sharpen else None,
                           chain=ChainParams(exposure=(0.3, 0.0), tonecurve=(0, curve)) if sharpen else None)
    want = hot_path.develop(raw, params)
    got, plans = develop_in_bands(hot_path, params, raw, world)
    worst = 0.0
    mag = np.maximum.reduce([np.abs(w_) for w_ in want])          # the pixel's largest channel, as in test_fullsize_gpu.py
    for g, w_, ch in zip(got, want, "RGB"):
        assert np.isfinite(g).all()
        err = np.abs(g - w_)
        lim = 1e-5 * mag + 0.02
        worst = max(worst, float((err / (mag + 0.02)).max()))
        assert (err <= lim).all(), "%s: %d of %d beyond 1e-5 of the pixel scale, worst %g (row %d)" % (
            ch, int((err > lim).sum()), g.size, float((err / (mag + 0.02)).max()), int(np.argmax((err > lim).any(axis=1))))
    print("\n[bands x%d%s] worst difference from the single-GPU frame %.3g of the pixel scale; plans %s" % (world, " + USM + chain" if sharpen else "", worst, plans))


def test_band_without_a_collective_is_refused(hot_path):
    import torch
    W, H = 640, 1208
    params = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=CAM2WORK,
                           denoise=DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15), fattal=None, wprof=PROPHOTO)
    plan = hot_path.band_plan(params, W, H, 0, 600)
    d = torch.zeros((H, W), dtype=torch.float32, device="cuda")
    o = [torch.zeros((H, W), dtype=torch.float32, device="cuda") for _ in range(3)]
    with pytest.raises(art_b200.HotPathError):
        hot_path.develop_band_dev(params, W, H, d.data_ptr(), W, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), W, plan)
    hot_path.sync()
    fat = DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=CAM2WORK,
                        denoise=None, fattal=(30, 20, 0), wprof=PROPHOTO)
    with pytest.raises(art_b200.HotPathError):      # the Poisson solve is a transform of the whole frame
        hot_path.develop_band_dev(fat, W, H, d.data_ptr(), W, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), W, plan)
