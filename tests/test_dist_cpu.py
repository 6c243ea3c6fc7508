"""N>1 host logic on CPU: world_size-2 gloo process group.  Checks the sharding arithmetic
(art_b200/dist.py) and the max-over-ranks reduction bench.py uses; no GPU, no compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import art_b200
from art_b200 import dist as adist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H = 5464
        out = {}
        for name, method in (("amaze", art_b200.BAYER_AMAZE), ("rcd", art_b200.BAYER_RCD)):
            band = adist.row_bands(H, world, method)[rank]
            rows = torch.zeros(H, dtype=torch.int32)
            rows[band["out"][0]:band["out"][1]] = 1
            dist.all_reduce(rows)                                  # every row produced exactly once
            out[name + "_cover"] = bool((rows == 1).all())
            out[name + "_band"] = band
        frames = adist.frames_for_rank(7, rank, world)
        cnt = torch.zeros(7, dtype=torch.int32)
        cnt[frames] = 1
        dist.all_reduce(cnt)
        out["frames_cover"] = bool((cnt == 1).all())
        out["max"] = adist.max_over_ranks(10.0 + rank, dist)      # the slowest rank defines the step time
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r]["amaze_cover"] and res[r]["rcd_cover"] and res[r]["frames_cover"]
        assert res[r]["max"] == 11.0
    # band cuts sit on the reference tile grids and the halo/mirror rows are inside the frame
    a0, a1 = res[0]["amaze_band"], res[1]["amaze_band"]
    assert a0["out"][0] == 0 and a0["out"][1] == a1["out"][0] and a1["out"][1] == 5464 and a0["out"][1] % 128 == 0
    assert a0["need"] == (0, a0["out"][1] + 16) and a1["need"] == (a1["out"][0] - 16, 5464)
    r0, r1 = res[0]["rcd_band"], res[1]["rcd_band"]
    assert (r0["out"][1] - 9) % 176 == 0 and r1["need"][0] == r1["out"][0] - 9


@pytest.mark.parametrize("H", [64, 300, 1536, 5464, 8192])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("method", [art_b200.BAYER_AMAZE, art_b200.BAYER_RCD])
def test_row_bands_partition(H, world, method):
    period, offset, halo = adist.band_grid(method)
    bands = adist.row_bands(H, world, method)
    assert len(bands) == world
    edges = [b["out"] for b in bands if b["out"][1] > b["out"][0]]
    assert edges[0][0] == 0 and edges[-1][1] == H
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0
    for (r0, r1) in edges:
        assert r0 == 0 or (r0 - offset) % period == 0
        assert r1 == H or (r1 - offset) % period == 0
    for b in bands:
        if b["out"][1] > b["out"][0]:
            assert 0 <= b["need"][0] <= b["out"][0] and b["out"][1] <= b["need"][1] <= H


# ---- one frame across GPUs: frame bands, band plans, and the MAD histogram all-reduce (ABI version 3) ----
def _band_params():
    import numpy as np
    from art_b200.api import DenoiseParams, DevelopParams
    from art_b200 import synth
    return DevelopParams(method=art_b200.BAYER_AMAZE, filters=synth.RGGB, mul=(1.9, 1.0, 1.6), do_clip=True, cam2work=np.eye(3),
                         denoise=DenoiseParams(luminance=30, luminanceDetail=50, chrominance=15), fattal=None, wprof=np.eye(3))


@pytest.mark.parametrize("H", [600, 1600, 5456, 8184])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_frame_bands_and_plans(H, world):
    from art_b200 import api
    params = _band_params()
    Hr, Wr = H + 8, 648                     # raw frame: the developed frame + the 4-pixel border either side
    bands = adist.frame_bands(H, world)
    assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == H
    for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
        assert a1 == b0 and a1 % 2 == 0
    for b0, b1 in bands:
        if b1 <= b0:
            continue
        p = api.band_plan(params, Wr, Hr, b0, b1, 200)
        assert (p.own_begin, p.own_end) == (b0, b1)
        assert p.band_begin % 50 == 0 and 0 <= p.band_begin <= b0 and b1 <= p.band_end <= H
        assert p.band_begin == 0 or b0 - p.band_begin >= 200          # at least the halo, rounded down to the block / decimation grid
        assert p.band_end == H or p.band_end - b1 == 200
        assert p.dm_begin % 128 == 0 and (p.dm_end % 128 == 0 or p.dm_end == Hr)            # AMaZE's tile grid
        assert p.dm_begin <= p.band_begin + 4 and p.dm_end >= p.band_end + 4
        assert 0 <= p.raw_begin <= max(0, p.dm_begin - 16) and min(Hr, p.dm_end + 16) <= p.raw_end <= Hr
    with pytest.raises(art_b200.HotPathError):
        api.band_plan(params, Wr, Hr, 1, H, 200)                      # an odd first row would split a subband row between two ranks
    with pytest.raises(art_b200.HotPathError):
        api.band_plan(params, Wr, Hr, 0, H, 10)                       # a halo shorter than the stages' reach


def _mad_worker(rank, world, port, q):
    """MadRgb over ranks: int32 histograms of the owned coefficient rows, summed with all_reduce, give the whole subband's median bin."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        H, W = 1000, 37
        coeff = (rng.standard_normal((H // 2, W)) * 300).astype(np.float32)        # a once-decimated subband of the frame, same on every rank
        b0, b1 = adist.frame_bands(H, world)[rank]
        own = coeff[b0 // 2:(b1 + 1) // 2]
        v = np.minimum(np.abs(own.astype(np.int32)), 65535)
        h = torch.from_numpy(np.bincount(v.ravel(), minlength=65536).astype(np.int32))
        dist.all_reduce(h)
        whole = np.bincount(np.minimum(np.abs(coeff.astype(np.int32)), 65535).ravel(), minlength=65536)
        q.put((rank, bool((h.numpy() == whole).all()), int(h.sum()) == coeff.size))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_mad_histograms():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(same and count for _, same, count in res)
