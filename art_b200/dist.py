"""Multi-GPU sharding arithmetic for the hot path (one process per GPU).

Two ways to use N GPUs:
  * replicas -- independent frames, one per rank at a time (the batch queue): `frames_for_rank`;
  * row bands -- one frame cut on the method's reference tile grid: `row_bands` gives every rank its
    output rows and the raw rows it must hold (band + halo + the mirror rows at the frame edges).
  * one developed frame per box -- `frame_bands` + HotPath.band_plan / develop_band_dev: every rank develops a band (+ halo) of the
    same frame; the only exchange is the int32 all-reduce of the wavelet subbands' MAD histograms (NCCL, inside the library).
torch.distributed is used for the barrier, for reducing the timing (`max_over_ranks`) and to hand the NCCL unique id to the ranks.
"""
from . import api


def frames_for_rank(n_frames, rank, world):
    """Round-robin assignment of a batch of independent frames."""
    return list(range(rank, n_frames, world))


def band_grid(method):
    """(period, offset, halo) of the method's reference tile grid in rows."""
    if method == api.BAYER_AMAZE:
        return 128, 0, 16        # amaze_demosaic_RT.cc L182-183: tiles at stride 128 from -16, 16-px margin
    if method == api.BAYER_RCD:
        return 176, 9, 9         # rcd_demosaic.cc L82-87, L305-316
    raise ValueError("unknown method %r" % (method,))


def row_bands(H, world, method):
    """Cut H rows into `world` contiguous bands on the tile grid, as evenly as the grid allows.

    Returns a list of dicts: out=(row_begin,row_end) the rows the rank produces, need=(lo,hi) the raw
    rows it reads (band + halo, plus [0,33) / [H-17,H) mirror rows when it touches the top / bottom).
    Ranks beyond the number of grid cells get an empty band.
    """
    period, offset, halo = band_grid(method)
    cuts = [0] + [r for r in range(offset + period, H, period) if r < H] + [H]
    if offset and cuts[1:2] and cuts[1] <= offset:
        cuts.pop(1)
    cells = len(cuts) - 1
    bands = []
    for k in range(world):
        c0 = cells * k // world
        c1 = cells * (k + 1) // world
        if c1 <= c0:
            bands.append({"out": (0, 0), "need": (0, 0)})
            continue
        r0, r1 = cuts[c0], cuts[c1]
        lo, hi = max(0, r0 - halo), min(H, r1 + halo)
        if r0 == 0:
            hi = max(hi, min(H, 33))
        if r1 == H:
            lo = min(lo, max(0, H - 17))
        bands.append({"out": (r0, r1), "need": (lo, hi)})
    return bands


def frame_bands(H, world, align=2):
    """One developed frame of H rows cut into `world` contiguous bands of owned rows, boundaries on multiples of `align`
    (2: subband rows of the once-decimated wavelet levels must not straddle two ranks).  Returns [(begin, end)] per rank; ranks beyond the
    number of cells get an empty band (begin == end)."""
    cells = max(1, H // align)
    cuts = [min(H, (cells * k // world) * align) for k in range(world)] + [H]
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def max_over_ranks(value, dist=None, device="cpu"):
    """MAX-reduce a python float over the process group (identity without one)."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
