#!/usr/bin/env python3
"""bench.py -- the hot path on synthetic Bayer / X-Trans frames, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload develop|c2|c3|c4|amaze|rcd]

A "step" is one pass of the hot path over one frame.  Workloads (BASELINE.json `configs`):
  develop  the metric -- "Mpixel/s end-to-end (demosaic+denoise+tonemap) on 45 MP Bayer": AMaZE, getImage gains + camera->working matrix (with
           the reference's border crop), RGB_denoise (luminance 30 / detail 50 / chrominance 15), Fattal (30 / 20) of an
           8192x5464 RGGB frame = configs[1]'s frame through the stages of the metric (art_hp_develop).  The default.
  c2       configs[2]: AMaZE + RGB_denoise + unsharp mask (radius 0.5, amount 200), 8192x5464
  c3       configs[3]: X-Trans 3-pass + chroma RGB_denoise + NL-means (50 / 80), 6240x4160
  c4       configs[4]: AMaZE + RGB_denoise + Fattal + the default colour chain (NEUTRAL film curve, saturation curve), 12288x8192
  amaze    configs[1] alone (demosaic only);  rcd  the configs[0]-style case
With N>1 (torchrun, one rank per GPU) every rank develops its own frames -- frames are independent objects (the batch queue), so there is no
data-path collective ("scaling": "weak"); torch.distributed carries the barrier and the max-over-ranks of the timed region only.
`--split frame --workload c2` (configs[2]) instead cuts ONE frame into row bands, one per rank (art_hp_develop_band_dev): each rank demosaics its
band on the frame's tile grid, carries `--halo` redundant rows either side through RGB_denoise and the unsharp mask, and the ranks exchange one thing,
the int32 MAD histograms of the wavelet subbands (ncclAllReduce inside the library, three per frame): "scaling": "strong", value = the frame's
pixels / the slowest rank's time.

Printed JSON (rank 0, one line): the driver contract plus
  roofline     dominant kernel: algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline the reference's own code (oracle/_ref, kind "reference") timed on this box's host cores on a bounded sample
  e2e          the same metric through the host-buffer C-ABI batch-queue call: pinned host planes in, the developed frame out in the
               reference's 16-bit wire format (Imagefloat::getScanline on the device), both copies inside the timed region
  clocks       nvidia-smi samples taken during the timed region
`--impl reference`: the reference's own functions (oracle/_ref) on ALL host threads, the full frame of the workload per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W45, H45 = 8192, 5464            # BASELINE configs[1..2]: 44.76 MP
BYTES_PER_PX = 16                # demosaic = 4 B read + 12 B written per pixel

# Algorithmic bytes per TILE pixel of each AMaZE kernel (DESIGN.md "AMaZE", table "planes per pass"):
# 4 B per full-resolution plane read or written, 2 B per half-resolution plane, once each.  A frame has
# ntiles * 160 * 160 tile pixels (the reference's 160x160 tiles at stride 128 cover every pixel 1.56x).
AMAZE_KERNEL_BYTES = {
    "k_fill": 4 * 0.64 + 8, "k_grad": 4 + 12, "k_dirinterp": 12 + 24, "k_hcd": 12 + 4, "k_vcd": 16 + 8,
    "k_hvwt": 24 + 2, "k_nyqtest": 8 + 0.5, "k_green": 18.5 + 10, "k_diag": 4 + 8, "k_rbdiag": 12 + 6,
    "k_rbint": 10 + 2, "k_greenrb": 20 + 6, "k_chroma": 2 + 2, "k_write": 10 + 12 * 0.64,
}
# DRAM bytes per launch from the committed ncu --set full capture (profiles/ncu_full_amaze_r1.csv,
# dram__bytes_read.sum + dram__bytes_write.sum at 8192x5464); None = not captured
NCU_TRAFFIC = {"k_dirinterp": 0.974235e9 + 1.551996e9, "k_rbdiag": 0.854833e9 + 0.355250e9, "k_grad": 0.282756e9 + 0.768683e9,
               "k_write": 0.565864e9 + 0.500666e9, "k_vcd": 1.110983e9 + 0.500444e9, "k_split": 0.081761e9 + 0.071985e9,
               "rcd_kernel": 181.060608e6 + 475.030528e6,
               # profiles/r1_develop_ncu_full.csv (isolated replays, cold L2; a 45 MB plane written by a shrink kernel can stay in the 126 MB L2)
               "k_dn_blocks": 358.347e6 + 1137.494e6, "k_fbox_h": 44.786e6 + 3.397e6, "k_fbox_v": 44.808e6 + 6.248e6,
               "k_sf_apply": 134.296e6 + 18.761e6, "k_sf_AB": 89.535e6 + 18.726e6, "k_mad_hist": 44.784e6, "k_wav_sy_sub": 358.56e6 + 137.405e6,
               "k_fat_dct_rows": 184.989e6 + 316.294e6, "k_fat_dct_solve": 370.647e6 + 327.235e6, "k_fat_dct_exp": 370.152e6 + 156.166e6,
               # profiles/r2_ncu_k_shrink_*_merged.txt, r2_ncu_k_mad_hist_all_merged.txt: the merged launches (a, b and L of a frame in one grid) at 8192x5464;
               # k_mad_hist_all has two launches per frame (15 luminance subbands: 0.67 GB, 30 chroma subbands: 1.34 GB), the mean is quoted
               "k_shrink_v": 6.030872e9 + 1.984292e9, "k_shrink_h": 2.124376e9 + 2.047802e9, "k_shrink_sf": 2.077135e9 + 1.971890e9,
               "k_mad_hist_all": (1.340461e9 + 7.640832e6) * 0.75}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ART_BENCH_WORKLOAD", "develop"), choices=["develop", "c2", "c3", "c4", "amaze", "rcd"])
    ap.add_argument("--method", default=None, choices=["amaze", "rcd"], help="alias: --workload amaze|rcd")
    ap.add_argument("--cpu-sample", default="2048x1366", help="frame of the bounded CPU sample (cpu_baseline leg of our arm)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: seconds of CPU work the timed steps may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split", default=os.environ.get("ART_BENCH_SPLIT", "replicas"), choices=["replicas", "frame"],
                    help="N > 1: replicas = one frame per rank (weak scaling, the default); frame = ONE frame cut into row bands across the ranks "
                         "(strong scaling, configs[2]: ncclAllReduce of the MAD histograms inside the library; workload c2)")
    ap.add_argument("--halo", type=int, default=200, help="--split frame: rows of redundant halo either side of a band")
    a = ap.parse_args()
    if a.method:
        a.workload = a.method
    a.method = "rcd" if a.workload == "rcd" else "amaze"
    dw, dh = {"c3": (6240, 4160), "c4": (12288, 8192)}.get(a.workload, (W45, H45))
    a.width, a.height = a.width or dw, a.height or dh
    return a


def tensor_peak():
    """TF32 dense peak for a kernel timed inside a long step: half the measured sustained bf16 matmul rate (MEASURED_PEAKS.json; Blackwell's dense TF32
    rate is half its bf16 rate), else half the profiling guide's 2250 TFLOP/s nominal figure."""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops_sustained"]) / 2.0, "measured (MEASURED_PEAKS.json: bf16_tflops_sustained / 2 for kind::tf32)"
    except Exception:
        return 1125.0, "fallback (B200_PROFILING.md: 2250 TFLOP/s bf16 nominal / 2)"


# kernels bounded by the tensor pipe, not HBM: TF32 flops per developed pixel.  k_dn_blocks: four 64^3 products per 64x64 block at stride 25, each
# run as three TF32 passes (3xTF32 split): 153 GFLOP of fp32-grade work = 459 GFLOP of tcgen05 kind::tf32 at 44.65 MP
TENSOR_KERNEL_FLOPS_PER_PX = {"k_dn_blocks": 459.0e9 / (8184.0 * 5456.0)}


def kernel_roofline(name, kernel_ms, geo, launches, hbm_peak):
    """roofline entry of one kernel: (bound, achieved, peak, unit, frac, algorithmic work per launch, description, peak source or None)"""
    if name in TENSOR_KERNEL_FLOPS_PER_PX:
        tp, how = tensor_peak()
        flops = TENSOR_KERNEL_FLOPS_PER_PX[name] * geo[2] * geo[3] / max(1, launches)
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return "tensor", ach, tp, "TFLOP/s", ach / tp, flops, "tcgen05 kind::tf32 flops of the block DCTs (3 passes per product)", how
    kb = kernel_bytes(name, *geo, launches=launches)
    if kb is None:
        return "hbm", None, hbm_peak, "GB/s", None, None, None, None
    ach = kb[0] / (kernel_ms * 1e-3) / 1e9
    return "hbm", ach, hbm_peak, "GB/s", ach / hbm_peak, kb[0], kb[1], None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_runner(method, raw, filters):
    """(callable() -> None, kind, threads).  Prefers the reference's own function bodies (oracle/_ref, the
    stock build exactly as the reference ships it).  Output planes are allocated once, outside the timed
    region (the reference allocates red/green/blue at load time, rawimagesource.cc L1458-1460), and the
    OpenMP thread count is the faster of {all hardware threads, half of them} on this box."""
    import numpy as np
    import oracle
    H, W = raw.shape
    out = [np.zeros((H, W), np.float32) for _ in range(3)]
    try:
        ref = oracle.ref(det=False)
        fn = ref.rcd if method == "rcd" else ref.amaze
        hw = os.cpu_count() or ref.max_threads()
        best = None
        for nt in sorted({hw, max(1, hw // 2)}, reverse=True):
            kw = {"nthreads": nt, "out": out}
            fn(raw, filters, **kw)                       # warm-up at this thread count
            t0 = time.perf_counter()
            fn(raw, filters, **kw)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, nt)
        nt = best[1]
        return (lambda: fn(raw, filters, nthreads=nt, out=out)), "reference", nt
    except (OSError, FileNotFoundError):
        port = oracle.port()
        fn = port.rcd if method == "rcd" else port.amaze
        return (lambda: fn(raw, filters)), "port", os.cpu_count() or 1


PROPHOTO = [[0.7976749, 0.1351917, 0.0313534], [0.2880402, 0.7118741, 0.0000857], [0.0, 0.0, 0.8252100]]   # iccmatrices.h xyz_prophoto
PROPHOTO_INV = [[1.3459433, -0.2556075, -0.0511118], [-0.5445989, 1.5081673, 0.0205351], [0.0, 0.0, 1.2118128]]
CAM2WORK = [[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]]      # a fixed camera->working matrix (synthetic camera)
MUL = (1.9, 1.0, 1.6)                                                           # rm, gm, bm of getImage for that camera
DN = dict(luminance=30.0, luminanceDetail=50.0, luminanceDetailThreshold=0, chrominance=15.0, chrominanceRedGreen=0.0,
          chrominanceBlueYellow=0.0, gamma=1.7, scale=1.0)
DN_CHROMA = dict(DN, luminance=0.0, luminanceDetail=0.0)
FATTAL = (30, 20, 0)                                                            # threshold, amount, satcontrol (procparams.cc L2074-2079)
PIPELINES = ("develop", "c2", "c3", "c4")
FILM_CURVE = [1, 0, 0, 0.11, 0.09, 0.32, 0.47, 0.66, 0.87, 1, 1]               # rtdata/profiles/Standard Film Curve.arp
FILM_SAT = [1, 0, 0.48, 0.34, 0.35, 1, 0.48, 0.35, 0.35]


def workload_text(wl, W, H):
    mp = W * H / 1e6
    return {
        "develop": "metric pipeline on configs[1]'s frame: AMaZE demosaic + gains/matrix (border crop 4) + RGB_denoise (lum 30, detail 50, chroma 15) + "
                   "Fattal (30/20), %dx%d synthetic RGGB (%.2f MP), art_hp_develop" % (W, H, mp),
        "c2": "configs[2]: AMaZE + gains/matrix + RGB_denoise (lum 30, detail 50, chroma 15) + unsharp mask (radius 0.5, amount 200), %dx%d synthetic RGGB "
              "(%.2f MP), art_hp_develop" % (W, H, mp),
        "c3": "configs[3]: X-Trans 3-pass demosaic + gains/matrix (border crop 7) + chroma RGB_denoise (15) + NL-means (50/80), %dx%d synthetic X-Trans "
              "(%.2f MP), art_hp_develop" % (W, H, mp),
        "c4": "configs[4]: AMaZE + gains/matrix + RGB_denoise + Fattal (30/20) + default colour chain (NEUTRAL Standard Film Curve, saturation curve), %dx%d "
              "synthetic RGGB (%.2f MP), art_hp_develop" % (W, H, mp),
        "amaze": "configs[1]: AMaZE demosaic, %dx%d synthetic RGGB (%.2f MP)" % (W, H, mp),
        "rcd": "configs[0]-style: RCD demosaic, %dx%d synthetic RGGB (%.2f MP)" % (W, H, mp),
    }[wl]


def make_raw(wl, W, H, seed):
    from art_b200 import synth
    if wl == "c3":
        return synth.xtrans_frame(W, H, synth.xtrans_matrix(1, 3), seed=seed)
    return synth.bayer_frame(W, H, synth.RGGB, seed=seed)


def develop_params(art_b200, wl):
    """DevelopParams of a pipeline workload (our arm).  The c4 tone curve comes from the committed golden fixture (the reference's own
    curve objects evaluated once, tests/golden/make_tone_golden.py): curves stay host-built, the hot path takes LUTs."""
    import numpy as np
    from art_b200 import synth
    from art_b200.api import ChainParams, DenoiseParams, DevelopParams, SharpenParams
    kw = dict(mul=MUL, do_clip=True, cam2work=CAM2WORK, wprof=PROPHOTO)
    if wl == "c3":
        return DevelopParams(method=art_b200.XTRANS_3PASS, xtrans=synth.xtrans_matrix(1, 3), rgb_cam=np.array(synth.XTRANS_RGB_CAM, np.float32),
                             denoise=DenoiseParams(**DN_CHROMA), nl_strength=50, nl_detail=80, **kw)
    kw.update(method=art_b200.BAYER_AMAZE, filters=0x94949494, initial_gain=1.0, border=4, denoise=DenoiseParams(**DN))
    if wl == "c2":
        return DevelopParams(sharpen=SharpenParams(radius=0.5, amount=200), **kw)
    if wl == "c4":
        z = np.load(os.path.join(ROOT, "tests", "golden", "tone_film.npz"))
        stages = [(1, z["poly_last"][:1], z["poly_last"][1:], 0, 0, 0), (0, None, None, 0, 0, 0)]
        return DevelopParams(fattal=FATTAL, chain=ChainParams(ws=PROPHOTO, iws=PROPHOTO_INV, tonecurve=(2, z["lut"]), stages=stages, satcurve=z["satlut"]), **kw)
    return DevelopParams(fattal=FATTAL, **kw)


# ---------------------------------------------------------------------------------------------------------------- roofline bookkeeping
# Algorithmic bytes per unit of the kernels that can dominate: fp32 planes, each plane a kernel must read or write
# counted once.  unit = what one LAUNCH covers.
def find_fast_dim(dim):
    v = dim - 1
    for sh in (1, 2, 4, 8, 16):
        v |= v >> sh
    d1 = v + 1
    for d in (d1 // 128 * 65, d1 // 64 * 33, d1 // 512 * 273, d1 // 16 * 9, d1 // 8 * 5, d1 // 16 * 11, d1 // 128 * 91,
              d1 // 4 * 3, d1 // 64 * 49, d1 // 16 * 13, d1 // 8 * 7, d1):
        if d >= dim:
            return d
    return dim


def kernel_bytes(name, Wr, Hr, W, H, nlev=5, launches=None):
    """(algorithmic bytes per launch, description) or None.  Wr x Hr: raw frame; W x H: the developed (cropped) frame.  launches = how often
    the kernel ran in a step: the shrink stage covers the 3 x 15 subbands of a frame in as many launches as the library chose (one since the
    channels were merged), so its bytes per launch are the frame's bytes over that count."""
    px, sub = W * H, ((W + 1) // 2) * ((H + 1) // 2)
    ntiles = ((Wr + 16 + 127) // 128) * ((Hr + 16 + 127) // 128)
    pad = (find_fast_dim(W) + 1) * (find_fast_dim(H) + 1)
    nsub = 3 * nlev
    nl = float(launches or 3)
    table = {
        # shrink stage, per frame (3 channels x 15 subbands): sf reads c (+ the L coefficient for a / b) and writes sf -> 8 + 12 + 12 B per coefficient
        "k_shrink_sf": (32.0 * sub * nsub / nl, "45 subbands of a frame / launches: coefficient (+ L coefficient for a, b) read, sf written"),
        "k_shrink_h": (3 * 8.0 * sub * nsub / nl, "45 subbands of a frame / launches: sf read, row sums written"),
        "k_shrink_v": (3 * 16.0 * sub * nsub / nl, "45 subbands of a frame / launches: row sums, sf, coefficient read, coefficient written"),
        "k_mad_hist_all": (3 * 4.0 * sub * nsub / nl, "45 subbands of a frame / launches, read once"),
        "k_dn_blocks": ((4 + 4 * (64.0 / 25.0) ** 2) * px, "residual read once + the windowed 64x64 blocks (stride 25) written"),
        "k_dn_gather": ((4 * (64.0 / 25.0) ** 2 + 8) * px, "blocks read, L read and written"),
        "k_dn_split": (24.0 * px, "3 planes read, 3 written"), "k_dn_merge": (24.0 * px, "3 planes read, 3 written"),
        "k_fat_dct_rows": (12.0 * pad, "padded plane: 4 read, 8 written"), "k_fat_dct_solve": (16.0 * pad, "8 read, 8 written"),
        "k_fat_dct_exp": (12.0 * pad, "8 read, 4 written"),
        "k_wav_sy_sub": (20.0 * px, "4 quarter-size planes read, 1 full written"), "k_wav_an_sub": (20.0 * px, "1 full plane read, 4 quarter-size written"),
        "k_wav_sy_haar": (20.0 * sub, "4 planes read, 1 written"), "k_wav_an_haar": (20.0 * sub, "1 plane read, 4 written"),
        "k_chain": (24.0 * px, "3 planes read, 3 written"), "k_scale_convert_crop": (24.0 * px, "3 planes read, 3 written"),
        "k_scanlines": (18.0 * px, "3 float planes read, 6 B/px written"),
        "k_nlm_tile": (44.0 * px, "SURVEY.md 8(d)"), "k_xtrans": (16.0 * Wr * Hr, "4 B read, 12 B written per pixel"),
        "rcd_kernel": (16.0 * Wr * Hr, "4 B read, 12 B written per pixel"),
    }
    if name in table:
        return table[name]
    if name in AMAZE_KERNEL_BYTES:
        return AMAZE_KERNEL_BYTES[name] * ntiles * 160 * 160, "tile pixels x planes per pass"
    return None


def roofline_top(per_step, kern, calls, geo, peak, n=8):
    out = []
    for k in sorted(per_step, key=lambda k: -per_step[k])[:n]:
        bound, ach, pk, unit, frac, work, what, how = kernel_roofline(k, kern[k], geo, calls[k], peak)
        e = {"kernel": k, "ms_per_step": round(per_step[k], 4), "launches_per_step": calls[k], "bound": bound, "achieved_GBps": None, "frac": None}
        if ach is not None and bound == "hbm":
            e.update(achieved_GBps=round(ach, 1), frac=round(frac, 4), traffic=NCU_TRAFFIC.get(k))
        elif ach is not None:
            e.update(achieved_TFLOPs=round(ach, 1), peak_TFLOPs=round(pk, 1), frac=round(frac, 4), traffic=NCU_TRAFFIC.get(k))
        out.append(e)
    return out


# ---------------------------------------------------------------------------------------------------------------- host-side placement
def use_all_host_threads(n=None):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm is the reference's OpenMP code on ALL host
    threads of the box, so the count is set explicitly (environment for a libgomp not loaded yet, omp_set_num_threads for one that is)."""
    import ctypes
    n = n or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def pin_rank_near_gpu(local, world):
    """One process per GPU: keep the rank's threads (and, by first touch, its pinned host buffers) on the CPUs of its GPU's NUMA node, and
    give every rank of that node its own slice of them, so eight ranks do not stack on the same cores.  Returns a description."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(path + "/numa_node").read())
        cl = open(path + "/local_cpulist").read().strip()
        cpus = []
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return {"numa_node": node, "pinned": False}
        per = max(1, len(cpus) // max(1, world))
        mine = cpus[(local * per) % len(cpus):][:per] or cpus
        os.sched_setaffinity(0, mine)
        return {"numa_node": node, "pinned": True, "cpus": "%d-%d" % (mine[0], mine[-1]), "ncpus": len(mine)}
    except Exception as ex:       # placement is an optimisation, never a requirement
        return {"pinned": False, "why": str(ex)[:120]}


# ---------------------------------------------------------------------------------------------------------------- the reference on the host
def cpu_pipeline_runner(wl, raw):
    """The workload's stages through the reference's OWN functions compiled in place (oracle/_ref, stock build), OpenMP on all host
    threads.  fftw3f is absent from this image, so the two FFTW call sites (64x64 block DCTs, 2-D REDFT00) run the oracle's double-precision
    stand-in, which is slower than FFTW would be.  Returns (callable, kind, threads)."""
    import ctypes
    import numpy as np
    import oracle
    from art_b200 import synth
    ncores = use_all_host_threads()
    ref = oracle.ref(det=False)
    use_all_host_threads()
    lib = ref.lib
    lib.artref_set_denoise_thread_limit(0)        # the reference's default: all OpenMP threads
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    Hr, Wr = raw.shape
    bd = 7 if wl == "c3" else 4                   # getImage crops the demosaiced frame by RawImageSource::border
    H, W = Hr - 2 * bd, Wr - 2 * bd
    out = [np.zeros((Hr, Wr), np.float32) for _ in range(3)]
    wp = np.array(PROPHOTO, np.float64)
    wpi = np.linalg.inv(wp)
    dn = DN_CHROMA if wl == "c3" else DN
    p = np.array([dn["luminance"], dn["luminanceDetail"], dn["luminanceDetailThreshold"], dn["chrominance"], dn["chrominanceRedGreen"],
                  dn["chrominanceBlueYellow"], dn["gamma"], dn["scale"]], np.float64)
    res = np.zeros(2, np.float32)
    xt = np.ascontiguousarray(synth.xtrans_matrix(1, 3), np.int32)
    cam = np.array(synth.XTRANS_RGB_CAM, np.float32)
    c1, c2, sat = np.array(FILM_CURVE, np.float64), np.array([0.0]), np.array(FILM_SAT, np.float64)
    wsf, iwsf = np.array(PROPHOTO, np.float32), np.array(PROPHOTO_INV, np.float32)
    thr = (ctypes.c_int * 4)(20, 80, 2000, 1200)
    P3 = lambda pl: [x.ctypes.data_as(fp) for x in pl]

    def run():
        if wl == "c3":
            assert lib.artref_xtrans(Wr, Hr, xt.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), cam.ctypes.data_as(fp), 3, 1, raw.ctypes.data_as(fp), *P3(out), 0, ncores) == 0
        else:
            ref.amaze(raw, synth.RGGB, nthreads=ncores, out=out)
        r, g, b = ref.scale_convert([np.ascontiguousarray(q[bd:Hr - bd, bd:Wr - bd]) for q in out], MUL, True, np.array(CAM2WORK, np.float64))
        assert lib.artref_rgb_denoise(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, p.ctypes.data_as(dp),
                                      wp.ctypes.data_as(dp), wpi.ctypes.data_as(dp), None, ctypes.c_float(0), None, None, None,
                                      res.ctypes.data_as(fp)) == 0
        if wl == "c3":       # Imagefloat::setMode(YUV) / NLMeans on Y / setMode(RGB), ipdenoise.cc L1173-1177
            w0, w1, w2 = [np.float32(v) for v in PROPHOTO[1]]
            Y = (r * w0 + g * w1) + b * w2
            u, v = Y - b, r - Y
            Y = np.ascontiguousarray(Y)
            assert lib.artref_nlmeans(Y.ctypes.data_as(fp), W, H, ctypes.c_float(65535.0), 50, 80, ctypes.c_float(1.0)) == 0
            r = v + Y
            b = Y - u
            g = (Y - w0 * r - w2 * b) / w1
        if wl in ("develop", "c4"):
            assert lib.artref_fattal(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, FATTAL[0], FATTAL[1], FATTAL[2],
                                     wp.ctypes.data_as(dp)) == 0
        if wl == "c2":
            assert lib.artref_usm(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, wp.ctypes.data_as(dp), ctypes.c_double(1.0),
                                  ctypes.c_double(20.0), ctypes.c_double(0.5), 200, thr, 0, 85, None) == 0
        if wl == "c4":
            assert lib.artref_tone_neutral(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, c1.ctypes.data_as(dp), len(c1),
                                           c2.ctypes.data_as(dp), len(c2), 0, ctypes.c_float(1.0), ctypes.c_double(1.0), wsf.ctypes.data_as(fp),
                                           iwsf.ctypes.data_as(fp), None) == 0
            lib.artref_tone_satcurve(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, sat.ctypes.data_as(dp), len(sat),
                                     c2.ctypes.data_as(dp), 1, ctypes.c_float(1.0), ctypes.c_double(1.0), wsf.ctypes.data_as(fp), iwsf.ctypes.data_as(fp))
    return run, "reference", ncores


def time_cpu_pipeline(wl, sample, seed, budget_s=20.0, max_runs=3):
    w, h = [int(v) for v in sample.lower().split("x")]
    raw = make_raw(wl, w, h, seed)
    run, kind, cores = cpu_pipeline_runner(wl, raw)
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < max_runs and (not ts or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": w * h / med / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d runs of the workload's stages on a %dx%d frame of the same synthetic scene (%.1f MP), median; reference functions compiled in "
                      "place, OpenMP on %d threads; FFTW call sites run the oracle's fp64 stand-in (fftw3f absent); the reference arm "
                      "(--impl reference) times the full frame" % (len(ts), w, h, w * h / 1e6, cores)}


def time_cpu(method, raw, filters, budget_s=12.0, max_runs=5):
    run, kind, cores = cpu_reference_runner(method, raw, filters)
    run()                                              # warm-up
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < max_runs and (not ts or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    H, W = raw.shape
    return {"value": W * H / med / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d full %dx%d frames after warm-up, median; OpenMP threads = %d (faster of all/half)" % (len(ts), W, H, cores)}


def reference_arm(args, config, W, H):
    """The reference's own CPU implementation of the workload on all host threads, the full frame per step.  Steps are bounded by
    --ref-budget seconds of CPU work (stated in the line); rank 0 alone runs it."""
    from art_b200 import synth
    wl = args.workload
    raw = make_raw(wl, W, H, 1002)
    if wl in PIPELINES:
        run, kind, cores = cpu_pipeline_runner(wl, raw)
    else:
        run, kind, cores = cpu_reference_runner(args.method, raw, synth.RGGB)
    t0 = time.perf_counter()
    run()                                                       # warm-up (page faults, OpenMP team start-up)
    first = time.perf_counter() - t0
    nsteps = max(1, min(args.steps, int(args.ref_budget / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(nsteps):
        run()
    dt = time.perf_counter() - t0
    val = nsteps * W * H / dt / 1e6
    cb = {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": kind,
          "sample": "%d timed steps (1 warm-up) of the FULL %dx%d frame through the reference's own functions compiled in place; OpenMP on %d threads "
                    "(set explicitly: torchrun exports OMP_NUM_THREADS=1); steps bounded by %.0f s of CPU work%s"
                    % (nsteps, W, H, cores, args.ref_budget,
                       "; FFTW call sites run the oracle's fp64 stand-in (fftw3f absent)" if wl in PIPELINES else "")}
    print(json.dumps({"impl": "reference", "metric": "Mpixel/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus,
                      "steps": nsteps, "warmup": 1, "ms_per_step": dt / nsteps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": config, "cpu_baseline": cb,
                      "note": "one CPU pipeline on the box's host threads; it does not multiply with --gpus",
                      "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def frame_split_arm(args, hp, dist, rank, world, local, W, H, wl, config, placement):
    """ONE frame across the ranks (configs[2]).  Every rank holds the same synthetic frame on the host, uploads the raw rows its band plan names,
    develops its band and downloads the rows it owns.  Device-resident `value`: band kernels only; `e2e`: upload + kernels + download per step."""
    import numpy as np
    import torch
    import art_b200
    from art_b200 import dist as adist
    if wl != "c2":
        print(json.dumps({"error": "--split frame is configs[2]: use --workload c2 (Fattal's solve and NL-means are not split)"}))
        return 2
    params = develop_params(art_b200, wl)
    Ho, Wo = params.out_shape(H, W)
    own = adist.frame_bands(Ho, world)[rank]
    plan = hp.band_plan(params, W, H, own[0], own[1], args.halo)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hp.set_stream(stream.cuda_stream)
    # the library's own communicator: rank 0 makes the id, torch.distributed hands it round
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(hp.comm_unique_id()), dtype=torch.uint8))
    if dist is not None:
        dist.broadcast(idt, 0)
    torch.cuda.synchronize()
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                      # NCCL's banner goes to stderr, stdout carries the ONE JSON line
    try:
        hp.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    raw = make_raw(wl, W, H, 1002)     # the SAME frame on every rank
    pitch = (W + 31) // 32 * 32
    opitch = (Wo + 31) // 32 * 32
    h_raw = torch.from_numpy(raw).pin_memory()
    d_raw = torch.full((H, pitch), float("nan"), dtype=torch.float32, device="cuda")
    d_out = [torch.zeros((Ho, opitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    h_out = [torch.zeros((own[1] - own[0], Wo), dtype=torch.float32).pin_memory() for _ in range(3)]

    def upload():
        d_raw[plan.raw_begin:plan.raw_end, :W].copy_(h_raw[plan.raw_begin:plan.raw_end], non_blocking=True)

    def step_dev():
        hp.develop_band_dev(params, W, H, d_raw.data_ptr(), pitch, d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), opitch, plan)

    def download():
        for c in range(3):
            h_out[c].copy_(d_out[c][own[0]:own[1], :Wo], non_blocking=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    upload()
    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    l0 = hp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms_per_step = max_ms(e0.elapsed_time(e1)) / args.steps
    launches = hp.launch_count() - l0
    # e2e: upload of the band's raw rows, kernels, download of the owned rows, every step, host wall clock
    for _ in range(2):
        upload(); step_dev(); download()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        upload(); step_dev(); download()
        torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_ms(wall) / args.steps
    clocks = sampler.stop() if sampler else None
    # check: the owned rows against this rank's own single-GPU development of the whole frame
    whole = [torch.zeros((Ho, opitch), dtype=torch.float32, device="cuda") for _ in range(3)]
    d_full = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
    d_full[:, :W] = torch.from_numpy(raw).cuda()
    hp.develop_dev(params, W, H, d_full.data_ptr(), pitch, whole[0].data_ptr(), whole[1].data_ptr(), whole[2].data_ptr(), opitch)
    torch.cuda.synchronize()
    worst = 0.0
    mag = torch.stack([w[own[0]:own[1], :Wo].abs() for w in whole]).amax(0)                # the pixel's largest channel (tests/test_fullsize_gpu.py)
    for c in range(3):
        a, b = d_out[c][own[0]:own[1], :Wo], whole[c][own[0]:own[1], :Wo]
        worst = max(worst, float(((a - b).abs() / (mag + 0.02)).max().item()))
    worst = max_ms(worst)
    # per-kernel device time of the band; every rank runs the pass (the all-reduces inside the step must match up), rank 0's is reported
    hp.profile_enable(True)
    for _ in range(args.steps):
        step_dev()
    prof = hp.profile_collect()
    hp.profile_enable(False)
    kern = {k: v[0] / args.steps for k, v in prof.items()}
    barrier()
    hp.comm_destroy()
    if rank == 0:
        peak, how = peaks()
        rows = plan.band_end - plan.band_begin
        step_bytes = 16 + 333 + 80 + 110
        step_achieved = step_bytes * W * H / (ms_per_step * 1e-3) / 1e9
        top = max(kern, key=lambda k: kern[k])
        config = dict(config)
        config["parallelism"] = ("row bands x%d of ONE frame: %d owned + up to %d halo rows per rank, demosaic on the frame's tile grid; collective = ncclAllReduce(int32, sum) "
                                 "of the wavelet subbands' MAD histograms (45 x 65536 counters per frame in two all-reduces: L, then a and b together), inside the library" % (world, own[1] - own[0], 2 * args.halo))
        out = {"metric": "Mpixel/s", "value": W * H / (ms_per_step * 1e-3) / 1e6, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config,
               "e2e": {"value": W * H / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": (plan.raw_end - plan.raw_begin) * W * 4,
                       "d2h_bytes_per_step": (own[1] - own[0]) * Wo * 12, "steps": args.steps, "host_memory": "pinned", "placement": placement,
                       "call": "per rank and step: upload of the band's raw rows, art_hp_develop_band_dev, download of the owned rows (float planes); bytes are rank 0's"},
               "gpu_launches": int(launches),
               "split": {"rank0_plan": repr(plan), "band_rows_rank0": rows, "owned_rows_rank0": own[1] - own[0],
                         "redundant_rows_fraction_rank0": 1.0 - (own[1] - own[0]) / rows,
                         "max_difference_from_single_gpu_frame_of_pixel_scale": worst,
                         "note": "difference measured on every rank's owned rows against its own art_hp_develop_dev of the whole frame, max over ranks; "
                                 "|a - b| / (largest channel of the pixel + 0.02); north_star's bound is 1e-4"},
               "roofline": {"bound": "hbm", "kernel": top, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": how,
                            "kernel_ms": kern[top], "note": "rank 0's band; see the single-GPU line of the same workload for per-kernel roofline fractions"},
               "step_roofline": {"achieved": step_achieved, "frac": step_achieved / peak, "unit": "GB/s", "bytes_per_pixel": step_bytes,
                                 "note": "whole frame over all ranks: SURVEY.md 8(d) ideal-fusion bytes per pixel / ms_per_step / n_gpus peaks", "n_gpus": world},
               "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])},
               "clocks": clocks}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = args.width, args.height
    wl = args.workload
    pipeline = wl in PIPELINES
    from art_b200 import synth
    filters = synth.RGGB
    config = {"workload": workload_text(wl, W, H), "frame": [W, H], "cfa": "X-Trans" if wl == "c3" else "RGGB", "frames_per_step_per_gpu": 1,
              "parallelism": "replicas x%d (independent frames, no collective)" % world,
              "l2": "per-step working set %.0f MB > 126 MB L2 (inputs larger than L2; no flush needed)" % (W * H * 16 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return 0
        use_all_host_threads()
        return reference_arm(args, config, W, H)

    # ---------------- our arm
    import numpy as np
    import torch
    import art_b200
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the hot path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    placement = pin_rank_near_gpu(local, world) if world > 1 else {"pinned": False, "why": "single rank"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator is created; stdout carries the ONE JSON line, so the
        # banner is sent to stderr: fd 1 points at fd 2 until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    hp = art_b200.HotPath(local)
    if args.split == "frame":
        return frame_split_arm(args, hp, dist, rank, world, local, W, H, wl, config, placement)
    method = art_b200.BAYER_RCD if args.method == "rcd" else art_b200.BAYER_AMAZE
    raw = make_raw(wl, W, H, 1002 + rank)

    # a real (non-default) stream: the library launches on it and the timing events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hp.set_stream(stream.cuda_stream)
    pitch = (W + 31) // 32 * 32
    d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
    d_raw[:, :W] = torch.from_numpy(raw).cuda()
    d_out = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]

    dparams = develop_params(art_b200, wl) if pipeline else None
    Ho, Wo = dparams.out_shape(H, W) if pipeline else (H, W)
    opitch = (Wo + 31) // 32 * 32

    def step_dev():
        if dparams is not None:
            hp.develop_dev(dparams, W, H, d_raw.data_ptr(), pitch, d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), opitch)
        else:
            hp.demosaic_bayer_dev(method, W, H, filters, d_raw.data_ptr(), pitch,
                                  d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), pitch, 1.0, 4)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    l0 = hp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = hp.launch_count() - l0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * W * H / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the host-buffer C-ABI call, pinned buffers, copies inside the timed region.  Pipelines go through the batch-queue form (what a
    #      batch of files calls): every step uploads its own CFA plane and downloads its own developed frame; the copies of frame k overlap the
    #      kernels of frames k-1 / k+1.  `packed`: the frame leaves as 16-bit interleaved scanlines (Imagefloat::getScanline on the device, the
    #      format the reference's writers take: 6 B/px over PCIe); `planes`: three float planes (12 B/px), reported beside it.
    def run_e2e(n, mode):
        if dparams is None:
            for _ in range(n):
                hp.demosaic_bayer(method, pins[0][0].array, filters, pins[0][1].array, pins[0][2].array, pins[0][3].array, 1.0, 4)
            return
        for k in range(n):
            if k >= 2:
                hp.develop_wait()
            sl = pins[k & 1]
            if mode == "packed":
                hp.develop_submit_packed(sl[0].array, dparams, packed[k & 1].array, 16, False)
            else:
                hp.develop_submit(sl[0].array, dparams, sl[1].array, sl[2].array, sl[3].array)
        while hp.develop_pending():
            hp.develop_wait()

    nslots = 2 if pipeline else 1
    pins = [[hp.pinned(H, W)] + [hp.pinned(Ho, Wo) for _ in range(3)] for _ in range(nslots)]
    packed = [hp.pinned(Ho, 3 * Wo, np.uint16) for _ in range(nslots)] if pipeline else None
    for sl in pins:
        sl[0].array[:] = raw
    e2e_steps = min(max(20, args.steps), 50)      # the pipeline fills and drains once (one upload + one download not hidden): amortised over >= 20 frames
    e2e = {}
    for mode in (("packed", "planes") if pipeline else ("planes",)):
        run_e2e(3, mode)
        barrier()
        t0 = time.perf_counter()
        run_e2e(e2e_steps, mode)               # returns when the last frame is in host memory
        wall = time.perf_counter() - t0
        barrier()
        ms2 = wall * 1e3
        if dist is not None:
            t = torch.tensor([ms2], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        e2e[mode] = world * W * H / (ms2 / e2e_steps * 1e-3) / 1e6
    # one synchronous call (upload, kernels, download back to back): the latency of a single frame
    p = pins[0]
    for rep in range(2):                   # the first call allocates the synchronous entry's own device planes
        t0 = time.perf_counter()
        if dparams is not None:
            hp.develop(p[0].array, dparams, p[1].array, p[2].array, p[3].array)
        else:
            hp.demosaic_bayer(method, p[0].array, filters, p[1].array, p[2].array, p[3].array, 1.0, 4)
        e2e_latency_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    checksum = float(p[2].array[Ho // 2, Wo // 2])

    # ---- per-kernel device time (CUDA events around every launch, on the launching stream), outside
    #      the timed regions above so the extra events do not perturb `value`
    kern, calls = {}, {}
    if rank == 0:
        hp.profile_enable(True)
        for _ in range(args.steps):
            step_dev()
        prof = hp.profile_collect()
        hp.profile_enable(False)
        kern = {k: v[0] / max(1, v[1]) for k, v in prof.items()}            # mean ms per launch
        calls = {k: v[1] // args.steps for k, v in prof.items()}            # launches per step

    if rank == 0:
        peak, how = peaks()
        geo = (W, H, Wo, Ho)
        # SURVEY.md 8(d) ideal-fusion bytes per pixel of the workload
        step_bytes = {"develop": 16 + 333 + 80 + 160, "c2": 16 + 333 + 80 + 110, "c3": 16 + 24 + 232 + 44, "c4": 16 + 333 + 80 + 160 + 24}.get(wl, BYTES_PER_PX)
        step_achieved = step_bytes * W * H / (ms_per_step * 1e-3) / 1e9
        per_step = {k: kern[k] * calls[k] for k in kern}                     # ms per step per kernel
        top = max((k for k in per_step if k != "memset_slabs"), key=lambda k: per_step[k])
        share = per_step[top] / sum(per_step.values())
        # per launch: bytes (or tensor flops) one launch moves / its mean duration
        rl_bound, achieved, rl_peak, rl_unit, rl_frac, rl_work, rl_what, rl_how = kernel_roofline(top, kern[top], geo, calls[top], peak)
        pipe_e2e = e2e.get("packed", e2e.get("planes"))
        out = {
            "metric": "Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": pipe_e2e, "unit": "Mpixel/s", "h2d_bytes_per_step": W * H * 4, "d2h_bytes_per_step": Wo * Ho * (6 if pipeline else 12),
                    "steps": e2e_steps, "host_memory": "pinned (art_hp_host_alloc)", "single_frame_latency_ms": e2e_latency_ms,
                    "float_planes_value": e2e.get("planes"), "float_planes_d2h_bytes_per_step": Wo * Ho * 12,
                    "placement": placement,
                    "call": ("art_hp_develop_submit_packed / art_hp_develop_wait: two frames in flight, every frame uploaded (float CFA plane) and "
                             "downloaded (16-bit interleaved scanlines, Imagefloat::getScanline on the device) inside the timed region (host wall clock "
                             "from the first submit to the last frame in host memory); float_planes_value = the same through art_hp_develop_submit "
                             "(three float planes out)")
                    if pipeline else "art_hp_demosaic_bayer (synchronous, banded copy/compute overlap inside the call)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": rl_bound, "kernel": top, "achieved": achieved, "peak": rl_peak, "unit": rl_unit,
                         "frac": rl_frac,
                         "traffic": NCU_TRAFFIC.get(top), "peak_source": rl_how or how,
                         "kernel_ms": kern[top], "launches_per_step": calls[top], "share_of_step": share,
                         ("algorithmic_flops_per_launch" if rl_bound == "tensor" else "algorithmic_bytes_per_launch"): rl_work, "what": rl_what,
                         "note": "dominant kernel by device time; achieved = algorithmic bytes per launch / mean CUDA-event "
                                 "duration of that kernel (events on the launching stream, separate pass of --steps steps)"},
            "step_roofline": {"achieved": step_achieved, "frac": step_achieved / peak, "unit": "GB/s",
                              "bytes_per_pixel": step_bytes,
                              "note": "whole step: SURVEY.md 8(d) ideal-fusion bytes per pixel / ms_per_step"},
            "roofline_top": roofline_top(per_step, kern, calls, geo, peak) if pipeline else None,
            "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
            "kernels_ms_sum": round(sum(per_step.values()), 4),
            "clocks": clocks, "checksum_green_center": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = time_cpu_pipeline(wl, args.cpu_sample, 1002) if pipeline else time_cpu(args.method, raw, filters)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                out["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
