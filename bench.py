#!/usr/bin/env python3
"""bench.py -- the hot path on synthetic 45 MP Bayer frames, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload develop|amaze|rcd]

A "step" is one pass of the hot path over one frame.  The default workload is the one BASELINE.json's
metric names -- "Mpixel/s end-to-end (demosaic+denoise+tonemap) on 45 MP Bayer": AMaZE demosaic, getImage
gains + camera->working matrix, RGB_denoise (luminance 30 / detail 50 / chrominance 15) and
Fattal tone mapping (threshold 30, amount 20) of an 8192x5464 synthetic RGGB frame, i.e. configs[1]'s frame
run through the stages of the metric (art_hp_develop).  `--workload amaze` is configs[1] alone (demosaic
only), `--workload rcd` the configs[0]-style case.  With N>1 (torchrun, one rank per GPU) every rank develops its own frame
-- frames are independent objects, so there is no data-path collective ("scaling": "weak"); the only
torch.distributed traffic is the barrier and the max-over-ranks of the timed region.

Printed JSON (rank 0, one line): the driver contract plus
  roofline     dominant kernel: algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline the reference's own code (oracle/_ref, kind "reference") or our C restatement
               (kind "port") timed on this box's host cores on a bounded sample
  e2e          the same metric through the host-buffer C-ABI call (art_hp_demosaic_bayer):
               pinned host planes in, pinned host planes out, both copies inside the timed region
  clocks       nvidia-smi samples taken during the timed region
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
               "k_fat_dct_rows": 184.989e6 + 316.294e6, "k_fat_dct_solve": 370.647e6 + 327.235e6, "k_fat_dct_exp": 370.152e6 + 156.166e6}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ART_BENCH_WORKLOAD", "develop"), choices=["develop", "amaze", "rcd"])
    ap.add_argument("--method", default=None, choices=["amaze", "rcd"], help="alias: --workload amaze|rcd")
    ap.add_argument("--cpu-sample", default="2048x1366", help="frame of the bounded CPU sample of the develop workload")
    ap.add_argument("--width", type=int, default=W45)
    ap.add_argument("--height", type=int, default=H45)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.method:
        a.workload = a.method
    a.method = "rcd" if a.workload == "rcd" else "amaze"
    return a


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
CAM2WORK = [[0.82, 0.15, 0.03], [0.07, 0.96, -0.03], [0.02, -0.10, 1.08]]      # a fixed camera->working matrix (synthetic camera)
MUL = (1.9, 1.0, 1.6)                                                           # rm, gm, bm of getImage for that camera
DN = dict(luminance=30.0, luminanceDetail=50.0, luminanceDetailThreshold=0, chrominance=15.0, chrominanceRedGreen=0.0,
          chrominanceBlueYellow=0.0, gamma=1.7, scale=1.0)
FATTAL = (30, 20, 0)                                                            # threshold, amount, satcontrol (procparams.cc L2074-2079)

# Algorithmic bytes per unit of the develop kernels that can dominate: fp32 planes, each plane a
# kernel must read or write counted once.  unit = what one launch covers.
DEVELOP_KERNEL_BYTES = {
    "k_fbox_h": (8, "subband coefficient"), "k_fbox_v": (8, "subband coefficient"),        # 4 R + 4 W
    "k_dn_blocks": (4 + 4 * (64.0 / 25.0) ** 2, "pixel"),   # residual read once + the windowed 64x64 blocks (stride 25) written
    "k_fat_dct_rows": (4 + 8, "padded pixel"), "k_fat_dct_solve": (8 + 8, "padded pixel"), "k_fat_dct_exp": (8 + 4, "padded pixel"),
    "k_wav_sy_sub": (16 + 4, "pixel"), "k_sf_apply": (12 + 4, "subband coefficient"), "k_mad_hist": (4, "subband coefficient"),
    "k_sf_L": (8 + 4, "subband coefficient"), "k_sf_AB": (12 + 4, "subband coefficient"),
}


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


def develop_units(name, W, H):
    if name in ("k_fbox_h", "k_fbox_v", "k_sf_apply", "k_mad_hist", "k_sf_L", "k_sf_AB"):
        return ((W + 1) // 2) * ((H + 1) // 2)
    if name.startswith("k_fat_dct"):
        return (find_fast_dim(W) + 1) * (find_fast_dim(H) + 1)
    return W * H


def roofline_top(per_step, kern, calls, W, H, peak, n=6):
    """The same algorithmic-bytes / CUDA-event-time figure for the n largest kernels of the step (the wavelet-shrink kernels of
    the three directions run on three concurrent streams, so their per-launch times include each other's interference)."""
    out = []
    ntiles = ((W + 16 + 127) // 128) * ((H + 16 + 127) // 128)
    for k in sorted(per_step, key=lambda k: -per_step[k])[:n]:
        ub = DEVELOP_KERNEL_BYTES.get(k)
        if ub is not None:
            units = develop_units(k, W, H)
            b = ub[0]
        elif k in AMAZE_KERNEL_BYTES:
            units, b = ntiles * 160 * 160, AMAZE_KERNEL_BYTES[k]
        else:
            out.append({"kernel": k, "ms_per_step": round(per_step[k], 4), "launches_per_step": calls[k], "achieved_GBps": None, "frac": None})
            continue
        ach = b * units / (kern[k] * 1e-3) / 1e9
        out.append({"kernel": k, "ms_per_step": round(per_step[k], 4), "launches_per_step": calls[k], "achieved_GBps": round(ach, 1),
                    "frac": round(ach / peak, 4)})
    return out


def develop_params(art_b200):
    from art_b200.api import DenoiseParams, DevelopParams
    return DevelopParams(method=art_b200.BAYER_AMAZE, filters=0x94949494, initial_gain=1.0, border=4, mul=MUL, do_clip=True,
                         cam2work=CAM2WORK, denoise=DenoiseParams(**DN), fattal=FATTAL, wprof=PROPHOTO)


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm is the reference's OpenMP code on ALL host
    threads of the box, so the count is set explicitly (environment for a libgomp not loaded yet, omp_set_num_threads for one that is)."""
    import ctypes
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def cpu_develop_runner(raw, filters):
    """The same stages through the reference's own functions compiled in place (oracle/_ref, stock build): AMaZE,
    getImage gains + matrix, RGB_denoise, ToneMapFattal02.  fftw3f is absent from this image, so the two FFTW call sites
    (64x64 block DCTs, 2-D REDFT00) run the oracle's double-precision stand-in, which is slower than FFTW would be."""
    import ctypes
    import numpy as np
    import oracle
    ncores = use_all_host_threads()
    ref = oracle.ref(det=False)
    use_all_host_threads()
    lib = ref.lib
    lib.artref_set_denoise_thread_limit(0)        # the reference's default: all OpenMP threads
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    Hr, Wr = raw.shape
    H, W = Hr - 8, Wr - 8                         # getImage crops the demosaiced frame by RawImageSource::border (4)
    out = [np.zeros((Hr, Wr), np.float32) for _ in range(3)]
    wp = np.array(PROPHOTO, np.float64)
    wpi = np.linalg.inv(wp)
    p = np.array([DN["luminance"], DN["luminanceDetail"], DN["luminanceDetailThreshold"], DN["chrominance"], DN["chrominanceRedGreen"],
                  DN["chrominanceBlueYellow"], DN["gamma"], DN["scale"]], np.float64)
    res = np.zeros(2, np.float32)

    def run():
        ref.amaze(raw, filters, nthreads=ncores, out=out)
        r, g, b = ref.scale_convert([np.ascontiguousarray(p[4:-4, 4:-4]) for p in out], MUL, True, np.array(CAM2WORK, np.float64))
        assert lib.artref_rgb_denoise(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, p.ctypes.data_as(dp),
                                      wp.ctypes.data_as(dp), wpi.ctypes.data_as(dp), None, ctypes.c_float(0), None, None, None,
                                      res.ctypes.data_as(fp)) == 0
        assert lib.artref_fattal(r.ctypes.data_as(fp), g.ctypes.data_as(fp), b.ctypes.data_as(fp), W, H, FATTAL[0], FATTAL[1], FATTAL[2],
                                 wp.ctypes.data_as(dp)) == 0
    return run, "reference", ncores


def time_cpu_develop(sample, filters, budget_s=20.0, max_runs=3):
    from art_b200 import synth
    w, h = [int(v) for v in sample.lower().split("x")]
    raw = synth.bayer_frame(w, h, filters, seed=1002)
    run, kind, cores = cpu_develop_runner(raw, filters)
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < max_runs and (not ts or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": w * h / med / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d runs of the same four stages on a %dx%d frame (%.1f MP, 1/%d of the workload's pixels), median; reference "
                      "functions compiled in place, OpenMP on %d threads; FFTW call sites run the oracle's fp64 stand-in (fftw3f absent)"
                      % (len(ts), w, h, w * h / 1e6, round(W45 * H45 / (w * h)), cores)}


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


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = args.width, args.height
    from art_b200 import synth
    filters = synth.RGGB
    workload = "configs[1]: AMaZE demosaic, %dx%d synthetic RGGB (%.2f MP)" % (W, H, W * H / 1e6)
    if args.workload == "develop":
        workload = ("metric pipeline on configs[1]'s frame: AMaZE demosaic + gains/matrix + RGB_denoise (lum 30, detail 50, chroma 15) + "
                    "Fattal (30/20), %dx%d synthetic RGGB (%.2f MP), art_hp_develop" % (W, H, W * H / 1e6))
    if args.workload == "rcd":
        workload = "configs[0]-style: RCD demosaic, %dx%d synthetic RGGB (%.2f MP)" % (W, H, W * H / 1e6)
    config = {"workload": workload, "frame": [W, H], "cfa": "RGGB", "frames_per_step_per_gpu": 1,
              "parallelism": "replicas x%d (independent frames, no collective)" % world,
              "l2": "per-step working set %.0f MB > 126 MB L2 (inputs larger than L2; no flush needed)" % (W * H * 16 / 1e6)}

    # ---------------- reference arm: the reference's own CPU code on this box's host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        use_all_host_threads()
        if args.workload == "develop":
            # bounded sample: the same four stages on a smaller frame of the same synthetic scene, at most a few steps
            sw, sh = [int(v) for v in args.cpu_sample.lower().split("x")]
            raw = synth.bayer_frame(sw, sh, filters, seed=1002)
            run, kind, cores = cpu_develop_runner(raw, filters)
            nsteps = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            run()
            first = time.perf_counter() - t0
            nwarm = 1 if first < 20 else 0
            t0 = time.perf_counter()
            for _ in range(nsteps):
                run()
            dt = time.perf_counter() - t0
            val = nsteps * sw * sh / dt / 1e6
            cb = {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": kind,
                  "sample": "%d steps (+%d warm-up) of the four stages on a %dx%d frame (1/%d of the workload's pixels; per-pixel rate "
                            "reported); reference functions compiled in place, OpenMP on %d threads; FFTW call sites run the oracle's "
                            "fp64 stand-in (fftw3f absent)" % (nsteps, nwarm + 0, sw, sh, round(W * H / (sw * sh)), cores)}
            print(json.dumps({"impl": "reference", "metric": "Mpixel/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus,
                              "steps": nsteps, "warmup": 1, "ms_per_step": dt / nsteps * 1e3,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": config, "cpu_baseline": cb,
                              "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        raw = synth.bayer_frame(W, H, filters, seed=1002)
        run, kind, cores = cpu_reference_runner(args.method, raw, filters)
        for _ in range(max(1, min(args.warmup, 2))):
            run()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run()
        dt = time.perf_counter() - t0
        val = args.steps * W * H / dt / 1e6
        cb = {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": kind,
              "sample": "%d full %dx%d frames (this run's steps); OpenMP threads = %d (faster of all/half)" % (args.steps, W, H, cores)}
        print(json.dumps({"impl": "reference", "metric": "Mpixel/s", "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---------------- our arm
    import torch
    import art_b200
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the hot path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
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
    method = art_b200.BAYER_RCD if args.method == "rcd" else art_b200.BAYER_AMAZE
    raw = synth.bayer_frame(W, H, filters, seed=1002 + rank)

    # a real (non-default) stream: the library launches on it and the timing events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    hp.set_stream(stream.cuda_stream)
    pitch = (W + 31) // 32 * 32
    d_raw = torch.zeros((H, pitch), dtype=torch.float32, device="cuda")
    d_raw[:, :W] = torch.from_numpy(raw).cuda()
    d_out = [torch.empty((H, pitch), dtype=torch.float32, device="cuda") for _ in range(3)]

    dparams = develop_params(art_b200) if args.workload == "develop" else None

    def step_dev():
        if dparams is not None:
            hp.develop_dev(dparams, W, H, d_raw.data_ptr(), pitch, d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), pitch)
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

    # ---- e2e: host-buffer C-ABI call, pinned planes, copies inside the timed region.  The develop workload goes through the
    #      batch-queue form (art_hp_develop_submit / _wait, what a batch of files calls): every step uploads its own CFA plane
    #      and downloads its own three planes; the copies of frame k overlap the kernels of frames k-1 / k+1.
    nslots = 2 if dparams is not None else 1
    Ho, Wo = dparams.out_shape(H, W) if dparams is not None else (H, W)
    pins = [[hp.pinned(H, W)] + [hp.pinned(Ho, Wo) for _ in range(3)] for _ in range(nslots)]
    for sl in pins:
        sl[0].array[:] = raw

    def run_e2e(n):
        if dparams is None:
            p = pins[0]
            for _ in range(n):
                hp.demosaic_bayer(method, p[0].array, filters, p[1].array, p[2].array, p[3].array, 1.0, 4)
            return
        for k in range(n):
            if k >= 2:
                hp.develop_wait()
            p = pins[k & 1]
            hp.develop_submit(p[0].array, dparams, p[1].array, p[2].array, p[3].array)
        while hp.develop_pending():
            hp.develop_wait()

    run_e2e(3)
    barrier()
    e2e_steps = min(max(20, args.steps), 50)      # the pipeline fills and drains once (one upload + one download not hidden): amortised over >= 20 frames
    t0 = time.perf_counter()
    run_e2e(e2e_steps)                     # returns when the last frame's planes are in host memory
    wall = time.perf_counter() - t0
    barrier()
    ms2 = wall * 1e3
    if dist is not None:
        t = torch.tensor([ms2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e_val = world * W * H / (ms2 / e2e_steps * 1e-3) / 1e6
    # one synchronous call (upload, kernels, download back to back): the latency of a single frame
    p = pins[0]
    for rep in range(2):                   # the first call allocates the synchronous entry's own device planes
        t0 = time.perf_counter()
        if dparams is not None:
            hp.develop(p[0].array, dparams, p[1].array, p[2].array, p[3].array)
        else:
            hp.demosaic_bayer(method, p[0].array, filters, p[1].array, p[2].array, p[3].array, 1.0, 4)
        e2e_latency_ms = (time.perf_counter() - t0) * 1e3
    pins = pins[0]
    clocks = sampler.stop() if sampler else None
    checksum = float(pins[2].array[Ho // 2, Wo // 2])

    # ---- per-kernel device time (CUDA events around every launch, on the launching stream), outside
    #      the timed regions above so the extra events do not perturb `value`
    kern = {}
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
        # SURVEY.md 8(d) ideal-fusion bytes per pixel: demosaic 16 (+ wavelet denoise 333 + denoise I/O and DCT 80 + Fattal 160)
        step_bytes = 16 + 333 + 80 + 160 if args.workload == "develop" else BYTES_PER_PX
        step_achieved = step_bytes * W * H / (ms_per_step * 1e-3) / 1e9
        per_step = {k: kern[k] * calls[k] for k in kern}                     # ms per step per kernel
        top = max((k for k in per_step if k != "memset_slabs"), key=lambda k: per_step[k])
        share = per_step[top] / sum(per_step.values())
        if args.workload == "develop":
            unit_bytes, unit_name = DEVELOP_KERNEL_BYTES.get(top, (AMAZE_KERNEL_BYTES.get(top), "tile pixel"))
            ntiles = ((W + 16 + 127) // 128) * ((H + 16 + 127) // 128)
            units = ntiles * 160 * 160 if unit_name == "tile pixel" else develop_units(top, W, H)
        elif args.workload == "rcd":
            units, unit_bytes, unit_name = W * H, BYTES_PER_PX, "pixel"
        else:
            ntiles = ((W + 16 + 127) // 128) * ((H + 16 + 127) // 128)
            units, unit_bytes, unit_name = ntiles * 160 * 160, AMAZE_KERNEL_BYTES.get(top), "tile pixel"
        if unit_bytes is None:
            achieved = None
        else:
            achieved = unit_bytes * units / (kern[top] * 1e-3) / 1e9      # per launch: bytes one launch moves / its mean duration
        out = {
            "metric": "Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_val, "unit": "Mpixel/s", "h2d_bytes_per_step": W * H * 4, "d2h_bytes_per_step": Wo * Ho * 12,
                    "steps": e2e_steps, "host_memory": "pinned (art_hp_host_alloc)", "single_frame_latency_ms": e2e_latency_ms,
                    "call": ("art_hp_develop_submit / art_hp_develop_wait: two frames in flight, every frame uploaded and downloaded "
                             "inside the timed region (host wall clock from the first submit to the last frame in host memory)")
                    if dparams is not None else "art_hp_demosaic_bayer (synchronous, banded copy/compute overlap inside the call)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": NCU_TRAFFIC.get(top), "peak_source": how,
                         "kernel_ms": kern[top], "launches_per_step": calls[top], "share_of_step": share,
                         "algorithmic_bytes_per_unit": unit_bytes, "unit_name": unit_name, "units_per_step": units,
                         "note": "dominant kernel by device time; achieved = algorithmic bytes per launch / mean CUDA-event "
                                 "duration of that kernel (events on the launching stream, separate pass of --steps steps)"},
            "step_roofline": {"achieved": step_achieved, "frac": step_achieved / peak, "unit": "GB/s",
                              "bytes_per_pixel": step_bytes,
                              "note": "whole step: SURVEY.md 8(d) ideal-fusion bytes per pixel / ms_per_step"},
            "roofline_top": roofline_top(per_step, kern, calls, W, H, peak) if args.workload == "develop" else None,
            "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
            "clocks": clocks, "checksum_green_center": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = time_cpu_develop(args.cpu_sample, filters) if args.workload == "develop" else time_cpu(args.method, raw, filters)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                out["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
