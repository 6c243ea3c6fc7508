#!/usr/bin/env python3
"""bench.py -- the hot path on synthetic 45 MP Bayer frames, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--method amaze|rcd]

A "step" is one pass of the hot path over one frame.  At N=1 the workload is BASELINE.json
configs[1]: AMaZE demosaic of an 8192x5464 synthetic RGGB frame (the configuration the metric is
quoted on that fits one GPU).  With N>1 (torchrun, one rank per GPU) every rank develops its own frame
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default=os.environ.get("ART_BENCH_METHOD", "amaze"), choices=["amaze", "rcd"])
    ap.add_argument("--width", type=int, default=W45)
    ap.add_argument("--height", type=int, default=H45)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


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


def cpu_reference_runner(method):
    """(callable(raw, filters) -> planes, kind, cores).  Prefers the reference's own bodies."""
    import oracle
    try:
        ref = oracle.ref(det=False)      # the stock reference, exactly as it ships
        fn = ref.rcd if method == "rcd" else ref.amaze
        return (lambda raw, f: fn(raw, f)), "reference", ref.max_threads()
    except Exception:
        port = oracle.port()
        fn = port.rcd if method == "rcd" else port.amaze
        return (lambda raw, f: fn(raw, f)), "port", os.cpu_count() or 1


def time_cpu(method, raw, filters, budget_s=12.0, max_runs=5):
    run, kind, cores = cpu_reference_runner(method)
    run(raw, filters)                                  # warm-up (page faults, OpenMP pool)
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < max_runs and (not ts or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        run(raw, filters)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    H, W = raw.shape
    return {"value": W * H / med / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d full %dx%d frames after 1 warm-up, median" % (len(ts), W, H)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = args.width, args.height
    from art_b200 import synth
    filters = synth.RGGB
    workload = "configs[1]: %s demosaic, %dx%d synthetic RGGB (%.2f MP)" % (args.method.upper() if args.method == "rcd" else "AMaZE", W, H, W * H / 1e6)
    if args.method == "rcd":
        workload = "configs[0]-style: RCD demosaic, %dx%d synthetic RGGB (%.2f MP)" % (W, H, W * H / 1e6)
    config = {"workload": workload, "frame": [W, H], "cfa": "RGGB", "frames_per_step_per_gpu": 1,
              "parallelism": "replicas x%d (independent frames, no collective)" % world,
              "l2": "per-step working set %.0f MB > 126 MB L2 (inputs larger than L2; no flush needed)" % (W * H * 16 / 1e6)}

    # ---------------- reference arm: the reference's own CPU code on this box's host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        raw = synth.bayer_frame(W, H, filters, seed=1002)
        run, kind, cores = cpu_reference_runner(args.method)
        for _ in range(max(1, min(args.warmup, 2))):
            run(raw, filters)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(raw, filters)
        dt = time.perf_counter() - t0
        val = args.steps * W * H / dt / 1e6
        cb = {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": kind,
              "sample": "%d full %dx%d frames (this run's steps)" % (args.steps, W, H)}
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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

    def step_dev():
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

    # ---- e2e: host-buffer C-ABI call, pinned planes, copies inside the timed region
    pins = [hp.pinned(H, W) for _ in range(4)]
    pins[0].array[:] = raw

    def step_e2e():
        hp.demosaic_bayer(method, pins[0].array, filters, pins[1].array, pins[2].array, pins[3].array, 1.0, 4)

    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    ms2 = max(e0.elapsed_time(e1), wall * 1e3)
    if dist is not None:
        t = torch.tensor([ms2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e_val = world * W * H / (ms2 / e2e_steps * 1e-3) / 1e6
    clocks = sampler.stop() if sampler else None
    checksum = float(pins[2].array[H // 2, W // 2])

    if rank == 0:
        peak, how = peaks()
        achieved = BYTES_PER_PX * W * H / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": "Mpixel/s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_val, "unit": "Mpixel/s", "h2d_bytes_per_step": W * H * 4, "d2h_bytes_per_step": W * H * 12,
                    "steps": e2e_steps, "host_memory": "pinned (art_hp_host_alloc)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": how,
                         "note": "algorithmic bytes (16 B/px) / mean device time of one step (all kernels of the step)"},
            "clocks": clocks, "checksum_green_center": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = time_cpu(args.method, raw, filters)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                out["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
