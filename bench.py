#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): temporal-median background over a uint8 frame stack.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                           # the reference's CPU path

One "step" = one pass of the hot path over one batch = the median of the whole synthetic
1920x1080 x 1000-frame stack (BASELINE.json configs[1]; SURVEY.md 8d seeds).  `value` is
megapixel-frames/s with the stack already resident in HBM (CUDA events on the launching stream,
max over ranks); `e2e` is the same metric through the C ABI's host-buffer interface
(cvvp_median_begin/push/finish) with the frames in pinned host memory, H2D and D2H inside the
timed region.  The stack (2.07 GB per GPU) is ~16x larger than L2, so no L2 flush is needed
between timed iterations.

N > 1 (one process per GPU under torchrun), default `--median-sharding frames`: weak scaling, every
GPU holds its own 1000-frame chunk and the job is the median of the 1000*N-frame stack, merged by
the two-round nibble-count exchange of csrc/median_shard.cu (counts are stored into the owner
rank's memory over NVLink by the counting kernels; see run_gpu_arm_sharded).  `--median-sharding
rows` instead splits the single C2 stack into row bands (shard_plan(): no collective on the data
path, one NCCL all_gather of the result bands; strong scaling).  The highlight section shards by
frame with no collective in both cases.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

WORKLOAD = dict(name="C2: 1920x1080 uint8 x 1000 frames, temporal median (BASELINE.json configs[1])",
                width=1920, height=1080, nframes=1000, seed=2, ndisks=30)
METRIC = "megapixel-frames/sec (temporal-median background, 1080p x 1000-frame stack)"
UNIT = "Mpx-frames/s"
FALLBACK_HBM_GBS = 6650.0


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_traffic(name: str = "median_ncu_summary.json"):
    """dram bytes per launch of a kernel from its committed ncu --set full capture (profiles/)."""
    p = REPO / "profiles" / name
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
        except Exception:
            pass
    return None


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while a timed region is active."""

    def __init__(self, device_index: int):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.active = False
        self._stop = False
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        self.h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        try:
                            self.h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                            break
                        except Exception:
                            continue
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    _REASONS = {
        "hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
        "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2,
        "display_clock_setting": 0x100,
    }

    def _loop(self):
        nv = self.nv
        while not self._stop:
            if self.active:
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append(mhz)
                    for name, bit in self._REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread:
            self._thread.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_near_gpu(device_index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the pinned host buffers it allocates
    afterwards come from that NUMA node (with 8 ranks streaming 55 GB/s each, remote pages halve the H2D rate).
    Returns the number of CPUs in the set (0: left unchanged)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def shard_plan(height: int, rank: int, world: int):
    """Row-band partition (the reference's own spatial sharding, cv_vid_frames_generator_algo.h:159-164, with
    horizontal bands instead of vertical strips so that every band is contiguous in memory).  The median is
    element-wise, so bands are independent: no collective on the data path."""
    base = height // world
    row0 = base * rank
    nrows = base if rank < world - 1 else height - row0
    return row0, nrows


# ------------------------------------------------------------------------------------------------
# CPU arm (reference implementation timed on host cores)
# ------------------------------------------------------------------------------------------------
def cpu_median_fn():
    """oracle/_ref (the reference's own class, kind='reference') when it was built, else the C port."""
    ref = REPO / "oracle" / "_ref" / "libcvvp_median_ref.so"
    port = REPO / "oracle" / "_build" / "libcvvp_oracle.so"
    if ref.exists():
        lib, name, kind = ctypes.CDLL(str(ref)), "cvvp_ref_median", "reference"
    elif port.exists():
        lib, name, kind = ctypes.CDLL(str(port)), "cvvp_oracle_median", "port"
    else:
        raise RuntimeError("neither oracle/_ref nor oracle/_build is built; run __graft_entry__.build()")
    fn = getattr(lib, name)
    fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                   ctypes.c_void_p]
    fn.restype = ctypes.c_int

    def run(frames: np.ndarray, nthreads: int) -> np.ndarray:
        n = frames.shape[0]
        nelem = int(np.prod(frames.shape[1:]))
        out = np.empty(nelem, np.uint8)
        rc = fn(frames.ctypes.data, n, nelem, frames.strides[0], 0, nthreads, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"{name} failed: {rc}")
        return out

    return run, kind


def host_sample_frames(rows: int, nframes: int | None = None) -> np.ndarray:
    """Rows [0, rows) of every frame of the workload, generated on the host (cvvidproc_b200/synth.py)."""
    from cvvidproc_b200 import synth

    w = WORKLOAD
    return synth.synth_frames(0, nframes or w["nframes"], w["width"], w["height"], w["seed"], w["ndisks"], row0=0, nrows=rows)


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    w = dict(WORKLOAD)
    if world > 1 and args.median_sharding == "frames":
        w["nframes"] = WORKLOAD["nframes"] * world  # the cuda arm's job at N GPUs: the 1000*N-frame stack
    cores = os.cpu_count() or 1
    run, kind = cpu_median_fn()
    # bounded sample: a band of rows of the SAME stack (all frames), sized from a calibration band so that
    # the whole run stays within ~2.5 minutes
    calib_rows = 8
    frames = host_sample_frames(calib_rows, w["nframes"])
    t0 = time.perf_counter()
    run(frames, cores)
    t_cal = max(time.perf_counter() - t0, 1e-4)
    per_row = t_cal / calib_rows
    budget = 150.0 / max(1, args.steps + args.warmup)
    rows = int(min(w["height"], max(calib_rows, budget / per_row)))
    rows = min(rows, max(calib_rows, 270 * WORKLOAD["nframes"] // w["nframes"]))  # host generation: ~0.1 s per row per 1000 frames
    frames = host_sample_frames(rows, w["nframes"])
    for _ in range(args.warmup):
        run(frames, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(frames, cores)
    dt = (time.perf_counter() - t0) / args.steps
    mpxf = rows * w["width"] * w["nframes"] / 1e6
    value = mpxf / dt
    sample = f"rows [0,{rows}) of all {w['nframes']} frames ({mpxf:.1f} Mpx-frames per step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if (world > 1 and args.median_sharding == "rows") else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "width": w["width"], "height": w["height"], "nframes": w["nframes"],
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_highlight:
        try:  # the reference's CPU highlight path on a small host-generated sample of the C3 stream
            from cvvidproc_b200 import synth

            hw = HL_WORKLOAD
            st = synth.synth_frames(0, 9, hw["width"], hw["height"], hw["seed"], hw["ndisks"])
            bg = np.sort(st, axis=0)[4]
            fr = synth.synth_frames(1000, 8, hw["width"], hw["height"], hw["seed"], hw["ndisks"])
            rate, done, _ = cpu_highlight_rate(fr, bg, cores, seconds=10.0)
            line["highlight"] = {"metric": "megapixel-frames/sec (per-frame highlight, 1080p)", "unit": UNIT, "value": rate,
                                 "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                                  "sample": f"{done} frames, cv2 restatement, background = median of 9 frames"}}
        except Exception as exc:
            line["highlight"] = {"value": None, "error": str(exc)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# highlight stage (second half of the BASELINE metric): reported as an extra object on the same line
# ------------------------------------------------------------------------------------------------
HL_WORKLOAD = dict(name="C3: 1920x1080 uint8 frames, per-frame highlight (BASELINE.json configs[2]), background = "
                        "device median of the stream's first 255 frames, canonical parameters",
                   width=1920, height=1080, seed=3, ndisks=30, frames_per_step=1024)


def cpu_highlight_rate(frames: np.ndarray, bg: np.ndarray, threads: int, seconds: float = 12.0):
    """cv2 restatement of highlight_objects_algo.cpp (oracle/highlight_oracle.py), one frame per task, cv2 internal
    threading off, `threads` Python threads (cv2 releases the GIL) -- the reference's frame-level data parallelism
    (cv_vid_objecttrack_helpers.cpp:72-84).  Returns (Mpx-frames/s, frames done, last mask)."""
    import concurrent.futures as cf

    import cv2

    from oracle import highlight_oracle as ho

    cv2.setNumThreads(1)
    p = ho.canonical_params(bg)
    n = frames.shape[0]
    done = 0
    last = None
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        while time.perf_counter() - t0 < seconds:
            res = list(ex.map(lambda i: ho.highlight_objects(frames[i % n].copy(), p), range(done, done + 2 * threads)))
            done += len(res)
            last = res[-1]
    dt = time.perf_counter() - t0
    return frames.shape[1] * frames.shape[2] * done / 1e6 / dt, done, last


def run_highlight_section(ctx, torch, dist, rank, local_rank, world, args, sampler, stream):
    """Frames are sharded over ranks by frame (no collective on the data path: frames are independent,
    highlight_objects_algo.h:82-85); each rank processes frames_per_step frames per step."""
    from cvvidproc_b200 import synth

    w = HL_WORKLOAD
    W, H, nfr = w["width"], w["height"], w["frames_per_step"]
    npix = W * H
    dev = f"cuda:{local_rank}"
    bgstack = torch.empty((255, npix), dtype=torch.uint8, device=dev)
    bg = torch.empty(npix, dtype=torch.uint8, device=dev)
    ctx.synth_frames_device(bgstack.data_ptr(), npix, W, H, 0, 255, w["seed"], w["ndisks"])
    ctx.median_device(bgstack.data_ptr(), 255, npix, npix, bg.data_ptr())
    ctx.synchronize()
    del bgstack
    bg_h = bg.cpu().numpy().reshape(H, W)
    cp = synth.CANONICAL_HIGHLIGHT  # the workload's parameters (the oracle is only used by the cpu_baseline leg)
    frames = torch.empty((nfr, npix), dtype=torch.uint8, device=dev)
    masks = torch.empty((nfr, npix), dtype=torch.uint8, device=dev)
    first = 1000 + rank * nfr  # this rank's frames of the stream
    ctx.synth_frames_device(frames.data_ptr(), npix, W, H, first, nfr, w["seed"], w["ndisks"])
    ctx.highlight_begin(bg_h, synth.canonical_struct_element(), cp["threshold"], cp["threshold_lo"], cp["threshold_hi"],
                        cp["min_size_hyst"], cp["min_size_threshold"], cp["width_border"])
    steps = max(3, min(args.steps, 10))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.highlight_device(frames.data_ptr(), nfr, npix, masks.data_ptr(), npix)
        barrier()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = True
        e0.record(stream)
        for _ in range(steps):
            ctx.highlight_device(frames.data_ptr(), nfr, npix, masks.data_ptr(), npix)
        e1.record(stream)
        barrier()
        sampler.active = False
        launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1) / steps
    # end to end from pinned host memory through cvvp_highlight_frames
    from cvvidproc_b200 import _cabi

    pin_in = _cabi.PinnedBuffer(nfr * npix)
    pin_out = _cabi.PinnedBuffer(nfr * npix)
    pin_in.array[:] = frames.cpu().numpy().reshape(-1)
    lib = _cabi.load()

    def e2e_step():
        rc = lib.cvvp_highlight_frames(ctx.handle, pin_in.array.ctypes.data, nfr, npix, pin_out.array.ctypes.data, npix)
        if rc != 0:
            raise RuntimeError(lib.cvvp_last_error(ctx.handle).decode())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    total_mpx = world * nfr * npix / 1e6
    same = bool(np.array_equal(pin_out.array.reshape(nfr, npix), masks.cpu().numpy()))
    out = {
        "metric": "megapixel-frames/sec (per-frame highlight, 1080p)", "unit": UNIT,
        "value": total_mpx / (ms * 1e-3), "ms_per_step": ms, "frames_per_step": world * nfr, "steps": steps,
        "gpu_launches": int(launches), "scaling": "weak", "sharding": "by frame, no collective",
        "e2e": {"value": total_mpx / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(nfr * npix) * world,
                "d2h_bytes_per_step": int(nfr * npix) * world, "ms_per_step": e2e_s * 1e3},
        "roofline": {"bound": "hbm", "achieved": 2.0 * nfr * npix / (ms * 1e-3) / 1e9, "peak": load_peaks()[0],
                     "unit": "GB/s", "frac": 2.0 * nfr * npix / (ms * 1e-3) / 1e9 / load_peaks()[0],
                     "traffic": load_traffic("highlight_ncu_summary.json") if (world == 1 and nfr == 1024) else None,
                     "kernel": "highlight_fused_kernel",
                     "note": "algorithmic bytes = frame in + mask out (2 B/px); one fused kernel launch per step"},
        "config": {"workload": w["name"]},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            host_frames = frames[:16].cpu().numpy().reshape(16, H, W)
            cores = os.cpu_count() or 1
            rate, done, last = cpu_highlight_rate(host_frames, bg_h, cores)
            idx = (done - 1) % 16
            same = same and bool(np.array_equal(last.reshape(-1), masks[idx].cpu().numpy()))
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{done} frames in ~12 s, cv2 4.x restatement of highlight_objects_algo.cpp, "
                                             f"{cores} threads x cv2.setNumThreads(1)"}
        except Exception as exc:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "unavailable",
                                   "sample": f"failed: {exc}"}
    out["parity_spot_check"] = same
    ctx.highlight_end()
    pin_in.close()
    pin_out.close()
    return out


def run_frame_source_section(ctx, torch, local_rank, args, stream):
    """The stage in front of both operators: decoded 1080p 3-channel frames resident in HBM ->
    grey frames (crop = whole frame, COLOR_RGB2GRAY; cv_vid_frames_generator_algo.h:141-156) by csrc/frames.cu.
    Algorithmic bytes = 3 B/px read + 1 B/px written.  CPU baseline = cv2.cvtColor on all host threads."""
    from cvvidproc_b200 import _cabi

    W, H, n = 1920, 1080, 96  # 597 MB of decoded frames: larger than the 126 MB L2
    dev = f"cuda:{local_rank}"
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    src = torch.randint(0, 256, (n, H * W * 3), dtype=torch.uint8, device=dev, generator=gen)
    dst = torch.empty((n, H * W), dtype=torch.uint8, device=dev)
    fmt = _cabi.FrameFormat.of((H, W, 3), _cabi.FRAMES_RGB2GRAY)
    steps = max(3, min(args.steps, 10))
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.frames_prepare_device(src.data_ptr(), n, H * W * 3, fmt, dst.data_ptr(), H * W)
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            ctx.frames_prepare_device(src.data_ptr(), n, H * W * 3, fmt, dst.data_ptr(), H * W)
        e1.record(stream)
        torch.cuda.synchronize()
        launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1) / steps
    mpx = n * W * H / 1e6
    peak = load_peaks()[0]
    achieved = 4.0 * n * W * H / (ms * 1e-3) / 1e9
    out = {
        "metric": "megapixel-frames/sec (frame source: 1080p decoded 3-channel -> grey)", "unit": UNIT,
        "value": mpx / (ms * 1e-3), "ms_per_step": ms, "frames_per_step": n, "steps": steps, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic("frames_ncu_summary.json"), "kernel": "frames_prepare_kernel",
                     "note": "algorithmic bytes = 3 B/px decoded frame in + 1 B/px prepared frame out"},
    }
    if not args.no_cpu_baseline:
        import concurrent.futures as cf

        import cv2

        cv2.setNumThreads(1)
        host = src[:16].cpu().numpy().reshape(16, H, W, 3)
        cores = os.cpu_count() or 1
        done, last = 0, None
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(max_workers=cores) as ex:
            while time.perf_counter() - t0 < 3.0:
                res = list(ex.map(lambda i: cv2.cvtColor(host[i % 16], cv2.COLOR_RGB2GRAY), range(done, done + 4 * cores)))
                done += len(res)
                last = res[-1]
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": done * W * H / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{done} frames in ~3 s, cv2.cvtColor(COLOR_RGB2GRAY), {cores} threads x cv2.setNumThreads(1)"}
        out["parity_spot_check"] = bool(np.array_equal(last.reshape(-1), dst[(done - 1) % 16].cpu().numpy()))
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank: int, local_rank: int, world: int):
    import torch

    from cvvidproc_b200 import _cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_near_gpu(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = WORKLOAD
    W, H, N = w["width"], w["height"], w["nframes"]
    row0, nrows = shard_plan(H, rank, world)
    nelem = nrows * W
    stride = (nelem + 127) // 128 * 128
    ctx = _cabi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # resident input: this rank's band of every frame, generated on the device
    stack = torch.empty((N, stride), dtype=torch.uint8, device=f"cuda:{local_rank}")
    out = torch.empty(stride, dtype=torch.uint8, device=f"cuda:{local_rank}")
    ctx.synth_frames_device(stack.data_ptr(), stride, W, H, 0, N, w["seed"], w["ndisks"], row0=row0, nrows=nrows)
    ctx.synchronize()
    gathered = None
    if world > 1:
        band_max = (H - (H // world) * (world - 1)) * W
        out_pad = torch.zeros(band_max, dtype=torch.uint8, device=f"cuda:{local_rank}")
        gathered = torch.empty(world * band_max, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def step():
        ctx.median_device(stack.data_ptr(), N, nelem, stride, out.data_ptr())
        if world > 1:
            out_pad[:nelem].copy_(out[:nelem], non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_pad)

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        launches0 = ctx.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern_evs = []
        sampler.active = True
        ev0.record(stream)
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.median_device(stack.data_ptr(), N, nelem, stride, out.data_ptr())
            b.record(stream)
            kern_evs.append((a, b))
            if world > 1:
                out_pad[:nelem].copy_(out[:nelem], non_blocking=True)
                dist.all_gather_into_tensor(gathered, out_pad)
        ev1.record(stream)
        barrier()
        sampler.active = False
        launches = ctx.launch_count - launches0
    total_ms = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_evs]))
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_max = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    job_mpxf = W * H * N / 1e6  # whole job, all ranks
    value = job_mpxf / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI: pinned host frames -> result in host memory
    pinned = _cabi.PinnedBuffer(N * nelem)
    host_frames = pinned.array.reshape(N, nelem)
    chunk = 50
    for i in range(0, N, chunk):  # fill the pinned buffer once (outside every timed region)
        host_frames[i : i + chunk] = stack[i : i + chunk, :nelem].cpu().numpy()
    host_out = np.empty(nelem, np.uint8)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ctx.median_begin(nelem, N)
        for i in range(0, N, 125):
            ctx.median_push_raw(host_frames[i].ctypes.data, min(125, N - i), nelem)
        ctx.median_finish(host_out)

    for _ in range(2):
        e2e_step()
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.active = False
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_value = job_mpxf / e2e_s
    sampler.stop()
    clocks = sampler.summary()
    # parity spot check of what was timed (cheap, outside the timed regions): e2e result == resident result
    same = bool(np.array_equal(host_out, out[:nelem].cpu().numpy()))

    # ---- roofline of the dominant (only) kernel
    peak, peak_src = load_peaks()
    algo_bytes = float(N) * nelem + nelem  # bytes one launch must move: every input byte once + the result
    achieved = algo_bytes / (kern_ms_max * 1e-3) / 1e9
    traffic = load_traffic() if world == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "median_pipe_kernel",
                "kernel_ms": kern_ms_max, "algorithmic_bytes_per_launch": algo_bytes}

    # ---- CPU baseline on rank 0 at N=1 only (bounded sample; reported, not the target)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            run, kind = cpu_median_fn()
            cores = os.cpu_count() or 1
            rows = 120
            sample_frames = np.ascontiguousarray(host_frames[:, : rows * W]).reshape(N, rows, W)
            want = run(sample_frames, cores)  # warm-up + parity of the sample against the GPU result
            same = same and bool(np.array_equal(want, host_out[: rows * W]))
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                run(sample_frames, cores)
            dt = (time.perf_counter() - t0) / reps
            cpu_baseline = {"value": rows * W * N / 1e6 / dt, "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"rows [0,{rows}) of all {N} frames, {reps} repetitions, all host threads"}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "unavailable",
                            "sample": f"failed: {exc}"}
    pinned.close()
    del stack

    highlight = None
    if not args.no_highlight:
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        try:
            highlight = run_highlight_section(ctx, torch, dist, rank, local_rank, world, args, sampler2, stream)
            highlight["clocks"] = sampler2.summary()
        finally:
            sampler2.stop()

    frame_source = None
    if world == 1 and not args.no_highlight:
        frame_source = run_frame_source_section(ctx, torch, local_rank, args, stream)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": w["name"], "width": W, "height": H, "nframes": N, "seed": w["seed"],
                       "ndisks": w["ndisks"], "sharding": "single GPU" if world == 1 else f"row bands x{world}, no data-path collective",
                       "l2": "input stack (2.07 GB / n_gpus) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(N * nelem),
                    "d2h_bytes_per_step": int(nelem), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity_spot_check": same,
            "highlight": highlight,
            "frame_source": frame_source,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# GPU arm, N > 1: frame-sharded median (BASELINE.json north_star: "the median shards over frame chunks")
# ------------------------------------------------------------------------------------------------
def run_gpu_arm_sharded(args, rank: int, local_rank: int, world: int):
    """Weak scaling: every rank holds its own 1000-frame chunk of the 1080p stream (frames [1000*rank, 1000*(rank+1))),
    so the job is the temporal median of a 1000*world-frame stack.  One step = the four phases of
    csrc/median_shard.cu with a one-element NCCL all-reduce as the barrier between them; the nibble counts travel by
    peer stores over NVLink from inside the counting kernels; every rank ends with the full result image."""
    import torch
    import torch.distributed as dist

    from cvvidproc_b200 import _cabi, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_near_gpu(local_rank)
    dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    W, H, N = w["width"], w["height"], w["nframes"]
    nelem = W * H
    ctx = _cabi.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    first, _ = sharded.frame_chunk(N * world, rank, world)
    stack = torch.empty((N, nelem), dtype=torch.uint8, device=dev)
    ctx.synth_frames_device(stack.data_ptr(), nelem, W, H, first, N, w["seed"], w["ndisks"])
    ctx.synchronize()
    job = sharded.ShardedMedian(ctx, nelem, rank, world)
    job.connect_processes()

    sampler = ClockSampler(local_rank)
    sampler.start()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def step():
        job.run(stack.data_ptr(), N, nelem)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_evs = []
    sampler.active = True
    ev0.record(stream)
    for _ in range(args.steps):
        evs = []
        for p in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            job.phase(p, stack.data_ptr(), N, nelem)
            b.record(stream)
            job.barrier()
            evs.append((a, b))
        phase_evs.append(evs)
    ev1.record(stream)
    barrier()
    sampler.active = False
    launches = ctx.launch_count - launches0
    total_ms = ev0.elapsed_time(ev1)
    phase_ms = [float(np.mean([evs[p][0].elapsed_time(evs[p][1]) for evs in phase_evs])) for p in range(4)]
    t = torch.tensor([total_ms] + phase_ms, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, phase_ms = float(t[0]), [float(v) for v in t[1:]]
    ms_per_step = total_ms / args.steps
    job_mpxf = W * H * N * world / 1e6
    value = job_mpxf / (ms_per_step * 1e-3)
    result_dev = ctx.copy_to_host(job.result_ptr(), nelem)

    # ---- end to end: this rank's frames start in pinned host memory, the full result ends in host memory
    pinned = _cabi.PinnedBuffer(N * nelem)
    host_t = torch.from_numpy(pinned.array).view(N, nelem)
    chunk = 50
    for i in range(0, N, chunk):
        host_t[i : i + chunk].copy_(stack[i : i + chunk])
    torch.cuda.synchronize()
    e2e_steps = max(3, min(args.steps, 10))
    host_out = None

    def e2e_step():
        nonlocal host_out
        with torch.cuda.stream(stream):
            for i in range(0, N, 125):
                stack[i : i + 125].copy_(host_t[i : i + 125], non_blocking=True)
        job.run(stack.data_ptr(), N, nelem)
        host_out = ctx.copy_to_host(job.result_ptr(), nelem)

    for _ in range(2):
        e2e_step()
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.active = False
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    sampler.stop()
    clocks = sampler.summary()

    # ---- parity spot check (outside the timed regions): every rank holds the same image, and on sampled elements it
    # is the order statistic of ALL ranks' frames (gathered through NCCL for the sample only)
    same = bool(np.array_equal(host_out, result_dev))
    cols = torch.arange(0, nelem, max(1, nelem // 4096), device=dev)
    mine = stack[:, cols].contiguous()
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    want = torch.sort(torch.cat(allc, 0), dim=0).values[(N * world) // 2].cpu().numpy()
    same = same and bool(np.array_equal(want, result_dev[cols.cpu().numpy()]))
    flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    same = bool(int(flag[0]))

    peak, peak_src = load_peaks()
    algo_bytes = float(N) * nelem + 32.0 * nelem  # one counting round: every input byte once + 32 B of counts per element
    k_ms = max(phase_ms[0], phase_ms[2])
    roofline = {"bound": "hbm", "achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                "kernel": "median_pipe_kernel (counting round 2, the slower of the two passes over the frames)",
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo_bytes,
                "phase_ms": {"count_hi": phase_ms[0], "pick_hi": phase_ms[1], "count_lo": phase_ms[2], "pick_lo": phase_ms[3]}}
    pinned.close()
    job.close()
    del stack

    highlight = None
    if not args.no_highlight:
        sampler2 = ClockSampler(local_rank)
        sampler2.start()
        try:
            highlight = run_highlight_section(ctx, torch, dist, rank, local_rank, world, args, sampler2, stream)
            highlight["clocks"] = sampler2.summary()
        finally:
            sampler2.stop()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": w["name"] + f"; N>1: every GPU holds its own {N}-frame chunk, the job is the median of "
                                               f"the {N * world}-frame stack",
                       "width": W, "height": H, "nframes": N * world, "frames_per_gpu": N, "seed": w["seed"],
                       "ndisks": w["ndisks"],
                       "sharding": f"frame chunks x{world}; two-round nibble-count exchange by NVLink peer stores "
                                   "(csrc/median_shard.cu), NCCL one-element all-reduce as the inter-phase barrier",
                       "l2": "input stack (2.07 GB per GPU) exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": job_mpxf / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(N * nelem) * world,
                    "d2h_bytes_per_step": int(nelem) * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": None,
            "parity_spot_check": same,
            "highlight": highlight,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["cuda", "reference"], default="cuda")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-highlight", action="store_true", help="skip the highlight-stage section")
    ap.add_argument("--median-sharding", choices=["frames", "rows"], default="frames",
                    help="N > 1: frame chunks with the NVLink count exchange (default, weak scaling) or row bands of "
                         "the single C2 stack with no data-path collective (strong scaling)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one process per GPU); running the single-GPU job",
              file=sys.stderr)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    elif world > 1 and args.median_sharding == "frames":
        run_gpu_arm_sharded(args, rank, local_rank, world)
    else:
        run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
