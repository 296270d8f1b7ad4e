#!/usr/bin/env python
"""bench.py — ORB frames/s of the B200-native ORB front-end (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step = one pass of the whole extractor (pyramid → FAST cells → quadtree → orientation → blur →
rBRIEF → output ordering) over one batch of synthetic 640×480 frames with the TUM1 settings
(nFeatures=1000, 8 levels, 1.2, FAST 20/7) — the configuration BASELINE.json's metric is quoted on.
`value`   : frames/s with the batch resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`     : frames/s through the public host-buffer call (orbx_extract_batch via the ORBextractor mirror):
            pinned host frames → H2D → pipeline → D2H keypoints+descriptors, every step.
`roofline`: the dominant kernel's algorithmic bytes per launch ÷ its CUDA-event duration vs measured HBM peak.
`cpu_baseline`: the CPU oracle (a port of the reference algorithm) on the host cores, bounded sample.
`match`   : Hamming kNN (k=2) Gpairs/s, 2000 queries × 10M descriptors, DB-sharded over the ranks.
For N>1 launch with torchrun (one rank per GPU); frames are sharded per rank with no collective.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LEVELS, SCALE, INI_TH, MIN_TH = 8, 1.2, 20, 7
# BASELINE.json configs (frame shape, nFeatures, default frames per GPU per step).  `tum1` is the configuration the
# headline metric is quoted on (640×480, 1000 keypoints, TUM1.yaml); the others are extra measurements.
WORKLOADS = {
    "tum1": (640, 480, 1000, 2048),
    "euroc": (752, 480, 1200, 512),
    "kitti": (1241, 376, 2000, 256),
    "4k": (3840, 2160, 8000, 64),
}
W_IMG, H_IMG, NFEAT = 640, 480, 1000
LEVEL_PX, SUM_P, B_ALG_FRAME, STAGE_BYTES = [], 0, 0, {}


def set_workload(name: str):
    """level sizes (fp32 cvRound of size·1/1.2^l) and the stage-streaming byte model."""
    global W_IMG, H_IMG, NFEAT, LEVEL_PX, SUM_P, B_ALG_FRAME, STAGE_BYTES
    W_IMG, H_IMG, NFEAT, batch = WORKLOADS[name]
    sf, LEVEL_PX = np.float32(1.0), []
    for l in range(LEVELS):
        inv = np.float32(1.0) / sf
        LEVEL_PX.append(int(np.rint(np.float32(W_IMG) * inv)) * int(np.rint(np.float32(H_IMG) * inv)))
        sf = np.float32(float(sf) * float(np.float32(SCALE)))
    SUM_P = sum(LEVEL_PX)
    B_ALG_FRAME = (sum(LEVEL_PX[:-1]) + sum(LEVEL_PX[1:])) + SUM_P + 2 * SUM_P + NFEAT * (749 + 512) + NFEAT * (32 + 28)
    STAGE_BYTES = {  # algorithmic bytes per frame of each stage (read once + write once)
        "pyramid": sum(LEVEL_PX[:-1]) + sum(LEVEL_PX[1:]),
        "fast_cells": SUM_P,
        "quadtree": 0,
        "assemble": NFEAT * 28,
        "blur": 2 * SUM_P,
        "orient_desc": NFEAT * (749 + 512) + NFEAT * 32,
    }
    return batch


set_workload("tum1")


def make_frames(count: int, first_index: int) -> np.ndarray:
    """`count` distinct corner-dense synthetic frames; frame i derives from seed (first_index+i)//8 by a roll."""
    from dani_slam_b200 import synth
    out = np.empty((count, H_IMG, W_IMG), np.uint8)
    cache = {}
    for i in range(count):
        gi = first_index + i
        seed, shift = divmod(gi, 8)
        if seed not in cache:
            cache[seed] = synth.throughput_frame(seed, W_IMG, H_IMG)
        out[i] = np.roll(cache[seed], (37 * shift, 53 * shift), axis=(0, 1)) if shift else cache[seed]
    return out


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------
# CPU arms (the oracle is test/baseline infrastructure: only this function may touch oracle/)
# ----------------------------------------------------------------------------------------------------
def cpu_extract_fps(frames: np.ndarray, threads: int, seconds: float, use_ref: bool):
    """Frames/s of the reference algorithm on the host: one frame per thread, `threads` threads."""
    from oracle import oracle
    kind = "port"
    make = lambda: oracle.Extractor(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH)  # noqa: E731
    if use_ref:
        try:
            from oracle import ref_binding
            if ref_binding.available():
                make = lambda: ref_binding.Extractor(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH)  # noqa: E731
                kind = "reference"
        except Exception:
            pass
    done = [0] * threads
    stop_at = [0.0]

    def worker(t):
        ex = make()
        i = t
        while time.perf_counter() < stop_at[0]:
            ex.extract(frames[i % len(frames)])
            done[t] += 1
            i += threads

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    stop_at[0] = t0 + seconds
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done), dt, kind


def bench_config(args, world, B):
    """The `config` object of the JSON line: the workload and its sharding, nothing measured — identical for this arm and for
    `--impl reference` (the driver compares the two)."""
    return {"workload": f"{args.workload}_{W_IMG}x{H_IMG}_nf{NFEAT}_8lv_fast20_7", "frames_per_gpu_per_step": B,
            "global_batch": world * B, "parallelism": f"frame-batch x{world} (no collective)",
            "l2": f"inputs {B * H_IMG * W_IMG / 1e6:.0f} MB + {B * SUM_P * 2 / 1e6:.0f} MB of pyramid planes touched per step > 126 MB L2 (no flush needed)"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    frames = make_frames(max(16, min(64, 2 * cores)), 0)
    per_step = 1.5  # seconds of wall clock per step: a bounded sample of the workload
    cpu_extract_fps(frames, cores, 1.0, True)  # page in / warm caches
    for _ in range(max(0, args.warmup - 1)):
        cpu_extract_fps(frames, cores, per_step, True)
    tot_frames, tot_t, kind = 0, 0.0, "port"
    for _ in range(args.steps):
        fps, n, dt, kind = cpu_extract_fps(frames, cores, per_step, True)
        tot_frames += n
        tot_t += dt
    value = tot_frames / tot_t
    line = {
        "impl": "reference", "metric": "orb_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, max(1, args.gpus), args.batch),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{tot_frames} frames of the bench workload over {args.steps} steps of {per_step}s "
                                   f"({tot_frames / max(args.steps, 1):.0f} frames per step), one frame per thread"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def measure_latency(orbx, device, frames, args):
    """orbx_extract — the call Frame::ExtractORB makes once per image (src/Frame.cc:420-427) — on ONE frame held in pageable host
    memory (a cv::Mat-like buffer), timed per call on the host clock around the C-ABI call: copy in, ~17 kernels, copy out, one
    synchronisation.  Mono: 1000 calls.  Stereo: two handles called from two threads at the same time (src/Frame.cc:124-127), the
    pair timed from the common start to the later finish."""
    import ctypes as C
    L = orbx.lib()
    cap = NFEAT + 2 * LEVELS + 8
    n_calls = args.latency_calls

    class One:
        def __init__(self, img):
            self.img = np.array(img, copy=True)                              # ordinary (pageable) memory
            self.h = L.orbx_create(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH, device, W_IMG, H_IMG, 1)
            if not self.h:
                raise SystemExit("orbx_create failed: " + L.orbx_last_error(None).decode())
            self.kps = np.zeros(cap, orbx.KP_DTYPE); self.desc = np.zeros((cap, 32), np.uint8)
            self.n, self.mono = C.c_int(0), C.c_int(0)
            self.args = (self.h, self.img.ctypes.data_as(C.c_void_p), H_IMG, W_IMG, W_IMG, None, 0, 0, 1000, self.kps.ctypes.data_as(C.c_void_p),
                         self.desc.ctypes.data_as(C.c_void_p), cap, C.byref(self.n), C.byref(self.mono))

        def call(self):
            rc = L.orbx_extract(*self.args)
            if rc != 0:
                raise SystemExit(f"orbx_extract failed: {rc} {L.orbx_last_error(self.h)}")

        def close(self):
            L.orbx_destroy(self.h)

    def pct(a):
        a = np.sort(np.asarray(a, np.float64)) / 1000.0                      # ns → µs
        return {"p50_us": float(a[len(a) // 2]), "p99_us": float(a[min(len(a) - 1, int(len(a) * 0.99))]), "min_us": float(a[0]),
                "mean_us": float(a.mean()), "calls": int(len(a))}

    left, right = One(frames[0]), One(frames[1 % len(frames)])
    for _ in range(20):
        left.call(); right.call()
    mono = []
    for _ in range(n_calls):
        t0 = time.perf_counter_ns(); left.call(); mono.append(time.perf_counter_ns() - t0)
    n_kp = left.n.value
    # stereo pair: two threads released together by a barrier per iteration
    n_pairs = max(50, n_calls // 2)
    bar = threading.Barrier(3)
    def side(o):
        for _ in range(n_pairs):
            bar.wait(); o.call(); bar.wait()
    ths = [threading.Thread(target=side, args=(o,)) for o in (left, right)]
    for t in ths:
        t.start()
    pair = []
    for _ in range(n_pairs):
        bar.wait(); t0 = time.perf_counter_ns(); bar.wait(); pair.append(time.perf_counter_ns() - t0)
    for t in ths:
        t.join()
    left.close(); right.close()
    out = {"api": "orbx_extract (one 640x480-class frame in pageable host memory → keypoints + descriptors in host memory), per call on the host clock",
           "workload": f"{W_IMG}x{H_IMG}_nf{NFEAT}", "keypoints": n_kp, "mono": pct(mono), "stereo_pair_two_threads": pct(pair),
           "note": "stereo: a Python thread barrier releases both callers, so the pair figure includes its wake-up jitter", "cpu_reference": None}
    if not args.no_cpu:
        from oracle import oracle as _orc
        kind, make = "port", (lambda: _orc.Extractor(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH))
        try:
            from oracle import ref_binding
            if ref_binding.available():
                kind, make = "reference", (lambda: ref_binding.Extractor(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH))
        except Exception:
            pass
        cx = make()
        cx.extract(frames[0])
        ts = []
        for i in range(12):
            t0 = time.perf_counter_ns(); cx.extract(frames[i % len(frames)]); ts.append(time.perf_counter_ns() - t0)
        out["cpu_reference"] = {"ms_per_frame_1_thread": float(np.median(ts)) / 1e6, "kind": kind, "sample": "median of 12 frames, one thread",
                                "cv2_primitives_ms_per_frame_1_thread": cv2_primitives_ms(frames),
                                "cv2_note": "the real OpenCV (cv2 wheel, SIMD builds) primitives of the path alone on one thread — 7 resizes, 8 borders, "
                                            "FAST-9 with NMS on every level at iniThFAST, 8 Gaussian blurs; no quadtree, orientation or descriptors. "
                                            "The reference arm runs the oracle's scalar restatements of these primitives under the reference's own "
                                            "ORBextractor.cc, so an OpenCV-linked build of the reference would sit between the two figures"}
    return out


def cv2_primitives_ms(frames, seconds=1.5):
    """Context for the CPU baseline: how long the REAL OpenCV primitives of the path take on one thread.  None without cv2."""
    try:
        import cv2
        return _cv2_primitives_ms(cv2, frames, seconds)
    except Exception:                                  # a reported context figure must never take the bench down
        return None


def _cv2_primitives_ms(cv2, frames, seconds):
    cv2.setNumThreads(1)
    fast = cv2.FastFeatureDetector_create(INI_TH, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    sizes, sf = [], np.float32(1.0)
    for _ in range(LEVELS):
        inv = np.float32(1.0) / sf
        sizes.append((int(np.rint(np.float32(W_IMG) * inv)), int(np.rint(np.float32(H_IMG) * inv))))
        sf = np.float32(float(sf) * float(np.float32(SCALE)))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds or n < 3:
        lv = [np.ascontiguousarray(frames[n % len(frames)])]
        for l in range(1, LEVELS):
            lv.append(cv2.resize(lv[-1], sizes[l], interpolation=cv2.INTER_LINEAR))
        for l in range(LEVELS):
            cv2.copyMakeBorder(lv[l], 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
            fast.detect(lv[l], None)
            cv2.GaussianBlur(lv[l], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        n += 1
    return 1000.0 * (time.perf_counter() - t0) / n


# ----------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank: int) -> None:
    """One process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned host buffers (first touch) and the
    copy submissions stay local to the GPU's PCIe root.  Best effort: does nothing when sysfs has no NUMA information."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        bus = bus[-12:] if len(bus) > 12 else bus                   # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            # one NUMA node (or none reported): give every rank its own contiguous slice of the allowed cores instead, so that the
            # copy-submitting threads of the ranks do not migrate over each other
            world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
            allowed = sorted(os.sched_getaffinity(0))
            if world > 1 and len(allowed) >= 2 * world:
                per = len(allowed) // world
                mine = set(allowed[local_rank * per:(local_rank + 1) * per])
                os.sched_setaffinity(0, mine)
                print(f"[bench] rank {local_rank}: single NUMA node, bound to cores {min(mine)}-{max(mine)}", file=sys.stderr)
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b_ = part.partition("-")
            cpus.update(range(int(a), int(b_ or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            print(f"[bench] rank {local_rank}: GPU {bus} on NUMA node {node}, bound to {len(cpus)} CPUs", file=sys.stderr)
    except Exception as e:  # noqa: BLE001
        print(f"[bench] NUMA binding skipped: {e}", file=sys.stderr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tum1", choices=sorted(WORKLOADS), help="tum1 = the headline configuration")
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: per workload)")
    ap.add_argument("--no-match", action="store_true", help="skip the Hamming kNN section")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--match-db", type=int, default=10_000_000)
    ap.add_argument("--no-bow", action="store_true", help="skip the bag-of-words section")
    ap.add_argument("--wc-input", action="store_true", help="host input frames in write-combined pinned memory (e2e / copy-ceiling experiment)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-frame latency section")
    ap.add_argument("--latency-calls", type=int, default=1000)
    ap.add_argument("--bow-levels", type=int, default=6, help="depth of the synthetic k=10 vocabulary (ORBvoc: 6)")
    args = ap.parse_args()
    default_batch = set_workload(args.workload)
    if args.batch <= 0:
        args.batch = default_batch
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)

    # rank 0's stdout must be the one JSON line, but NCCL prints its version banner there when the first communicator is
    # created: everything written to fd 1 goes to stderr until the line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as td
    from dani_slam_b200 import orbx, sharded

    if not torch.cuda.is_available() or orbx.lib().orbx_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist_on:
            td.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if not dist_on:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if not dist_on:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t.item())

    B, K, Wm = args.batch, args.steps, args.warmup
    cap = NFEAT + 2 * LEVELS + 8            # a level can overshoot its quota by at most 2
    frames = make_frames(B, rank * B)
    h_frames = torch.from_numpy(frames).pin_memory()
    d_frames = h_frames.to(dev)
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_mono = torch.zeros(B, dtype=torch.int32, device=dev)
    ex = orbx.ORBextractor(NFEAT, SCALE, LEVELS, INI_TH, MIN_TH, device=local_rank, max_width=W_IMG, max_height=H_IMG, max_batch=B)
    ex.cap = cap
    stream = torch.cuda.ExternalStream(ex.stream(), device=dev)
    torch.cuda.synchronize(dev)

    def step_device():
        ex.extract_batch_device(d_frames.data_ptr(), H_IMG * W_IMG, B, H_IMG, W_IMG, W_IMG, d_kps.data_ptr(), d_desc.data_ptr(),
                                cap, d_n.data_ptr(), d_mono.data_ptr())

    # ---------------- device-resident throughput (`value`) ----------------
    for _ in range(Wm):
        step_device()
    ex.sync()
    n_host = d_n.cpu().numpy()
    if n_host.min() < NFEAT or n_host.max() > NFEAT + 2 * LEVELS:
        raise SystemExit(f"sanity gate failed: keypoints per frame {n_host.min()}..{n_host.max()} (expected {NFEAT}..{NFEAT + 16})")
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ex.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step_device()
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = ex.launch_count() - launches0
    frames_total = world * B * K
    value = frames_total / (ms_dev / 1000.0)

    # ---------------- end to end through the host-buffer API (`e2e`) ----------------
    h_imgs = h_frames.numpy()
    if args.wc_input:
        import ctypes as _C
        wc_ptr = orbx.lib().orbx_host_alloc_wc(h_imgs.nbytes)
        if not wc_ptr:
            raise SystemExit("orbx_host_alloc_wc failed")
        _C.memmove(wc_ptr, h_imgs.ctypes.data, h_imgs.nbytes)
        h_imgs = np.ctypeslib.as_array((_C.c_uint8 * h_imgs.nbytes).from_address(wc_ptr)).reshape(h_imgs.shape)
    ptrs_kps = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
    ptrs_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    out_k = ptrs_kps.numpy().view(np.uint8).reshape(B, cap * 28).view(orbx.KP_DTYPE)
    out_d = ptrs_desc.numpy()
    out_n = np.zeros(B, np.int32)
    out_m = np.zeros(B, np.int32)
    import ctypes as C
    img_ptrs = (C.c_void_p * B)(*[h_imgs.ctypes.data + b * H_IMG * W_IMG for b in range(B)])

    def step_e2e():
        rc = ex.L.orbx_extract_batch(ex.h, img_ptrs, B, H_IMG, W_IMG, W_IMG, None, 0, 0, 0, out_k.ctypes.data_as(C.c_void_p),
                                     out_d.ctypes.data_as(C.c_void_p), cap, out_n.ctypes.data_as(C.c_void_p), out_m.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise SystemExit(f"orbx_extract_batch failed: {rc} {ex.L.orbx_last_error(ex.h)}")

    for _ in range(Wm):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    torch.cuda.synchronize(dev)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()   # sampled across both timed regions (device-resident and end-to-end)
    e2e_value = frames_total / t_e2e
    n_avg = float(out_n.mean())
    h2d = B * H_IMG * W_IMG
    d2h = int(B * cap * 60 + 8 * B)        # the call copies cap-strided keypoint (28 B) and descriptor (32 B) rows + n_out, mono_index

    # transfer ceiling of the host path: the same copies on the same streams with the same chunking, no kernels
    def step_copy():
        rc = ex.L.orbx_copy_only_batch(ex.h, img_ptrs, B, H_IMG, W_IMG, W_IMG, out_k.ctypes.data_as(C.c_void_p), out_d.ctypes.data_as(C.c_void_p), cap)
        if rc != 0:
            raise SystemExit(f"orbx_copy_only_batch failed: {rc} {ex.L.orbx_last_error(ex.h)}")

    for _ in range(2):
        step_copy()
    barrier()
    Kc = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Kc):
        step_copy()
    torch.cuda.synchronize(dev)
    t_copy = max_over_ranks(time.perf_counter() - t0)
    barrier()
    copy_ceiling = world * B * Kc / t_copy

    # ---------------- single-frame latency of the drop-in call (`latency`; BASELINE config 1, rank 0 only) ----------------
    latency = None
    if rank == 0 and not args.no_latency:
        latency = measure_latency(orbx, local_rank, frames, args)

    # ---------------- per-stage device time → roofline of the dominant kernel ----------------
    ex.set_profiling(True)
    ex.stage_ms(reset=True)
    for _ in range(5):
        step_device()
    ex.sync()
    stage_ms, calls = ex.stage_ms()
    ex.set_profiling(False)
    per_call = {k: v / max(calls, 1) for k, v in stage_ms.items()}
    dominant = max(per_call, key=per_call.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom_bytes = STAGE_BYTES[dominant] * B
    dom_gbs = dom_bytes / (per_call[dominant] / 1000.0) / 1e9 if per_call[dominant] > 0 else 0.0
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))   # ncu dram bytes per frame
        traffic = prof.get(dominant) * B if (prof.get(dominant) is not None and args.workload == "tum1") else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": dom_gbs, "peak": peak_gbs, "unit": "GB/s",
                "frac": dom_gbs / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": per_call[dominant],
                "stage_ms_per_batch": per_call, "stage_share": {k: v / max(sum(per_call.values()), 1e-9) for k, v in per_call.items()}}
    pipeline_gbs = B_ALG_FRAME * (value / world) / 1e9
    # instruction-issue view of the same kernel: thread instructions per frame from the committed ncu capture (profiles/issue.json,
    # smsp__thread_inst_executed.sum per frame) × frames per launch ÷ the launch time measured here, against 128 lanes/clk/SM
    roofline_issue = None
    try:
        iss = json.load(open(os.path.join(ROOT, "profiles", "issue.json")))
        if args.workload == "tum1" and iss.get(dominant) and per_call[dominant] > 0:
            clk = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
            peak_i = 148 * 128 * clk
            ach = float(iss[dominant]) * B / (per_call[dominant] / 1000.0)
            roofline_issue = {"bound": "issue", "kernel": dominant, "achieved": ach, "peak": peak_i, "unit": "thread-instructions/s", "frac": ach / peak_i,
                              "thread_inst_per_frame": float(iss[dominant]), "thread_inst_per_pixel": float(iss[dominant]) / SUM_P,
                              "peak_source": "148 SMs x 4 schedulers x 32 lanes x median SM clock of this run",
                              "source": iss.get("source")}
    except Exception:
        pass
    roofline_pipeline = {"bound": "hbm", "achieved": pipeline_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": pipeline_gbs / peak_gbs,
                         "algorithmic_bytes_per_frame": B_ALG_FRAME}

    # ---------------- Hamming kNN, DB-sharded (`match`) ----------------
    match = None
    if not args.no_match:
        nq, ndb = 2000, args.match_db
        lo, hi = sharded.shard_bounds(ndb, rank, world)
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        d_db = torch.randint(0, 256, (hi - lo, 32), dtype=torch.uint8, device=dev, generator=g)
        gq = torch.Generator(device=dev)
        gq.manual_seed(99)
        d_q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=gq)
        sm = sharded.CudaShardedMatcher(local_rank)
        mstream = torch.cuda.ExternalStream(sm.m.stream(), device=dev)
        group = None

        if dist_on:
            sm.init_comm()                       # orbx_comm over NCCL: the collective runs inside the C ABI (orbx_knn2_sharded)

        def step_match():
            if dist_on:
                return sm.knn2_cabi(d_q, d_db, lo)
            rec = sm.local_top2(d_q, d_db, lo)
            return rec[0], rec[1]

        for _ in range(2):
            step_match()
        barrier()
        Km = max(3, min(K, 5))
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(torch.cuda.current_stream(dev))
        for _ in range(Km):
            idx, dist = step_match()
        m1.record(torch.cuda.current_stream(dev))
        m1.synchronize()
        barrier()
        ms_m = max_over_ranks(m0.elapsed_time(m1))
        gpairs = nq * ndb * Km / (ms_m / 1000.0) / 1e9
        # parity of the sharded result on this hardware: rank 0 rebuilds the whole database from the per-rank seeds, scans it
        # unsharded on its own GPU, and every rank's merged (idx, dist) must equal that scan bit for bit
        parity = None
        if dist_on:
            full = []
            if rank == 0:
                for r in range(world):
                    rlo, rhi = sharded.shard_bounds(ndb, r, world)
                    gr = torch.Generator(device=dev)
                    gr.manual_seed(1234 + r)
                    full.append(torch.randint(0, 256, (rhi - rlo, 32), dtype=torch.uint8, device=dev, generator=gr))
                rec = sm.local_top2(d_q, torch.cat(full), 0)
                ref_rec = torch.stack([rec[0], rec[1]]).contiguous()
                del full
            else:
                ref_rec = torch.empty((2, nq, 2), dtype=torch.int32, device=dev)
            td.broadcast(ref_rec, src=0)
            same = torch.tensor([int(torch.equal(ref_rec[0], idx) and torch.equal(ref_rec[1], dist))], dtype=torch.int32, device=dev)
            td.all_reduce(same, op=td.ReduceOp.MIN)
            parity = bool(same.item())
        # the shipped kernel for this size is the tensor-core GEMM (tcgen05.mma.kind::i8 over 0/1 bytes: 2·256 integer ops per pair); the POPC
        # kernel is timed beside it on rank 0's shard as the cross-check and the previous round's figure
        tc_used = sm.m.tc_launches() > 0
        popc_gpairs = None
        if not dist_on:
            os.environ["ORBX_KNN_POPC"] = "1"
            smp = sharded.CudaShardedMatcher(local_rank)
            del os.environ["ORBX_KNN_POPC"]
            rec_p = smp.local_top2(d_q, d_db, lo)
            torch.cuda.synchronize(dev)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(torch.cuda.current_stream(dev))
            for _ in range(2):
                rec_p = smp.local_top2(d_q, d_db, lo)
            p1.record(torch.cuda.current_stream(dev))
            p1.synchronize()
            popc_gpairs = nq * ndb * 2 / (p0.elapsed_time(p1) / 1000.0) / 1e9
            popc_same = bool(torch.equal(rec_p[0], idx) and torch.equal(rec_p[1], dist))
        pk = {}
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "pipe_peaks.json")))
        except Exception:
            pass
        popc_peak = float(pk.get("popc_lanes_per_clk_per_sm", 16.0)) * float(pk.get("sm_count", 148)) * 1e6 * float(clocks.get("sm_max_mhz") or 1965.0)
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        tops = 2 * 256 * gpairs * 1e9 / world / 1e12                    # integer multiply-adds counted as 2 ops, per GPU
        match = {"metric": "hamming_knn2_gpairs_per_s", "value": gpairs, "unit": "Gpairs/s", "nq": nq, "ndb": ndb,
                 "scaling": "strong", "ms_per_step": ms_m / Km, "steps": Km,
                 "kernel": "k_knn2_tc (tcgen05.mma.kind::i8, TMEM accumulators, fused top-2 epilogue)" if tc_used else "k_knn2_partial (POPC)",
                 "collective": "orbx_knn2_sharded (C ABI): one ncclAllGather of the packed per-shard top-2 record (nq*2*2 int32 = 32 KB per rank)" if dist_on else None,
                 "parity": parity,
                 "parity_note": "sharded result of every rank == rank 0's unsharded scan of the whole DB (indices and distances)" if dist_on else
                                "single GPU: see cpu_baseline.sample and tests/test_match_gpu.py::test_config4_full_size_vs_oracle",
                 "popc_per_s": 8 * gpairs * 1e9, "cpu_baseline": None,
                 "roofline": {"bound": "tensor", "achieved": tops, "peak": 4500.0, "unit": "TOP/s (int8) per GPU", "frac": tops / 4500.0,
                              "peak_source": "nominal dense 8-bit tensor rate of B200 (B200_PROFILING.md: 4.5 POP/s; MEASURED_PEAKS.json holds no 8-bit "
                                             "figure)",
                              "frac_of_2x_measured_bf16": tops / (2 * bf16_peak),
                              "measured_bf16_tflops": bf16_peak,
                              "note": "algorithmic work: one 256-term dot product per pair = 512 integer ops.  The cuBLAS bf16 GEMM of MEASURED_PEAKS.json "
                                      "reaches 0.73 of ITS nominal rate on this pool; twice that figure is given beside the nominal 8-bit peak.  ncu "
                                      "(profiles/ncu_knn_tc_r02.txt): tensor pipe active for about 0.8 of the kernel"},
                 "popc_kernel": None if popc_gpairs is None else {
                     "value": popc_gpairs, "unit": "Gpairs/s", "same_result": popc_same,
                     "roofline": {"bound": "int-popc", "issued": 5 * popc_gpairs * 1e9, "achieved": 8 * popc_gpairs * 1e9, "peak": popc_peak,
                                  "unit": "POPC/s per GPU", "frac_issued": 5 * popc_gpairs * 1e9 / popc_peak, "frac": 8 * popc_gpairs * 1e9 / popc_peak,
                                  "peak_source": "profiles/pipe_peaks.json (register-resident POPC microbenchmark, lanes/clk/SM) x 148 SMs x sm_max_mhz",
                                  "note": "`issued` = the 5 POPC.32 per pair the carry-save kernel executes (the honest pipe utilisation); `achieved` = "
                                          "the algorithmic 8 POPC.32 per pair of SURVEY.md §8(d)"}}}
        if world == 1 and not args.no_cpu:
            # the oracle's popcount kNN on all host cores over a DB slice (the scan is linear in the DB length)
            from oracle import oracle as _orc
            cores = os.cpu_count() or 1
            h_q = d_q.cpu().numpy()
            slice_rows = min(ndb, 400_000)
            h_db = d_db[:slice_rows].cpu().numpy()
            t0 = time.time(); reps = 0
            while time.time() - t0 < 4.0:
                cidx, cdist = _orc.knn2(h_q, h_db, nthreads=cores)
                reps += 1
            dt = time.time() - t0
            grec = sm.local_top2(d_q, d_db[:slice_rows].contiguous(), 0)
            torch.cuda.synchronize(dev)
            same = bool(np.array_equal(grec[0].cpu().numpy(), cidx) and np.array_equal(grec[1].cpu().numpy(), cdist))
            match["cpu_baseline"] = {"value": nq * slice_rows * reps / dt / 1e9, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                     "sample": f"{nq} queries x {slice_rows} DB rows, {reps} passes in {dt:.1f}s on {cores} threads (oracle popcount kNN); "
                                               f"GPU result on the same slice identical: {same}"}
        del d_db

    # ---------------- bag-of-words transform of the extracted descriptors (`bow`) ----------------
    bow = None
    if not args.no_bow:
        from dani_slam_b200 import synth as _synth
        voc = _synth.vocabulary_complete(10, args.bow_levels, seed=7)
        vdev = orbx.ORBVocabulary(local_rank).from_nodes(voc)
        i32 = torch.int32
        b_word = torch.zeros(B * cap, dtype=i32, device=dev); b_node = torch.zeros_like(b_word)
        b_ids = torch.zeros_like(b_word); b_vals = torch.zeros(B * cap, dtype=torch.float64, device=dev)
        b_fn = torch.zeros_like(b_word); b_fi = torch.zeros_like(b_word); b_fo = torch.zeros(B * (cap + 1), dtype=i32, device=dev)
        b_nb = torch.zeros(B, dtype=i32, device=dev); b_nf = torch.zeros(B, dtype=i32, device=dev)
        vstream = torch.cuda.ExternalStream(vdev.stream(), device=dev)

        def step_bow():
            vdev.transform_batch_device(d_desc.data_ptr(), cap * 32, d_n.data_ptr(), B, cap, 4, b_word.data_ptr(), b_node.data_ptr(), b_ids.data_ptr(),
                                        b_vals.data_ptr(), b_nb.data_ptr(), b_fn.data_ptr(), b_fo.data_ptr(), b_fi.data_ptr(), b_nf.data_ptr())

        torch.cuda.synchronize(dev)
        for _ in range(3):
            step_bow()
        vdev.sync()
        barrier()
        Kb = max(3, min(K, 10))
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(vstream)
        for _ in range(Kb):
            step_bow()
        w1.record(vstream)
        w1.synchronize()
        barrier()
        ms_b = max_over_ranks(w0.elapsed_time(w1))
        n_desc = float(d_n.sum().item())
        feats = sum_over_ranks(n_desc) * Kb / (ms_b / 1000.0)
        bow = {"metric": "bow_descriptors_per_s", "value": feats, "unit": "descriptors/s", "frames_per_s": world * B * Kb / (ms_b / 1000.0),
               "ms_per_step": ms_b / Kb, "steps": Kb, "scaling": "weak",
               "vocabulary": f"synthetic complete tree k=10 L={args.bow_levels} ({len(voc['parent']) + 1} nodes, {int(voc['is_leaf'].sum())} words), TF-IDF, L1, levelsup 4",
               "words_per_frame": float(b_nb.float().mean().item()), "cpu_baseline": None}
        if world == 1 and not args.no_cpu:
            from oracle import oracle as _orc
            from concurrent.futures import ThreadPoolExecutor
            cores = os.cpu_count() or 1
            ov = _orc.Vocabulary(voc=voc)
            h_desc = d_desc[: min(B, 4 * cores)].cpu().numpy(); h_n = d_n[: min(B, 4 * cores)].cpu().numpy()
            t0 = time.time(); done = 0
            with ThreadPoolExecutor(cores) as pool:
                while time.time() - t0 < 5.0:
                    list(pool.map(lambda b_: ov.transform(h_desc[b_, : h_n[b_]], 4), range(len(h_n))))
                    done += int(h_n.sum())
            dt = time.time() - t0
            bow["cpu_baseline"] = {"value": done / dt, "unit": "descriptors/s", "cores": cores, "kind": "port",
                                   "sample": f"{done} descriptors in {dt:.1f}s, one frame per thread on {cores} threads (oracle/bow_oracle.cpp)"}
        del vdev

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        fps, n, dt, kind = cpu_extract_fps(frames[: min(B, 64)], cores, 12.0, False)
        cpu_baseline = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                        "sample": f"{n} frames of the same workload in {dt:.1f}s, one frame per thread on {cores} threads"}

    if rank == 0:
        line = {
            "metric": "orb_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": bench_config(args, world, B),
            "keypoints_per_frame": n_avg,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "orbx_extract_batch (host pinned buffers, H2D + D2H inside the timed region)" + (", write-combined input" if args.wc_input else ""),
                    "copy_ceiling_frames_per_s": copy_ceiling,
                    "copy_ceiling_note": "orbx_copy_only_batch: the same bytes over the same streams and chunks with no kernel launched, all ranks at once; "
                                         f"{(h2d + d2h) * copy_ceiling / B / 1e9:.1f} GB/s over PCIe in total"},
            "latency": latency,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_issue": roofline_issue,
            "roofline_pipeline": roofline_pipeline,
            "cpu_baseline": cpu_baseline,
            "match": match,
            "bow": bow,
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if dist_on:
        td.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
