#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json: synthetic 1080p frames/s through ORB extract + match.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU restatement of the reference on host cores)

One "step" = one pass of the hot path over one batch of B synthetic 1080p frames per GPU: ORBextractor (1000
features, 8 levels, scale 1.2, FAST 20/7) on every frame, then SearchByProjection of every frame against its
predecessor (th=15, retry at 30).  Workload = BASELINE.json configs[1] (the single-GPU 1080p extractor config)
batched so the kernels see inputs larger than L2.

JSON line keys: see the task contract.  `value` = frames/s with the frames already resident in HBM; `e2e` = the
same through the C-ABI from pinned HOST frames with host<->device copies timed; `roofline` = the FAST-9 score
kernel against the measured HBM copy peak; `cpu_baseline` = the oracle (CPU port of the reference) on host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1920, 1080, 1000
PYRAMID_PX = 6_419_321                      # sum of the 8 level sizes of a 1080p frame
# SURVEY.md 8(d), K2 unfused: pyramid read once + u8 score map written once -- the HBM work the fused kernel replaces
# (and the denominator round 1's k_fast_score was quoted on)
FAST_K2_BYTES_PER_FRAME = 2 * PYRAMID_PX
# SURVEY.md 8(d), K2+K3 fused ("6,419,321 B read + 8 B x candidates written"; here 4 B per candidate slot + 4 B per cell count)
N_CELLS = 6257
WORKLOAD = ("ORBextractor 1000 keypoints, 1920x1080 synthetic frames, 8-level pyramid + SearchByProjection vs previous "
            "frame (BASELINE configs[1], batched)")
# pgb_orb_run_stage ids: 0 pyramid, 1 FAST score + cell NMS (fused k_fast_cells2), 3 octree, 4 orientation + descriptor;
# 2 = the round-1 unfused pair (k_fast_score -> k_cells) recomputing the same candidates, timed for the A/B only
STAGES = {"pyramid": 0, "fast_cells": 1, "octree": 3, "orient_desc": 4}
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_fast_cells_traffic.json")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per FRAME of the dominant kernel, read from the summary that
    tools/ncu_traffic.py wrote from the committed `ncu --set full` capture (None when there is no capture)."""
    try:
        t = json.load(open(TRAFFIC_JSON))
        return (float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])) / float(t["frames_per_launch"]), t.get("source")
    except Exception:
        return None, None


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (between begin() and end()).  NVML is polled from
    a thread every 2 ms (what nvidia-smi reads, without its start-up latency: a 10-step timed region lasts ~30 ms);
    if NVML cannot be loaded, `nvidia-smi -lms 20` is used and the samples are filtered by arrival time."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples, self.windows, self.p, self.nv, self.max_mhz = [], [], None, None, None   # samples: (t, mhz, reasons)
        self.stop_flag = False
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]
            get(self.h); nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.nv, self.source = (nv, get, bits), "nvml, 2 ms poll"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.source = "nvidia-smi -lms 20"
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                       str(index), "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll(self):
        nv, get, bits = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)); r = int(get(self.h))
                self.samples.append((time.monotonic(), mhz, [n for b, n in bits if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.p.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 6 and c[0].replace(".", "").isdigit():
                if c[1].replace(".", "").isdigit():
                    self.max_mhz = max(self.max_mhz or 0.0, float(c[1]))
                self.samples.append((time.monotonic(), float(c[0]), [n for n, v in zip(self.NAMES, c[2:6]) if v.lower().startswith("active")]))

    def wait_ready(self, timeout=5.0):
        t0 = time.monotonic()
        while not self.samples and time.monotonic() - t0 < timeout and (self.nv or self.p):
            time.sleep(0.01)

    def begin(self):
        self._t0 = time.monotonic()

    def end(self):
        self.windows.append((self._t0, time.monotonic()))

    def stop(self):
        if not self.nv and not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.stop_flag = True
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                pass
        slack = 0.0 if self.nv else 0.02                                  # a piped nvidia-smi line arrives up to one period late
        inside = [s for s in self.samples if any(a <= s[0] <= b + slack for a, b in self.windows)]
        sm = [s[1] for s in inside]
        reasons = sorted({n for s in inside for n in s[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": self.source,
                "window_ms": round(1e3 * sum(b - a for a, b in self.windows), 2)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and, by first touch, its pinned frame buffers) to the NUMA node the GPU hangs off:
    with 8 ranks each pulling 133 MB per step over PCIe, remote-socket host memory is the first e2e bottleneck."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def cpu_frames(n):
    from pilotguru_b200 import synth
    fr = np.stack([synth.frame(t) for t in range(n)])
    fl = np.array([synth.flow(t) if t > 0 else (0, 0) for t in range(n)], np.float32)
    return fr, fl


def cpu_run(n_frames, threads, frames=None, flows=None):
    """The oracle (CPU port of the reference path, `-O3 -march=native` build made on this host) on `threads` host threads
    over n_frames synthetic frames: extraction of every frame, then the match of EVERY consecutive pair.  Returns
    (seconds, keypoints per frame, matches per frame [entry 0 = -1], build description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    build = O.native_lib()[1]
    if frames is None:
        frames, flows = cpu_frames(min(n_frames, 16))
    distinct = len(frames)
    fr, fl = frames, flows
    if n_frames > distinct:
        # ping-pong over the distinct frames (0..d-1, d-2..0, 1..): consecutive frames stay consecutive, so every pair is
        # an ordinary small-flow pair; going backwards the flow is minus the forward flow of the later frame
        idx, sign = [], []
        i, step = 0, 1
        for _ in range(n_frames):
            idx.append(i); sign.append(step)
            if i + step < 0 or i + step >= distinct:
                step = -step
            i += step
        idx = np.array(idx)
        fr = frames[idx]
        fl = np.zeros((n_frames, 2), np.float32)
        for k in range(1, n_frames):
            a, b = idx[k - 1], idx[k]
            fl[k] = flows[b] if b > a else (-flows[a] if b < a else 0)
    else:
        fr, fl = frames[:n_frames], flows[:n_frames]
    O.bench_extract_match(fr[:min(threads, n_frames)], fl[:min(threads, n_frames)], threads, native=True)  # warm: page in, spin threads
    sec, _, _, nk, nm = O.bench_extract_match(fr, fl, threads, per_frame=True, native=True)
    return sec, nk, nm, build


def cv2_orb_baseline(frames, threads):
    """A second CPU data point: OpenCV's OWN optimised ORB + brute-force Hamming matcher on the same frames --
    cv2.ORB_create(1000, 1.2, 8).detectAndCompute per frame, cv2.BFMatcher(NORM_HAMMING).match per consecutive pair, one
    frame per host thread (cv2.setNumThreads(1)).  NOT the reference's algorithm (no 30-px cell grid with threshold retry,
    no octree distribution, no projection windows): it bounds from above what a SIMD implementation of a comparable
    pipeline does on these cores, so the GPU/CPU ratio against the scalar port has an honest companion.  None if cv2 is absent."""
    try:
        import cv2
    except Exception:
        return None
    import concurrent.futures as cf
    cv2.setNumThreads(1)
    n = len(frames)

    def work(rng):
        orb = cv2.ORB_create(nfeatures=NFEAT, scaleFactor=1.2, nlevels=8)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        prev = None
        for i in rng:
            k, d = orb.detectAndCompute(frames[i], None)
            if prev is not None and d is not None and prev is not None:
                bf.match(prev, d)
            prev = d
        return len(rng)

    chunks = [range(a, min(a + (n + threads - 1) // threads, n)) for a in range(0, n, (n + threads - 1) // threads)]
    work(range(0, min(2, n)))                                             # warm
    t0 = time.time()
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, chunks))
    dt = time.time() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": threads, "kind": "cv2.ORB + BFMatcher (OpenCV %s), a different, SIMD-optimised pipeline" % cv2.__version__,
            "sample": f"{n} synthetic 1080p frames, {dt:.1f} s wall"}


def calibration_cpu_baseline(d, hz, cores, budget_s=20.0):
    """CPU baseline of the calibration leg: the oracle's window loop (fit_motion.cc:156-293) on all host cores -- one
    independent 60-s slice of the recording per thread (12 windows of 500 L-BFGS iterations each) -- in the LITERAL
    sequential restatement of velocity.cc (what the reference computes) and in the contract arithmetic the GPU runs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import concurrent.futures as cf
    import oracle_lib as O
    n_imu = int(60 * hz) + 1

    def slice_(k):
        i0 = k * n_imu
        t0, t1 = d["gyro_t"][i0], d["gyro_t"][i0 + n_imu - 1]
        g = (d["gps_t"] >= t0) & (d["gps_t"] <= t1)
        return dict(gyro=d["gyro"][i0:i0 + n_imu], gyro_t=d["gyro_t"][i0:i0 + n_imu], acc=d["acc"][i0:i0 + n_imu],
                    acc_t=d["acc_t"][i0:i0 + n_imu], gps_v=d["gps_v"][g], gps_t=d["gps_t"][g])

    n_slices = min(cores, (len(d["gyro_t"]) - 1) // n_imu)
    out = {}
    for mode, name in ((0, "literal"), (1, "contract")):
        t0 = time.time()
        with cf.ThreadPoolExecutor(n_slices) as ex:
            res = list(ex.map(lambda k: O.fit_motion(slice_(k), mode=mode), range(n_slices)))
        dt = time.time() - t0
        out[name] = sum(len(r["iters"]) for r in res) / dt
        if dt > budget_s:
            pass
    return {"value": out["literal"], "unit": "windows/s", "cores": n_slices, "kind": "port",
            "contract_arithmetic_windows_per_s": out["contract"],
            "sample": f"{n_slices} slices of 60 s @ {hz:.0f} Hz (12 windows x 500 L-BFGS iterations each), one per host thread, oracle/liboracle.so"}


def calibration_leg(device, seconds, hz, cpu=True):
    """BASELINE configs[3]: fit_motion's velocity calibration over `seconds` of `hz` IMU + 1 Hz GPS (all sliding
    windows, 500 L-BFGS iterations each) through the C-ABI, host arrays in, host arrays out (secondary metric)."""
    import ctypes as C
    import torch
    from pilotguru_b200 import calibration as cal, synth
    from pilotguru_b200._lib import check, lib
    d = synth.imu_gps(seconds, hz)
    t0 = time.time()
    imu = cal.ImuSeries(d["gyro"], d["gyro_t"], d["acc"], d["acc_t"], device=device)
    t_up = time.time() - t0
    cal.fit_windows(imu, d["gps_v"][:200], d["gps_t"][:200])                # warm-up (module load, allocations)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        t0 = time.time()
        r = cal.fit_windows(imu, d["gps_v"], d["gps_t"])
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    sweep, solve, speeds = C.c_float(), C.c_float(), C.c_float()
    n_iv = C.c_int64()
    check(lib().pgb_imu_last_kernel_ms(imu._h, C.byref(sweep), C.byref(solve), C.byref(speeds), C.byref(n_iv)))
    n_win = len(r["iters"])
    covered = int((r["speed_cnt"] > 0).sum())
    per_window = min(40, len(d["gps_v"])) - 1
    intervals = n_win * per_window * hz                                    # IMU intervals a per-window sweep would touch per evaluation
    imu.close()
    peak, peak_src = measured_peaks()
    sweep_bytes = 64 * n_iv.value                                          # two {x,y,z,t} 32-B records per IMU interval
    out = {"workload": f"fit_motion {seconds:.0f} s @ {hz:.0f} Hz IMU + 1 Hz GPS, window 40 / step 5, 500 L-BFGS iterations",
           "windows": n_win, "windows_per_s": n_win / best, "seconds_per_fit": best, "upload_s": t_up,
           "imu_events_covered": covered, "lbfgs_iterations_total": int(np.abs(r["iters"]).sum()),
           "imu_intervals_per_window_pass": intervals, "dtype": "f64",
           "kernel_ms": {"k_imu_sweep": sweep.value, "k_imu_solve": solve.value, "k_imu_speeds": speeds.value},
           # every IMU sub-interval is swept ONCE per fit (the per-GPS-interval records are window-independent), not once per
           # window per evaluation as in the reference: the kernel is a latency-bound fp64 recurrence, far from the HBM roofline
           "roofline": {"kernel": "k_imu_sweep", "bound": "hbm", "achieved": sweep_bytes / (sweep.value * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": sweep_bytes / (sweep.value * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                        "algorithmic_bytes_per_launch": sweep_bytes, "imu_intervals": int(n_iv.value),
                        "note": "64 B per IMU sub-interval (SURVEY 8d), each swept once; one thread per GPS interval walks ~500 dependent "
                                "fp64 quaternion/matrix steps: latency-bound, not bandwidth-bound"}}
    if cpu:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        out["cpu_baseline"] = calibration_cpu_baseline(d, hz, cores)
    return out


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = max(32, min(8 * cores, 256))  # ~10-20 s of CPU work per step at ~14 frames/s/core
    frames, flows = cpu_frames(16)
    build = ""
    for _ in range(min(args.warmup, 1)):
        cpu_run(per_step, cores, frames, flows)
    tot_s, tot_f = 0.0, 0
    for _ in range(args.steps):
        s, _, _, build = cpu_run(per_step, cores, frames, flows)
        tot_s += s; tot_f += per_step
    v = tot_f / tot_s
    line = {"impl": "reference", "metric": "1080p frames/sec ORB extract+match", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            # the same keys as the GPU arm's config (the workload is the same; a step here is a bounded sample of it)
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": per_step, "global_frames_per_step": per_step,
                       "l2": "n/a (CPU arm)", "schedule": f"one contiguous block of frames per host thread, {cores} threads",
                       "host_numa_node_rank0": None, "parallelism": "host cores only"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "build": build,
                             "sample": f"{per_step} frames per step x {args.steps} steps (extract every frame + match every consecutive pair), oracle on {cores} host threads"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line goes to the process's original stdout; everything else that writes to fd 1 (NCCL prints its
    version banner there) was diverted to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--one-extractor", action="store_true",
                    help="A/B: all resident steps on ONE extractor handle (default: consecutive steps alternate over two handles / streams)")
    ap.add_argument("--no-calibration", action="store_true", help="skip the fit_motion (BASELINE configs[3]) leg")
    ap.add_argument("--calib-seconds", type=float, default=3600.0)
    ap.add_argument("--calib-hz", type=float, default=500.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pilotguru_b200 import launch_count, synth
    from pilotguru_b200.dist import BoundaryExchange, FeatureExchange, PgbComm
    from pilotguru_b200.matcher import ORBmatcher
    from pilotguru_b200.orb import ORBextractor

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    # ---- synthetic input: B distinct frames per rank (rank r holds frames r*B .. r*B+B-1 of the sequence)
    t0 = rank * B
    frames_np = np.stack([synth.frame(t0 + i) for i in range(B)])
    host_frames = torch.from_numpy(frames_np).pin_memory()
    flows_np = np.array([synth.flow(t0 + i) for i in range(B)], np.float32)  # flow into frame i from frame i-1
    dev_frames = host_frames.cuda(non_blocking=False)

    ex = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B, device=local)
    cap = ex.cap
    stream = torch.cuda.ExternalStream(ex.stream)
    mt = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, device=local, stream=ex.stream)
    xch = FeatureExchange(1, 0, B, cap, device=torch.device("cuda", local))       # the rank's feature region (slot 0 = predecessor)
    # N > 1: the C-ABI communicator (NCCL inside libpgb200.so) and a side stream on which the boundary exchange and the
    # one pair that depends on it run while the main stream matches the other B-1 pairs
    comm = PgbComm(local, rank, world) if world > 1 else None
    side = torch.cuda.Stream() if world > 1 else None
    mt_side = ORBmatcher(0.9, True, max_feats=cap, max_batch=1, device=local, stream=side.cuda_stream) if world > 1 else None
    bx = BoundaryExchange(comm, xch) if world > 1 else None
    ev_feat, ev_x = torch.cuda.Event(), torch.cuda.Event()
    flow_dev = torch.from_numpy(flows_np).cuda()
    match = torch.full((B, cap), -1, dtype=torch.int32, device="cuda")
    nmatch = torch.zeros(B, dtype=torch.int32, device="cuda")
    sf = ex.GetScaleFactors()
    # slot 0 of the exchange buffer = the block's predecessor frame t0-1 (rank r>0 also receives it every step from
    # its left neighbour through the all-gather; rank 0 keeps this one)
    pred = torch.from_numpy(synth.frame(t0 - 1)[None].copy()).cuda()
    torch.cuda.synchronize()
    ex.extract_ptr(pred.data_ptr(), 3, 1, W, H, W, W * H, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0), cap)
    ex.check()

    def step(frames_ptr, where):
        """extract B frames into slots 1..B of the rank's feature region, exchange the block boundary, match B pairs."""
        ex.extract_ptr(frames_ptr, where, B, W, H, W, W * H, xch.kps_ptr(1), xch.desc_ptr(1), xch.counts_ptr(1), cap)
        if world == 1:
            mt.match_consecutive_ptr(B, cap, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0), flow_dev.data_ptr(),
                                     float(W), float(H), 15.0, sf, match.data_ptr(), nmatch.data_ptr())
            return
        # side stream: pack my last frame -> ONE NCCL all-gather of the boundary records (pgb_allgather_feats) -> unpack the
        # left neighbour's into slot 0 -> match pair 0 (the only pair that needs it).  Main stream: pairs 1..B-1 meanwhile.
        ev_feat.record(stream)
        side.wait_event(ev_feat)
        bx.issue(side.cuda_stream)
        mt_side.match_consecutive_ptr(1, cap, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0), flow_dev.data_ptr(),
                                      float(W), float(H), 15.0, sf, match.data_ptr(), nmatch.data_ptr())
        ev_x.record(side)
        mt.match_consecutive_ptr(B - 1, cap, xch.kps_ptr(1), xch.desc_ptr(1), xch.counts_ptr(1), flow_dev.data_ptr() + 8,
                                 float(W), float(H), 15.0, sf, match.data_ptr() + 4 * cap, nmatch.data_ptr() + 4)
        stream.wait_event(ev_x)

    # ---- the timed resident step: extraction on the extractor's stream, matching on a second stream, two feature regions.
    # The matcher's kernels are latency-bound (a few hundred CTAs with sequential sections: 1.4 us/frame when nothing runs
    # beside them); on their own stream they run under the NEXT batch's extraction kernels, which is how a streaming caller
    # would drive the two handles.  Step k extracts into region k & 1 and waits for the matcher to be done with that region
    # (step k - 2); at N > 1 the boundary exchange (pack, ONE all-gather, unpack) sits on the matcher stream in front of the
    # match.  Results are identical to the serial step (checked below).
    mstream = torch.cuda.Stream(priority=-1)   # the latency-bound matcher kernels get their CTAs scheduled ahead of the throughput kernels they run under
    mtM = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, device=local, stream=mstream.cuda_stream)
    xchB = FeatureExchange(1, 0, B, cap, device=torch.device("cuda", local))
    ex.extract_ptr(pred.data_ptr(), 3, 1, W, H, W, W * H, xchB.kps_ptr(0), xchB.desc_ptr(0), xchB.counts_ptr(0), cap)
    ex.check()
    regions = [xch, xchB]
    bxs = [BoundaryExchange(comm, r_) for r_ in regions] if world > 1 else None
    matchP = [torch.full((B, cap), -1, dtype=torch.int32, device="cuda") for _ in range(2)]
    nmatchP = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in range(2)]
    ev_featP = [torch.cuda.Event(), torch.cuda.Event()]
    ev_doneP = [torch.cuda.Event(), torch.cuda.Event()]
    pipe_k = [0]

    exP = [ex, None]                                                         # --two-extractors: a second extractor handle (own stream, own scratch)
    streamP = [stream, None]

    def step_pipelined(frames_ptr, where):
        r = pipe_k[0] & 1
        pipe_k[0] += 1
        x = regions[r]
        if exP[1] is not None:                                               # extraction of consecutive steps alternates over two handles / streams
            e_, s_ = exP[r], streamP[r]
            s_.wait_event(ev_doneP[r])
            with torch.cuda.stream(s_):
                e_.extract_ptr(frames_ptr, where, B, W, H, W, W * H, x.kps_ptr(1), x.desc_ptr(1), x.counts_ptr(1), cap)
                ev_featP[r].record(s_)
            mstream.wait_event(ev_featP[r])
            if world > 1:
                bxs[r].issue(mstream.cuda_stream)
            mtM.match_consecutive_ptr(B, cap, x.kps_ptr(0), x.desc_ptr(0), x.counts_ptr(0), flow_dev.data_ptr(),
                                      float(W), float(H), 15.0, sf, matchP[r].data_ptr(), nmatchP[r].data_ptr())
            ev_doneP[r].record(mstream)
            return
        stream.wait_event(ev_doneP[r])                                       # the matcher has finished with this region (two steps ago)
        ex.extract_ptr(frames_ptr, where, B, W, H, W, W * H, x.kps_ptr(1), x.desc_ptr(1), x.counts_ptr(1), cap)
        ev_featP[r].record(stream)
        mstream.wait_event(ev_featP[r])
        if world > 1:
            bxs[r].issue(mstream.cuda_stream)                                # boundary record all-gather (C-ABI, NCCL)
        mtM.match_consecutive_ptr(B, cap, x.kps_ptr(0), x.desc_ptr(0), x.counts_ptr(0), flow_dev.data_ptr(),
                                  float(W), float(H), 15.0, sf, matchP[r].data_ptr(), nmatchP[r].data_ptr())
        ev_doneP[r].record(mstream)

    def drain_pipeline():
        stream.wait_event(ev_doneP[0]); stream.wait_event(ev_doneP[1])

    def timed(fn, n, drain=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        if drain:
            drain()                                                          # the timed region ends when the last match has finished
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.cuda.stream(stream):
        dev_step = lambda: step(dev_frames.data_ptr(), ORBextractor.IN_DEVICE | ORBextractor.OUT_DEVICE)
        pipe_step = lambda: step_pipelined(dev_frames.data_ptr(), ORBextractor.IN_DEVICE | ORBextractor.OUT_DEVICE)

        class E2ESet:
            """Everything one in-flight e2e step owns: extractor + matcher handles (one CUDA stream), the exchange
            region, pinned result buffers.  Two sets alternate so that the tail of step k (last chunk's kernels,
            matcher, D2H) overlaps the H2D of step k+1 -- the double buffering any streaming caller would use."""

            def __init__(self, ex_, mt_, xch_, comm_):
                self.ex, self.mt, self.xch = ex_, mt_, xch_
                self.bx = BoundaryExchange(comm_, xch_) if world > 1 else None   # each in-flight set has its own communicator
                self.stream = torch.cuda.ExternalStream(ex_.stream)
                self.d2h = torch.cuda.Stream()
                self.ev_feat = torch.cuda.Event()
                self.match = torch.full((B, cap), -1, dtype=torch.int32, device="cuda")
                self.nmatch = torch.zeros(B, dtype=torch.int32, device="cuda")
                self.h_counts = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
                self.h_nmatch = torch.zeros(B, dtype=torch.int32).pin_memory()
                self.h_kps = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory()
                self.h_desc = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
                self.h_match = torch.zeros((B, cap), dtype=torch.int32).pin_memory()

            def issue(self):
                x = self.xch
                with torch.cuda.stream(self.stream):
                    self.ex.extract_ptr(host_frames.data_ptr(), ORBextractor.OUT_DEVICE, B, W, H, W, W * H, x.kps_ptr(1),
                                        x.desc_ptr(1), x.counts_ptr(1), cap)   # H2D of the frames inside the call
                    if self.bx:
                        self.bx.issue(self.stream.cuda_stream)                 # boundary record all-gather (C-ABI, NCCL)
                    self.ev_feat.record(self.stream)
                    with torch.cuda.stream(self.d2h):                          # D2H of the features while the matcher runs
                        self.d2h.wait_event(self.ev_feat)
                        self.h_counts.copy_(x.counts_view(), non_blocking=True)
                        self.h_kps.copy_(x.kps_view()[1:], non_blocking=True)
                        self.h_desc.copy_(x.desc_view()[1:], non_blocking=True)
                    self.mt.match_consecutive_ptr(B, cap, x.kps_ptr(0), x.desc_ptr(0), x.counts_ptr(0), flow_dev.data_ptr(),
                                                  float(W), float(H), 15.0, sf, self.match.data_ptr(), self.nmatch.data_ptr())
                    self.h_match.copy_(self.match, non_blocking=True)          # D2H of the matches
                    self.h_nmatch.copy_(self.nmatch, non_blocking=True)

            def wait(self):                                                    # the caller holds every result of the step
                self.d2h.synchronize()
                self.stream.synchronize()

        ex2 = ORBextractor(NFEAT, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B, device=local)
        mt2 = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, device=local, stream=ex2.stream)
        xch2 = FeatureExchange(1, 0, B, cap, device=torch.device("cuda", local))
        ex2.extract_ptr(pred.data_ptr(), 3, 1, W, H, W, W * H, xch2.kps_ptr(0), xch2.desc_ptr(0), xch2.counts_ptr(0), cap)
        ex2.check()
        comm2 = PgbComm(local, rank, world) if world > 1 else None
        sets = [E2ESet(ex, mt, xch, comm), E2ESet(ex2, mt2, xch2, comm2)]

        def e2e_run(n):
            """n e2e steps, two in flight; returns after the results of all of them are on the host."""
            for k in range(n):
                sets[k & 1].issue()
                if k:
                    sets[(k - 1) & 1].wait()
            sets[(n - 1) & 1].wait()

        def timed_e2e(n):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            e2e_run(n)
            torch.cuda.synchronize()
            e1.record(stream)
            e1.synchronize()
            ms_ = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
            return float(ms_.item())

        sampler = ClockSampler(local) if rank == 0 else None              # started before the warm-up: ready when timing starts
        for _ in range(Wm):
            dev_step()
        ex.check()
        ms_serial = timed(dev_step, K)                                       # everything on one stream (round 1's step), for context
        ex.check()
        nm_dev = nmatch.cpu().numpy().copy(); cnt_dev = xch.counts_view().cpu().numpy().copy()
        if not args.one_extractor:
            exP[1] = ex2; streamP[1] = torch.cuda.ExternalStream(ex2.stream)
        for _ in range(Wm):
            pipe_step()
        drain_pipeline()
        ex.check()
        l0 = launch_count()
        if sampler:
            sampler.wait_ready()
            sampler.begin()
        ms = timed(pipe_step, K, drain_pipeline)
        if sampler:
            sampler.end()
        launches = launch_count() - l0
        ex.check(); torch.cuda.synchronize()
        for r_ in range(2):                                                  # the pipelined steps produced what the serial step produces
            assert np.array_equal(nmatchP[r_].cpu().numpy(), nm_dev) and np.array_equal(regions[r_].counts_view().cpu().numpy(), cnt_dev) \
                and torch.equal(matchP[r_], match), "pipelined and serial resident steps disagree"
        cand_total = sum(len(ex.candidates(l, frame=0)) for l in range(8)) * B   # candidates handed to the octree (frame 0 x B)
        # ---- N > 1: rank r's pair 0 (its first frame against the LEFT NEIGHBOUR's last frame, received through the
        # exchange) must equal a local recomputation with the true predecessor frame t0 - 1 extracted here
        parity = {"checked": False}
        if world > 1:
            got_m = match[0].clone(); got_n = nmatch[0:1].clone()
            ex.extract_ptr(pred.data_ptr(), 3, 1, W, H, W, W * H, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0), cap)
            m1 = torch.full((1, cap), -1, dtype=torch.int32, device="cuda"); n1 = torch.zeros(1, dtype=torch.int32, device="cuda")
            mt.match_consecutive_ptr(1, cap, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0), flow_dev.data_ptr(),
                                     float(W), float(H), 15.0, sf, m1.data_ptr(), n1.data_ptr())
            ex.check()
            ok = torch.tensor([int(torch.equal(m1[0], got_m) and torch.equal(n1, got_n) and int(n1.item()) >= 20)], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            parity = {"checked": True, "cross_rank_pair0_equals_local_recompute": bool(ok.item())}
            if not ok.item():
                raise SystemExit("bench.py: a rank's boundary pair differs from the local recomputation with the true predecessor")
            dev_step()                                                      # restore the exchanged state
            torch.cuda.synchronize()

        e2e_run(3)
        if sampler:
            sampler.begin()
        ms_e2e = timed_e2e(K)
        if sampler:
            sampler.end()
        clocks = sampler.stop() if sampler else None
        ex.check(); ex2.check()
        # context for the e2e number: the bare host->device transfer of one step's frames (pinned, same stream)
        ms_h2d = timed(lambda: dev_frames.copy_(host_frames, non_blocking=True), 5) / 5
        for st_ in sets:
            assert np.array_equal(st_.h_nmatch.numpy(), nm_dev) and np.array_equal(st_.h_counts.numpy(), cnt_dev), \
                "e2e (host frames) and device-resident runs disagree"

        # ---- per-stage times and the FAST kernel roofline (same resident batch, events on the launching stream)
        stage_us = {}
        for name, which in list(STAGES.items()) + [("unfused_fast_score_plus_cell_nms", 2)]:
            for _ in range(2):
                ex.run_stage(which)
            reps = 10
            t = timed(lambda: ex.run_stage(which), reps)
            stage_us[name] = 1e3 * t / reps / B
        ex.check()
        unfused_us = stage_us.pop("unfused_fast_score_plus_cell_nms")
        match_fn = lambda: mt.match_consecutive_ptr(B, cap, xch.kps_ptr(0), xch.desc_ptr(0), xch.counts_ptr(0),
                                                    flow_dev.data_ptr(), float(W), float(H), 15.0, sf,
                                                    match.data_ptr(), nmatch.data_ptr())
        match_fn()
        stage_us["match"] = 1e3 * timed(match_fn, 10) / 10 / B
    frames_total = B * world
    value = frames_total * K / (ms * 1e-3)
    e2e = frames_total * K / (ms_e2e * 1e-3)
    peak, peak_src = measured_peaks()
    fast_s = stage_us["fast_cells"] * 1e-6 * B
    achieved = FAST_K2_BYTES_PER_FRAME * B / fast_s / 1e9
    cand_per_frame = float(cand_total) / B
    fused_bytes = PYRAMID_PX + 4 * (N_CELLS + cand_per_frame)
    traffic_pf, traffic_src = ncu_traffic()
    h2d = int(host_frames.numel())
    s0 = sets[0]
    d2h = int(s0.h_counts.numel() * 4 + s0.h_kps.numel() * 4 + s0.h_desc.numel() + s0.h_match.numel() * 4 + s0.h_nmatch.numel() * 4)
    # Teardown in dependency order: torch's pinned-host allocator records an event on every stream a block was used
    # on when the block is freed, so the pinned tensors must go before the handles that own those streams.
    torch.cuda.synchronize()
    for st_ in sets:
        st_.h_counts = st_.h_nmatch = st_.h_kps = st_.h_desc = st_.h_match = None
    # BASELINE configs[1] as the reference calls it: ONE 1080p frame through ORBextractor::operator() with host arrays in and out
    # (upload, seven pyramid launches, FAST, octree, descriptors, download of keypoints + descriptors); wall clock, median of 20
    single_us = None
    if rank == 0:
        one = np.ascontiguousarray(frames_np[0])
        lat = []
        for i in range(23):
            t0_ = time.perf_counter()
            k1_, d1_ = ex(one)
            lat.append(time.perf_counter() - t0_)
        single_us = 1e6 * float(np.median(lat[3:]))
        assert len(k1_) == int(cnt_dev[1]), "single-frame call and batched call disagree on frame 0"
    s0 = st_ = host_frames = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    sets.clear()
    for h_ in (mtM, mt2, ex2, mt, ex) + ((mt_side,) if mt_side else ()):
        h_.close()
    if comm:
        torch.cuda.synchronize()
        comm2.close()
        comm.close()

    if rank == 0:
        line = {"metric": "1080p frames/sec ORB extract+match", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": ms / K, "ms_per_step_single_stream": ms_serial / K,
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "frames_per_gpu_per_step": B, "global_frames_per_step": frames_total,
                           "l2": f"inputs larger than L2: {B * W * H / 1e6:.0f} MB of frames + {B * 6.4:.0f} MB pyramid per step",
                           "schedule": ("two steps in flight: consecutive steps alternate over two extractor handles (own streams and scratch), "
                                        "matching (and at N > 1 the boundary all-gather) on a third stream with two feature regions -- "
                                        "the latency-bound kernels of one step (octree, matcher) run under the other step's "
                                        "throughput-bound ones (the single-stream time of the same step is reported as ms_per_step_single_stream)")
                                       if not args.one_extractor else "one extractor handle; matcher on a second stream",
                           "host_numa_node_rank0": numa,
                           "parallelism": (f"frames sharded over {world} GPU(s); one NCCL all-gather (pgb_allgather_feats, C-ABI) of the "
                                           f"block-boundary keypoint/descriptor records per step, overlapped with the matcher") if world > 1 else "1 GPU"},
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / K, "steps_in_flight": 2, "h2d_only_ms_per_step": ms_h2d,
                        "h2d_only_gbs": h2d / (ms_h2d * 1e-3) / 1e9,
                        # every rank copying its pinned frames at the same time, ONE plain cudaMemcpyAsync each, no kernels
                        # running: what the box's host side can deliver (per rank / all ranks together)
                        "h2d_box_limit_gbs": h2d / (ms_h2d * 1e-3) / 1e9, "h2d_box_limit_gbs_all_ranks": world * h2d / (ms_h2d * 1e-3) / 1e9},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": "k_fast_cells2 (FAST-9 score + per-cell NMS, iniThFAST pass then minThFAST for the empty cells, fused; 4-band + 5-band launches)", "bound": "hbm",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic_pf * B if traffic_pf else None, "peak_source": peak_src,
                             "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": FAST_K2_BYTES_PER_FRAME * B,
                             "algorithmic_bytes_note": "SURVEY 8(d) K2 figure (pyramid read + u8 score map written = 12,838,642 B/frame): "
                                                       "the HBM work of the unfused score kernel this launch replaces; round 1 was quoted on it",
                             "fused_formulation": {"algorithmic_bytes_per_launch": fused_bytes * B,
                                                   "achieved": fused_bytes * B / fast_s / 1e9,
                                                   "frac": fused_bytes * B / fast_s / 1e9 / peak,
                                                   "note": "SURVEY 8(d) fused figure: pyramid read once + 4 B per candidate slot and per cell count; "
                                                           "the score map is never written"},
                             "us_per_launch": stage_us["fast_cells"] * B,
                             "unfused_pair_us_per_frame": unfused_us},
                "stage_us_per_frame": stage_us,
                "keypoints_per_frame": float(cnt_dev[1:].mean()), "matches_per_frame": float(nm_dev.mean()),
                "single_frame_call_us": single_us}
        line["parity_checked"] = parity["checked"]
        line["parity"] = parity
        if not args.no_cpu_baseline and world == 1:
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            n = max(64, min(16 * cores, 512))                              # ~15-30 s of CPU work (oracle: ~14 frames/s/core)
            # the CPU leg runs the SAME frames the GPU leg held (frames 0..B-1, then ping-pong): per-frame keypoint counts
            # and per-pair match counts of the first B frames are asserted against the GPU's, outside the timed regions
            sec, nk, nm, build = cpu_run(n, cores, frames_np, flows_np)
            m = min(n, B)
            same = bool(np.array_equal(nk[:m], cnt_dev[1:m + 1]) and np.array_equal(nm[1:m], nm_dev[1:m]))
            line["parity"].update({"checked": True, "frames_compared_with_cpu_leg": m, "keypoint_counts_equal": bool(np.array_equal(nk[:m], cnt_dev[1:m + 1])),
                                   "match_counts_equal": bool(np.array_equal(nm[1:m], nm_dev[1:m]))})
            line["parity_checked"] = True
            if not same:
                raise SystemExit(f"bench.py: GPU and CPU legs disagree: keypoints {cnt_dev[1:m + 1][:8]} vs {nk[:8]}, matches {nm_dev[1:m][:8]} vs {nm[1:m][:8]}")
            line["cpu_baseline_cv2_orb"] = cv2_orb_baseline(frames_np[:min(len(frames_np), 8 * cores)], cores)
            line["cpu_baseline"] = {"value": n / sec, "unit": "frames/s", "cores": cores, "kind": "port", "build": build,
                                    "sample": f"{n} synthetic 1080p frames (extract every frame + match every consecutive pair), oracle on {cores} host threads ({sec:.1f} s wall)"}
        if not args.no_calibration and world == 1:
            line["calibration"] = calibration_leg(local, args.calib_seconds, args.calib_hz, cpu=not args.no_cpu_baseline)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
