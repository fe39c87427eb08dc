#!/usr/bin/env python3
"""Calibration leg (BASELINE configs[3]: 1 h @ 500 Hz IMU + 1 Hz GPS, 720 windows x 500 L-BFGS iterations) by itself:
   PGB_IMU_TIMING=1 python tools/gpu_calib_profile.py            -> host/kernel phase breakdown on stderr
   ncu --metrics gpu__time_duration.sum ... python tools/gpu_calib_profile.py --once   -> launch list of the k_imu_* kernels
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3600.0)
ap.add_argument("--hz", type=float, default=500.0)
ap.add_argument("--once", action="store_true")
ap.add_argument("--iters", type=int, default=500)
args = ap.parse_args()

from pilotguru_b200 import calibration as cal, synth  # noqa: E402

d = synth.imu_gps(args.seconds, args.hz)
imu = cal.ImuSeries(d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
if not args.once:
    cal.fit_windows(imu, d["gps_v"][:200], d["gps_t"][:200], max_iterations=args.iters)
for rep in range(1 if args.once else 3):
    t0 = time.time()
    r = cal.fit_windows(imu, d["gps_v"], d["gps_t"], max_iterations=args.iters)
    print(f"fit {rep}: {1e3 * (time.time() - t0):.2f} ms, {len(r['iters'])} windows, {int(abs(r['iters']).sum())} iterations",
          file=sys.stderr)
imu.close()
