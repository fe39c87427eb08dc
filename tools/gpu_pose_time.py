#!/usr/bin/env python3
"""Wall time of pgb_pose_optimization for a batch of synthetic PnP scenes (host buffers in, host buffers out) and the
largest pose difference against the sequential oracle on a few of them."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import pose_util as U
from pilotguru_b200.optimizer import PoseOptimization

B, N = int(os.environ.get("POSE_BATCH", 256)), 1000
scenes = [U.scene(1000 + i, n=N) for i in range(B)]
T0 = np.stack([s["T0"] for s in scenes]); xy = np.stack([s["xy"] for s in scenes]); oc = np.stack([s["octave"] for s in scenes])
X = np.stack([s["Xw"] for s in scenes]); has = np.stack([s["has"] for s in scenes])
PoseOptimization(T0, xy, oc, X, has, U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)
t0 = time.perf_counter()
for _ in range(5):
    ni, T, out = PoseOptimization(T0, xy, oc, X, has, U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)
dt = (time.perf_counter() - t0) / 5
t1 = time.perf_counter()
worst = 0.0
for i in range(8):
    s = scenes[i]
    on, oT, oout, _ = O.pose_optimization(s["T0"], s["xy"], s["octave"], s["Xw"], s["has"], U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)
    assert on == ni[i] and np.array_equal(oout, out[i])
    worst = max(worst, float(np.abs(oT - T[i]).max()))
cpu = (time.perf_counter() - t1) / 8
print(f"pose_optimization: {B} frames x {N} features ({int(has.sum() / B)} edges/frame): {dt * 1e3:.2f} ms per batch = {dt / B * 1e6:.1f} us/frame "
      f"(host buffers, copies included); oracle 1 thread {cpu * 1e3:.2f} ms/frame; max |T_gpu - T_oracle| = {worst:.2e}")
