#!/usr/bin/env python3
"""Generates tests/golden/cv2_pca.npz: cv2.PCACompute2 (cv2 4.13, build container) on seeded n x 3 matrices --
the pin for the oracle's restatement of cv::PCA as rotation.cc:55 and track_image_sequence.cc:27-28 call it.
Run from the repo root:  python tests/golden/make_golden_pca.py"""
import os
import numpy as np
import cv2
HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20260102)
out = {}
for k in range(24):
    n = int(rng.integers(3, 300))
    rows = rng.normal(size=(n, 3)) * rng.uniform(0.01, 3, size=3) * 10.0 ** rng.uniform(-5, 1) + rng.normal(size=3)
    if k % 4 == 0:
        rows[:, 2] = 0.5 * rows[:, 0] + 1e-9 * rng.normal(size=n)   # nearly planar (the trajectory case)
    if k % 6 == 1:
        rows[:, 1] = 1e-4 * rng.normal(size=n)                       # one dominant axis (the steering case)
    mean, vec, val = cv2.PCACompute2(rows, mean=None)
    out[f"rows{k}"] = rows; out[f"vec{k}"] = vec; out[f"val{k}"] = val.ravel(); out[f"mean{k}"] = mean.ravel()
np.savez_compressed(os.path.join(HERE, "cv2_pca.npz"), **out)
print("wrote cv2_pca.npz", cv2.__version__)
