#!/usr/bin/env python3
"""Generates tests/golden/*.npz: outputs of the OpenCV primitives the reference calls (cv2 4.13 in the build
container) on small seeded inputs, plus whole-pipeline outputs of the pinned oracle on seeded synthetic frames.
Run from the repo root in the build container:  python tests/golden/make_golden.py
The .npz files are committed; tests compare the oracle (CPU) and the CUDA path (GPU box) against them."""
import os, sys
import numpy as np
import cv2
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from pilotguru_b200 import synth

rng = np.random.default_rng(20260101)
out = {}
# --- cv2 primitives
img = synth.frame(2, w=320, h=240)
noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
out["img"] = img; out["noise"] = noise
out["resize_img_267x200"] = cv2.resize(img, (267, 200), interpolation=cv2.INTER_LINEAR)
out["resize_noise_109x81"] = cv2.resize(noise, (109, 81), interpolation=cv2.INTER_LINEAR)
for name, im in (("img", img), ("noise", noise)):
    for th in (7, 20):
        det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kp = det.detect(im)
        out[f"fast_{name}_th{th}"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kp], np.int32).reshape(-1, 3)
    out[f"blur_{name}"] = cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
yy = rng.integers(-70000, 70000, 4000).astype(np.float32); xx = rng.integers(-70000, 70000, 4000).astype(np.float32)
yy[:4] = [0, 0, 1, -1]; xx[:4] = [0, 1, 0, 0]
out["atan_y"] = yy; out["atan_x"] = xx
out["atan_deg"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(yy, xx)], np.float32)
orb = cv2.ORB_create(nfeatures=400, nlevels=1, edgeThreshold=31, patchSize=31, fastThreshold=20)
kp = orb.detect(img, None)
kp, d = orb.compute(img, kp)
out["orb_xy"] = np.array([[int(round(k.pt[0])), int(round(k.pt[1]))] for k in kp], np.int32)
out["orb_angle"] = np.array([k.angle for k in kp], np.float32)
out["orb_desc"] = d
g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
out["orb_blurred_float_path"] = cv2.sepFilter2D(img, cv2.CV_8U, g, g, borderType=cv2.BORDER_REFLECT_101)
np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)

# --- whole pipeline (oracle, pinned above) on seeded frames: the fixtures the GPU path is compared with
pipe = {}
orc = O.OrbOracle(500, 1.2, 8, 20, 7)
frames = [synth.frame(t, w=640, h=480) for t in range(3)]
sf = orc.tables()["scale"]
prev = None
for t, f in enumerate(frames):
    k, d = orc.extract(f)
    pipe[f"kps{t}"] = k; pipe[f"desc{t}"] = d
    if prev is not None:
        fl = synth.flow(t, w=640, h=480)
        n, m = O.match_consecutive(prev[0], prev[1], k, d, fl, 640, 480, 15.0, sf)
        pipe[f"match{t}"] = m; pipe[f"nmatch{t}"] = np.int32(n)
    prev = (k, d)
np.savez_compressed(os.path.join(HERE, "orb_pipeline_640x480.npz"), **pipe)
print("golden written:", {k: v.shape for k, v in pipe.items() if hasattr(v, "shape")})
