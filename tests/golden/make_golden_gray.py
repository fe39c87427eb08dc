#!/usr/bin/env python3
"""Generates tests/golden/cv2_gray.npz: cv2.cvtColor (cv2 4.13, build container) RGB/BGR/RGBA/BGRA -> GRAY and cv2.flip on a
small seeded image -- the pin for the oracle's restatement of the frame feed (Tracking.cc:243-258,
image_sequence_reader.cc:163-175).  Run from the repo root:  python tests/golden/make_golden_gray.py"""
import os
import numpy as np
import cv2
HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20260103)
rgb = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
rgba = rng.integers(0, 256, (37, 45, 4), dtype=np.uint8)
out = {"rgb": rgb, "rgba": rgba,
       "rgb2gray": cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY), "bgr2gray": cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY),
       "rgba2gray": cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY), "bgra2gray": cv2.cvtColor(rgba, cv2.COLOR_BGRA2GRAY),
       "flip_v_rgb2gray": cv2.cvtColor(cv2.flip(rgb, 0), cv2.COLOR_RGB2GRAY),
       "flip_h_rgb2gray": cv2.cvtColor(cv2.flip(rgb, 1), cv2.COLOR_RGB2GRAY),
       "flip_vh_rgb2gray": cv2.cvtColor(cv2.flip(rgb, -1), cv2.COLOR_RGB2GRAY)}
np.savez_compressed(os.path.join(HERE, "cv2_gray.npz"), **out)
print("wrote cv2_gray.npz", cv2.__version__)
