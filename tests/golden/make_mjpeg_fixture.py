#!/usr/bin/env python3
"""Writes tests/golden/mjpeg_256x192.avi: 5 synthetic colour frames (the survey's generator, three shifted copies as the
channels) as a Motion-JPEG AVI through cv2.VideoWriter (OpenCV's FFmpeg backend, yuvj420p).  ~30 KB.  Run here once; the
file is the fixture of tests/test_video_demux.py (CPU) and tests/test_gpu_video.py."""
import os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pilotguru_b200 import synth

w, h, n = 256, 192, 5
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mjpeg_256x192.avi")
vw = cv2.VideoWriter(out, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (w, h), True)
assert vw.isOpened()
for t in range(n):
    g = synth.frame(t, w=w, h=h)
    vw.write(np.stack([g, np.roll(g, 3, axis=1), 255 - np.roll(g, 5, axis=0)], axis=2))   # BGR
vw.release()
print(out, os.path.getsize(out), "bytes")
