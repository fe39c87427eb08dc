"""Synthetic inputs for the BASELINE.json configs.

Frames: a 2400x1400 canvas of random axis-aligned rectangles + noise, blurred,
cropped to 1920x1080 along a ping-pong path so consecutive frames have a known
integer flow.  IMU/GPS: planar-car model.  Pure numpy (no cv2) so the same
generator runs on the GPU box.
"""
from __future__ import annotations

import numpy as np

CANVAS_W, CANVAS_H = 2400, 1400
FRAME_W, FRAME_H = 1920, 1080
N_RECTS = 2430


def _gauss_blur_f32(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(np.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2 * sigma * sigma))
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i in range(2 * r + 1):
        out += k[i] * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i in range(2 * r + 1):
        out2 += k[i] * p[i:i + img.shape[0], :]
    return out2


_canvas_cache: dict = {}


def canvas(seed: int = 1234, w: int = CANVAS_W, h: int = CANVAS_H,
           n_rects: int | None = None) -> np.ndarray:
    key = (seed, w, h, n_rects)
    if key in _canvas_cache:
        return _canvas_cache[key]
    if n_rects is None:
        n_rects = int(round(N_RECTS * (w * h) / (CANVAS_W * CANVAS_H)))
    rng = np.random.default_rng(seed)
    c = np.full((h, w), 128.0, dtype=np.float32)
    for _ in range(n_rects):
        x0 = int(rng.integers(0, w)); y0 = int(rng.integers(0, h))
        rw = int(rng.integers(8, 160)); rh = int(rng.integers(8, 160))
        val = int(rng.integers(20, 236))
        c[y0:min(y0 + rh, h), x0:min(x0 + rw, w)] = val
    c += rng.normal(0.0, 2.0, size=c.shape).astype(np.float32)
    c = _gauss_blur_f32(c, 1.0)
    out = np.clip(np.rint(c), 0, 255).astype(np.uint8)
    _canvas_cache[key] = out
    return out


def _pp(a: int, p: int) -> int:
    return p - abs((a % (2 * p)) - p)


def frame_origin(t: int, w: int = FRAME_W, h: int = FRAME_H,
                 cw: int = CANVAS_W, ch: int = CANVAS_H) -> tuple[int, int]:
    px = cw - w - 32
    py = ch - h - 32
    return 16 + _pp(2 * t, px), 16 + _pp(t, py)


def frame(t: int, seed: int = 1234, w: int = FRAME_W, h: int = FRAME_H,
          cw: int | None = None, ch: int | None = None) -> np.ndarray:
    """Frame t (uint8 HxW).  Flow from frame t-1 to t is -(origin(t)-origin(t-1))."""
    if cw is None:
        cw = w + 480
    if ch is None:
        ch = h + 320
    c = canvas(seed, cw, ch)
    x0, y0 = frame_origin(t, w, h, cw, ch)
    f = c[y0:y0 + h, x0:x0 + w].astype(np.float32)
    f += np.random.default_rng(seed + 1 + t).normal(0.0, 1.0, size=f.shape).astype(np.float32)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def flow(t: int, w: int = FRAME_W, h: int = FRAME_H,
         cw: int | None = None, ch: int | None = None) -> tuple[int, int]:
    """Pixel displacement of scene content from frame t-1 to frame t."""
    if cw is None:
        cw = w + 480
    if ch is None:
        ch = h + 320
    a = frame_origin(t - 1, w, h, cw, ch)
    b = frame_origin(t, w, h, cw, ch)
    return a[0] - b[0], a[1] - b[1]


def imu_gps(duration_s: float, imu_hz: float, seed: int = 11, gps_offset_s: float = 0.5,
            interleaved: bool = False):
    """Planar-car IMU+GPS traces.

    Returns dict with gyro (N,3) f64, gyro_t (N,) i64 usec, acc (N,3), acc_t, gps_v (M,), gps_t (M,).
    """
    rng = np.random.default_rng(seed)
    n = int(round(duration_s * imu_hz)) + 1
    dt_us = int(round(1e6 / imu_hz))
    t_us = np.arange(n, dtype=np.int64) * dt_us
    t = t_us * 1e-6
    speed = 8 + 4 * np.sin(0.15 * t) + 2 * np.sin(0.5 * t)
    dspeed = 4 * 0.15 * np.cos(0.15 * t) + 2 * 0.5 * np.cos(0.5 * t)
    yaw = 0.25 * np.sin(0.2 * t)
    yaw_rate = 0.25 * 0.2 * np.cos(0.2 * t)
    # device frame: x forward, y left, z up; gravity reaction +9.81 on z
    acc = np.stack([dspeed, speed * yaw_rate, np.full(n, 9.81)], axis=1)
    acc += np.array([0.15, -0.1, 0.2])
    acc += rng.normal(0, 0.05, size=acc.shape)
    gyro = np.stack([np.zeros(n), np.zeros(n), yaw_rate], axis=1)
    gyro += rng.normal(0, 0.01, size=gyro.shape)
    gyro[:, :2] += rng.normal(0, 0.002, size=(n, 2))
    acc_t = t_us.copy()
    if interleaved:
        acc_t = acc_t + dt_us // 2
    m = int(np.floor(duration_s - gps_offset_s)) + 1
    gps_t = (np.arange(m, dtype=np.int64) * 1_000_000 + int(gps_offset_s * 1e6))
    gps_t = gps_t[gps_t <= t_us[-1]]
    gts = gps_t * 1e-6
    gps_v = 8 + 4 * np.sin(0.15 * gts) + 2 * np.sin(0.5 * gts) + rng.normal(0, 0.1, size=gts.shape)
    return dict(gyro=np.ascontiguousarray(gyro), gyro_t=t_us, acc=np.ascontiguousarray(acc),
                acc_t=acc_t, gps_v=np.ascontiguousarray(gps_v), gps_t=gps_t)


def write_imu_gps_json(d, out_dir: str):
    """Writes rotations.json / accelerations.json / locations.json in the PilotGuru Recorder format
    (mobile/android/README.md:22-98, field names include/io/json_converters.hpp:10-35) for an imu_gps() dict."""
    import json, os
    os.makedirs(out_dir, exist_ok=True)
    paths = {k: os.path.join(out_dir, k + ".json") for k in ("rotations", "accelerations", "locations")}
    def xyz(a, t):
        return [{"x": float(r[0]), "y": float(r[1]), "z": float(r[2]), "time_usec": int(u)} for r, u in zip(a, t)]
    json.dump({"rotations": xyz(d["gyro"], d["gyro_t"])}, open(paths["rotations"], "w"))
    json.dump({"accelerations": xyz(d["acc"], d["acc_t"])}, open(paths["accelerations"], "w"))
    json.dump({"locations": [{"lat": 0.0, "lon": 0.0, "accuracy_m": 5.0, "speed_m_s": float(v), "time_usec": int(u)}
                             for v, u in zip(d["gps_v"], d["gps_t"])]}, open(paths["locations"], "w"), indent=1)
    return paths
