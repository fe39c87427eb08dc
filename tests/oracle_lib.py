"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (never imported by pilotguru_b200)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cc", ".h")) or f == "Makefile"]
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        try:
            so = build()
        except Exception:
            if not os.path.exists(so):
                raise
        _LIB = C.CDLL(so)
        _LIB.pgo_orb_create.restype = C.c_void_p
        _LIB.pgo_fast_atan2.restype = C.c_float
        _LIB.pgo_fast_atan2.argtypes = [C.c_float, C.c_float]
        _LIB.pgo_ic_angle.restype = C.c_float
    return _LIB


_NATIVE = None


def native_lib():
    """(CDLL, build description) of the `-O3 -march=native` build of the same sources, compiled on THIS host (bench.py's
    CPU legs: SURVEY.md 8d's flags).  Falls back to the portable build when the host has no compiler."""
    global _NATIVE
    if _NATIVE is None:
        so = os.path.join(ORACLE_DIR, "_native", "liboracle.so")
        try:
            subprocess.run(["make", "-C", ORACLE_DIR, "native"], check=True, capture_output=True, timeout=300)
            l = C.CDLL(so)
            _NATIVE = (l, "g++ -O3 -march=native -ffp-contract=off, built on this host")
        except Exception as e:  # noqa: BLE001
            _NATIVE = (lib(), f"g++ -O3 -march=x86-64-v3 -ffp-contract=off (portable build; native build failed: {type(e).__name__})")
    return _NATIVE


def ptr(a, t):
    return a.ctypes.data_as(t)


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().pgo_resize_linear(ptr(src, u8p), src.shape[1], src.shape[0], ptr(dst, u8p), dw, dh)
    return dst


def fast(img: np.ndarray, th: int, nms: bool = True) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = lib().pgo_fast(ptr(img, u8p), img.shape[1], img.shape[0], th, int(nms), ptr(out, i32p), cap)
    assert n >= 0
    return out[:n].copy()


def fast_score_map(img: np.ndarray, min_th: int) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty_like(img)
    lib().pgo_fast_score_map(ptr(img, u8p), img.shape[1], img.shape[0], min_th, ptr(out, u8p))
    return out


def gaussian_blur7(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.empty_like(img)
    lib().pgo_gaussian_blur7(ptr(img, u8p), img.shape[1], img.shape[0], ptr(out, u8p))
    return out


def fast_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().pgo_fast_atan2_many(ptr(y, f32p), ptr(x, f32p), ptr(out, f32p), y.size)
    return out


def ic_angle(img: np.ndarray, cx: int, cy: int) -> float:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    return float(lib().pgo_ic_angle(ptr(img, u8p), img.shape[1], img.shape[0], int(cx), int(cy)))


def orb_descriptor(blurred: np.ndarray, cx: int, cy: int, angle: float) -> np.ndarray:
    blurred = np.ascontiguousarray(blurred, dtype=np.uint8)
    d = np.empty(32, np.uint8)
    lib().pgo_orb_descriptor(ptr(blurred, u8p), blurred.shape[1], blurred.shape[0], int(cx), int(cy),
                             C.c_float(angle), ptr(d, u8p))
    return d


def distribute_octree(xyr: np.ndarray, minX, maxX, minY, maxY, N) -> np.ndarray:
    xyr = np.ascontiguousarray(xyr, np.int32)
    keep = np.empty(N + 8, np.int32)
    n = lib().pgo_distribute_octree(ptr(xyr, i32p), len(xyr), minX, maxX, minY, maxY, N, ptr(keep, i32p), len(keep))
    assert n >= 0
    return keep[:n].copy()


class OrbOracle:
    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.l = lib()
        self.h = C.c_void_p(self.l.pgo_orb_create(nfeatures, C.c_float(scale), nlevels, ini_th, min_th))
        assert self.h
        self.nlevels = nlevels
        self.nfeatures = nfeatures

    def __del__(self):
        if getattr(self, "h", None):
            self.l.pgo_orb_destroy(self.h)
            self.h = None

    def tables(self):
        L = self.nlevels
        s = np.empty(L, np.float32); inv = np.empty(L, np.float32); s2 = np.empty(L, np.float32)
        is2 = np.empty(L, np.float32); n = np.empty(L, np.int32); um = np.empty(16, np.int32)
        self.l.pgo_orb_tables(self.h, ptr(s, f32p), ptr(inv, f32p), ptr(s2, f32p), ptr(is2, f32p), ptr(n, i32p),
                              ptr(um, i32p))
        return dict(scale=s, inv_scale=inv, sigma2=s2, inv_sigma2=is2, n_per_level=n, umax=um)

    def level_size(self, w, h, level):
        lw = C.c_int(); lh = C.c_int()
        self.l.pgo_orb_level_size(self.h, w, h, level, C.byref(lw), C.byref(lh))
        return lw.value, lh.value

    def extract(self, gray: np.ndarray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        cap = self.nfeatures + 2 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.l.pgo_orb_extract(self.h, ptr(gray, u8p), gray.shape[1], gray.shape[0], C.c_size_t(gray.shape[1]),
                                   kps.ctypes.data_as(C.c_void_p), ptr(desc, u8p), cap)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy()

    def level(self, level):
        w = C.c_int(); h = C.c_int()
        assert self.l.pgo_orb_get_level(self.h, level, None, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        self.l.pgo_orb_get_level(self.h, level, ptr(out, u8p), C.byref(w), C.byref(h))
        return out

    def candidates(self, level):
        n = self.l.pgo_orb_get_candidates(self.h, level, None, 0)
        out = np.empty((max(n, 1), 3), np.int32)
        self.l.pgo_orb_get_candidates(self.h, level, ptr(out, i32p), n)
        return out[:n]

    def level_keypoints(self, level):
        n = self.l.pgo_orb_get_level_keypoints(self.h, level, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.l.pgo_orb_get_level_keypoints(self.h, level, out.ctypes.data_as(C.c_void_p), n)
        return out[:n]

    def stage_times(self, reset=True):
        t = np.zeros(6)
        self.l.pgo_orb_stage_times(self.h, ptr(t, f64p), int(reset))
        return t


def descriptor_distance(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().pgo_descriptor_distance(ptr(a, u8p), ptr(b, u8p)))


def search_by_projection(cur_kps, cur_desc, q_uv, q_octave, q_angle, q_desc, q_valid, bounds, th, scale_factors,
                         check_ori=True):
    cur_kps = np.ascontiguousarray(cur_kps, KP_DTYPE); cur_desc = np.ascontiguousarray(cur_desc, np.uint8)
    q_uv = np.ascontiguousarray(q_uv, np.float32); q_octave = np.ascontiguousarray(q_octave, np.int32)
    q_angle = np.ascontiguousarray(q_angle, np.float32); q_desc = np.ascontiguousarray(q_desc, np.uint8)
    q_valid = np.ascontiguousarray(q_valid, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    m = np.full(max(len(cur_kps), 1), -1, np.int32)
    bd = np.full(max(len(q_octave), 1), -1, np.int32)
    n = lib().pgo_search_by_projection(cur_kps.ctypes.data_as(C.c_void_p), ptr(cur_desc, u8p), len(cur_kps),
                                       ptr(q_uv, f32p), ptr(q_octave, i32p), ptr(q_angle, f32p), ptr(q_desc, u8p),
                                       ptr(q_valid, u8p), len(q_octave), C.c_float(bounds[0]), C.c_float(bounds[1]),
                                       C.c_float(bounds[2]), C.c_float(bounds[3]), C.c_float(th), ptr(sf, f32p),
                                       len(sf), int(check_ori), ptr(m, i32p), ptr(bd, i32p))
    return n, m[:len(cur_kps)], bd[:len(q_octave)]


def match_consecutive(prev_kps, prev_desc, cur_kps, cur_desc, flow, max_x, max_y, th, scale_factors):
    prev_kps = np.ascontiguousarray(prev_kps, KP_DTYPE); prev_desc = np.ascontiguousarray(prev_desc, np.uint8)
    cur_kps = np.ascontiguousarray(cur_kps, KP_DTYPE); cur_desc = np.ascontiguousarray(cur_desc, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    m = np.full(max(len(cur_kps), 1), -1, np.int32)
    n = lib().pgo_match_consecutive(prev_kps.ctypes.data_as(C.c_void_p), ptr(prev_desc, u8p), len(prev_kps),
                                    cur_kps.ctypes.data_as(C.c_void_p), ptr(cur_desc, u8p), len(cur_kps),
                                    C.c_float(flow[0]), C.c_float(flow[1]), C.c_float(max_x), C.c_float(max_y),
                                    C.c_float(th), ptr(sf, f32p), len(sf), ptr(m, i32p))
    return n, m[:len(cur_kps)]


def bench_extract_match(frames: np.ndarray, flows: np.ndarray, nthreads: int, nfeatures=1000, th=15.0, per_frame=False,
                        native=False):
    """Timed CPU baseline: returns (seconds, total_keypoints, total_matches[, keypoints per frame, matches per frame]).
    Every consecutive pair (t-1, t) is matched; matches_per_frame[0] = -1."""
    frames = np.ascontiguousarray(frames, np.uint8); flows = np.ascontiguousarray(flows, np.float32)
    n, h, w = frames.shape
    l = native_lib()[0] if native else lib()
    l.pgo_bench_extract_match.restype = C.c_double
    tk = C.c_int64(); tm = C.c_int64()
    nk = np.zeros(n, np.int32); nm = np.zeros(n, np.int32)
    s = l.pgo_bench_extract_match(ptr(frames, u8p), n, w, h, ptr(flows, f32p), nfeatures, C.c_float(1.2), 8, 20, 7,
                                  C.c_float(th), nthreads, C.byref(tk), C.byref(tm), ptr(nk, i32p), ptr(nm, i32p))
    if per_frame:
        return float(s), tk.value, tm.value, nk, nm
    return float(s), tk.value, tm.value


class CalibOracle:
    """AccelerometerCalibrator restatement (literal) + the contract-header ("core") evaluation on the host."""

    def __init__(self, gps_v, gps_t, gyro, gyro_t, acc, acc_t):
        self.l = lib()
        self.l.pgo_calib_create.restype = C.c_void_p
        self.l.pgo_calib_eval.restype = C.c_double
        self.l.pgo_calib_eval_core.restype = C.c_double
        for f in ("pgo_calib_merged_count", "pgo_calib_num_intervals", "pgo_calib_integrate", "pgo_calib_integrate_core"):
            getattr(self.l, f).restype = C.c_int64
        a = lambda x, t: np.ascontiguousarray(x, t)
        self.gps_v, self.gps_t = a(gps_v, np.float64), a(gps_t, np.int64)
        self.gyro, self.gyro_t, self.acc, self.acc_t = a(gyro, np.float64), a(gyro_t, np.int64), a(acc, np.float64), a(acc_t, np.int64)
        self.h = C.c_void_p(self.l.pgo_calib_create(ptr(self.gps_v, f64p), ptr(self.gps_t, i64p), len(self.gps_v),
                                                    ptr(self.gyro, f64p), ptr(self.gyro_t, i64p), C.c_int64(len(self.gyro_t)),
                                                    ptr(self.acc, f64p), ptr(self.acc_t, i64p), C.c_int64(len(self.acc_t))))
        if not self.h:
            raise ValueError("pgo_calib_create failed (CHECK failure in the reference)")

    def __del__(self):
        if getattr(self, "h", None):
            self.l.pgo_calib_destroy(self.h); self.h = None

    def merged(self):
        n = self.l.pgo_calib_merged_count(self.h)
        t = np.empty(n, np.int64); gi = np.empty(n, np.int64); ai = np.empty(n, np.int64)
        self.l.pgo_calib_merged_events(self.h, ptr(t, i64p), ptr(gi, i64p), ptr(ai, i64p))
        return t, gi, ai

    def intervals(self):
        n = self.l.pgo_calib_num_intervals(self.h)
        r = [np.empty(n, np.int64) for _ in range(4)]
        self.l.pgo_calib_intervals(self.h, *[ptr(x, i64p) for x in r])
        return r  # ref_idx, merged_idx, start, end

    def eval(self, x, core=False):
        x = np.ascontiguousarray(x, np.float64); g = np.zeros(9)
        f = (self.l.pgo_calib_eval_core if core else self.l.pgo_calib_eval)(self.h, ptr(x, f64p), ptr(g, f64p))
        return float(f), g

    def minimize(self, x0=None, max_iterations=500, epsilon=1e-5, mode="literal"):
        x = np.zeros(9) if x0 is None else np.array(x0, np.float64)
        fx = C.c_double(); ne = C.c_int()
        fn = {"literal": self.l.pgo_calib_minimize, "core": self.l.pgo_calib_minimize_core,
              "literal_driver_core_eval": self.l.pgo_calib_minimize_literal_driver_core_eval}[mode]
        it = fn(self.h, ptr(x, f64p), C.byref(fx), max_iterations, C.c_double(epsilon), C.byref(ne))
        return it, x, fx.value, ne.value

    def integrate(self, x, core=False):
        x = np.ascontiguousarray(x, np.float64)
        cap = self.l.pgo_calib_merged_count(self.h) + 8
        idx = np.empty(cap, np.int64); sp = np.empty(cap); q = np.empty((cap, 4)); v = np.empty((cap, 3)); d = np.empty(cap, np.int64)
        if core:
            n = self.l.pgo_calib_integrate_core(self.h, ptr(x, f64p), C.c_int64(cap), ptr(idx, i64p), ptr(sp, f64p), ptr(v, f64p), ptr(d, i64p))
        else:
            n = self.l.pgo_calib_integrate(self.h, ptr(x, f64p), C.c_int64(cap), ptr(idx, i64p), ptr(sp, f64p), ptr(q, f64p), ptr(v, f64p), ptr(d, i64p))
        assert n >= 0
        return idx[:n].copy(), sp[:n].copy(), q[:n].copy(), v[:n].copy(), d[:n].copy()


def smooth_time_series(values, times, target, sigma):
    v = np.ascontiguousarray(values, np.float64); t = np.ascontiguousarray(times, np.float64)
    tt = np.ascontiguousarray(target, np.float64); out = np.empty(len(tt))
    lib().pgo_smooth_time_series(ptr(v, f64p), ptr(t, f64p), C.c_int64(len(v)), ptr(tt, f64p), C.c_int64(len(tt)),
                                 C.c_double(sigma), ptr(out, f64p))
    return out


def fit_motion(d, batch_size=40, shift_step=5, max_iters=500, sigma=0.003, mode=0):
    """The window loop of fit_motion.cc:156-273 on a synth.imu_gps() dict. mode 0 literal, 1 contract/core."""
    l = lib()
    l.pgo_fit_motion.restype = C.c_int64
    gv = np.ascontiguousarray(d["gps_v"], np.float64); gt = np.ascontiguousarray(d["gps_t"], np.int64)
    gy = np.ascontiguousarray(d["gyro"], np.float64); gyt = np.ascontiguousarray(d["gyro_t"], np.int64)
    ac = np.ascontiguousarray(d["acc"], np.float64); act = np.ascontiguousarray(d["acc_t"], np.int64)
    cap = len(gyt) + len(act) + 8
    nwin = (len(gv) + shift_step - 1) // shift_step
    idx = np.empty(cap, np.int64); ts = np.empty(cap, np.int64); avg = np.empty(cap); sm = np.empty(cap)
    xo = np.zeros((nwin, 9)); it = np.zeros(nwin, np.int32); fx = np.zeros(nwin); ne = C.c_int64()
    n = l.pgo_fit_motion(ptr(gv, f64p), ptr(gt, i64p), len(gv), ptr(gy, f64p), ptr(gyt, i64p), C.c_int64(len(gyt)),
                         ptr(ac, f64p), ptr(act, i64p), C.c_int64(len(act)), batch_size, shift_step, max_iters,
                         C.c_double(sigma), mode, C.c_int64(cap), ptr(idx, i64p), ptr(ts, i64p), ptr(avg, f64p),
                         ptr(sm, f64p), ptr(xo, f64p), ptr(it, i32p), ptr(fx, f64p), C.byref(ne))
    if n < 0:
        raise RuntimeError(f"pgo_fit_motion failed: {n}")
    return dict(idx=idx[:n].copy(), t_usec=ts[:n].copy(), avg=avg[:n].copy(), smoothed=sm[:n].copy(), x=xo, iters=it,
                fx=fx, n_evals=ne.value)


REF_SO = os.path.join(ORACLE_DIR, "_ref", "libpilotguru_ref.so")
_REF = None


def ref_lib():
    """oracle/_ref/libpilotguru_ref.so -- the reference's own sources compiled where they lie (oracle/Makefile `_ref`) --
    or None when it is neither built nor buildable (no /root/reference, e.g. a GPU box that got no prebuilt copy)."""
    global _REF
    if _REF is None:
        if os.path.isdir("/root/reference"):
            try:
                subprocess.run(["make", "-C", ORACLE_DIR, "_ref"], check=True, capture_output=True)
            except Exception:
                pass
        _REF = C.CDLL(REF_SO) if os.path.exists(REF_SO) else False
    return _REF or None


def ref_fit_motion(d, batch_size=40, shift_step=5, max_iters=500, sigma=0.003):
    """ComputeAndSaveForwardVelocitiesFromImu (src/fit_motion.cc:156-293) compiled from the REFERENCE'S file (oracle/_ref):
    (event timestamps, smoothed velocities).  None when oracle/_ref is unavailable."""
    l = ref_lib()
    if l is None:
        return None
    l.pgr_fit_motion.restype = C.c_int64
    a = lambda x, t: np.ascontiguousarray(x, t)
    gv, gt, gy, gyt, ac, act = a(d["gps_v"], np.float64), a(d["gps_t"], np.int64), a(d["gyro"], np.float64), a(d["gyro_t"], np.int64), a(d["acc"], np.float64), a(d["acc_t"], np.int64)
    cap = len(gyt) + len(act) + 8
    vertical = np.array([0.0, 0.0, 1.0])
    rt = np.empty(cap, np.int64); rs = np.empty(cap); rf = np.zeros(3)
    n = l.pgr_fit_motion(ptr(gv, f64p), ptr(gt, i64p), C.c_int64(len(gt)), ptr(gy, f64p), ptr(gyt, i64p), C.c_int64(len(gyt)),
                         ptr(ac, f64p), ptr(act, i64p), C.c_int64(len(act)), ptr(vertical, f64p), C.c_int64(batch_size),
                         C.c_int64(shift_step), C.c_int64(max_iters), C.c_double(sigma), C.c_double(5.0), C.c_double(0.2),
                         ptr(rt, i64p), ptr(rs, f64p), C.c_int64(cap), ptr(rf, f64p))
    if n < 0:
        raise RuntimeError(f"pgr_fit_motion failed: {n}")
    return rt[:n].copy(), rs[:n].copy()


def forward_axis_sum(d, x_all, batch_size=40, shift_step=5, mode=0, min_vel=5.0, min_rot=0.2):
    """total_velocity_local of fit_motion.cc:172-173,223-248 for per-window solutions x_all; returns (sum[3], windows used)."""
    l = lib()
    gv = np.ascontiguousarray(d["gps_v"], np.float64); gt = np.ascontiguousarray(d["gps_t"], np.int64)
    gy = np.ascontiguousarray(d["gyro"], np.float64); gyt = np.ascontiguousarray(d["gyro_t"], np.int64)
    ac = np.ascontiguousarray(d["acc"], np.float64); act = np.ascontiguousarray(d["acc_t"], np.int64)
    xa = np.ascontiguousarray(x_all, np.float64); out = np.zeros(3); used = C.c_int32()
    rc = l.pgo_forward_axis_sum(ptr(gv, f64p), ptr(gt, i64p), len(gv), ptr(gy, f64p), ptr(gyt, i64p), C.c_int64(len(gyt)),
                                ptr(ac, f64p), ptr(act, i64p), C.c_int64(len(act)), batch_size, shift_step, ptr(xa, f64p),
                                mode, C.c_double(min_vel), C.c_double(min_rot), ptr(out, f64p), C.byref(used))
    if rc:
        raise RuntimeError(f"pgo_forward_axis_sum failed: {rc}")
    return out, used.value


def pca3(rows):
    """cv::PCA(rows, noArray(), DATA_AS_ROW) of an n x 3 matrix: (eigenvectors 3x3, eigenvalues, mean)."""
    r = np.ascontiguousarray(rows, np.float64); ev = np.zeros(9); ew = np.zeros(3); mu = np.zeros(3)
    lib().pgo_pca3(ptr(r, f64p), C.c_int64(len(r)), ptr(ev, f64p), ptr(ew, f64p), ptr(mu, f64p))
    return ev.reshape(3, 3), ew, mu


def principal_rotation_axes(gyro, gyro_t, interval_usec=500000):
    """GetPrincipalRotationAxes (rotation.cc:16-57), literal: (axes 3x3, PCA input rows)."""
    l = lib()
    l.pgo_principal_rotation_axes.restype = C.c_int64
    g = np.ascontiguousarray(gyro, np.float64); t = np.ascontiguousarray(gyro_t, np.int64)
    axes = np.zeros(9); rows = np.zeros((len(t), 3))
    n = l.pgo_principal_rotation_axes(ptr(g, f64p), ptr(t, i64p), C.c_int64(len(t)), C.c_int64(interval_usec), ptr(axes, f64p),
                                      ptr(rows, f64p), C.c_int64(len(t)))
    if n < 0:
        raise ValueError("fewer than 3 rotation integration intervals")
    return axes.reshape(3, 3), rows[:n].copy()


def angular_velocities_around_axis(gyro, axis):
    g = np.ascontiguousarray(gyro, np.float64); a = np.ascontiguousarray(axis, np.float64); out = np.empty(len(g))
    lib().pgo_angular_velocities_around_axis(ptr(g, f64p), C.c_int64(len(g)), ptr(a, f64p), ptr(out, f64p))
    return out


def time_averaged_values(values, times_usec, frame_times_usec):
    """annotate_frames.cc:59-72 (TimeAveragedValue per frame interval), literal: (values, valid)."""
    v = np.ascontiguousarray(values, np.float64); t = np.ascontiguousarray(times_usec, np.int64)
    ft = np.ascontiguousarray(frame_times_usec, np.int64)
    out = np.full(max(len(ft) - 1, 0), np.nan); ok = np.zeros(max(len(ft) - 1, 0), np.uint8)
    rc = lib().pgo_time_averaged_values(ptr(v, f64p), ptr(t, i64p), C.c_int64(len(v)), ptr(ft, i64p), C.c_int64(len(ft)),
                                        ptr(out, f64p), ptr(ok, u8p))
    if rc:
        raise ValueError("the reference would CHECK-fail on this input")
    return out, ok.astype(bool)


def search_for_initialization(k1, d1, k2, d2, prev_matched, window_size, bounds, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:407-522), literal: (nmatches, vnMatches12, updated vbPrevMatched)."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2)
    d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
    pm = np.ascontiguousarray(prev_matched, np.float32).copy(); m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = lib().pgo_search_for_initialization(k1.ctypes.data_as(C.c_void_p), ptr(d1, u8p), len(k1), k2.ctypes.data_as(C.c_void_p),
                                            ptr(d2, u8p), len(k2), ptr(pm, f32p), int(window_size), C.c_float(bounds[0]),
                                            C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]), C.c_float(nnratio),
                                            int(check_ori), ptr(m12, i32p))
    return n, m12[:len(k1)].copy(), pm


def search_map_points(kps, desc, has_mp, proj_xy, track_level, view_cos, mp_desc, in_view, mp_observed, bounds, th,
                      scale_factors, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:46-131), literal."""
    kps = np.ascontiguousarray(kps); desc = np.ascontiguousarray(desc, np.uint8); has_mp = np.ascontiguousarray(has_mp, np.uint8)
    uv = np.ascontiguousarray(proj_xy, np.float32); lv = np.ascontiguousarray(track_level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32); qd = np.ascontiguousarray(mp_desc, np.uint8)
    iv = np.ascontiguousarray(in_view, np.uint8); ob = np.ascontiguousarray(mp_observed, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32); mo = np.full(max(len(kps), 1), -1, np.int32)
    n = lib().pgo_search_map_points(kps.ctypes.data_as(C.c_void_p), ptr(desc, u8p), len(kps), ptr(has_mp, u8p), ptr(uv, f32p),
                                    ptr(lv, i32p), ptr(vc, f32p), ptr(qd, u8p), ptr(iv, u8p), ptr(ob, u8p), len(lv),
                                    C.c_float(bounds[0]), C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]),
                                    C.c_float(th), ptr(sf, f32p), C.c_float(nnratio), ptr(mo, i32p))
    return n, mo[:len(kps)].copy()


def search_by_bow(kf_desc, kf_angle, kf_has_mp, kf_featvec, f_desc, f_angle, f_featvec, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW (ORBmatcher.cc:161-290), literal; feature vectors as (node_ids, starts, indices)."""
    kd = np.ascontiguousarray(kf_desc, np.uint8); ka = np.ascontiguousarray(kf_angle, np.float32)
    kh = np.ascontiguousarray(kf_has_mp, np.uint8); fd = np.ascontiguousarray(f_desc, np.uint8)
    fa = np.ascontiguousarray(f_angle, np.float32)
    kn, ks, ki = [np.ascontiguousarray(a, t) for a, t in zip(kf_featvec, (np.uint32, np.int32, np.uint32))]
    fn, fs, fi = [np.ascontiguousarray(a, t) for a, t in zip(f_featvec, (np.uint32, np.int32, np.uint32))]
    mo = np.full(max(len(fd), 1), -1, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib().pgo_search_by_bow(vp(kd), vp(ka), vp(kh), vp(kn), vp(ks), vp(ki), len(kn), vp(fd), vp(fa), len(fd), vp(fn), vp(fs),
                                vp(fi), len(fn), C.c_float(nnratio), int(check_ori), vp(mo))
    return n, mo[:len(fd)].copy()


def pose_optimization(Tcw, kp_xy, kp_octave, mp_xyz, has_mp, inv_level_sigma2, fx, fy, cx, cy):
    """Optimizer::PoseOptimization (Optimizer.cc:239-451), mono; returns (n_inliers, Tcw_out, outlier, outlier_after_round[4])."""
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16); xy = np.ascontiguousarray(kp_xy, np.float32)
    oc = np.ascontiguousarray(kp_octave, np.int32); X = np.ascontiguousarray(mp_xyz, np.float32)
    hm = np.ascontiguousarray(has_mp, np.uint8); s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    n = len(oc)
    To = np.zeros(16, np.float32); out = np.zeros(max(n, 1), np.uint8); rounds = np.full((4, max(n, 1)), 255, np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    r = lib().pgo_pose_optimization(vp(T), vp(xy), vp(oc), vp(X), vp(hm), n, vp(s2), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                                    C.c_float(cy), vp(To), vp(out), vp(rounds))
    return int(r), To.reshape(4, 4), out[:n].copy(), rounds[:, :n].copy()


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:259-324), literal; -1 for an empty set."""
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    return int(lib().pgo_distinctive_descriptor(ptr(d, u8p) if len(d) else None, len(d)))


def to_gray(img, rgb_order=True, vflip=False, hflip=False, formula=0):
    """cv::flip + cvtColor to gray of one (h, w[, c]) uint8 image."""
    a = np.ascontiguousarray(img, np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    out = np.empty((h, w), np.uint8)
    lib().pgo_to_gray(ptr(a, u8p), w, h, c, int(rgb_order), int(vflip), int(hflip), int(formula), ptr(out, u8p))
    return out


def rotate_like_reader(img, rotate_degrees):
    """The rotation VideoImageSequenceSource::fetchNext applies (src/io/image_sequence_reader.cc:186-207), in numpy:
    90 -> cv::flip(raw.t(), 0), 180 -> cv::flip(raw, -1), 270 -> cv::flip(raw.t(), 1).  tests/test_oracle_feed.py pins it
    against cv2.transpose / cv2.flip."""
    a = np.asarray(img)
    r = rotate_degrees % 360
    if r == 0:
        return a
    if r == 180:
        return a[::-1, ::-1]
    t = a.transpose(1, 0, 2) if a.ndim == 3 else a.T
    return t[::-1] if r == 90 else t[:, ::-1]
