"""Host-side mirror of pilotguru::AccelerometerCalibrator (include/calibration/velocity.hpp:38-76) and of
fit_motion's velocity pipeline (src/fit_motion.cc:156-293) over the libpgb200 C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, np_ptr


def _f64(a): return np.ascontiguousarray(a, np.float64)
def _i64(a): return np.ascontiguousarray(a, np.int64)


class ImuSeries:
    """The gyro + accelerometer recording on the device (AccelerometerCalibrator's two sensor references), merged
    once (MergedTimeSeries, align_time_series.cc:29-143)."""

    def __init__(self, gyro_xyz, gyro_t, acc_xyz, acc_t, device: int = 0, stream=None):
        self._h = None
        g, gt, a, at = _f64(gyro_xyz), _i64(gyro_t), _f64(acc_xyz), _i64(acc_t)
        if g.shape != (len(gt), 3) or a.shape != (len(at), 3):
            raise ValueError("sensor arrays must be (n, 3) with matching timestamp vectors")
        h = lib().pgb_imu_create(device, np_ptr(g), np_ptr(gt), len(gt), np_ptr(a), np_ptr(at), len(at), stream)
        if not h:
            raise _lib.PgbError(-1, _lib.last_error())
        self._h = C.c_void_p(h)
        self.device = device

    def close(self):
        if self._h:
            lib().pgb_imu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def merged_events(self):
        """ImuTimes(): (effective time usec, gyro index, accel index) per merged event."""
        n = int(lib().pgb_imu_merged_count(self._h))
        t = np.empty(n, np.int64); gi = np.empty(n, np.int64); ai = np.empty(n, np.int64)
        check(lib().pgb_imu_merged_events(self._h, np_ptr(t), np_ptr(gi), np_ptr(ai)))
        return t, gi, ai


class AccelerometerCalibrator:
    """``AccelerometerCalibrator(reference_velocities, rotation_velocities, accelerations)`` (velocity.cc:29-39).

    ``calibrator(x)`` -> (loss, grad) is ``operator()(x, grad)`` (velocity.cc:182-193); x = (g, h, v0), 9 entries."""

    def __init__(self, gps_v, gps_t, imu: ImuSeries):
        self.imu = imu
        self.gps_v, self.gps_t = _f64(gps_v), _i64(gps_t)
        check(lib().pgb_imu_set_window(imu._h, np_ptr(self.gps_v), np_ptr(self.gps_t), len(self.gps_v)))

    def num_intervals(self) -> int:
        return int(lib().pgb_imu_window_intervals(self.imu._h))

    def eval(self, x):
        x = _f64(x)
        if x.shape != (9,):
            raise ValueError("x must have exactly 9 entries")   # CHECK_EQ(in.size(), 9), velocity.cc:47
        loss = C.c_double(); grad = np.zeros(9)
        check(lib().pgb_imu_eval(self.imu._h, np_ptr(x), C.byref(loss), np_ptr(grad)))
        return loss.value, grad

    __call__ = eval

    def minimize(self, x0=None, max_iterations: int = 500, epsilon: float = 1e-5):
        """LBFGSSolver::minimize(calibrator, x, fx) with fit_motion's parameters (fit_motion.cc:167-197)."""
        x = np.zeros(9) if x0 is None else _f64(x0).copy()
        fx = C.c_double(); it = C.c_int()
        check(lib().pgb_imu_minimize(self.imu._h, np_ptr(x), C.byref(fx), C.byref(it), max_iterations, epsilon))
        return it.value, x, fx.value

    def IntegrateTrajectory(self, g, h, v0):
        """velocity.cc:199-256: dict of arrays keyed like the reference's map<size_t, MotionIntegrationOutcome>."""
        x = _f64(np.concatenate([g, h, v0]))
        cap = int(lib().pgb_imu_merged_count(self.imu._h)) + 8
        idx = np.empty(cap, np.int64); sp = np.empty(cap); q = np.empty((cap, 4)); v = np.empty((cap, 3)); d = np.empty(cap, np.int64)
        n = C.c_int64()
        check(lib().pgb_imu_integrate(self.imu._h, np_ptr(x), cap, np_ptr(idx), np_ptr(sp), np_ptr(q), np_ptr(v), np_ptr(d),
                                      C.byref(n)))
        n = n.value
        return dict(index=idx[:n].copy(), speed=sp[:n].copy(), orientation=q[:n].copy(), velocity=v[:n].copy(),
                    duration_usec=d[:n].copy())


def smooth_time_series(values, times, target_times, sigma: float, device: int = 0):
    """SmoothTimeSeries (smoothing.cc:56-98)."""
    v, t, tt = _f64(values), _f64(times), _f64(target_times)
    out = np.empty(len(tt))
    check(lib().pgb_smooth_time_series(device, np_ptr(v), np_ptr(t), len(v), np_ptr(tt), len(tt), sigma, np_ptr(out)))
    return out


def time_averaged_values(values, times_usec, frame_times_usec, device: int = 0):
    """annotate_frames.cc:59-72 over TimeSeries::TimeAveragedValue (time_series.hpp:129-189): (values, valid) for
    frames 1..n-1, each averaged over (t[i-1], t[i]]."""
    v, t, ft = _f64(values), _i64(times_usec), _i64(frame_times_usec)
    out = np.full(max(len(ft) - 1, 0), np.nan); ok = np.zeros(max(len(ft) - 1, 0), np.uint8)
    check(lib().pgb_time_averaged_values(device, np_ptr(v), np_ptr(t), len(v), np_ptr(ft), len(ft), np_ptr(out), np_ptr(ok)))
    return out, ok.astype(bool)


def num_windows(n_gps: int, shift_step: int) -> int:
    return int(lib().pgb_imu_num_windows(n_gps, shift_step))


def principal_rotation_axes(gyro_xyz, gyro_t, integration_interval_usec: int = 500000, device: int = 0):
    """GetPrincipalRotationAxes (src/calibration/rotation.cc:16-57): 3x3 eigenvector rows (row 0 = the vehicle's
    vertical axis as fit_motion.cc:326-333 uses it) and the number of integration intervals."""
    g, t = _f64(gyro_xyz), _i64(gyro_t)
    axes = np.zeros(9); n = C.c_int64()
    check(lib().pgb_principal_rotation_axes(device, np_ptr(g), np_ptr(t), len(t), integration_interval_usec, np_ptr(axes),
                                            C.byref(n)))
    return axes.reshape(3, 3), n.value


def angular_velocities_around_axis(gyro_xyz, axis, device: int = 0):
    """GetAngularVelocitiesAroundAxisDirect (rotation.cc:103-119)."""
    g, a = _f64(gyro_xyz), _f64(axis)
    out = np.empty(len(g))
    check(lib().pgb_angular_velocities_around_axis(device, np_ptr(g), len(g), np_ptr(a), np_ptr(out)))
    return out


def forward_axis(fwd_sum, vertical_axis):
    """fit_motion.cc:281-283: remove the vertical component of the summed device-frame velocities, normalise."""
    f = _f64(fwd_sum).copy(); v = _f64(vertical_axis)
    f -= v * v.dot(f)
    return f / (np.sqrt((f * f).sum()) + 1e-5)


def fit_windows(imu: ImuSeries, gps_v, gps_t, batch_size=40, shift_step=5, max_iterations=500, epsilon=1e-5,
                first_window=0, n_windows=-1, fwd_min_velocity=None, fwd_min_rotation_rad=0.2):
    """All sliding windows of fit_motion.cc:179-221 at once, every L-BFGS on the device.  Returns per-merged-event
    (sum of |v| over covering windows, count) plus per-window x / fx / iterations for the selected shard."""
    gv, gt = _f64(gps_v), _i64(gps_t)
    M = int(lib().pgb_imu_merged_count(imu._h))
    nw_all = num_windows(len(gv), shift_step)
    nw = nw_all - first_window if n_windows < 0 else n_windows
    ssum = np.zeros(M); scnt = np.zeros(M, np.int32)
    x = np.zeros((max(nw, 1), 9)); fx = np.zeros(max(nw, 1)); it = np.zeros(max(nw, 1), np.int32)
    if fwd_min_velocity is None:
        check(lib().pgb_imu_fit_windows(imu._h, np_ptr(gv), np_ptr(gt), len(gv), batch_size, shift_step, max_iterations,
                                        epsilon, first_window, n_windows, np_ptr(ssum), np_ptr(scnt), np_ptr(x),
                                        np_ptr(fx), np_ptr(it)))
        return dict(speed_sum=ssum, speed_cnt=scnt, x=x[:nw], fx=fx[:nw], iters=it[:nw])
    fwd = np.zeros(3); used = C.c_int32()
    check(lib().pgb_imu_fit_windows_fwd(imu._h, np_ptr(gv), np_ptr(gt), len(gv), batch_size, shift_step, max_iterations,
                                        epsilon, first_window, n_windows, np_ptr(ssum), np_ptr(scnt), np_ptr(x),
                                        np_ptr(fx), np_ptr(it), fwd_min_velocity, fwd_min_rotation_rad, np_ptr(fwd),
                                        C.byref(used)))
    return dict(speed_sum=ssum, speed_cnt=scnt, x=x[:nw], fx=fx[:nw], iters=it[:nw], fwd_sum=fwd, fwd_windows=used.value)


def forward_velocities(gyro_xyz, gyro_t, acc_xyz, acc_t, gps_v, gps_t, batch_size=40, shift_step=5, max_iterations=500,
                       post_smoothing_sigma_sec=0.003, device: int = 0, shards=None):
    """ComputeAndSaveForwardVelocitiesFromImu's velocity output (fit_motion.cc:156-279): returns
    (time_usec[], smoothed speed[], averaged speed[], per-window x).  `shards`: optional list of
    (first_window, n_windows) evaluated separately and summed -- what window-sharded ranks do."""
    imu = ImuSeries(gyro_xyz, gyro_t, acc_xyz, acc_t, device=device)
    try:
        parts = shards if shards is not None else [(0, -1)]
        tot = None; cnt = None; xs = []
        for fw, nw in parts:
            r = fit_windows(imu, gps_v, gps_t, batch_size, shift_step, max_iterations, 1e-5, fw, nw)
            tot = r["speed_sum"] if tot is None else tot + r["speed_sum"]
            cnt = r["speed_cnt"] if cnt is None else cnt + r["speed_cnt"]
            xs.append(r["x"])
        mt, _, _ = imu.merged_events()
    finally:
        imu.close()
    covered = cnt > 0
    t_usec = mt[covered]
    avg = tot[covered] / cnt[covered]
    t_sec = (t_usec - t_usec[0]).astype(np.float64) * 1e-6 if len(t_usec) else np.zeros(0)
    sm = smooth_time_series(avg, t_sec, t_sec, post_smoothing_sigma_sec, device) if len(avg) else avg
    return t_usec, sm, avg, np.concatenate(xs) if xs else np.zeros((0, 9))


def smoke_check():
    """Tiny calibration smoke used by __graft_entry__.smoke(): one evaluation + a short solve on the device."""
    from . import synth
    d = synth.imu_gps(12, 100)
    imu = ImuSeries(d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    cal = AccelerometerCalibrator(d["gps_v"], d["gps_t"], imu)
    loss, grad = cal(np.zeros(9))
    assert np.isfinite(loss) and np.all(np.isfinite(grad)) and loss > 0
    it, x, fx = cal.minimize(max_iterations=50)
    assert fx < loss
    imu.close()
