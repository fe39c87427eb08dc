"""GPU parity tests of the fit_motion path beyond the velocity calibration (tests/test_gpu_calib.py): principal
rotation axis + steering (src/calibration/rotation.cc:16-57,103-119), forward-axis evidence
(src/fit_motion.cc:223-248,281-283), and the C++ `fit_motion` drop-in binary end to end on BASELINE configs[0]
(60 s synthetic 100 Hz IMU + 1 Hz GPS JSON), all against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rotation_axes_and_steering():
    from pilotguru_b200 import calibration as cal
    d = synth.imu_gps(60, 100)
    axes, n = cal.principal_rotation_axes(d["gyro"], d["gyro_t"], 500000)
    oaxes, rows = O.principal_rotation_axes(d["gyro"], d["gyro_t"], 500000)
    assert n == len(rows) == 120
    assert np.max(np.abs(axes - oaxes)) <= 1e-9                          # same vectors, same (OpenCV) signs
    st = cal.angular_velocities_around_axis(d["gyro"], oaxes[0])
    ost = O.angular_velocities_around_axis(d["gyro"], oaxes[0])
    assert np.max(np.abs(st - ost)) <= 1e-15
    with pytest.raises(Exception):
        cal.principal_rotation_axes(d["gyro"][:100], d["gyro_t"][:100], 500000)   # < 3 intervals: CHECK_GE
    with pytest.raises(Exception):
        cal.angular_velocities_around_axis(d["gyro"], [0.0, 0.0, 2.0])            # axis not normalised: CHECK_LT


def test_forward_axis_sum_matches_oracle():
    from pilotguru_b200 import calibration as cal
    d = synth.imu_gps(60, 100)
    imu = cal.ImuSeries(d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    r = cal.fit_windows(imu, d["gps_v"], d["gps_t"], max_iterations=80, fwd_min_velocity=5.0, fwd_min_rotation_rad=0.02)
    imu.close()
    fm = O.fit_motion(d, max_iters=80, mode=1)
    assert np.array_equal(r["x"], fm["x"])                                # per-window solutions: bit-exact (contract)
    s, used = O.forward_axis_sum(d, fm["x"], mode=1, min_vel=5.0, min_rot=0.02)
    assert r["fwd_windows"] == used and used > 0
    assert np.max(np.abs(r["fwd_sum"] - s)) <= 1e-9 * np.max(np.abs(s))
    # sharded over "ranks": the per-shard sums add up
    imu = cal.ImuSeries(d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    a = cal.fit_windows(imu, d["gps_v"], d["gps_t"], max_iterations=80, first_window=0, n_windows=5, fwd_min_velocity=5.0,
                        fwd_min_rotation_rad=0.02)
    b = cal.fit_windows(imu, d["gps_v"], d["gps_t"], max_iterations=80, first_window=5, n_windows=-1, fwd_min_velocity=5.0,
                        fwd_min_rotation_rad=0.02)
    imu.close()
    assert a["fwd_windows"] + b["fwd_windows"] == used
    assert np.max(np.abs(a["fwd_sum"] + b["fwd_sum"] - s)) <= 1e-9 * np.max(np.abs(s))


def test_fit_motion_binary_end_to_end(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "pilotguru_b200", "host")], check=True, capture_output=True)
    d = synth.imu_gps(60, 100)                                            # BASELINE configs[0]
    paths = synth.write_imu_gps_json(d, str(tmp_path))
    out = {k: str(tmp_path / (k + ".json")) for k in ("velocities", "steering", "forward")}
    p = subprocess.run([os.path.join(ROOT, "pilotguru_b200", "host", "fit_motion"),
                        "--rotations_json", paths["rotations"], "--accelerations_json=" + paths["accelerations"],
                        "--locations_json", paths["locations"], "--velocities_out_json", out["velocities"],
                        "--steering_out_json", out["steering"], "--forward_axis_out_json", out["forward"],
                        "--forward_axis_inference_min_rotation_rad=0.02", "--logtostderr"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stderr.count("Sliding window optimization") == 12            # 59 GPS samples, step 5
    vel = json.load(open(out["velocities"]))["velocities"]
    fm = O.fit_motion(d, mode=1)                                          # contract evaluation on the host
    assert [e["time_usec"] for e in vel] == fm["t_usec"].tolist()
    sp = np.array([e["speed_m_s"] for e in vel])
    assert np.max(np.abs(sp - fm["smoothed"]) / np.abs(fm["smoothed"])) <= 1e-6   # the north-star gate
    lit = O.fit_motion(d, mode=0)                                         # literal restatement of the reference
    print("fit_motion binary vs literal restatement: max rel dev of smoothed speeds",
          float(np.max(np.abs(sp - lit["smoothed"]) / np.abs(lit["smoothed"]))))
    st = json.load(open(out["steering"]))["steering"]
    oaxes, _ = O.principal_rotation_axes(d["gyro"], d["gyro_t"], 500000)
    ost = O.angular_velocities_around_axis(d["gyro"], oaxes[0])
    assert [e["time_usec"] for e in st] == d["gyro_t"].tolist()
    assert np.max(np.abs(np.array([e["angular_velocity"] for e in st]) - ost)) <= 1e-9
    fw = json.load(open(out["forward"]))["forward_axis"]
    s, used = O.forward_axis_sum(d, fm["x"], mode=1, min_vel=5.0, min_rot=0.02)
    f = s - oaxes[0] * oaxes[0].dot(s); f /= np.sqrt((f * f).sum()) + 1e-5
    assert used > 0 and np.max(np.abs(np.array([fw["x"], fw["y"], fw["z"]]) - f)) <= 1e-6
    assert fw["x"] > 0.9


def test_time_averaged_values_and_annotate_frames_binary(tmp_path):
    from pilotguru_b200 import calibration as cal
    rng = np.random.default_rng(4)
    t = np.cumsum(rng.integers(1500, 2500, 20000)).astype(np.int64) + 1_000_000
    v = np.sin(t * 2e-6) * 5 + rng.normal(0, 0.2, len(t))
    ft = (np.arange(0, 1300) * 33_333 + 600_000).astype(np.int64)
    out, ok = cal.time_averaged_values(v, t, ft)
    oout, ook = O.time_averaged_values(v, t, ft)
    assert np.array_equal(ok, ook) and ok.any() and (~ok).any()
    assert np.array_equal(out[ok], oout[ok])                               # same fp64 operations in the same order
    with pytest.raises(Exception):
        cal.time_averaged_values(v, t, np.array([t[5], int(t[-1])], np.int64))   # the reference's CHECK
    # the binary: smoothing + annotation, frame ids and values
    subprocess.run(["make", "-C", os.path.join(ROOT, "pilotguru_b200", "host")], check=True, capture_output=True)
    fj, sj, oj = tmp_path / "frames.json", tmp_path / "velocities.json", tmp_path / "out.json"
    fj.write_text(json.dumps({"frames": [{"frame_id": int(i), "time_usec": int(x)} for i, x in enumerate(ft)]}))
    sj.write_text(json.dumps({"velocities": [{"speed_m_s": float(a), "time_usec": int(b)} for a, b in zip(v, t)]}))
    p = subprocess.run([os.path.join(ROOT, "pilotguru_b200", "host", "annotate_frames"), "--frames_json", str(fj), "--in_json", str(sj),
                        "--json_root_element_name=velocities", "--json_value_name=speed_m_s", "--out_json", str(oj),
                        "--smoothing_sigma=0.01"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.load(open(oj))["velocities"]
    ts = (t - t[0]) * 1e-6
    sm = O.smooth_time_series(v, ts, ts, 0.01)
    want, wok = O.time_averaged_values(sm, t, ft)
    assert [e["frame_id"] for e in res] == (np.nonzero(wok)[0] + 1).tolist()
    got = np.array([e["speed_m_s"] for e in res])
    assert np.max(np.abs(got - want[wok])) <= 1e-12 * np.max(np.abs(want[wok]))
