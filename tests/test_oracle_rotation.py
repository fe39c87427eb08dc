"""CPU tests of the oracle's rotation / forward-axis restatement (src/calibration/rotation.cc:16-57,103-119,
src/fit_motion.cc:223-248,281-283) and of its cv::PCA restatement, pinned against cv2 golden vectors."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth


def test_pca_matches_cv2_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_pca.npz"))
    k = 0
    while f"rows{k}" in g:
        vec, val, mean = O.pca3(g[f"rows{k}"])
        assert np.max(np.abs(vec - g[f"vec{k}"])) <= 1e-12, k            # eigenvectors INCLUDING OpenCV's signs
        assert np.allclose(val, g[f"val{k}"], rtol=1e-10, atol=1e-12 * g[f"val{k}"].max()), k
        assert np.allclose(mean, g[f"mean{k}"], rtol=1e-13, atol=1e-15), k
        k += 1
    assert k >= 20


def test_pca_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(50):
        rows = rng.normal(size=(int(rng.integers(3, 200)), 3)) * rng.uniform(0.1, 2, size=3)
        vec, _, _ = O.pca3(rows)
        _, e, _ = cv2.PCACompute2(rows, mean=None)
        assert np.max(np.abs(vec - e)) <= 1e-12


def _rows_numpy(gyro, t, interval):
    """independent restatement of the interval integration (rotation.cc:20-44) with numpy quaternions"""
    rows = []; q = np.array([1.0, 0, 0, 0]); cur = 0
    for i in range(1, len(t)):
        dur = int(t[i] - t[i - 1]); cur += dur
        w = gyro[i]; rate = np.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]); half = rate * (dur * 1e-6) * 0.5
        s = np.sin(half) / (rate + 1e-30)
        d = np.array([np.cos(half), w[0] * s, w[1] * s, w[2] * s])
        q = np.array([q[0] * d[0] - q[1] * d[1] - q[2] * d[2] - q[3] * d[3], q[0] * d[1] + q[1] * d[0] + q[2] * d[3] - q[3] * d[2],
                      q[0] * d[2] + q[2] * d[0] + q[3] * d[1] - q[1] * d[3], q[0] * d[3] + q[3] * d[0] + q[1] * d[2] - q[2] * d[1]])
        if cur >= interval:
            rows.append(q[1:].copy()); q = np.array([1.0, 0, 0, 0]); cur = 0
    return np.array(rows)


def test_principal_axes_and_steering():
    d = synth.imu_gps(30, 100)
    axes, rows = O.principal_rotation_axes(d["gyro"], d["gyro_t"], 500000)
    ref = _rows_numpy(d["gyro"], d["gyro_t"], 500000)
    assert rows.shape == ref.shape == (60, 3) and np.max(np.abs(rows - ref)) <= 1e-15
    # the synthetic car only yaws: the principal rotation axis is +-z of the device frame
    assert abs(abs(axes[0, 2]) - 1.0) < 1e-3 and abs(np.linalg.norm(axes[0]) - 1.0) < 1e-12
    assert np.allclose(axes @ axes.T, np.eye(3), atol=1e-12)
    st = O.angular_velocities_around_axis(d["gyro"], axes[0])
    assert np.allclose(st, d["gyro"] @ axes[0] / np.linalg.norm(axes[0]), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        O.principal_rotation_axes(d["gyro"][:100], d["gyro_t"][:100], 500000)   # 0.99 s: fewer than 3 intervals


def test_forward_axis_points_forward():
    d = synth.imu_gps(60, 100)
    fm = O.fit_motion(d, max_iters=60, mode=1)
    s, used = O.forward_axis_sum(d, fm["x"], mode=1, min_vel=5.0, min_rot=0.02)
    s0, used0 = O.forward_axis_sum(d, fm["x"], mode=0, min_vel=5.0, min_rot=0.02)
    assert used == used0 and used > 0
    assert np.max(np.abs(s - s0)) <= 1e-7 * np.max(np.abs(s0))        # contract vs literal trajectory integration
    axes, _ = O.principal_rotation_axes(d["gyro"], d["gyro_t"])
    f = s - axes[0] * axes[0].dot(s); f /= np.linalg.norm(f) + 1e-5
    assert f[0] > 0.9                                                  # the generator drives along +x of the device
    # a rotation threshold no window reaches leaves the sum empty
    z, used_none = O.forward_axis_sum(d, fm["x"], mode=1, min_vel=5.0, min_rot=3.0)
    assert used_none == 0 and not z.any()
