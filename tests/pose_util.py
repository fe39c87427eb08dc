"""Synthetic PnP scenes for the PoseOptimization tests: map points in a frustum, a true camera pose, keypoints =
projections + octave-dependent noise, a fraction of gross outliers, and a perturbed initial pose (what the motion model
hands to Optimizer::PoseOptimization)."""
import numpy as np

FX, FY, CX, CY = 1000.0, 1000.0, 960.0, 540.0
INV_SIGMA2 = (1.0 / np.cumprod(np.r_[1.0, np.full(7, 1.2, np.float32)]).astype(np.float32) ** 2).astype(np.float32)


def rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def scene(seed, n=400, n_mp=None, outlier_frac=0.2, noise=1.0, perturb=(0.02, 0.15)):
    rng = np.random.default_rng(seed)
    R = rodrigues(rng.normal(0, 0.2, 3)); t = rng.normal(0, 1.0, 3)
    z = rng.uniform(4, 30, n)
    u = rng.uniform(0, 1920, n); v = rng.uniform(0, 1080, n)
    Xc = np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], axis=1)
    Xw = ((Xc - t) @ R).astype(np.float32)                      # Xc = R Xw + t
    octave = rng.integers(0, 8, n).astype(np.int32)
    Xc32 = Xw.astype(np.float64) @ R.T + t
    xy = np.stack([Xc32[:, 0] / Xc32[:, 2] * FX + CX, Xc32[:, 1] / Xc32[:, 2] * FY + CY], axis=1)
    xy += rng.normal(0, noise, (n, 2)) * (1.2 ** octave)[:, None]
    bad = rng.uniform(size=n) < outlier_frac
    xy[bad] = np.stack([rng.uniform(0, 1920, bad.sum()), rng.uniform(0, 1080, bad.sum())], axis=1)
    has = np.ones(n, np.uint8)
    if n_mp is not None:
        has[:] = 0; has[rng.permutation(n)[:n_mp]] = 1
    else:
        has[rng.uniform(size=n) < 0.25] = 0
    T_true = np.eye(4); T_true[:3, :3] = R; T_true[:3, 3] = t
    dR = rodrigues(rng.normal(0, perturb[0], 3)); dt = rng.normal(0, perturb[1], 3)
    T0 = np.eye(4); T0[:3, :3] = dR @ R; T0[:3, 3] = dR @ t + dt
    return dict(T0=T0.astype(np.float32), T_true=T_true, xy=xy.astype(np.float32), octave=octave, Xw=Xw, has=has, bad=bad)
