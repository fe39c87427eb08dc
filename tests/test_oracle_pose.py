"""CPU tests of the PoseOptimization oracle (Optimizer.cc:239-451 + the g2o Levenberg-Marquardt it drives).  The
reference holds no vectors for it; pins: recovery of the true pose, the inlier/outlier split, and an independent scipy
least-squares fit of the last round's problem."""
import numpy as np
from scipy.optimize import least_squares

import oracle_lib as O
import pose_util as U


def _run(S):
    return O.pose_optimization(S["T0"], S["xy"], S["octave"], S["Xw"], S["has"], U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)


def test_pose_is_recovered_and_outliers_are_found():
    for seed in range(6):
        S = U.scene(seed)
        n_in, T, out, rounds = _run(S)
        has = S["has"].astype(bool)
        assert n_in == int((has & (out == 0)).sum())
        assert (out[~has] == 0).all()
        # gross outliers are rejected, clean observations are kept
        assert out[has & S["bad"]].mean() > 0.95 and out[has & ~S["bad"]].mean() < 0.08
        dR = T[:3, :3].astype(np.float64) @ S["T_true"][:3, :3].T
        ang = np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
        assert ang < 0.1 and np.abs(T[:3, 3] - S["T_true"][:3, 3]).max() < 0.05
        assert np.array_equal(T[3], [0, 0, 0, 1])
        assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)


def test_last_round_agrees_with_scipy_least_squares():
    """Round 4 runs without the robust kernel on the inliers of round 3, from the initial pose: its fixed point is the
    plain weighted least-squares optimum over that set, which scipy finds independently."""
    S = U.scene(11)
    n_in, T, out, rounds = _run(S)
    sel = S["has"].astype(bool) & (rounds[2] == 0)
    X = S["Xw"][sel].astype(np.float64); xy = S["xy"][sel].astype(np.float64); w = np.sqrt(U.INV_SIGMA2[S["octave"][sel]].astype(np.float64))
    T0 = S["T0"].astype(np.float64)

    def res(p):
        R = U.rodrigues(p[:3]) @ T0[:3, :3]; t = U.rodrigues(p[:3]) @ T0[:3, 3] + p[3:]
        Xc = X @ R.T + t
        r = np.stack([xy[:, 0] - (Xc[:, 0] / Xc[:, 2] * U.FX + U.CX), xy[:, 1] - (Xc[:, 1] / Xc[:, 2] * U.FY + U.CY)], axis=1)
        return (r * w[:, None]).ravel()
    sol = least_squares(res, np.zeros(6), xtol=1e-14, ftol=1e-14, gtol=1e-14)
    R = U.rodrigues(sol.x[:3]) @ T0[:3, :3]; t = U.rodrigues(sol.x[:3]) @ T0[:3, 3] + sol.x[3:]
    assert np.abs(T[:3, :3] - R).max() < 2e-6 and np.abs(T[:3, 3] - t).max() < 2e-5


def test_edge_cases():
    S = U.scene(3, n=50, n_mp=2)                       # fewer than 3 correspondences: pose untouched, returns 0
    n_in, T, out, rounds = _run(S)
    assert n_in == 0 and np.array_equal(T, S["T0"]) and (out == 0).all() and (rounds == 255).all()
    S = U.scene(4, n=60, n_mp=8, outlier_frac=0.0)     # fewer than 10 edges: a single round (Optimizer.cc:438-439)
    n_in, T, out, rounds = _run(S)
    assert (rounds[0] != 255).all() and (rounds[1:] == 255).all() and n_in >= 6
    S = U.scene(5, n=0)
    n_in, T, out, rounds = _run(S)
    assert n_in == 0 and np.array_equal(T, S["T0"])
    # already at the optimum: starting from the answer returns (almost) the same pose
    S = U.scene(6, outlier_frac=0.1)
    n1, T1, o1, _ = _run(S)
    S2 = dict(S); S2["T0"] = T1
    n2, T2, o2, _ = _run(S2)
    assert np.abs(T2 - T1).max() < 1e-5 and abs(n2 - n1) <= 1


def test_degenerate_inputs_terminate():
    """Points behind the camera, a rank-deficient scene, exact data and z = 0: every loop of the restated LM is bounded, so
    the call returns; NaN chi2 compares false against the threshold exactly like the reference's `if(chi2>chi2Mono[it])`."""
    S = U.scene(21, n=300, outlier_frac=0.1)
    X = S["Xw"].copy(); X[:20] *= -1
    n_in, T, out, _ = O.pose_optimization(S["T0"], S["xy"], S["octave"], X, S["has"], U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)
    assert np.isfinite(T).all() and out[:20][S["has"][:20] > 0].mean() > 0.8 and n_in > 100
    S = U.scene(22, n=200, outlier_frac=0.0, noise=0.0, perturb=(0.0, 0.0))      # exact data, start at the truth
    n_in, T, out, _ = _run(S)
    assert n_in == int(S["has"].sum()) and np.abs(T - S["T0"]).max() < 1e-6
    S = U.scene(23, n=50, outlier_frac=0.0)                                       # all correspondences identical
    X = np.tile(S["Xw"][:1], (50, 1)); xy = np.tile(S["xy"][:1], (50, 1))
    n_in, T, out, _ = O.pose_optimization(S["T0"], xy, S["octave"], X, S["has"], U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)
    assert np.isfinite(T).all()
    n_in, T, out, _ = O.pose_optimization(np.eye(4, dtype=np.float32), S["xy"], S["octave"], np.zeros_like(S["Xw"]), S["has"],
                                          U.INV_SIGMA2, U.FX, U.FY, U.CX, U.CY)   # z = 0: NaN errors
    assert n_in == int(S["has"].sum()) and (out == 0).all()
