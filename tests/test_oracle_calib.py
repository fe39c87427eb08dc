"""CPU tests of the calibration oracle: the reference's doc-comment example, an independent numpy restatement of
eval(), finite-difference sanity of the (deliberately inexact) gradient structure, agreement between the literal
restatement and the arithmetic-contract ("core") evaluation, and the two L-BFGS codings."""
import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth


def test_merge_doc_example():
    """include/interpolation/align_time_series.hpp:17-26."""
    c = O.CalibOracle([1.0, 1.0], [0, 100], np.zeros((5, 3)), [1, 3, 4, 6, 7], np.zeros((5, 3)), [2, 3, 4, 5, 6])
    t, gi, ai = c.merged()
    assert t.tolist() == [2, 3, 4, 5, 6]
    assert list(zip(gi.tolist(), ai.tolist())) == [(0, 0), (1, 1), (2, 2), (2, 3), (3, 4)]


def test_intervals_structure():
    d = synth.imu_gps(12, 100)
    c = O.CalibOracle(d["gps_v"], d["gps_t"], d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    ref, m, s, e = c.intervals()
    assert ref.min() == 1                      # the first reference interval is always empty
    assert np.all(e > s) and np.all(s[1:] >= s[:-1])
    # per GPS interval the pieces tile (gps[r-1], gps[r]] exactly
    for r in range(1, len(d["gps_t"])):
        sel = ref == r
        assert s[sel][0] == d["gps_t"][r - 1] and e[sel][-1] == d["gps_t"][r]
        assert np.all(s[sel][1:] == e[sel][:-1])
    # interleaved sensors: merged events alternate and partial intervals appear at GPS boundaries
    d2 = synth.imu_gps(6, 100, interleaved=True)
    c2 = O.CalibOracle(d2["gps_v"], d2["gps_t"], d2["gyro"], d2["gyro_t"], d2["acc"], d2["acc_t"])
    t, gi, ai = c2.merged()
    assert len(t) == 2 * len(d2["gyro_t"]) - 2


def numpy_eval(d, x):
    """velocity.cc:41-180 restated with numpy quaternion helpers, independently of oracle/pgo_calib.cc."""
    g, h, v = x[0:3].copy(), x[3:6].copy(), x[6:9].copy()
    def qmul(a, b):
        return np.array([a[0]*b[0]-a[1]*b[1]-a[2]*b[2]-a[3]*b[3], a[0]*b[1]+a[1]*b[0]+a[2]*b[3]-a[3]*b[2],
                         a[0]*b[2]+a[2]*b[0]+a[3]*b[1]-a[1]*b[3], a[0]*b[3]+a[3]*b[0]+a[1]*b[2]-a[2]*b[1]])
    def rot(q):
        w, x_, y, z = q
        return np.array([[1-2*(y*y+z*z), 2*(x_*y-z*w), 2*(x_*z+y*w)], [2*(x_*y+z*w), 1-2*(x_*x_+z*z), 2*(y*z-x_*w)],
                         [2*(x_*z-y*w), 2*(y*z+x_*w), 1-2*(x_*x_+y*y)]])
    gt, gv = d["gps_t"], d["gps_v"]
    it = d["gyro_t"]                      # aligned sensors: merged events == samples
    q = np.array([1.0, 0, 0, 0]); Wm = np.zeros((3, 3)); tau = 0; loss = 0.0; grad = np.zeros(9)
    k = int(np.searchsorted(it, gt[0], side="right"))
    for r in range(1, len(gt)):
        D = np.zeros(3); dref = 0.0; outs = []
        last = gt[r - 1]
        while True:
            end = it[k] if (k < len(it) and it[k] <= gt[r]) else gt[r]
            if end > last:
                if k >= len(it):
                    break
                dt = (end - last) * 1e-6
                w = d["gyro"][k]; a = d["acc"][k]
                v = v + (rot(q) @ (a + h) + g) * dt
                n = np.linalg.norm(w); ht = n * dt * 0.5
                q = qmul(q, np.concatenate([[np.cos(ht)], w * (np.sin(ht) / (n + 1e-30))]))
                D += dt * v; dref += dt * gv[r]; outs.append((q.copy(), end - last))
            last = end
            if k < len(it) and it[k] <= gt[r]:
                k += 1
            else:
                break
        dn = np.linalg.norm(D); e = dn - dref; loss += e * e
        dL = 2 * e * D / (dn + 1e-5)
        for qq, du in outs:
            isec = du * 1e-6; tau += du; ts = tau * 1e-6
            grad[0:3] += ts * isec * dL
            Wm += rot(qq) * isec
            grad[3:6] += isec * (Wm.T @ dL)
            grad[6:9] += isec * dL
    T = tau * 1e-6
    return loss / T, grad / T


def test_eval_against_numpy_restatement():
    d = synth.imu_gps(8, 100)
    c = O.CalibOracle(d["gps_v"], d["gps_t"], d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.normal(0, 2, 9)
        f, g = c.eval(x)
        fn, gn = numpy_eval(d, x)
        assert f == pytest.approx(fn, rel=1e-11) and np.allclose(g, gn, rtol=1e-9, atol=1e-12)


def test_v0_gradient_is_the_true_derivative():
    """Of the three gradient blocks only d/dv0 is an exact derivative (the g and h blocks use inclusive cumulative
    time and post-step rotations): a central difference must reproduce it."""
    d = synth.imu_gps(8, 100)
    c = O.CalibOracle(d["gps_v"], d["gps_t"], d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    x = np.array([0.1, -0.2, -9.6, 0.05, 0.1, -0.1, 7.0, 0.5, 0.2])
    _, g = c.eval(x)
    for i in (6, 7, 8):
        xp, xm = x.copy(), x.copy(); xp[i] += 1e-6; xm[i] -= 1e-6
        fd = (c.eval(xp)[0] - c.eval(xm)[0]) / 2e-6
        assert g[i] == pytest.approx(fd, rel=1e-5)


@pytest.mark.parametrize("interleaved", [False, True])
def test_literal_vs_contract_evaluation(interleaved):
    d = synth.imu_gps(45, 100, interleaved=interleaved)
    c = O.CalibOracle(d["gps_v"][:40], d["gps_t"][:40], d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    rng = np.random.default_rng(1)
    for k in range(4):
        x = rng.normal(0, 1, 9) if k else np.zeros(9)
        f0, g0 = c.eval(x); f1, g1 = c.eval(x, core=True)
        assert abs(f0 - f1) <= 1e-12 * abs(f0)
        assert np.max(np.abs(g0 - g1)) <= 1e-11 * np.max(np.abs(g0))
    i0, s0, q0, v0, d0 = c.integrate(x); i1, s1, _, v1, d1 = c.integrate(x, core=True)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert np.max(np.abs(s0 - s1) / s0) < 1e-11


def test_two_lbfgs_codings_agree_bitwise():
    d = synth.imu_gps(30, 100)
    c = O.CalibOracle(d["gps_v"], d["gps_t"], d["gyro"], d["gyro_t"], d["acc"], d["acc_t"])
    a = c.minimize(mode="core", max_iterations=120)
    b = c.minimize(mode="literal_driver_core_eval", max_iterations=120)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
    it, x, fx, ne = c.minimize(mode="literal", max_iterations=120)
    assert fx < c.eval(np.zeros(9))[0] * 1e-3        # the optimiser actually optimises
    assert ne >= it


def test_det_sincos_matches_libm():
    """The contract's deterministic sincos against libm over the range RotationMotionToQuaternion can see."""
    import ctypes as C
    l = O.lib()
    if not hasattr(l, "pgo_det_sincos"):
        pytest.skip("not exported")
    l.pgo_det_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-10, 10, 20000), rng.uniform(-1e-3, 1e-3, 5000), [0.0, np.pi / 4, -np.pi / 4, 1e5]])
    worst = 0.0
    for v in xs:
        s = C.c_double(); c = C.c_double()
        l.pgo_det_sincos(float(v), C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(v)), abs(c.value - np.cos(v)))
    assert worst < 4e-16


def test_smoothing_properties():
    t = np.arange(0, 2.0, 0.01)
    const = O.smooth_time_series(np.full(len(t), 3.25), t, t, 0.003)
    assert np.allclose(const, 3.25, rtol=0, atol=1e-12)          # weights sum to 1
    v = np.sin(t * 3)
    sm = O.smooth_time_series(v, t, t, 0.003)
    assert np.max(np.abs(sm - v)) < 0.02                         # sigma << sample spacing: nearly the identity
    wide = O.smooth_time_series(v, t, t, 0.2)
    assert np.std(wide) < np.std(v)


def test_fit_motion_chaos_is_documented_not_hidden():
    """literal and contract evaluations agree to 1e-12 per call, yet 500 L-BFGS iterations on the
    flat g-vs-h valley amplify that to ~1e-2 in the final speeds.  The gate for the CUDA path is therefore bit
    equality with the contract evaluation; this test records the literal-vs-contract deviation."""
    d = synth.imu_gps(60, 100)
    lit = O.fit_motion(d, mode=0); core = O.fit_motion(d, mode=1)
    assert np.array_equal(lit["idx"], core["idx"]) and len(lit["idx"]) == 5900
    dev = np.max(np.abs(lit["smoothed"] - core["smoothed"]) / np.abs(lit["smoothed"]))
    assert dev < 0.1
    assert (lit["iters"] == 500).sum() >= 8      # most windows hit the iteration cap, as the survey observed


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.mark.parametrize("iters", [5, 10, 20])
def test_contract_velocities_within_1e6_of_literal_and_reference_source(iters):
    """BASELINE north_star tolerance (1e-6 relative on the calibrated velocities), asserted where it is definable: below the
    objective's chaos horizon, i.e. with the per-window L-BFGS capped at 5 / 10 / 20 iterations, the
    arithmetic contract the CUDA kernels implement (mode 1; the GPU equals it bit for bit, tests/test_gpu_calib.py) is within
    1e-6 of the literal sequential restatement (mode 0) AND of the reference's own fit_motion.cc window loop compiled in
    place (oracle/_ref), on BASELINE configs[0] (60 s, 100 Hz, 12 windows)."""
    d = synth.imu_gps(60, 100)
    lit = O.fit_motion(d, max_iters=iters, mode=0)
    con = O.fit_motion(d, max_iters=iters, mode=1)
    assert np.array_equal(lit["t_usec"], con["t_usec"])
    assert _rel(con["smoothed"], lit["smoothed"]) <= 1e-6
    ref = O.ref_fit_motion(d, max_iters=iters)
    if ref is not None:
        assert np.array_equal(ref[0], con["t_usec"])
        assert _rel(con["smoothed"], ref[1]) <= 1e-6 and _rel(lit["smoothed"], ref[1]) <= 1e-6


def test_500_iteration_envelope():
    """At fit_motion's real setting (500 iterations) the final velocities are chaotic in the last bits of the objective: the
    literal restatement and the reference's own source -- two renderings of the SAME sequential arithmetic -- already differ
    by ~1e-2.  That spread is the reference's reproducibility envelope; the contract (= the GPU) has to sit inside a small
    multiple of it.  Both numbers are printed (DESIGN.md section 5 quotes them)."""
    d = synth.imu_gps(60, 100)
    lit = O.fit_motion(d, max_iters=500, mode=0)
    con = O.fit_motion(d, max_iters=500, mode=1)
    dev = _rel(con["smoothed"], lit["smoothed"])
    print(f"500 iterations: contract vs literal {dev:.3e}")
    ref = O.ref_fit_motion(d, max_iters=500)
    if ref is None:
        pytest.skip("oracle/_ref unavailable")
    env = _rel(lit["smoothed"], ref[1])
    dev_ref = _rel(con["smoothed"], ref[1])
    print(f"500 iterations: literal vs reference source {env:.3e} (envelope), contract vs reference source {dev_ref:.3e}")
    assert env > 1e-6, "the envelope collapsed: the 1e-6 bar would be definable at 500 iterations -- tighten this test"
    assert dev <= 3 * env and dev_ref <= 3 * env
