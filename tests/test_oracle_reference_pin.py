"""The one piece of the path whose REAL reference source compiles here -- src/interpolation/align_time_series.cc
(MergedTimeSeries, MakeInterpolationIntervals; its only external dependency is glog's CHECK macros, supplied by
oracle/ref_shims) -- built where it lies into oracle/_ref/ (`make -C oracle _ref`) and run against the oracle's
restatement on the same inputs.  Everything downstream of these indices (the calibration objective, the GPU kernels)
is tested against the oracle, so this pins row a20 of SURVEY section 8 to the reference itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libpilotguru_ref.so")
i64p = C.POINTER(C.c_int64)


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], check=True, capture_output=True)
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    l = C.CDLL(SO)
    for f in ("pgr_merge_two", "pgr_make_interpolation_intervals", "pgr_num_reference_rows"):
        getattr(l, f).restype = C.c_int64
    return l


def _p(a):
    return a.ctypes.data_as(i64p)


def _ref_merge(l, a, b):
    a = np.ascontiguousarray(a, np.int64); b = np.ascontiguousarray(b, np.int64)
    cap = len(a) + len(b) + 4
    ia, ib, t = (np.empty(cap, np.int64) for _ in range(3))
    n = l.pgr_merge_two(_p(a), C.c_int64(len(a)), _p(b), C.c_int64(len(b)), _p(ia), _p(ib), _p(t), C.c_int64(cap))
    return ia[:n], ib[:n], t[:n]


def _ref_intervals(l, ref_t, interp_t):
    ref_t = np.ascontiguousarray(ref_t, np.int64); interp_t = np.ascontiguousarray(interp_t, np.int64)
    cap = len(ref_t) + len(interp_t) + 8
    out = [np.empty(cap, np.int64) for _ in range(4)]
    rows = l.pgr_num_reference_rows(_p(ref_t), C.c_int64(len(ref_t)), _p(interp_t), C.c_int64(len(interp_t)))
    per = np.zeros(max(rows, 1), np.int64)
    n = l.pgr_make_interpolation_intervals(_p(ref_t), C.c_int64(len(ref_t)), _p(interp_t), C.c_int64(len(interp_t)),
                                           *[_p(x) for x in out], C.c_int64(cap), _p(per))
    assert n <= cap
    return [x[:n] for x in out], per[:rows]


def test_doc_comment_example(ref):
    ia, ib, t = _ref_merge(ref, [1, 3, 4, 6, 7], [2, 3, 4, 5, 6])          # align_time_series.hpp:17-26
    assert list(zip(ia.tolist(), ib.tolist(), t.tolist())) == [(0, 0, 2), (1, 1, 3), (2, 2, 4), (2, 3, 5), (3, 4, 6)]


@pytest.mark.parametrize("case", ["shared", "interleaved", "jitter", "sparse_gps", "c1"])
def test_oracle_indices_equal_the_reference(ref, case):
    rng = np.random.default_rng(17)
    if case == "c1":
        d = synth.imu_gps(60.0, 100.0, seed=11)
        gt, at, pt = d["gyro_t"], d["acc_t"], d["gps_t"][:40]
        gyro, acc, gv = d["gyro"], d["acc"], d["gps_v"][:40]
    else:
        n = 4000
        gt = np.cumsum(rng.integers(1500, 2600, n)).astype(np.int64)
        if case == "shared":
            at = gt.copy()
        elif case == "interleaved":
            at = gt[:-1] + (np.diff(gt) // 2)
        else:
            at = np.unique(np.cumsum(rng.integers(900, 4100, n)).astype(np.int64) + int(gt[0]) // 2)
        step = 1_000_000 if case != "sparse_gps" else 2_700_000
        pt = np.arange(int(max(gt[0], at[0])) + 300_000, int(min(gt[-1], at[-1])) - 300_000, step, dtype=np.int64)[:40]
        gyro = rng.normal(0, 0.1, (len(gt), 3)); acc = rng.normal(0, 1, (len(at), 3)); gv = rng.uniform(5, 15, len(pt))
    orc = O.CalibOracle(gv, pt, gyro, gt, acc, at)
    ot, ogi, oai = orc.merged()
    ria, rib, rt = _ref_merge(ref, gt, at)
    assert np.array_equal(ogi, ria) and np.array_equal(oai, rib) and np.array_equal(ot, rt) and len(rt) > 100
    o_ref, o_m, o_s, o_e = orc.intervals()
    (r_ref, r_m, r_s, r_e), per = _ref_intervals(ref, pt, rt)
    assert len(per) == len(pt) and per[0] == 0                               # the first reference interval is always empty
    assert np.array_equal(o_ref, r_ref) and np.array_equal(o_m, r_m) and np.array_equal(o_s, r_s) and np.array_equal(o_e, r_e)
    assert len(r_ref) > 100 and (r_e >= r_s).all()
