"""CPU tests of the oracle's TimeAveragedValue restatement (include/interpolation/time_series.hpp:103-225,
src/annotate_frames.cc:59-72) against an independent exact-integration formula in numpy."""
import numpy as np
import pytest

import oracle_lib as O


def _exact(values, t, a, b):
    """integral of the piecewise-linear interpolant over [a, b] / (b - a), by dense trapezoids on the merged knots"""
    ts = t * 1e-6
    knots = np.unique(np.concatenate([[a * 1e-6, b * 1e-6], ts[(t > a) & (t < b)]]))
    vals = np.interp(knots, ts, values)
    return float(np.sum(0.5 * (vals[1:] + vals[:-1]) * np.diff(knots)) / ((b - a) * 1e-6))


def test_time_averaged_values_match_exact_integral():
    rng = np.random.default_rng(2)
    t = np.cumsum(rng.integers(1500, 2500, 4000)).astype(np.int64) + 1_000_000          # ~500 Hz, jittered
    v = np.sin(t * 2e-6) * 5 + rng.normal(0, 0.2, len(t))
    ft = (np.arange(0, 300) * 33_333 + 600_000).astype(np.int64)                         # starts before, ends after
    out, ok = O.time_averaged_values(v, t, ft)
    assert len(out) == len(ft) - 1
    cover = (ft[:-1] >= t[0]) & (ft[1:] <= t[-1])
    assert np.array_equal(ok, cover) and cover.any() and (~cover).any() and np.isnan(out[~cover]).all()
    for i in np.nonzero(cover)[0]:
        assert abs(out[i] - _exact(v, t, ft[i], ft[i + 1])) <= 1e-9 * max(1.0, abs(out[i]))
    # a frame interval inside one series interval, and frames hitting event times exactly
    out2, ok2 = O.time_averaged_values(v, t, np.array([t[10] + 100, t[10] + 300, t[11], t[13], t[13] + 1], np.int64))
    assert ok2.all()
    for i, (a, b) in enumerate([(t[10] + 100, t[10] + 300), (t[10] + 300, t[11]), (t[11], t[13]), (t[13], t[13] + 1)]):
        assert abs(out2[i] - _exact(v, t, a, b)) <= 1e-9


def test_time_averaged_values_fatal_cases():
    t = np.array([0, 1000, 2000, 3000], np.int64); v = np.array([1.0, 2.0, 3.0, 4.0])
    with pytest.raises(ValueError):
        O.time_averaged_values(v, t, np.array([500, 500], np.int64))       # CHECK_GT(end, start)
    with pytest.raises(ValueError):
        O.time_averaged_values(v, t, np.array([500, 3000], np.int64))      # ends ON the last event: LinearInterpolate CHECK
    out, ok = O.time_averaged_values(v, t, np.array([500, 2999], np.int64))
    assert ok.all() and abs(out[0] - _exact(v, t, 500, 2999)) < 1e-12
