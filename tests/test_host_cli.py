"""CPU tests of the host-side C++ mirrors of the reference CLIs (pilotguru_b200/host): flag handling and CHECK
behaviour of fit_motion (src/fit_motion.cc:47-104,303-313) and the JSON wire format (src/io/json_converters.cc:172-202).
No compute call is made here (there is no GPU and no CPU fallback)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pilotguru_b200", "host")


@pytest.fixture(scope="module")
def host_bins():
    if not os.path.exists(os.path.join(ROOT, "pilotguru_b200", "libpgb200.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "pilotguru_b200", "csrc")], check=True, capture_output=True)
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    return HOST


def run(bin_dir, name, *args):
    return subprocess.run([os.path.join(bin_dir, name), *args], capture_output=True, text=True, timeout=120)


def test_fit_motion_requires_its_inputs(host_bins):
    p = run(host_bins, "fit_motion")
    assert p.returncode == -6 and "Check failed: !rotations_json.empty()" in p.stderr      # CHECK -> abort
    p = run(host_bins, "fit_motion", "--rotations_json=a", "--accelerations_json=b", "--locations_json=c",
            "--locations_batch_size=3", "--locations_shift_step=5")
    assert p.returncode == -6 and "locations_batch_size >= locations_shift_step" in p.stderr
    p = run(host_bins, "fit_motion", "--rotations_json=a", "--accelerations_json=b", "--locations_json=c",
            "--optimization_iters", "0")
    assert p.returncode == -6 and "optimization_iters > 0" in p.stderr
    p = run(host_bins, "fit_motion", "--no_such_flag=1")
    assert p.returncode == 1 and "unknown command line flag 'no_such_flag'" in p.stderr    # gflags behaviour
    p = run(host_bins, "fit_motion", "--rotations_json=/nonexistent/r.json", "--accelerations_json=b", "--locations_json=c")
    assert p.returncode == -6 and "cannot open JSON file" in p.stderr


def test_json_round_trip(host_bins, tmp_path):
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.normal(0, 10, 50), [0.0, 1.0, -2.5e-310, 1e300, 123456789.125]])
    ts = (np.arange(len(vals), dtype=np.int64) * 10007 + 1_700_000_000_000_000)             # epoch microseconds
    src = {"locations": [{"lat": 1.5, "lon": -2.0, "nested": {"a": [1, 2, {"b": 'x"y'}]}, "speed_m_s": float(v),
                          "time_usec": int(t), "accuracy_m": 3} for v, t in zip(vals, ts)], "other": [1, 2, 3]}
    fin, fout = tmp_path / "in.json", tmp_path / "out.json"
    fin.write_text(json.dumps(src, indent=2))
    p = run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "velocities", "speed_m_s")
    assert p.returncode == 0, p.stderr
    text = fout.read_text()
    out = json.loads(text)
    assert list(out.keys()) == ["velocities"] and len(out["velocities"]) == len(vals)
    assert [e["time_usec"] for e in out["velocities"]] == ts.tolist()
    assert np.array_equal(np.array([e["speed_m_s"] for e in out["velocities"]]), vals)      # exact round trip
    # dump(2) layout: two-space indent, alphabetical keys inside a record (speed_m_s < time_usec)
    assert text.startswith('{\n  "velocities": [\n    {\n      "speed_m_s": ')
    # "steering" records put angular_velocity before time_usec
    p = run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "steering", "angular_velocity")
    assert p.returncode == 0 and '{\n      "angular_velocity": ' in fout.read_text()
    # missing field / empty table are fatal (nlohmann type_error / CHECK(!empty))
    fin.write_text(json.dumps({"locations": [{"time_usec": 1}]}))
    assert run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "v", "s").returncode == -6
    fin.write_text(json.dumps({"locations": []}))
    p = run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "v", "s")
    assert p.returncode == -6 and "is empty" in p.stderr
