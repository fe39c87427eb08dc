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


def _np_finish(poses, sigma):
    """numpy restatement of track_image_sequence.cc:63-109 (+ smoothing.cc:11-46, horizontal_flatten.cc:7-64)"""
    t = poses[:, 2:5].copy(); q = poses[:, 5:9].copy()
    if sigma > 0:
        ks = 4 * sigma + 1
        x = np.arange(ks) - (ks - 1) / 2
        k = np.exp(-0.5 * x * x / (sigma * sigma)); k /= k.sum()
        idx = np.clip(np.arange(len(q))[:, None] + np.arange(ks)[None, :] - ks // 2, 0, len(q) - 1)
        q = (q[idx] * k[None, :, None]).sum(axis=1)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    import oracle_lib as O
    vec, val, _ = O.pca3(t)
    plane = vec[:2]
    w, v = q[:, 0], q[:, 1:]
    z = np.array([0.0, 0.0, 1.0])
    uv = 2 * np.cross(v, z)
    d = z + w[:, None] * uv + np.cross(v, uv)
    dirs = d @ plane.T
    turn = np.zeros(len(dirs))
    for i in range(1, len(dirs)):
        p, c = dirs[i - 1], dirs[i]
        cs = p.dot(c) / np.linalg.norm(p) / np.linalg.norm(c)
        turn[i] = np.arccos(cs) * (1.0 if p[0] * c[1] - p[1] * c[0] > 0 else -1.0)
    return plane, dirs, turn, q, val


@pytest.mark.parametrize("sigma", [-1, 2])
def test_trajectory_postprocessing(host_bins, tmp_path, sigma):
    rng = np.random.default_rng(7)
    n = 60
    s = np.linspace(0, 3, n)
    yaw = 0.6 * np.sin(s) + 0.02 * rng.normal(size=n)
    pos = np.stack([np.cumsum(np.sin(yaw)), 1e-4 * rng.normal(size=n), np.cumsum(np.cos(yaw))], axis=1)
    quat = np.stack([np.cos(yaw / 2), np.zeros(n), np.sin(yaw / 2), np.zeros(n)], axis=1)
    ts = (np.arange(n) * 33333).astype(np.int64)
    poses = np.concatenate([ts[:, None].astype(float), np.arange(n)[:, None].astype(float), pos, quat], axis=1)
    fin, fout = tmp_path / "poses.json", tmp_path / "traj.json"
    fin.write_text(json.dumps({"poses": [[int(r[0]), int(r[1])] + [float(x) for x in r[2:]] for r in poses]}))
    p = run(host_bins, "trajectory_selftest", str(fin), str(sigma), str(fout))
    assert p.returncode == 0, p.stderr
    out = json.loads(fout.read_text())
    plane, dirs, turn, q, val = _np_finish(poses, sigma)
    assert sorted(out.keys()) == ["plane", "trajectory"] and len(out["trajectory"]) == n
    assert np.max(np.abs(np.array(out["plane"]) - plane)) <= 1e-12
    e0 = out["trajectory"][0]
    assert sorted(e0.keys()) == ["angular_velocity", "frame_id", "is_lost", "planar_direction", "pose", "time_usec"]
    assert e0["angular_velocity"] == 0 and e0["is_lost"] is False
    for i, e in enumerate(out["trajectory"]):
        assert e["frame_id"] == i and e["time_usec"] == int(ts[i])
        assert np.max(np.abs(np.array(e["planar_direction"]) - dirs[i])) <= 1e-12
        r = e["pose"]["rotation"]
        assert np.max(np.abs(np.array([r["w"], r["x"], r["y"], r["z"]]) - q[i])) <= 1e-12
        assert np.array_equal(np.array(e["pose"]["translation"]), pos[i])
        if i:
            assert abs(e["angular_velocity"] - turn[i] / ((ts[i] - ts[i - 1]) * 1e-6 + 1e-10)) <= 1e-8 * max(1.0, abs(e["angular_velocity"]))
    # a trajectory with real vertical motion is dropped (3rd eigenvalue gate) and nothing is written
    pos[:, 1] = 5.0 * np.sin(3 * s)
    poses[:, 2:5] = pos
    fin.write_text(json.dumps({"poses": [[int(r[0]), int(r[1])] + [float(x) for x in r[2:]] for r in poses]}))
    fout.unlink()
    p = run(host_bins, "trajectory_selftest", str(fin), str(sigma), str(fout))
    assert p.returncode == 3 and "3rd eigenvalue was too large" in p.stderr and not fout.exists()


def test_optical_trajectories_flags(host_bins):
    p = run(host_bins, "optical_trajectories")
    assert p.returncode == -6 and "Check failed: !vocabulary_file.empty()" in p.stderr
    p = run(host_bins, "optical_trajectories", "--vocabulary_file=v", "--camera_settings=/nonexistent.yml", "--in_video=video.mp4")
    assert p.returncode == -6 and "raw:<path>:<w>x<h>" in p.stderr and "synth:<canvas>" in p.stderr


def test_json_number_forms(host_bins, tmp_path):
    """The reader's number path (std::from_chars): exponents, negative zero, integer literals for a real field, an int64
    beyond 2^53 kept exact, a real literal for the integer field, odd whitespace."""
    text = ('{ "locations" :\t[\n'
            ' {"speed_m_s": 1E+2, "time_usec": 9007199254740993},\n'
            ' {"time_usec":2,"speed_m_s":-0.0},\n'
            ' {"speed_m_s": 3, "time_usec": 3},\n'
            ' {"speed_m_s": 2.5e-3 , "time_usec": 1.5e3 , "extra": [1, {"a": null}, true, "s\\"q"]},\n'
            ' {"speed_m_s": 0.1, "time_usec": -7}\n'
            ']}')
    fin, fout = tmp_path / "in.json", tmp_path / "out.json"
    fin.write_text(text)
    p = run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "v", "speed_m_s")
    assert p.returncode == 0, p.stderr
    out = json.loads(fout.read_text())["v"]
    assert [e["time_usec"] for e in out] == [9007199254740993, 2, 3, 1500, -7]
    vals = [e["speed_m_s"] for e in out]
    assert vals == [100.0, 0.0, 3.0, 0.0025, 0.1] and np.signbit(vals[1])
    fin.write_text('{"locations": [{"speed_m_s": abc, "time_usec": 1}]}')
    assert run(host_bins, "json_selftest", str(fin), "locations", "speed_m_s", str(fout), "v", "s").returncode == -6


def test_json_writer_prints_percent_17g(host_bins, tmp_path):
    """The writer's number text is exactly printf's %.17g (+ ".0" for integral values, like nlohmann's dump)."""
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.normal(0, 1, 300) * 10.0 ** rng.integers(-300, 300, 300), [0.0, -0.0, 1.0, -3.0, 1e22, 1e-5, 123456.0,
                           5e-324, 1.7976931348623157e308, 0.1, 1 / 3]])
    ts = np.arange(len(vals), dtype=np.int64) - 5
    fin, fout = tmp_path / "in.json", tmp_path / "out.json"
    fin.write_text(json.dumps({"t": [{"v": float(v), "time_usec": int(t)} for v, t in zip(vals, ts)]}))
    p = run(host_bins, "json_selftest", str(fin), "t", "v", str(fout), "root", "v")
    assert p.returncode == 0, p.stderr
    lines = fout.read_text().splitlines()
    got = [l.split(": ", 1)[1].rstrip(",") for l in lines if l.strip().startswith('"v"')]
    want = []
    for v in vals:
        s = "%.17g" % v
        want.append(s if any(c in s for c in ".e") else s + ".0")
    assert got == want
    assert [int(l.split(": ", 1)[1].rstrip(",")) for l in lines if l.strip().startswith('"time_usec"')] == ts.tolist()


def test_json_reader_differential_fuzz(host_bins, tmp_path):
    """Random documents (nested junk values, escapes in strings and keys, varying whitespace and separators) through
    the C++ pull parser and back: same columns as Python's json gives."""
    import random
    rnd = random.Random(1234)

    def junk(depth=0):
        k = rnd.randrange(7 if depth < 3 else 4)
        if k == 0: return rnd.uniform(-1e6, 1e6)
        if k == 1: return rnd.randrange(-10**12, 10**12)
        if k == 2: return rnd.choice([True, False, None])
        if k == 3: return "".join(rnd.choice('ab"\\/ \t{}[],:é') for _ in range(rnd.randrange(6)))
        if k == 4: return [junk(depth + 1) for _ in range(rnd.randrange(4))]
        return {"".join(rnd.choice('xy"\\z') for _ in range(1 + rnd.randrange(4))): junk(depth + 1) for _ in range(rnd.randrange(4))}

    for case in range(40):
        n = rnd.randrange(1, 30)
        rows = []
        for i in range(n):
            rec = {"junk%d" % j: junk() for j in range(rnd.randrange(3))}
            rec["val"] = rnd.choice([rnd.uniform(-1, 1) * 10.0 ** rnd.randrange(-20, 20), float(rnd.randrange(-5, 5)), rnd.randrange(-1000, 1000)])
            rec["time_usec"] = rnd.randrange(-10**15, 10**15)
            items = list(rec.items()); rnd.shuffle(items)
            rows.append(dict(items))
        doc = {"before": junk(), "table": rows, "after": junk()}
        items = list(doc.items()); rnd.shuffle(items)
        text = json.dumps(dict(items), indent=rnd.choice([None, 0, 1, 3]), separators=rnd.choice([(",", ":"), (", ", ": "), (" ,\t", " :\n")]),
                          ensure_ascii=rnd.choice([True, False]))
        fin, fout = tmp_path / f"in{case}.json", tmp_path / f"out{case}.json"
        fin.write_text(text, encoding="utf-8")
        p = run(host_bins, "json_selftest", str(fin), "table", "val", str(fout), "r", "val")
        assert p.returncode == 0, (case, p.stderr, text[:300])
        out = json.loads(fout.read_text())["r"]
        assert [e["time_usec"] for e in out] == [r["time_usec"] for r in rows], case
        assert [e["val"] for e in out] == [float(r["val"]) for r in rows], case
