"""GPU end-to-end test of the `optical_trajectories` drop-in binary in flow-tracking mode (pilotguru_b200/host):
synthetic frames with a known integer flow -> ORB extraction + projection matching on the B200 through the C-ABI
(host buffers, the reference's calling convention) -> trajectory JSON in the reference's schema
(src/io/json_converters.cc:37-96).  The per-frame keypoint/match counts the binary reports are checked against the
oracle's extraction + matching of the same frames."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_optical_trajectories_binary(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "pilotguru_b200", "host")], check=True, capture_output=True)
    w, h, n = 640, 480, 21
    frames = np.stack([synth.frame(t, w=w, h=h) for t in range(n)])
    raw = tmp_path / "frames.gray"
    frames.tofile(raw)
    settings = tmp_path / "settings.yml"
    # the fork's settings format: underscore keys, as src/calibrate.cc:504-544 writes them and Tracking.cc:52-135 reads them
    settings.write_text("%YAML:1.0\nCamera_fx: 1.2e+03\nCamera_fy: 1.2e+03\nCamera_cx: 320.\nCamera_cy: 240.\nCamera_k1: 0.\nCamera_k2: 0.\n"
                        "Camera_p1: 0.\nCamera_p2: 0.\nCamera_fps: 25.\nCamera_RGB: 1\nORBextractor_nFeatures: 500\n"
                        "ORBextractor_scaleFactor: 1.2\nORBextractor_nLevels: 8\nORBextractor_iniThFAST: 20\nORBextractor_minThFAST: 7\n"
                        "Viewer_KeyFrameSize: 5.0000000000000003e-02\nViewer_PointSize: 2\n")
    p = subprocess.run([os.path.join(ROOT, "pilotguru_b200", "host", "optical_trajectories"), "--vocabulary_file=unused.txt",
                        "--camera_settings", str(settings), "--out_dir", str(tmp_path), f"--in_video=raw:{raw}:{w}x{h}",
                        "--novisualize", "--batch=8", "--logtostderr"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.load(open(tmp_path / "trajectory-0.json"))
    assert not (tmp_path / "trajectory-1.json").exists()                  # tracking never lost on this sequence
    tr = out["trajectory"]
    assert len(tr) == n and np.array(out["plane"]).shape == (2, 3)
    assert [e["frame_id"] for e in tr] == list(range(n)) and [e["time_usec"] for e in tr] == [int(round(i * 1e6 / 25.0)) for i in range(n)]
    # translation = minus the accumulated image flow (x -> x, y -> z); the synthetic flow is an exact integer shift
    flow = np.array([synth.flow(t, w=w, h=h) if t else (0, 0) for t in range(n)], float)
    want = -np.cumsum(flow, axis=0)
    got = np.array([[e["pose"]["translation"][0], e["pose"]["translation"][2]] for e in tr])
    assert np.max(np.abs(got - want)) <= 0.51, (got - want)
    assert all(e["pose"]["translation"][1] == 0 for e in tr) and not any(e["is_lost"] for e in tr)
    # the counts the binary reports equal the oracle's on the same frames (extraction + zero-velocity-guess matching)
    orc = O.OrbOracle(500, 1.2, 8, 20, 7)
    feats = [orc.extract(f) for f in frames]
    sf = orc.tables()["scale"]
    tot_k = sum(len(k) for k, _ in feats); tot_m = 0
    vflow = np.zeros(2, np.float32)
    for t in range(1, n):
        (pk, pd), (ck, cd) = feats[t - 1], feats[t]
        uv = np.stack([pk["x"] + vflow[0], pk["y"] + vflow[1]], axis=1).astype(np.float32)
        nm, mo, _ = O.search_by_projection(ck, cd, uv, pk["octave"], pk["angle"], pd, np.ones(len(pk), np.uint8), (0, w, 0, h), 15.0, sf)
        if nm < 20:
            nm, mo, _ = O.search_by_projection(ck, cd, uv, pk["octave"], pk["angle"], pd, np.ones(len(pk), np.uint8), (0, w, 0, h), 30.0, sf)
        tot_m += nm
    line = [l for l in p.stderr.splitlines() if "keypoints/frame" in l][-1]
    assert f"{tot_k / n:.1f} keypoints/frame" in line and f"{tot_m / (n - 1):.1f} matches/frame" in line, (line, tot_k / n, tot_m / (n - 1))


def test_optical_trajectories_colour_and_flip_input(tmp_path):
    """raw24 input with --vertical_flip: the device feed (flip + cvtColor) in front of the extractor gives the same
    trajectory as feeding the equivalent gray frames."""
    host = os.path.join(ROOT, "pilotguru_b200", "host")
    subprocess.run(["make", "-C", host], check=True, capture_output=True)
    w, h, n = 320, 240, 9
    gray = np.stack([synth.frame(t, w=w, h=h) for t in range(n)])
    rgb = np.repeat(gray[:, ::-1, :, None], 3, axis=3)                    # R = G = B, stored upside down
    (tmp_path / "s.yml").write_text("%YAML:1.0\nCamera.fps: 30.0\nCamera.RGB: 1\nORBextractor.nFeatures: 300\n")
    gray.tofile(tmp_path / "g.raw"); np.ascontiguousarray(rgb).tofile(tmp_path / "c.raw")
    outs = []
    for spec, extra, sub in ((f"raw:{tmp_path / 'g.raw'}:{w}x{h}", [], "a"), (f"raw24:{tmp_path / 'c.raw'}:{w}x{h}", ["--vertical_flip"], "b")):
        os.makedirs(tmp_path / sub)
        p = subprocess.run([os.path.join(host, "optical_trajectories"), "--vocabulary_file=x", "--camera_settings", str(tmp_path / "s.yml"),
                            "--out_dir", str(tmp_path / sub), "--in_video=" + spec] + extra, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-1500:]
        outs.append(json.load(open(tmp_path / sub / "trajectory-0.json")))
    # (R*4899 + G*9617 + B*1868 + 8192) >> 14 with R = G = B = v is v exactly, so both runs see identical gray frames
    assert outs[0] == outs[1] and len(outs[0]["trajectory"]) == n


def test_annotation_pipeline_binaries(tmp_path):
    """BASELINE configs[4] in miniature: the three drop-in binaries chained as python/preprocess_all.py chains the
    reference's -- optical_trajectories (frames -> trajectory JSON), fit_motion (IMU + GPS -> velocities / steering),
    annotate_frames (velocities and steering -> per-frame labels) -- on synthetic data, checked for consistency."""
    host = os.path.join(ROOT, "pilotguru_b200", "host")
    subprocess.run(["make", "-C", host], check=True, capture_output=True)
    w, h, n, fps = 640, 480, 31, 30.0
    np.stack([synth.frame(t, w=w, h=h) for t in range(n)]).tofile(tmp_path / "frames.gray")
    (tmp_path / "settings.yml").write_text("%YAML:1.0\nCamera.fps: 30.0\nORBextractor.nFeatures: 500\nORBextractor.scaleFactor: 1.2\n"
                                           "ORBextractor.nLevels: 8\nORBextractor.iniThFAST: 20\nORBextractor.minThFAST: 7\n")
    run = lambda *a: subprocess.run(list(a), capture_output=True, text=True, timeout=600)
    p = run(os.path.join(host, "optical_trajectories"), "--vocabulary_file=unused", "--camera_settings", str(tmp_path / "settings.yml"),
            "--out_dir", str(tmp_path), f"--in_video=raw:{tmp_path / 'frames.gray'}:{w}x{h}", "--rotation_smooth_sigma=2")
    assert p.returncode == 0, p.stderr[-1500:]
    traj = json.load(open(tmp_path / "trajectory-0.json"))["trajectory"]
    assert len(traj) == n
    d = synth.imu_gps(60, 100)
    paths = synth.write_imu_gps_json(d, str(tmp_path))
    p = run(os.path.join(host, "fit_motion"), "--rotations_json", paths["rotations"], "--accelerations_json", paths["accelerations"],
            "--locations_json", paths["locations"], "--velocities_out_json", str(tmp_path / "velocities.json"),
            "--steering_out_json", str(tmp_path / "steering.json"), "--optimization_iters=100")
    assert p.returncode == 0, p.stderr[-1500:]
    vel = json.load(open(tmp_path / "velocities.json"))["velocities"]
    # frames.json: the trajectory's frames placed inside the IMU recording (frame 0 at t = 10 s)
    t0 = 10_000_000
    frames = [{"frame_id": e["frame_id"], "time_usec": t0 + e["time_usec"]} for e in traj]
    (tmp_path / "frames.json").write_text(json.dumps({"frames": frames}))
    for root, val, src in (("velocities", "speed_m_s", "velocities.json"), ("steering", "angular_velocity", "steering.json")):
        p = run(os.path.join(host, "annotate_frames"), "--frames_json", str(tmp_path / "frames.json"), "--in_json", str(tmp_path / src),
                "--json_root_element_name", root, "--json_value_name", val, "--out_json", str(tmp_path / f"frame_{root}.json"))
        assert p.returncode == 0, p.stderr[-1500:]
        lab = json.load(open(tmp_path / f"frame_{root}.json"))[root]
        assert [e["frame_id"] for e in lab] == list(range(1, n))            # every frame after the first is covered
    # the per-frame speed is the time average of the velocity series over the frame interval: bounded by its extremes
    lab = json.load(open(tmp_path / "frame_velocities.json"))["velocities"]
    vt = np.array([e["time_usec"] for e in vel]); vv = np.array([e["speed_m_s"] for e in vel])
    for e, fr_prev, fr in zip(lab, frames[:-1], frames[1:]):
        seg = vv[(vt >= fr_prev["time_usec"] - 20_000) & (vt <= fr["time_usec"] + 20_000)]
        assert seg.min() - 1e-9 <= e["speed_m_s"] <= seg.max() + 1e-9
    gps_mean = float(np.mean(d["gps_v"][9:13]))
    assert abs(np.mean([e["speed_m_s"] for e in lab]) - gps_mean) < 1.5          # m/s: the calibrated speed tracks GPS


def _run(tmp_path, settings_text, spec, extra=(), sub="o"):
    host = os.path.join(ROOT, "pilotguru_b200", "host")
    subprocess.run(["make", "-C", host], check=True, capture_output=True)
    os.makedirs(tmp_path / sub, exist_ok=True)
    (tmp_path / "settings.yml").write_text(settings_text)
    return subprocess.run([os.path.join(host, "optical_trajectories"), "--vocabulary_file=x", "--camera_settings", str(tmp_path / "settings.yml"),
                           "--out_dir", str(tmp_path / sub), "--in_video=" + spec, "--logtostderr"] + list(extra),
                          capture_output=True, text=True, timeout=900)


def test_settings_keys_of_this_fork_are_honoured_and_missing_keys_fail(tmp_path):
    """ORBextractor_nFeatures (underscore: src/calibrate.cc:527, Tracking.cc:131) must reach the extractor -- round 1 read the
    upstream dotted spelling only and silently fell back to 1000 features -- and a settings file without any ORB key is an
    error, not a run on defaults."""
    w, h, n = 640, 480, 4
    np.stack([synth.frame(t, w=w, h=h) for t in range(n)]).tofile(tmp_path / "f.gray")
    spec = f"raw:{tmp_path / 'f.gray'}:{w}x{h}"
    base = "%YAML:1.0\nCamera_fps: 30.\nCamera_RGB: 1\nORBextractor_scaleFactor: 1.2\nORBextractor_nLevels: 8\nORBextractor_iniThFAST: 20\nORBextractor_minThFAST: 7\n"
    per_frame = {}
    for nf in (300, 900):
        p = _run(tmp_path, base + f"ORBextractor_nFeatures: {nf}\n", spec, sub=f"n{nf}")
        assert p.returncode == 0, p.stderr[-1500:]
        line = [l for l in p.stderr.splitlines() if "keypoints/frame" in l][-1]
        per_frame[nf] = float(line.split(" frames, ")[1].split(" keypoints/frame")[0])
    assert 280 < per_frame[300] < 340 and 850 < per_frame[900] < 960, per_frame
    p = _run(tmp_path, "%YAML:1.0\nCamera_fx: 500.\nCamera_fps: 30.\n", spec, sub="bad")
    assert p.returncode != 0 and "ORBextractor_" in p.stderr
    p = _run(tmp_path, "%YAML:1.0\nCamera.fps: 30.0\nORBextractor.nFeatures: 300\n", spec, sub="dotted")      # upstream spelling: accepted, warned
    assert p.returncode == 0 and "dotted keys" in p.stderr


def _device_count():
    import torch
    return torch.cuda.device_count()


def test_synthetic_device_source_is_deterministic(tmp_path):
    """synth: source (frames rendered on the device from a canvas file, BASELINE configs[2]/[4] stand-in for a decoder): two
    runs with different batch sizes give the identical trajectory file."""
    cw, ch, w, h, n = 1120, 800, 640, 480, 37
    synth.canvas(1234, cw, ch).tofile(tmp_path / "canvas.gray")
    spec = f"synth:{tmp_path / 'canvas.gray'}:{cw}x{ch}:{n}:{w}x{h}"
    st = "%YAML:1.0\nCamera_fps: 30.\nORBextractor_nFeatures: 500\n"
    a = _run(tmp_path, st, spec, ["--batch=16"], "a"); b = _run(tmp_path, st, spec, ["--batch=5"], "b")
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-800:], b.stderr[-800:])
    ja, jb = open(tmp_path / "a" / "trajectory-0.json").read(), open(tmp_path / "b" / "trajectory-0.json").read()
    assert ja == jb and len(json.loads(ja)["trajectory"]) == n
    # (an angular_velocity may be null: identical consecutive headings can give a rotation cosine that rounds above 1,
    # whose acos is NaN in the reference too, horizontal_flatten.cc:56-61, and nlohmann dumps NaN as null)
    tr = json.loads(ja)["trajectory"]
    flow = np.array([synth.flow(t, w=w, h=h, cw=cw, ch=ch) if t else (0, 0) for t in range(n)], float)
    got = np.array([[e["pose"]["translation"][0], e["pose"]["translation"][2]] for e in tr])
    assert np.max(np.abs(got + np.cumsum(flow, axis=0))) <= 0.51


@pytest.mark.skipif("_device_count() < 2")
def test_frames_sharded_over_two_gpus_equal_one_gpu(tmp_path):
    """--num_gpus 2 (BASELINE configs[2] in small): contiguous blocks, one NCCL all-gather of the block-boundary feature
    records through the C-ABI, boundary pair matched afterwards -- byte-identical trajectory and identical keypoint / match
    totals to the single-GPU run, for a frame count that does not divide evenly."""
    cw, ch, w, h, n = 1120, 800, 640, 480, 45
    synth.canvas(1234, cw, ch).tofile(tmp_path / "canvas.gray")
    spec = f"synth:{tmp_path / 'canvas.gray'}:{cw}x{ch}:{n}:{w}x{h}"
    st = "%YAML:1.0\nCamera_fps: 30.\nORBextractor_nFeatures: 500\n"
    one = _run(tmp_path, st, spec, ["--batch=8"], "one"); two = _run(tmp_path, st, spec, ["--batch=8", "--num_gpus=2"], "two")
    assert one.returncode == 0 and two.returncode == 0, (one.stderr[-800:], two.stderr[-1500:])
    assert open(tmp_path / "one" / "trajectory-0.json").read() == open(tmp_path / "two" / "trajectory-0.json").read()
    tot = lambda p: [l for l in p.stderr.splitlines() if "totals:" in l][-1].split("totals:")[1]
    assert tot(one) == tot(two)
