#!/usr/bin/env python3
"""BASELINE configs[4] end to end (and configs[2] with --only-frames): a synthetic drive of `--minutes` minutes -- 30 fps
1080p video + 500 Hz IMU + 1 Hz GPS -- through the three drop-in binaries on `--gpus` B200s:

    optical_trajectories --num_gpus N   frames -> ORB extract + match (frames sharded, one NCCL all-gather) -> trajectory JSON
                                        (per-frame heading: planar_direction / angular_velocity, json_converters.cc:37-96)
    fit_motion --num_gpus N             IMU + GPS -> calibrated forward velocities + steering (windows sharded)
    annotate_frames (x2)                velocities / steering -> one value per video frame (annotate_frames.cc:31-74)

Mirrors the reference's chain python/preprocess_all.py:20-28.  The video is the SURVEY 8(d) synthetic sequence rendered on
the device (`synth:` source: 54 000 raw 1080p frames would be 112 GB).  Writes a timing summary (JSON) to --out and checks
that every frame received a velocity and a heading.  Run under gpurun --gpus N.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from pilotguru_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--minutes", type=float, default=30.0)
ap.add_argument("--fps", type=float, default=30.0)
ap.add_argument("--imu-hz", type=float, default=500.0)
ap.add_argument("--frames", type=int, default=0, help="override the frame count (configs[2]: 10000)")
ap.add_argument("--only-frames", action="store_true", help="configs[2]: just optical_trajectories")
ap.add_argument("--compare-one-gpu", action="store_true", help="also run optical_trajectories on 1 GPU and compare outputs")
ap.add_argument("--work", default="/tmp/pgb_c5")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c5_summary.json"))
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--fit-gpus", type=int, default=1, help="devices fit_motion shards its windows over")
args = ap.parse_args()

HOST = os.path.join(ROOT, "pilotguru_b200", "host")
os.makedirs(args.work, exist_ok=True)
secs = args.minutes * 60.0
n_frames = args.frames or int(round(secs * args.fps))
summary = {"config": f"{args.minutes:g} min, {n_frames} frames 1920x1080 @ {args.fps:g} fps, IMU {args.imu_hz:g} Hz, GPS 1 Hz", "gpus": args.gpus,
           "stages_s": {}}


def run(name, cmd, n_gpus=None):
    """n_gpus: the devices the binary uses.  The process is shown only those (CUDA_VISIBLE_DEVICES): the driver initialises
    every VISIBLE device when a process starts (measured on the 8-GPU box: ~6 s for a binary that then works 1 s on one GPU)."""
    t0 = time.time()
    env = dict(os.environ)
    if n_gpus is not None:
        env["CUDA_VISIBLE_DEVICES"] = ",".join(str(i) for i in range(n_gpus))
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    dt = time.time() - t0
    summary["stages_s"][name] = round(dt, 3)
    if p.returncode != 0:
        print(p.stderr[-3000:], file=sys.stderr)
        raise SystemExit(f"{name} failed ({p.returncode})")
    return p


t0 = time.time()
canvas = os.path.join(args.work, "canvas.gray")
synth.canvas().tofile(canvas)
settings = os.path.join(args.work, "settings.yml")
open(settings, "w").write("%YAML:1.0\nCamera_fps: {:g}\nCamera_RGB: 1\nORBextractor_nFeatures: 1000\nORBextractor_scaleFactor: 1.2\n"
                          "ORBextractor_nLevels: 8\nORBextractor_iniThFAST: 20\nORBextractor_minThFAST: 7\n".format(args.fps))
summary["stages_s"]["generate canvas + settings"] = round(time.time() - t0, 3)

spec = f"synth:{canvas}:{synth.CANVAS_W}x{synth.CANVAS_H}:{n_frames}:{synth.FRAME_W}x{synth.FRAME_H}"


def optical(n_gpus, sub):
    out = os.path.join(args.work, sub)
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(out):
        os.remove(os.path.join(out, f))
    p = run(f"optical_trajectories ({n_gpus} GPU)", [os.path.join(HOST, "optical_trajectories"), "--vocabulary_file=unused", "--camera_settings", settings,
                                                    "--out_dir", out, "--in_video=" + spec, f"--num_gpus={n_gpus}", f"--batch={args.batch}", "--logtostderr"], n_gpus)
    line = [l for l in p.stderr.splitlines() if "extract+match:" in l][-1]
    return out, line


traj_dir, line = optical(args.gpus, "traj")
summary["optical_trajectories"] = line.split("I ", 1)[-1]
if args.compare_one_gpu:
    one_dir, line1 = optical(1, "traj1")
    summary["optical_trajectories_1gpu"] = line1.split("I ", 1)[-1]
    same = sorted(os.listdir(one_dir)) == sorted(os.listdir(traj_dir)) and all(
        open(os.path.join(one_dir, f)).read() == open(os.path.join(traj_dir, f)).read() for f in os.listdir(one_dir))
    summary["sharded_equals_one_gpu"] = bool(same)
    tot = lambda l: l.split("totals:")[1].strip()
    summary["totals_equal"] = tot(line) == tot(line1)
    if not (same and summary["totals_equal"]):
        json.dump(summary, open(args.out, "w"), indent=1)
        raise SystemExit("sharded run differs from the single-GPU run")

segs = sorted(f for f in os.listdir(traj_dir) if f.startswith("trajectory-"))
traj = []
for f in segs:
    traj += json.load(open(os.path.join(traj_dir, f)))["trajectory"]
summary["trajectory_frames"] = len(traj)
summary["trajectory_segments"] = len(segs)

if not args.only_frames:
    t0 = time.time()
    d = synth.imu_gps(secs, args.imu_hz)
    paths = synth.write_imu_gps_json(d, args.work)
    frames_json = os.path.join(args.work, "frames.json")
    # (+137 us: a frame timestamp EXACTLY equal to the last velocity timestamp trips the reference's own CHECK in
    # TimeSeries::LinearInterpolate, time_series.hpp:210-213, which annotate_frames reproduces; recorded data never ties)
    json.dump({"frames": [{"frame_id": i, "time_usec": int(round(i * 1e6 / args.fps)) + 137} for i in range(n_frames)]}, open(frames_json, "w"))
    summary["stages_s"]["generate IMU/GPS/frames JSON (python)"] = round(time.time() - t0, 3)
    vel, steer, fwd = (os.path.join(args.work, n) for n in ("velocities.json", "steering.json", "forward.json"))
    # the calibration of a 30-minute drive is 2 ms of kernels: one GPU (BASELINE configs[3] names one), so that the process
    # pays for one context instead of eight (--fit-gpus N shards the windows anyway)
    run(f"fit_motion ({args.fit_gpus} GPU)", [os.path.join(HOST, "fit_motion"), "--rotations_json", paths["rotations"], "--accelerations_json", paths["accelerations"],
                                         "--locations_json", paths["locations"], "--velocities_out_json", vel, "--steering_out_json", steer,
                                         "--forward_axis_out_json", fwd, f"--num_gpus={args.fit_gpus}"], args.fit_gpus)
    fv, fs = os.path.join(args.work, "frame_velocities.json"), os.path.join(args.work, "frame_steering.json")
    run("annotate_frames (velocity)", [os.path.join(HOST, "annotate_frames"), "--frames_json", frames_json, "--in_json", vel,
                                       "--json_root_element_name=velocities", "--json_value_name=speed_m_s", "--out_json", fv], 1)
    run("annotate_frames (steering)", [os.path.join(HOST, "annotate_frames"), "--frames_json", frames_json, "--in_json", steer,
                                       "--json_root_element_name=steering", "--json_value_name=angular_velocity", "--out_json", fs], 1)
    jv, js = json.load(open(fv)), json.load(open(fs))
    lv, ls = jv[next(iter(jv))], js[next(iter(js))]
    summary["frames_with_velocity"] = len(lv)
    summary["frames_with_steering"] = len(ls)
    speeds = np.array([e["speed_m_s"] for e in lv])
    tsec = (np.array([e["frame_id"] for e in lv]) - 0.5) / args.fps          # a frame's label averages over (t[i-1], t[i]]
    truth = 8 + 4 * np.sin(0.15 * tsec) + 2 * np.sin(0.5 * tsec)
    summary["velocity_rmse_vs_generating_model_m_s"] = float(np.sqrt(np.mean((speeds - truth) ** 2)))
    # one record per video frame: velocity + heading (the deliverable of configs[4])
    by_id = {e["frame_id"]: e for e in lv}
    merged = [{"frame_id": e["frame_id"], "time_usec": e["time_usec"], "planar_direction": e["planar_direction"],
               "angular_velocity": e["angular_velocity"], "speed_m_s": by_id.get(e["frame_id"], {}).get("speed_m_s")} for e in traj]
    json.dump({"frames": merged}, open(os.path.join(args.work, "per_frame_velocity_heading.json"), "w"))
    summary["per_frame_records"] = len(merged)
    summary["per_frame_records_with_velocity"] = sum(1 for m in merged if m["speed_m_s"] is not None)
    assert len(traj) == n_frames, (len(traj), n_frames)
    assert summary["per_frame_records_with_velocity"] >= n_frames - 2 * int(args.fps) - 2   # the first / last second lie outside the GPS windows
summary["total_s"] = round(sum(summary["stages_s"].values()), 3)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(summary, open(args.out, "w"), indent=1)
print(json.dumps(summary, indent=1))
