"""GPU tests of the decode part of the frame feed (pgb_video_read_rgb: Motion-JPEG AVI -> nvJPEG -> RGB24 in device memory
-> rotate / flip / gray -> extractor), the stand-in for VideoImageSequenceSource + Tracking::GrabImageMonocular
(src/io/image_sequence_reader.cc:74-208, thirdparty/orb-slam2/src/Tracking.cc:243-258).

Tolerance: a JPEG decoder is not bit-defined (IDCT precision and chroma upsampling are the implementation's choice), so the
frames are compared with cv2's decode of the same file -- FFmpeg's mjpeg decoder + swscale, the libraries the reference
itself links -- within 3 grey levels on the GRAY image (what the extractor sees; the BT.601 recombination cancels the
chroma-upsampling differences) and within 1.0 mean absolute difference per channel on RGB."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import video
from pilotguru_b200.orb import ORBextractor

pytestmark = pytest.mark.gpu


def test_mjpeg_frames_match_ffmpeg_decode_and_feed_the_extractor(golden_dir):
    cv2 = pytest.importorskip("cv2")
    path = os.path.join(golden_dir, "mjpeg_256x192.avi")
    src = video.VideoImageSequenceSource(path)
    cap = cv2.VideoCapture(path)
    ref = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        ref.append(bgr[..., ::-1].copy())
    assert len(ref) == src.n_frames == 5
    rgb, ts = src.read_rgb(0, src.n_frames)
    assert np.allclose(ts, np.arange(5) / 25.0, rtol=0, atol=1e-12)          # pts * time_base (image_sequence_reader.cc:153-155)
    for i in range(5):
        d = np.abs(rgb[i].astype(np.int32) - ref[i].astype(np.int32))
        assert d.mean() < 1.0, (i, d.mean())
        ga, gb = O.to_gray(rgb[i], formula=0), O.to_gray(ref[i], formula=0)
        assert np.abs(ga.astype(np.int32) - gb.astype(np.int32)).max() <= 3, i
    # hasNext / next walk the file like the reference's source; frame ids count from 1
    ids = []
    while src.hasNext():
        img, t, fid = src.next()
        assert np.array_equal(img, rgb[fid - 1]) and t == ts[fid - 1]
        ids.append(fid)
    assert ids == [1, 2, 3, 4, 5]
    # decode -> gray on the device -> extractor == oracle on the same gray frame (bit-exact from the gray frame on)
    gray, _ = src.read_gray_device(1, 2, vertical_flip=True)
    g = gray.cpu().numpy()
    assert np.array_equal(g[0], O.to_gray(rgb[1], vflip=True, formula=0))
    ex = ORBextractor(200, 1.2, 8, 20, 7, max_width=256, max_height=192, max_batch=2)
    kps, desc, counts = ex.extract_batch(g)
    ok_, od_ = O.OrbOracle(200, 1.2, 8, 20, 7).extract(g[1])
    n = counts[1]
    assert n == len(ok_) and np.array_equal(kps[1, :n], ok_) and np.array_equal(desc[1, :n], od_) and n > 50
    ex.close()
    src.close()


def test_optical_trajectories_reads_a_motion_jpeg_file(tmp_path, golden_dir):
    """The drop-in binary on a video FILE (the reference's calling convention: --in_video <file>): the run on the AVI equals the
    run on raw gray frames produced by the library's own decode -> gray path (same frames in, byte-identical trajectory), the
    time stamps come from the container's frame rate, and two GPU-less facts hold: 5 frames, 25 fps."""
    import json
    import subprocess
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(ROOT, "pilotguru_b200", "host")], check=True, capture_output=True)
    path = os.path.join(golden_dir, "mjpeg_256x192.avi")
    src = video.VideoImageSequenceSource(path)
    gray, _ = src.read_gray_device(0, src.n_frames)
    src.close()
    raw = tmp_path / "frames.gray"
    gray.cpu().numpy().tofile(raw)
    settings = tmp_path / "settings.yml"
    settings.write_text("%YAML:1.0\nCamera_fps: 30.\nCamera_RGB: 1\nORBextractor_nFeatures: 200\nORBextractor_scaleFactor: 1.2\n"
                        "ORBextractor_nLevels: 8\nORBextractor_iniThFAST: 20\nORBextractor_minThFAST: 7\n")
    outs = []
    for tag, spec in (("avi", path), ("raw", f"raw:{raw}:256x192")):
        d = tmp_path / tag
        d.mkdir()
        p = subprocess.run([os.path.join(ROOT, "pilotguru_b200", "host", "optical_trajectories"), "--vocabulary_file=unused.txt",
                            "--camera_settings", str(settings), "--out_dir", str(d), "--in_video", spec, "--novisualize", "--batch=4",
                            "--logtostderr"], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.load(open(d / "trajectory-0.json")))
    a, r = outs
    assert len(a["trajectory"]) == 5
    assert [e["time_usec"] for e in a["trajectory"]] == [int(round(i * 1e6 / 25.0)) for i in range(5)]      # the container's 25 fps
    assert [e["time_usec"] for e in r["trajectory"]] == [int(round(i * 1e6 / 30.0)) for i in range(5)]      # Camera_fps for raw frames
    assert [e["pose"] for e in a["trajectory"]] == [e["pose"] for e in r["trajectory"]] and a["plane"] == r["plane"]
