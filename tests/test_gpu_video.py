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
from pilotguru_b200 import synth, video
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


def test_frames_without_huffman_tables_get_the_standard_ones(tmp_path):
    """Capture hardware writes Motion-JPEG frames without DHT segments ("AVI1"): the decoder must put the typical tables of
    ITU-T T.81 Annex K.3 in front of the scan, as libavcodec's mjpeg decoder does.  Frames are encoded here with those tables
    (libjpeg's defaults, optimisation off), their DHT segments are cut out, the rest goes into an AVI: the decode must equal the
    decode of the untouched frames (same decoder, same coefficients: bit for bit)."""
    import struct
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    w, h, n = 160, 120, 3

    def chunk(cc, body):
        return cc + struct.pack("<I", len(body)) + body + (b"\x00" if len(body) & 1 else b"")

    def avi(frames):
        avih = struct.pack("<14I", 40000, 0, 0, 0x10, len(frames), 0, 1, 0, w, h, 0, 0, 0, 0)
        vids = chunk(b"strh", b"vids" + b"MJPG" + struct.pack("<10I", 0, 0, 0, 1, 25, 0, len(frames), 0, 0, 0) + bytes(8)) + \
            chunk(b"strf", struct.pack("<IiiHH4sIiiII", 40, w, h, 1, 24, b"MJPG", w * h * 3, 0, 0, 0, 0))
        hdrl = chunk(b"LIST", b"hdrl" + chunk(b"avih", avih) + chunk(b"LIST", b"strl" + vids))
        return chunk(b"RIFF", b"AVI " + hdrl + chunk(b"LIST", b"movi" + b"".join(chunk(b"00dc", f) for f in frames)))

    def strip_dht(j):
        out, i = bytearray(j[:2]), 2
        while j[i + 1] != 0xDA:
            L = struct.unpack(">H", j[i + 2:i + 4])[0]
            if j[i + 1] != 0xC4:
                out += j[i:i + 2 + L]
            i += 2 + L
        return bytes(out + j[i:])

    full = []
    for t in range(n):
        img = np.clip(synth.frame(t, w=w, h=h)[..., None].astype(np.int32) + rng.integers(-20, 20, (h, w, 3)), 0, 255).astype(np.uint8)
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_OPTIMIZE, 0])
        full.append(bytes(enc))
    bare = [strip_dht(j) for j in full]
    assert all(b"\xff\xc4" not in b[:b.index(b"\xff\xda")] and len(b) == len(j) - 432 for b, j in zip(bare, full))
    outs = []
    for name, frames in (("full.avi", full), ("bare.avi", bare)):
        p = tmp_path / name
        p.write_bytes(avi(frames))
        src = video.VideoImageSequenceSource(str(p))
        assert src.n_frames == n and (src.width, src.height) == (w, h)
        outs.append(src.read_rgb(0, n)[0])
        src.close()
    assert np.array_equal(outs[0], outs[1])
    ref = cv2.imdecode(np.frombuffer(full[1], np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
    assert np.abs(O.to_gray(outs[1][1], formula=0).astype(np.int32) - O.to_gray(np.ascontiguousarray(ref), formula=0).astype(np.int32)).max() <= 3
