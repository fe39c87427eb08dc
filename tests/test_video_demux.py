"""CPU tests of the container side of the video frame source (pgb_video_open / pgb_video_info / pgb_video_frame_span: no GPU,
no decode): the frame index of the C++ RIFF walker against an independent Python walker of the same file, the stream facts
against cv2.VideoCapture (OpenCV's FFmpeg backend = the libavformat the reference links, image_sequence_reader.cc:74-120),
and the refusal of what this build does not demux / decode."""
import os
import struct

import numpy as np
import pytest

from pilotguru_b200 import video
from pilotguru_b200._lib import PgbError


def _walk(d, off, end, out, in_movi=False):
    while off + 8 <= end:
        cc, sz = d[off:off + 4], struct.unpack("<I", d[off + 4:off + 8])[0]
        if cc in (b"RIFF", b"LIST"):
            kind = d[off + 8:off + 12]
            _walk(d, off + 12, off + 8 + sz, out, in_movi or kind == b"movi")
        elif in_movi and cc[2:] in (b"dc", b"db") and sz > 0:
            out.append((off + 8, sz))
        off += 8 + sz + (sz & 1)


def test_frame_index_matches_an_independent_walker_and_ffmpeg(golden_dir):
    path = os.path.join(golden_dir, "mjpeg_256x192.avi")
    src = video.VideoImageSequenceSource(path)
    d = open(path, "rb").read()
    want = []
    _walk(d, 0, len(d), want)
    assert src.n_frames == len(want) == 5
    assert [src.frame_span(i) for i in range(src.n_frames)] == want
    for off, size in want:
        assert d[off:off + 2] == b"\xff\xd8" and d[off + size - 2:off + size] == b"\xff\xd9"     # every chunk is one JPEG image
    assert (src.width, src.height, src.rotate_degrees) == (256, 192, 0) and abs(src.fps - 25.0) < 1e-12
    assert src.hasNext()
    cv2 = pytest.importorskip("cv2")
    cap = cv2.VideoCapture(path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == src.n_frames
    assert abs(cap.get(cv2.CAP_PROP_FPS) - src.fps) < 1e-9
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (src.width, src.height)
    src.close()


def test_truncated_and_foreign_files_are_refused_or_cut(tmp_path, golden_dir):
    d = open(os.path.join(golden_dir, "mjpeg_256x192.avi"), "rb").read()
    with pytest.raises(PgbError, match="not a RIFF AVI"):
        p = tmp_path / "x.mp4"
        p.write_bytes(b"\x00\x00\x00\x18ftypmp42" + bytes(64))
        video.VideoImageSequenceSource(str(p))
    with pytest.raises(PgbError, match="cannot open"):
        video.VideoImageSequenceSource(str(tmp_path / "missing.avi"))
    # another codec in the stream header: refused with its FOURCC
    other = bytearray(d)
    i = other.index(b"strf")
    other[i + 8 + 16:i + 8 + 20] = b"H264"
    j = other.index(b"strh")
    other[j + 8 + 4:j + 8 + 8] = b"H264"
    p = tmp_path / "h264.avi"
    p.write_bytes(bytes(other))
    with pytest.raises(PgbError, match="'H264' is not decoded"):
        video.VideoImageSequenceSource(str(p))
    # a recording cut in the middle of the fourth frame keeps the three complete ones (the RIFF sizes still claim the full file)
    src = video.VideoImageSequenceSource(os.path.join(golden_dir, "mjpeg_256x192.avi"))
    off4, _ = src.frame_span(3)
    src.close()
    p = tmp_path / "cut.avi"
    p.write_bytes(d[:off4 + 100])
    cut = video.VideoImageSequenceSource(str(p))
    assert cut.n_frames == 3
    cut.close()


def _chunk(cc, body):
    return cc + struct.pack("<I", len(body)) + body + (b"\x00" if len(body) & 1 else b"")


def test_rec_lists_second_stream_and_opendml_segments(tmp_path, golden_dir):
    """Container shapes the fixture does not have, rebuilt around its JPEG frames: an audio stream in FRONT of the video stream
    (the frames are then '01dc' chunks: VideoStreamIndexOrDie picks the first VIDEO stream, image_sequence_reader.cc:63-71),
    frames grouped in 'rec ' lists, an empty chunk (a dropped frame) and a second RIFF 'AVIX' segment (OpenDML) with more frames."""
    d = open(os.path.join(golden_dir, "mjpeg_256x192.avi"), "rb").read()
    spans = []
    _walk(d, 0, len(d), spans)
    jpegs = [d[o:o + s] for o, s in spans]
    avih = struct.pack("<14I", 40000, 0, 0, 0x10, 6, 0, 2, 0, 256, 192, 0, 0, 0, 0)
    auds = _chunk(b"strh", b"auds" + bytes(4) + struct.pack("<10I", 0, 0, 0, 1, 8000, 0, 0, 0, 0, 0) + bytes(8)) + _chunk(b"strf", bytes(18))
    vids = _chunk(b"strh", b"vids" + b"MJPG" + struct.pack("<10I", 0, 0, 0, 1001, 30000, 0, 6, 0, 0, 0) + bytes(8)) + \
        _chunk(b"strf", struct.pack("<IiiHH4sIiiII", 40, 256, 192, 1, 24, b"MJPG", 256 * 192 * 3, 0, 0, 0, 0))
    hdrl = _chunk(b"LIST", b"hdrl" + _chunk(b"avih", avih) + _chunk(b"LIST", b"strl" + auds) + _chunk(b"LIST", b"strl" + vids))
    rec0 = _chunk(b"LIST", b"rec " + _chunk(b"00wb", bytes(31)) + _chunk(b"01dc", jpegs[0]) + _chunk(b"01dc", b""))
    rec1 = _chunk(b"LIST", b"rec " + _chunk(b"01dc", jpegs[1]) + _chunk(b"00wb", bytes(8)) + _chunk(b"01db", jpegs[2]))
    movi0 = _chunk(b"LIST", b"movi" + rec0 + rec1)
    seg0 = _chunk(b"RIFF", b"AVI " + hdrl + _chunk(b"JUNK", bytes(13)) + movi0 + _chunk(b"idx1", bytes(32)))
    seg1 = _chunk(b"RIFF", b"AVIX" + _chunk(b"LIST", b"movi" + _chunk(b"01dc", jpegs[3]) + _chunk(b"01dc", jpegs[4])))
    p = tmp_path / "shapes.avi"
    p.write_bytes(seg0 + seg1)
    src = video.VideoImageSequenceSource(str(p))
    assert src.n_frames == 5 and (src.width, src.height) == (256, 192) and abs(src.fps - 30000 / 1001) < 1e-12
    blob = p.read_bytes()
    for i in range(5):
        off, size = src.frame_span(i)
        assert blob[off:off + size] == jpegs[i], i
    src.close()
    # no video stream at all: the reference's "Inspected all the streams, but no video stream found"
    only_audio = _chunk(b"RIFF", b"AVI " + _chunk(b"LIST", b"hdrl" + _chunk(b"avih", avih) + _chunk(b"LIST", b"strl" + auds)) + _chunk(b"LIST", b"movi" + _chunk(b"00wb", bytes(16))))
    q = tmp_path / "audio.avi"
    q.write_bytes(only_audio)
    with pytest.raises(PgbError, match="no video stream found"):
        video.VideoImageSequenceSource(str(q))
