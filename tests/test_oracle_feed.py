"""CPU tests of the oracle's frame-feed restatement (cv::flip + cvtColor to gray: image_sequence_reader.cc:163-175,
Tracking.cc:243-258) against cv2 golden vectors (OpenCV >= 3 formula) and the published OpenCV 2.4 formula."""
import os

import numpy as np

import oracle_lib as O


def test_to_gray_matches_cv2_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv2_gray.npz"))
    assert np.array_equal(O.to_gray(g["rgb"], rgb_order=True, formula=1), g["rgb2gray"])
    assert np.array_equal(O.to_gray(g["rgb"], rgb_order=False, formula=1), g["bgr2gray"])
    assert np.array_equal(O.to_gray(g["rgba"], rgb_order=True, formula=1), g["rgba2gray"])
    assert np.array_equal(O.to_gray(g["rgba"], rgb_order=False, formula=1), g["bgra2gray"])
    assert np.array_equal(O.to_gray(g["rgb"], vflip=True, formula=1), g["flip_v_rgb2gray"])
    assert np.array_equal(O.to_gray(g["rgb"], hflip=True, formula=1), g["flip_h_rgb2gray"])
    assert np.array_equal(O.to_gray(g["rgb"], vflip=True, hflip=True, formula=1), g["flip_vh_rgb2gray"])


def test_to_gray_opencv2_formula():
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    r, gch, b = (img[..., i].astype(np.int64) for i in range(3))
    want = ((r * 4899 + gch * 9617 + b * 1868 + 8192) >> 14).astype(np.uint8)   # OpenCV 2.4 RGB2Gray<uchar>, yuv_shift = 14
    assert np.array_equal(O.to_gray(img, formula=0), want)
    gray = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    assert np.array_equal(O.to_gray(gray, vflip=True), gray[::-1])
    assert np.array_equal(O.to_gray(gray, hflip=True), gray[:, ::-1])


def test_reader_rotation_matches_cv2():
    """O.rotate_like_reader restates the switch of VideoImageSequenceSource::fetchNext (image_sequence_reader.cc:186-207) with
    numpy views; here it is held to the reference's own calls, cv::flip / Mat::t(), through cv2."""
    cv2 = __import__("pytest").importorskip("cv2")
    rng = np.random.default_rng(12)
    for shape in ((37, 61, 3), (48, 64), (1, 5, 3)):
        raw = rng.integers(0, 256, shape, dtype=np.uint8)
        want = {0: raw, 90: cv2.flip(cv2.transpose(raw), 0), 180: cv2.flip(raw, -1), 270: cv2.flip(cv2.transpose(raw), 1)}
        for deg, w in want.items():
            w = w.reshape(O.rotate_like_reader(raw, deg).shape)   # cv2 drops a trailing unit axis
            assert np.array_equal(O.rotate_like_reader(raw, deg), w), (shape, deg)
            assert np.array_equal(O.rotate_like_reader(raw, deg + 360), w)
