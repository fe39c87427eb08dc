"""CPU tests: the oracle's OpenCV-primitive restatements against (a) committed golden vectors produced by cv2 4.13
(tests/golden/make_golden.py) and (b) cv2 itself when it is importable.  These pins are what makes the oracle a
trustworthy stand-in for the un-buildable reference."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "cv2_primitives.npz"))


def test_resize_golden(G):
    assert np.array_equal(O.resize_linear(G["img"], 267, 200), G["resize_img_267x200"])
    assert np.array_equal(O.resize_linear(G["noise"], 109, 81), G["resize_noise_109x81"])


@pytest.mark.parametrize("name", ["img", "noise"])
@pytest.mark.parametrize("th", [7, 20])
def test_fast_golden(G, name, th):
    got = O.fast(G[name], th)
    assert np.array_equal(got, G[f"fast_{name}_th{th}"])  # positions, row-major order, responses


@pytest.mark.parametrize("name", ["img", "noise"])
def test_blur_golden(G, name):
    assert np.array_equal(O.gaussian_blur7(G[name]), G[f"blur_{name}"])


def test_fast_atan2_golden(G):
    got = O.fast_atan2(G["atan_y"], G["atan_x"])
    assert np.array_equal(got, G["atan_deg"])
    assert O.fast_atan2(np.zeros(1, np.float32), np.zeros(1, np.float32))[0] == 0.0


def test_ic_angle_and_rbrief_golden(G):
    img = G["img"]
    for (x, y), a in zip(G["orb_xy"], G["orb_angle"]):
        assert np.float32(O.ic_angle(img, x, y)) == a
    bl = G["orb_blurred_float_path"]
    for (x, y), a, d in zip(G["orb_xy"], G["orb_angle"], G["orb_desc"]):
        assert np.array_equal(O.orb_descriptor(bl, x, y, float(a)), d)


def test_tables_match_reference_constants():
    t = O.OrbOracle(1000, 1.2, 8, 20, 7).tables()
    assert t["n_per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 61]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    o = O.OrbOracle(1000, 1.2, 8, 20, 7)
    sizes = [o.level_size(1920, 1080, l) for l in range(8)]
    assert sizes == [(1920, 1080), (1600, 900), (1333, 750), (1111, 625), (926, 521), (772, 434), (643, 362), (536, 301)]


def test_pipeline_golden_regression(golden_dir):
    P = np.load(os.path.join(golden_dir, "orb_pipeline_640x480.npz"))
    orc = O.OrbOracle(500, 1.2, 8, 20, 7)
    k, d = orc.extract(synth.frame(1, w=640, h=480))
    assert np.array_equal(k, P["kps1"]) and np.array_equal(d, P["desc1"])
    # structural properties of operator(): level-major order, quotas, integer level coordinates
    assert np.all(np.diff(k["octave"]) >= 0)
    npl = orc.tables()["n_per_level"]
    cnt = np.bincount(k["octave"], minlength=8)
    assert np.all(cnt <= npl + 2)
    lvl0 = k[k["octave"] == 0]
    assert np.all(lvl0["x"] == np.round(lvl0["x"])) and np.all(lvl0["x"] >= 19) and np.all(lvl0["x"] < 640 - 19)


def test_octree_properties():
    rng = np.random.default_rng(3)
    # distinct random points with random responses
    pts = set()
    while len(pts) < 3000:
        pts.add((int(rng.integers(3, 1885)), int(rng.integers(3, 1045))))
    xy = np.array(sorted(pts, key=lambda p: (p[1] // 31, p[0] // 31, p[1], p[0])), np.int32)
    xyr = np.concatenate([xy, rng.integers(7, 200, (len(xy), 1)).astype(np.int32)], axis=1)
    keep = O.distribute_octree(xyr, 16, 1904, 16, 1064, 217)
    assert 217 <= len(keep) <= 219 and len(set(keep.tolist())) == len(keep)
    # fewer points than requested: every point survives on its own
    few = xyr[:50]
    keep = O.distribute_octree(few, 16, 1904, 16, 1064, 217)
    assert sorted(keep.tolist()) == list(range(50))
    assert len(O.distribute_octree(xyr[:0], 16, 1904, 16, 1064, 217)) == 0


def test_empty_and_flat_images():
    orc = O.OrbOracle(300, 1.2, 8, 20, 7)
    k, d = orc.extract(np.full((240, 320), 77, np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)


cv2 = pytest.importorskip("cv2")


def test_primitives_against_cv2_live():
    rng = np.random.default_rng(0)
    for (w, h, dw, dh) in [(1920, 1080, 1600, 900), (643, 362, 536, 301), (301, 207, 251, 173)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(O.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    f = synth.frame(5, w=800, h=600)
    for th in (7, 20):
        det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        for _ in range(25):  # cell-sized views, like ComputeKeyPointsOctTree's
            x = int(rng.integers(0, 760)); y = int(rng.integers(0, 560))
            cw = int(rng.integers(7, 38)); ch = int(rng.integers(7, 38))
            c = np.ascontiguousarray(f[y:y + ch, x:x + cw])
            ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(c)], np.int32).reshape(-1, 3)
            assert np.array_equal(ref, O.fast(c, th))
    img = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    assert np.array_equal(O.gaussian_blur7(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
