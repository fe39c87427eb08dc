"""GPU parity tests of the CUDA ORB extractor, through the C-ABI, against the oracle and the golden fixtures.
Bar: bit-exact pyramids, FAST score maps, candidate lists, keypoints (all 7 fields) and descriptors."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

pytestmark = pytest.mark.gpu

KPF = ["x", "y", "size", "angle", "response", "octave", "class_id"]


def _mk(nf, w, h, batch=1):
    from pilotguru_b200.orb import ORBextractor
    return ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=batch)


def _assert_same(gk, gd, ok, od):
    assert len(gk) == len(ok)
    for f in KPF:
        assert np.array_equal(gk[f], ok[f]), f
    assert np.array_equal(gd, od)


@pytest.mark.parametrize("case", ["synth1080", "noise", "odd", "lowtexture", "gradient"])
def test_extract_stage_by_stage(case):
    rng = np.random.default_rng(9)
    if case == "synth1080":
        img, nf = synth.frame(7), 1000
    elif case == "noise":
        img, nf = rng.integers(0, 256, (480, 640), dtype=np.uint8), 1000
    elif case == "odd":
        img, nf = synth.frame(3, w=701, h=403), 777
    elif case == "lowtexture":   # almost everything needs the minThFAST retry; many empty cells
        img = np.clip(synth.frame(2, w=640, h=480).astype(np.int32) // 8 + 100, 0, 255).astype(np.uint8); nf = 500
    else:                        # smooth ramp with a few saturated blocks: exercises ties and flat regions
        yy, xx = np.mgrid[0:400, 0:600]
        img = ((xx + yy) // 4 % 256).astype(np.uint8); img[100:140, 200:260] = 255; img[300:330, 50:90] = 0; nf = 300
    h, w = img.shape
    ex = _mk(nf, w, h)
    orc = O.OrbOracle(nf, 1.2, 8, 20, 7)
    ok, od = orc.extract(img)
    gk, gd = ex(img)
    for l in range(8):
        lvl = orc.level(l)
        assert np.array_equal(ex.image_pyramid(l), lvl), f"pyramid level {l}"
        assert np.array_equal(ex.score_map(l), O.fast_score_map(lvl, 7)), f"score map level {l}"
        assert np.array_equal(ex.candidates(l), orc.candidates(l)), f"candidates level {l}"
        assert np.array_equal(ex.blurred_level(l), O.gaussian_blur7(lvl)), f"blur level {l}"
    _assert_same(gk, gd, ok, od)
    ex.close()


def test_batch_matches_single_and_golden(golden_dir):
    P = np.load(os.path.join(golden_dir, "orb_pipeline_640x480.npz"))
    frames = np.stack([synth.frame(t, w=640, h=480) for t in range(3)])
    ex = _mk(500, 640, 480, batch=3)
    kps, desc, counts = ex.extract_batch(frames)
    for t in range(3):
        n = counts[t]
        _assert_same(kps[t, :n], desc[t, :n], P[f"kps{t}"], P[f"desc{t}"])
    # the same frames one at a time through operator()
    for t in range(3):
        gk, gd = ex(frames[t])
        _assert_same(gk, gd, P[f"kps{t}"], P[f"desc{t}"])
    ex.close()


def test_getters_and_levels():
    ex = _mk(1000, 1920, 1080)
    t = O.OrbOracle(1000, 1.2, 8, 20, 7).tables()
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == pytest.approx(1.2)
    assert np.array_equal(ex.GetScaleFactors(), t["scale"]) and np.array_equal(ex.GetInverseScaleFactors(), t["inv_scale"])
    assert np.array_equal(ex.GetScaleSigmaSquares(), t["sigma2"]) and np.array_equal(ex.GetInverseScaleSigmaSquares(), t["inv_sigma2"])
    assert ex.features_per_level().tolist() == t["n_per_level"].tolist()
    assert [ex.level_size(1920, 1080, l) for l in range(8)][-1] == (536, 301)
    ex.close()


def test_edge_cases():
    from pilotguru_b200 import PgbError
    ex = _mk(300, 320, 240, batch=2)
    k, d = ex(np.zeros((0, 0), np.uint8))            # empty image: silent return (ORBextractor.cc:1045)
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((240, 320), 50, np.uint8))     # flat image: no keypoints, descriptors released
    assert len(k) == 0
    with pytest.raises(ValueError):
        ex(np.zeros((240, 320), np.float32))         # CV_8UC1 assert (ORBextractor.cc:1049)
    with pytest.raises(PgbError):
        ex(np.zeros((480, 640), np.uint8))           # larger than the handle's capacity: loud failure
    with pytest.raises(PgbError):
        ex.extract_batch(np.zeros((3, 240, 320), np.uint8))  # more frames than max_batch
    # smaller than capacity is fine and still exact
    img = synth.frame(4, w=300, h=200)
    gk, gd = ex(img)
    ok, od = O.OrbOracle(300, 1.2, 8, 20, 7).extract(img)
    _assert_same(gk, gd, ok, od)
    ex.close()


def test_size_independent_properties_full_size_batch():
    """At BASELINE size (1080p, 1000 features, batch of 8): idempotence across calls, batch position independence,
    level-major order and quota bounds -- properties that do not need the oracle."""
    frames = np.stack([synth.frame(t) for t in range(8)])
    ex = _mk(1000, 1920, 1080, batch=8)
    k1, d1, c1 = ex.extract_batch(frames)
    k2, d2, c2 = ex.extract_batch(frames[::-1].copy())
    assert np.array_equal(c1, c2[::-1])
    for t in range(8):
        n = c1[t]
        assert np.array_equal(k1[t, :n], k2[7 - t, :n]) and np.array_equal(d1[t, :n], d2[7 - t, :n])
        assert np.all(np.diff(k1[t, :n]["octave"]) >= 0)
        assert np.all(np.bincount(k1[t, :n]["octave"], minlength=8) <= ex.features_per_level() + 2)
        assert 900 < n <= ex.cap
    ex.close()


def test_host_frame_pipeline_matches_resident():
    """PGB_OUT_DEVICE (host frames copied in chunks 4, 8, 16, ..., 2, 2, 1 on a copy stream while three compute
    streams alternate over the chunks) must give exactly what the device-resident call gives, for an odd batch size,
    and keep doing so when the call is repeated (the copy stream must not overtake the previous call's kernels)."""
    import torch
    from pilotguru_b200.orb import ORBextractor
    n, w, h = 37, 640, 480
    frames = np.stack([synth.frame(t, w=w, h=h) for t in range(n)])
    ex = _mk(500, w, h, batch=n)
    cap = ex.cap
    host = torch.from_numpy(frames).pin_memory()
    dev = host.cuda()
    outs = []
    for where, ptr in ((ORBextractor.IN_DEVICE | ORBextractor.OUT_DEVICE, dev.data_ptr()), (ORBextractor.OUT_DEVICE, host.data_ptr()),
                       (ORBextractor.OUT_DEVICE, host.data_ptr())):
        kps = torch.zeros((n, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
        counts = torch.zeros(n, dtype=torch.int32, device="cuda")
        ex.extract_ptr(ptr, where, n, w, h, w, w * h, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), cap)
        ex.check()
        outs.append((kps.cpu().numpy().view(np.uint32), desc.cpu().numpy(), counts.cpu().numpy()))
    for o in outs[1:]:
        assert np.array_equal(o[2], outs[0][2]) and (outs[0][2] > 300).all()
        for t in range(n):
            c = outs[0][2][t]
            assert np.array_equal(o[0][t, :c], outs[0][0][t, :c]) and np.array_equal(o[1][t, :c], outs[0][1][t, :c])
    ex.close()


def test_frame_feed_matches_oracle_and_feeds_the_extractor(golden_dir):
    """pgb_frames_to_gray (cv::flip + cvtColor) against the oracle for every layout, and its output through the extractor."""
    from pilotguru_b200.orb import frames_to_gray
    rng = np.random.default_rng(9)
    g = np.load(os.path.join(golden_dir, "cv2_gray.npz"))
    assert np.array_equal(frames_to_gray(g["rgb"][None], formula=1)[0], g["rgb2gray"])          # the cv2 pin, on the GPU
    assert np.array_equal(frames_to_gray(g["rgba"][None], rgb_order=False, formula=1)[0], g["bgra2gray"])
    for c in (1, 3, 4):
        for w, h in ((64, 48), (61, 37), (1, 1), (5, 3)):
            fr = rng.integers(0, 256, (3, h, w, c), dtype=np.uint8)
            for rgbo, vf, hf, fm in ((1, 0, 0, 0), (0, 1, 0, 1), (1, 0, 1, 1), (0, 1, 1, 0)):
                got = frames_to_gray(fr, bool(rgbo), bool(vf), bool(hf), fm)
                for i in range(3):
                    assert np.array_equal(got[i], O.to_gray(fr[i], bool(rgbo), bool(vf), bool(hf), fm)), (c, w, h, rgbo, vf, hf, fm)
    # the container's `rotate` metadata (image_sequence_reader.cc:186-207) ahead of the flips and the conversion
    from pilotguru_b200.orb import frames_to_gray_rotated
    for c in (1, 3, 4):
        for w, h in ((64, 48), (61, 37), (1, 1), (5, 3), (33, 130)):
            fr = rng.integers(0, 256, (2, h, w, c), dtype=np.uint8)
            for deg in (0, 90, 180, 270, 450):
                for rgbo, vf, hf, fm in ((1, 0, 0, 0), (0, 1, 0, 1), (1, 1, 1, 0)):
                    got = frames_to_gray_rotated(fr, deg, bool(rgbo), bool(vf), bool(hf), fm)
                    for i in range(2):
                        want = O.to_gray(np.ascontiguousarray(O.rotate_like_reader(fr[i], deg)), bool(rgbo), bool(vf), bool(hf), fm)
                        assert got[i].shape == want.shape and np.array_equal(got[i], want), (c, w, h, deg, rgbo, vf, hf, fm)
    with pytest.raises(Exception, match="Unsupported rotation angle"):
        frames_to_gray_rotated(rng.integers(0, 256, (1, 4, 4, 3), dtype=np.uint8), 45)
    # colour frame -> gray on the device -> extractor == extractor on the oracle's gray
    base = synth.frame(2, w=320, h=240)
    rgb = np.stack([base, np.roll(base, 1, axis=1), np.roll(base, 2, axis=0)], axis=2)
    gray = frames_to_gray(rgb[None], formula=0)[0]
    ex = _mk(300, 320, 240)
    gk, gd = ex(gray)
    ok, od = O.OrbOracle(300, 1.2, 8, 20, 7).extract(O.to_gray(rgb, formula=0))
    _assert_same(gk, gd, ok, od)
    ex.close()


def test_resident_batch_chunked_over_streams_matches_single_pass(monkeypatch):
    """With PGB_RES_CHUNK set, device-resident batches of >= 2 chunks are cut into chunks that alternate over the handle's
    three compute streams (off by default: it measured slower).  Same keypoints and descriptors as the single-pass
    schedule, call after call, including a ragged last chunk."""
    import torch
    from pilotguru_b200.orb import ORBextractor
    n, w, h = 23, 640, 480
    dev = torch.from_numpy(np.stack([synth.frame(t, w=w, h=h) for t in range(n)])).cuda()
    outs = []
    for chunk in ("0", "4", "5"):
        monkeypatch.setenv("PGB_RES_CHUNK", chunk)
        ex = _mk(500, w, h, batch=n)
        cap = ex.cap
        for rep in range(2):
            kps = torch.zeros((n, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
            counts = torch.zeros(n, dtype=torch.int32, device="cuda")
            ex.extract_ptr(dev.data_ptr(), ORBextractor.IN_DEVICE | ORBextractor.OUT_DEVICE, n, w, h, w, w * h, kps.data_ptr(),
                           desc.data_ptr(), counts.data_ptr(), cap)
            ex.check()
            outs.append((kps.cpu().numpy().view(np.uint32), desc.cpu().numpy(), counts.cpu().numpy()))
        ex.close()
    for o in outs[1:]:
        assert np.array_equal(o[2], outs[0][2]) and (outs[0][2] > 300).all()
        for t in range(n):
            c = outs[0][2][t]
            assert np.array_equal(o[0][t, :c], outs[0][0][t, :c]) and np.array_equal(o[1][t, :c], outs[0][1][t, :c])


@pytest.mark.parametrize("size", [(1920, 1080, 1000), (701, 403, 777), (330, 250, 200)])
def test_fused_fast_cells_equals_the_unfused_pair(monkeypatch, size):
    """The hot path's fused kernel (k_fast_cells: score -> per-cell NMS -> candidates, no score map in HBM) against the
    round-1 pair k_fast_score -> k_cells on the same resident frames: identical per-cell candidate lists on every level
    (class A tiles: cells <= 32 px; class B: the small levels with larger cells), then identical keypoints / descriptors
    from an extractor that runs the unfused pair end to end (PGB_UNFUSED=1)."""
    w, h, nf = size
    rng = np.random.default_rng(5)
    f0 = synth.frame(0, w=w, h=h)
    low = (f0.astype(np.float32) * 0.3 + 90).astype(np.uint8)                                  # most corners between minTh and iniTh: the second tier
    mixed = f0.copy()                                                                          # textured | weak noise | strong noise, cell by cell
    mixed[:, w // 3:2 * w // 3] = 128 + rng.integers(-9, 10, (h, 2 * w // 3 - w // 3))
    mixed[:, 2 * w // 3:] = rng.integers(0, 256, (h, w - 2 * w // 3), dtype=np.uint8)
    frames = np.stack([f0, synth.frame(1, w=w, h=h),
                       rng.integers(0, 256, (h, w), dtype=np.uint8),                           # a noise frame (list overflow: bitmap path)
                       low, mixed])
    nfr = len(frames)
    ex = _mk(nf, w, h, batch=nfr)
    k1, d1, c1 = ex.extract_batch(frames)
    fused = [[ex.candidates(l, frame=f) for l in range(8)] for f in range(nfr)]
    ex.run_stage(2); ex.check()                                                              # recompute the slots with the unfused pair
    for f in range(nfr):
        for l in range(8):
            assert np.array_equal(ex.candidates(l, frame=f), fused[f][l]), (f, l)
    assert sum(len(c) for c in fused[0]) > 100
    ex.close()
    monkeypatch.setenv("PGB_UNFUSED", "1")
    ex2 = _mk(nf, w, h, batch=nfr)
    k2, d2, c2 = ex2.extract_batch(frames)
    ex2.close()
    assert np.array_equal(c1, c2)
    for f in range(nfr):
        assert np.array_equal(k1[f, :c1[f]], k2[f, :c1[f]]) and np.array_equal(d1[f, :c1[f]], d2[f, :c1[f]])


def test_fused_fast_cells_fuzz_sizes_and_thresholds():
    """Differential fuzz of the two-tier fused kernel (and the generic single-pass one) against the unfused pair: random level
    geometries (cells of every class: <= 32 x 32, 33-40 px rows, wider than 32 px), scale factors, level counts and
    iniThFAST / minThFAST pairs (equal thresholds, minThFAST = 0, a very high iniThFAST that sends every cell to the second
    tier), on textured, low-contrast and part-noise frames.  Candidates per level must be identical."""
    from pilotguru_b200.orb import ORBextractor
    rng = np.random.default_rng(2024)
    cases = [(1280, 720, 1.2, 8, 20, 7), (811, 607, 1.2, 6, 20, 20), (640, 360, 1.15, 7, 60, 5), (333, 421, 1.3, 4, 12, 0),
             (1024, 300, 1.2, 5, 255, 30), (260, 260, 1.1, 3, 9, 3), (1919, 517, 1.25, 8, 35, 10), (600, 800, 1.2, 8, 20, 7)]
    checked = 0
    for (w, h, sf, nl, ini, mn) in cases:
        base = synth.frame(int(rng.integers(0, 50)), w=min(w, 1920), h=min(h, 1080)) if w <= 1920 and h <= 1080 else None
        if base is None or base.shape != (h, w):
            big = np.tile(synth.frame(3), (2, 1))
            base = np.ascontiguousarray(big[:h, :w])
        low = (base.astype(np.float32) * 0.25 + 100).astype(np.uint8)
        part = base.copy()
        part[h // 2:, : w // 2] = rng.integers(0, 256, (h - h // 2, w // 2), dtype=np.uint8)
        part[: h // 3, w // 2:] = 77
        frames = np.stack([base, low, part])
        ex = ORBextractor(300, sf, nl, ini, mn, max_width=w, max_height=h, max_batch=3)
        ex.extract_batch(frames)
        fused = [[ex.candidates(l, frame=f) for l in range(nl)] for f in range(3)]
        ex.run_stage(2); ex.check()
        for f in range(3):
            for l in range(nl):
                assert np.array_equal(ex.candidates(l, frame=f), fused[f][l]), (w, h, sf, nl, ini, mn, f, l)
                checked += len(fused[f][l])
        ex.close()
    assert checked > 20000
