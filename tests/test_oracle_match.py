"""CPU tests of the matcher oracle: DescriptorDistance against numpy popcount, SearchByProjection against an
independent brute-force restatement (no grid), and the golden pipeline fixture."""
import os

import numpy as np

import oracle_lib as O
from pilotguru_b200 import synth


def test_descriptor_distance_popcount():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8); b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        assert O.descriptor_distance(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())
    assert O.descriptor_distance(a[0], a[0]) == 0
    assert O.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def brute_force(cur_k, cur_d, uv, octv, ang, qd, bounds, th, sf):
    """SearchByProjection restated without the grid: candidate set by the window/octave test, ties by
    (cell x, cell y, index), greedy in query order, rotation histogram filter."""
    minX, maxX, minY, maxY = bounds
    invW = np.float32(64) / np.float32(maxX - minX); invH = np.float32(48) / np.float32(maxY - minY)
    posX = np.round((cur_k["x"] - np.float32(minX)) * invW).astype(int)
    posY = np.round((cur_k["y"] - np.float32(minY)) * invH).astype(int)
    ingrid = (posX >= 0) & (posX < 64) & (posY >= 0) & (posY < 48)
    match = np.full(len(cur_k), -1, np.int32); bins = {}
    for i in range(len(uv)):
        u, v = uv[i]
        if u < minX or u > maxX or v < minY or v > maxY:
            continue
        r = np.float32(th) * sf[octv[i]]
        ok = ingrid & (np.abs(cur_k["x"] - u) < r) & (np.abs(cur_k["y"] - v) < r)
        ok &= (cur_k["octave"] >= octv[i] - 1) & (cur_k["octave"] <= octv[i] + 1) & (match < 0)
        idx = np.nonzero(ok)[0]
        if len(idx) == 0:
            continue
        dist = np.unpackbits(cur_d[idx] ^ qd[i], axis=1).sum(axis=1)
        order = np.lexsort((idx, posY[idx], posX[idx], dist))
        j = idx[order[0]]
        if dist[order[0]] <= 100:
            match[j] = i
            rot = np.float32(ang[i]) - cur_k["angle"][j]
            if rot < 0:
                rot += np.float32(360)
            b = int(np.floor(np.float32(rot * np.float32(1.0 / 30)) + 0.5))
            bins.setdefault(0 if b == 30 else b, []).append(j)
    sizes = sorted(((len(v), -k) for k, v in bins.items()), reverse=True)
    keep = [-k for _, k in sizes[:3]]
    if len(sizes) > 1 and sizes[1][0] < 0.1 * sizes[0][0]:
        keep = keep[:1]
    elif len(sizes) > 2 and sizes[2][0] < 0.1 * sizes[0][0]:
        keep = keep[:2]
    for k, v in bins.items():
        if k not in keep:
            match[v] = -1
    return int((match >= 0).sum()), match


def test_search_by_projection_vs_bruteforce():
    orc = O.OrbOracle(500, 1.2, 8, 20, 7)
    sf = orc.tables()["scale"]
    k0, d0 = orc.extract(synth.frame(0, w=640, h=480)); k1, d1 = orc.extract(synth.frame(1, w=640, h=480))
    fl = synth.flow(1, w=640, h=480)
    uv = np.stack([k0["x"] + np.float32(fl[0]), k0["y"] + np.float32(fl[1])], axis=1).astype(np.float32)
    for th in (15.0, 30.0):
        n, m, _ = O.search_by_projection(k1, d1, uv, k0["octave"], k0["angle"], d0, np.ones(len(k0), np.uint8),
                                         (0, 640, 0, 480), th, sf)
        bn, bm = brute_force(k1, d1, uv, k0["octave"], k0["angle"], d0, (0, 640, 0, 480), th, sf)
        assert n == bn and np.array_equal(m, bm)
        assert n > 100
        idx = np.nonzero(m >= 0)[0]
        dx = k1["x"][idx] - k0["x"][m[idx]]
        assert abs(np.median(dx) - fl[0]) < 1.0   # matches follow the known synthetic flow


def test_match_edge_cases():
    sf = np.array([1.0, 1.2], np.float32)
    kz = np.zeros(0, O.KP_DTYPE); dz = np.zeros((0, 32), np.uint8)
    n, m, _ = O.search_by_projection(kz, dz, np.zeros((0, 2), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32),
                                     dz, np.zeros(0, np.uint8), (0, 640, 0, 480), 15.0, sf)
    assert n == 0 and len(m) == 0
    # two queries competing for one target: the earlier query wins, the later finds nothing else
    k = np.zeros(1, O.KP_DTYPE); k["x"] = 100; k["y"] = 100
    d = np.zeros((1, 32), np.uint8)
    uv = np.array([[100, 100], [101, 100]], np.float32)
    n, m, bd = O.search_by_projection(k, d, uv, np.zeros(2, np.int32), np.zeros(2, np.float32),
                                      np.zeros((2, 32), np.uint8), np.ones(2, np.uint8), (0, 640, 0, 480), 15.0, sf)
    assert n == 1 and m[0] == 0 and bd.tolist() == [0, 256]
    # a keypoint whose grid column rounds to 64 is never indexed (Frame.cc:388-392), so it cannot be matched
    k["x"] = 637.0
    n, m, _ = O.search_by_projection(k, d, np.array([[637, 100]], np.float32), np.zeros(1, np.int32),
                                     np.zeros(1, np.float32), np.zeros((1, 32), np.uint8), np.ones(1, np.uint8),
                                     (0, 640, 0, 480), 15.0, sf)
    assert n == 0


def test_match_golden(golden_dir):
    P = np.load(os.path.join(golden_dir, "orb_pipeline_640x480.npz"))
    sf = O.OrbOracle(500, 1.2, 8, 20, 7).tables()["scale"]
    for t in (1, 2):
        n, m = O.match_consecutive(P[f"kps{t-1}"], P[f"desc{t-1}"], P[f"kps{t}"], P[f"desc{t}"],
                                   synth.flow(t, w=640, h=480), 640, 480, 15.0, sf)
        assert n == int(P[f"nmatch{t}"]) and np.array_equal(m, P[f"match{t}"])
