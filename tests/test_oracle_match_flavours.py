"""CPU tests of the oracle's SearchForInitialization (ORBmatcher.cc:407-522) and SearchByProjection(Frame, MapPoints)
(ORBmatcher.cc:46-131) restatements: invariants of the reference's algorithms plus a brute-force numpy restatement of
the per-query candidate sets (no grid)."""
import numpy as np

import oracle_lib as O
from pilotguru_b200 import synth


def _feats(t, w=640, h=480):
    orc = O.OrbOracle(500, 1.2, 8, 20, 7)
    return orc.extract(synth.frame(t, w=w, h=h)), orc.tables()["scale"]


def _dist(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_search_for_initialization_invariants():
    (k1, d1), _ = _feats(0)
    (k2, d2), _ = _feats(3)
    bounds = (0.0, 640.0, 0.0, 480.0)
    pm = np.stack([k1["x"], k1["y"]], axis=1)
    n, m12, pm2 = O.search_for_initialization(k1, d1, k2, d2, pm, 100, bounds, nnratio=0.9)
    sel = np.nonzero(m12 >= 0)[0]
    assert n == len(sel) and n > 20
    assert len(set(m12[sel].tolist())) == n                                     # a target is held by one query only
    assert (k1["octave"][sel] == 0).all() and (k2["octave"][m12[sel]] == 0).all()  # level 0 on both sides
    for i in sel:
        j = m12[i]
        assert _dist(d1[i], d2[j]) <= 50                                         # TH_LOW
        assert abs(k2["x"][j] - pm[i, 0]) < 100 and abs(k2["y"][j] - pm[i, 1]) < 100
        assert pm2[i, 0] == k2["x"][j] and pm2[i, 1] == k2["y"][j]               # vbPrevMatched updated
    rest = np.setdiff1d(np.arange(len(k1)), sel)
    assert np.array_equal(pm2[rest], pm[rest])
    # the true flow between the frames is recovered by most matches
    fl = np.array(synth.flow(1, w=640, h=480)) + np.array(synth.flow(2, w=640, h=480)) + np.array(synth.flow(3, w=640, h=480))
    dxy = np.stack([k2["x"][m12[sel]] - k1["x"][sel], k2["y"][m12[sel]] - k1["y"][sel]], axis=1)
    assert (np.abs(dxy - fl).max(axis=1) <= 1.0).mean() > 0.8
    # without the orientation filter there are at least as many matches
    n2, _, _ = O.search_for_initialization(k1, d1, k2, d2, pm, 100, bounds, nnratio=0.9, check_ori=False)
    assert n2 >= n
    # empty inputs
    n0, m0, _ = O.search_for_initialization(k1[:0], d1[:0], k2, d2, pm[:0], 100, bounds)
    assert n0 == 0 and len(m0) == 0


def test_search_map_points_against_brute_force():
    (k, d), sf = _feats(5)
    rng = np.random.default_rng(8)
    bounds = (0.0, 640.0, 0.0, 480.0)
    # map points = the frame's own keypoints, jittered, with slightly corrupted descriptors
    sel = rng.permutation(len(k))[:300]
    uv = np.stack([k["x"][sel], k["y"][sel]], axis=1) + rng.normal(0, 1.5, (300, 2)).astype(np.float32)
    lv = np.clip(k["octave"][sel] + rng.integers(0, 2, 300), 0, 7).astype(np.int32)
    vc = rng.uniform(0.99, 1.0, 300).astype(np.float32)
    vc[::9] = np.float32(0.998)   # float32(0.998) > the double literal 0.998 the reference compares with: r = 2.5
    qd = d[sel].copy(); flip = rng.integers(0, 32, 300); qd[np.arange(300), flip] ^= rng.integers(0, 256, 300).astype(np.uint8)
    iv = (rng.uniform(size=300) > 0.1).astype(np.uint8); ob = (rng.uniform(size=300) > 0.3).astype(np.uint8)
    has = (rng.uniform(size=len(k)) > 0.9).astype(np.uint8)
    th = 3.0
    n, mo = O.search_map_points(k, d, has, uv, lv, vc, qd, iv, ob, bounds, th, sf, nnratio=0.8)
    # brute force replay (candidate order: cell x, cell y, index)
    invW = np.float32(64) / np.float32(640); invH = np.float32(48) / np.float32(480)
    posX = np.round(k["x"] * invW).astype(int); posY = np.round(k["y"] * invH).astype(int)
    ingrid = (posX >= 0) & (posX < 64) & (posY >= 0) & (posY < 48)
    observed = has.astype(bool).copy(); want = np.full(len(k), -1, np.int32); nm = 0
    for i in range(300):
        if not iv[i]:
            continue
        r = np.float32(2.5 if float(vc[i]) > 0.998 else 4.0) * np.float32(th) * sf[lv[i]]
        ok = ingrid & (np.abs(k["x"] - uv[i, 0]) < r) & (np.abs(k["y"] - uv[i, 1]) < r) & (k["octave"] >= lv[i] - 1) & (k["octave"] <= lv[i])
        # GetFeaturesInArea's cell clipping can only drop keypoints outside the +-r box, never inside: no extra test
        idx = np.nonzero(ok & ~observed)[0]
        if len(idx) == 0:
            continue
        dist = np.unpackbits(d[idx] ^ qd[i], axis=1).sum(axis=1)
        order = np.lexsort((idx, posY[idx], posX[idx], dist))
        b = order[0]
        if dist[b] > 100:
            continue
        if len(order) > 1:
            b2 = order[1]
            if k["octave"][idx[b]] == k["octave"][idx[b2]] and dist[b] > np.float32(0.8) * np.float32(dist[b2]):
                continue
        want[idx[b]] = i; observed[idx[b]] = bool(ob[i]); nm += 1
    assert n == nm and n > 100 and np.array_equal(mo, want)
    assert (mo[has.astype(bool)] == -1).all()


def test_distinctive_descriptor_against_numpy():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 7, 20, 65):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.stack([base ^ (rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8))
                      for _ in range(n)])
        dist = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
        med = np.sort(dist, axis=1)[:, int(0.5 * (n - 1))]
        assert O.distinctive_descriptor(d) == int(np.argmin(med))          # argmin: first index wins ties
    assert O.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == -1


def test_search_by_bow_against_python_restatement():
    """SearchByBoW (ORBmatcher.cc:161-290): the literal merge-loop oracle against an independent dict-based restatement,
    plus the invariants of the reference's algorithm."""
    import bow_util as B
    from pilotguru_b200.matcher import featvec_csr
    for (t0, t1, seed), (ratio, ori) in zip(((0, 1, 1), (2, 5, 2), (3, 3, 3)), ((0.7, True), (0.9, True), (0.75, False))):
        P = B.problem(t0, t1, seed)
        n, mo = O.search_by_bow(P["kf_desc"], P["kf_angle"], P["kf_has"], featvec_csr(P["kf_fv"]), P["f_desc"], P["f_angle"],
                                featvec_csr(P["f_fv"]), nnratio=ratio, check_ori=ori)
        pn, pm = B.python_search_by_bow(P, ratio, ori)
        assert n == pn and np.array_equal(mo, pm) and n > 40
        sel = np.nonzero(mo >= 0)[0]
        assert len(sel) == n and P["kf_has"][mo[sel]].all()
        node_of_f = {i: k for k, v in P["f_fv"].items() for i in v}
        node_of_k = {i: k for k, v in P["kf_fv"].items() for i in v}
        for jf in sel:
            assert node_of_f[jf] == node_of_k[mo[jf]]                            # matches never leave a vocabulary node
            assert _dist(P["kf_desc"][mo[jf]], P["f_desc"][jf]) <= 50            # TH_LOW
    # disjoint vocabularies, empty feature vectors
    P = B.problem(0, 1, 4)
    kfv = {k + 100000: v for k, v in P["kf_fv"].items()}
    n, mo = O.search_by_bow(P["kf_desc"], P["kf_angle"], P["kf_has"], featvec_csr(kfv), P["f_desc"], P["f_angle"], featvec_csr(P["f_fv"]))
    assert n == 0 and (mo == -1).all()
    n, mo = O.search_by_bow(P["kf_desc"], P["kf_angle"], P["kf_has"], featvec_csr({}), P["f_desc"], P["f_angle"], featvec_csr(P["f_fv"]))
    assert n == 0 and (mo == -1).all()
