"""GPU parity tests of the CUDA projection matcher, through the C-ABI, against the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

pytestmark = pytest.mark.gpu


def test_descriptor_distance():
    from pilotguru_b200.matcher import ORBmatcher
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (4096, 32), dtype=np.uint8); b = rng.integers(0, 256, (4096, 32), dtype=np.uint8)
    got = ORBmatcher.DescriptorDistance(a, b)
    assert np.array_equal(got, np.unpackbits(a ^ b, axis=1).sum(axis=1))
    assert ORBmatcher.DescriptorDistance(a[0], a[0]) == 0
    assert ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def _queries(k0, fl):
    return np.stack([k0["x"] + np.float32(fl[0]), k0["y"] + np.float32(fl[1])], axis=1).astype(np.float32)


@pytest.mark.parametrize("th", [15.0, 30.0, 60.0])
@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_projection_parity(golden_dir, th, check_ori):
    from pilotguru_b200.matcher import ORBmatcher
    P = np.load(os.path.join(golden_dir, "orb_pipeline_640x480.npz"))
    sf = O.OrbOracle(500, 1.2, 8, 20, 7).tables()["scale"]
    m = ORBmatcher(0.9, check_ori, max_feats=600)
    for t in (1, 2):
        k0, d0, k1, d1 = P[f"kps{t-1}"], P[f"desc{t-1}"], P[f"kps{t}"], P[f"desc{t}"]
        uv = _queries(k0, synth.flow(t, w=640, h=480))
        valid = np.ones(len(k0), np.uint8); valid[::7] = 0          # some queries without a map point
        on, om, _ = O.search_by_projection(k1, d1, uv, k0["octave"], k0["angle"], d0, valid, (0, 640, 0, 480), th, sf,
                                           check_ori)
        gn, gm = m.SearchByProjection(k1, d1, uv, k0["octave"], k0["angle"], d0, valid, (0.0, 640.0, 0.0, 480.0), th, sf)
        assert gn == on and np.array_equal(gm, om)
    m.close()


def test_greedy_exclusion_and_overflow_paths():
    """Crowded scene: hundreds of identical-position targets inside every window force the sorted-list overflow
    and the sequential re-sweep; result must still equal the reference's greedy loop."""
    from pilotguru_b200.matcher import ORBmatcher
    rng = np.random.default_rng(4)
    n = 900
    k = np.zeros(n, O.KP_DTYPE)
    k["x"] = rng.integers(300, 340, n).astype(np.float32); k["y"] = rng.integers(200, 240, n).astype(np.float32)
    k["octave"] = rng.integers(0, 3, n); k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    d = np.tile(base, (n, 1)); flip = rng.integers(0, 32, n); d[np.arange(n), flip] ^= rng.integers(0, 256, n).astype(np.uint8)
    q = 700
    uv = np.stack([rng.uniform(300, 340, q), rng.uniform(200, 240, q)], axis=1).astype(np.float32)
    qo = rng.integers(0, 3, q).astype(np.int32); qa = rng.uniform(0, 360, q).astype(np.float32)
    qd = np.tile(base, (q, 1)); qd[np.arange(q), rng.integers(0, 32, q)] ^= rng.integers(0, 256, q).astype(np.uint8)
    sf = np.array([1.0, 1.2, 1.44], np.float32)
    on, om, _ = O.search_by_projection(k, d, uv, qo, qa, qd, np.ones(q, np.uint8), (0, 640, 0, 480), 15.0, sf)
    m = ORBmatcher(0.9, True, max_feats=1000)
    gn, gm = m.SearchByProjection(k, d, uv, qo, qa, qd, np.ones(q, np.uint8), (0.0, 640.0, 0.0, 480.0), 15.0, sf)
    assert gn == on and np.array_equal(gm, om)
    assert on > 100
    m.close()


def test_empty_inputs():
    from pilotguru_b200.matcher import ORBmatcher
    m = ORBmatcher(0.9, True, max_feats=64)
    sf = np.array([1.0, 1.2], np.float32)
    kz = np.zeros(0, O.KP_DTYPE); dz = np.zeros((0, 32), np.uint8)
    n, mm = m.SearchByProjection(kz, dz, np.zeros((0, 2), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32), dz,
                                 np.zeros(0, np.uint8), (0.0, 640.0, 0.0, 480.0), 15.0, sf)
    assert n == 0 and len(mm) == 0
    m.close()


def test_extract_then_match_consecutive_device_resident(golden_dir):
    """The bench path: frames -> extract (device buffers) -> pgb_match_consecutive, all on the GPU, checked
    against the golden pipeline fixture."""
    import torch
    from pilotguru_b200.matcher import ORBmatcher
    from pilotguru_b200.orb import ORBextractor
    P = np.load(os.path.join(golden_dir, "orb_pipeline_640x480.npz"))
    B = 3
    frames = torch.from_numpy(np.stack([synth.frame(t, w=640, h=480) for t in range(B)])).cuda()
    ex = ORBextractor(500, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=B)
    cap = ex.cap
    kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.extract_ptr(frames.data_ptr(), 3, B, 640, 480, 640, 640 * 480, kps.data_ptr(), desc.data_ptr(),
                   counts.data_ptr(), cap)
    ex.check()
    m = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, stream=ex.stream)
    flow = torch.tensor([synth.flow(t, w=640, h=480) for t in range(1, B)], dtype=torch.float32, device="cuda")
    match = torch.full((B - 1, cap), -2, dtype=torch.int32, device="cuda")
    nm = torch.zeros(B - 1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    m.match_consecutive_ptr(B - 1, cap, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), flow.data_ptr(), 640.0, 480.0,
                            15.0, ex.GetScaleFactors(), match.data_ptr(), nm.data_ptr())
    ex.check()
    c = counts.cpu().numpy(); mm = match.cpu().numpy(); nmh = nm.cpu().numpy()
    kh = kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28).copy().view(O.KP_DTYPE).reshape(B, cap)
    for t in range(B):
        assert np.array_equal(kh[t, :c[t]], P[f"kps{t}"])
    for t in (1, 2):
        assert nmh[t - 1] == int(P[f"nmatch{t}"])
        assert np.array_equal(mm[t - 1, :c[t]], P[f"match{t}"])
    m.close(); ex.close()


def test_match_consecutive_at_the_baseline_config_with_retry_branches():
    """The BENCHMARKED matcher path at the BASELINE size: pgb_match_consecutive over a batch of 1080p frames with 1000
    features (cap = the extractor's), every pair compared with the oracle (= SearchByProjection at th, Tracking.cc:876-883's
    retry at 2*th when fewer than 20 matches, both pinned to the reference's own function bodies).  The batch contains the
    three outcomes of that retry: pairs that pass at th = 15; pairs that have < 20 matches at 15 and recover at 30 (only
    low-octave features + a flow that is 24 px off: windows of 15 / 18 px miss, 30 / 36 px hit); pairs that fail at both
    thresholds (descriptors replaced by noise)."""
    import torch
    from pilotguru_b200.matcher import ORBmatcher
    from pilotguru_b200.orb import ORBextractor
    W, H, NF, B = 1920, 1080, 1000, 16
    orc = O.OrbOracle(NF, 1.2, 8, 20, 7)
    sf = orc.tables()["scale"]
    feats = [orc.extract(synth.frame(t)) for t in range(B + 1)]
    flows = np.array([synth.flow(t) for t in range(1, B + 1)], np.float32)
    rng = np.random.default_rng(77)
    low = lambda kd: (kd[0][kd[0]["octave"] <= 1], kd[1][kd[0]["octave"] <= 1])
    for t in (5, 6, 7):                                  # pairs 4..7 touch a frame reduced to octaves 0 and 1
        feats[t] = low(feats[t])
    flows[5] += np.float32([24, 0]); flows[6] += np.float32([0, -24])      # pairs 5, 6: low octaves only AND a 24 px flow error
    feats[11] = (feats[11][0], rng.integers(0, 256, feats[11][1].shape, dtype=np.uint8))   # pairs 10, 11: noise descriptors
    flows[13] += np.float32([500, 300])                                     # pair 13: windows land nowhere near
    ex = ORBextractor(NF, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
    cap = ex.cap
    ex.close()
    kps = np.zeros((B + 1, cap), O.KP_DTYPE); desc = np.zeros((B + 1, cap, 32), np.uint8); cnt = np.zeros(B + 1, np.int32)
    for t, (k, d) in enumerate(feats):
        kps[t, :len(k)] = k; desc[t, :len(k)] = d; cnt[t] = len(k)
    want_n, want_m, n15 = [], [], []
    for p in range(B):
        (k0, d0), (k1, d1) = feats[p], feats[p + 1]
        n, m = O.match_consecutive(k0, d0, k1, d1, flows[p], float(W), float(H), 15.0, sf)
        want_n.append(n); want_m.append(m)
        uv = _queries(k0, flows[p])
        n15.append(O.search_by_projection(k1, d1, uv, k0["octave"], k0["angle"], d0, np.ones(len(k0), np.uint8), (0, W, 0, H), 15.0, sf)[0])
    n15 = np.array(n15); want_n = np.array(want_n)
    assert ((n15 >= 20)).sum() >= 8                                         # ordinary pairs
    assert ((n15 < 20) & (want_n >= 20)).sum() >= 2, (n15, want_n)          # recovered by the 2*th retry
    assert ((n15 < 20) & (want_n < 20)).sum() >= 2, (n15, want_n)           # fail at both thresholds
    tk = torch.from_numpy(kps.view(np.uint8).reshape(B + 1, cap, 28).copy()).cuda()
    td = torch.from_numpy(desc).cuda(); tc = torch.from_numpy(cnt).cuda(); tf = torch.from_numpy(flows).cuda()
    match = torch.full((B, cap), -2, dtype=torch.int32, device="cuda"); nm = torch.zeros(B, dtype=torch.int32, device="cuda")
    m = ORBmatcher(0.9, True, max_feats=cap, max_batch=B)
    torch.cuda.synchronize()
    for _ in range(2):                                                      # second call: scratch state left by the first must not matter
        m.match_consecutive_ptr(B, cap, tk.data_ptr(), td.data_ptr(), tc.data_ptr(), tf.data_ptr(), float(W), float(H), 15.0, sf,
                                match.data_ptr(), nm.data_ptr())
        torch.cuda.synchronize()
        gm = match.cpu().numpy(); gn = nm.cpu().numpy()
        assert np.array_equal(gn, want_n), (gn, want_n)
        for p in range(B):
            assert np.array_equal(gm[p, :cnt[p + 1]], want_m[p]), p
    m.close()


def _feats_small(t, w=640, h=480):
    orc = O.OrbOracle(500, 1.2, 8, 20, 7)
    return orc.extract(synth.frame(t, w=w, h=h)), orc.tables()["scale"]


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_for_initialization_matches_oracle(check_ori):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:407-522): vnMatches12, the count and the updated
    vbPrevMatched are bit-exact against the literal restatement."""
    from pilotguru_b200.matcher import ORBmatcher
    (k1, d1), _ = _feats_small(0)
    bounds = (0.0, 640.0, 0.0, 480.0)
    m = ORBmatcher(0.9, check_ori, max_feats=600)
    for t2, win in ((1, 100), (4, 100), (2, 20)):
        (k2, d2), _ = _feats_small(t2)
        pm = np.stack([k1["x"], k1["y"]], axis=1)
        on, om, opm = O.search_for_initialization(k1, d1, k2, d2, pm, win, bounds, nnratio=0.9, check_ori=check_ori)
        gn, gm, gpm = m.SearchForInitialization(k1, d1, k2, d2, pm, win, bounds)
        assert gn == on and on > 10 and np.array_equal(gm, om) and np.array_equal(gpm, opm)
    gn, gm, _ = m.SearchForInitialization(k1[:0], d1[:0], k2, d2, np.zeros((0, 2), np.float32), 100, bounds)
    assert gn == 0 and len(gm) == 0


def test_search_map_points_matches_oracle():
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:46-131)."""
    from pilotguru_b200.matcher import ORBmatcher
    (k, d), sf = _feats_small(5)
    rng = np.random.default_rng(8)
    bounds = (0.0, 640.0, 0.0, 480.0)
    nq = 400
    sel = rng.integers(0, len(k), nq)
    uv = (np.stack([k["x"][sel], k["y"][sel]], axis=1) + rng.normal(0, 1.5, (nq, 2))).astype(np.float32)
    lv = np.clip(k["octave"][sel] + rng.integers(0, 2, nq), 0, 7).astype(np.int32)
    vc = rng.uniform(0.99, 1.0, nq).astype(np.float32)
    vc[::9] = np.float32(0.998)   # float32(0.998) > the double literal 0.998 the reference compares with: r = 2.5
    qd = d[sel].copy(); qd[np.arange(nq), rng.integers(0, 32, nq)] ^= rng.integers(0, 256, nq).astype(np.uint8)
    iv = (rng.uniform(size=nq) > 0.1).astype(np.uint8); ob = (rng.uniform(size=nq) > 0.3).astype(np.uint8)
    has = (rng.uniform(size=len(k)) > 0.9).astype(np.uint8)
    for th, ratio in ((1.0, 0.8), (3.0, 0.8), (5.0, 0.6)):
        m = ORBmatcher(ratio, True, max_feats=600)
        on, om = O.search_map_points(k, d, has, uv, lv, vc, qd, iv, ob, bounds, th, sf, nnratio=ratio)
        gn, gm = m.SearchByProjectionMapPoints(k, d, has, uv, lv, vc, qd, iv, ob, bounds, th, sf)
        assert gn == on and on > 50 and np.array_equal(gm, om)


def test_distinctive_descriptors_match_oracle():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:259-324) for a batch of map points."""
    from pilotguru_b200.matcher import ComputeDistinctiveDescriptors
    rng = np.random.default_rng(12)
    sets = []
    for n in [1, 2, 3, 5, 0, 8, 31, 32, 33, 64, 100, 256, 4, 17]:
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        sets.append(np.stack([base ^ (rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8))
                              for _ in range(n)]) if n else np.zeros((0, 32), np.uint8))
    got = ComputeDistinctiveDescriptors(sets)
    want = [O.distinctive_descriptor(s) for s in sets]
    assert got.tolist() == want
    with pytest.raises(Exception):
        ComputeDistinctiveDescriptors([rng.integers(0, 256, (257, 32), dtype=np.uint8)])   # capacity: loud, never truncated


def test_search_by_bow_matches_oracle():
    """ORBmatcher::SearchByBoW (ORBmatcher.cc:161-290): vpMapPointMatches and the count, bit-exact against the literal
    restatement, for small and large vocabulary nodes, shuffled node lists, with and without the orientation filter."""
    import bow_util as B
    from pilotguru_b200 import PgbError
    from pilotguru_b200.matcher import ORBmatcher, featvec_csr
    total = 0
    for (t0, t1, seed, cell), (ratio, ori) in zip(((0, 1, 1, 80), (2, 5, 2, 80), (3, 3, 3, 40), (1, 2, 4, 1000)),
                                                  ((0.7, True), (0.9, True), (0.75, False), (0.8, True))):
        P = B.problem(t0, t1, seed, cell=cell)
        m = ORBmatcher(ratio, ori, max_feats=600)
        on, om = O.search_by_bow(P["kf_desc"], P["kf_angle"], P["kf_has"], featvec_csr(P["kf_fv"]), P["f_desc"], P["f_angle"],
                                 featvec_csr(P["f_fv"]), nnratio=ratio, check_ori=ori)
        gn, gm = m.SearchByBoW(P["kf_desc"], P["kf_angle"], P["kf_has"], P["kf_fv"], P["f_desc"], P["f_angle"], P["f_fv"])
        assert gn == on and np.array_equal(gm, om)
        total += on
        m.close()
    assert total > 200
    m = ORBmatcher(0.7, True, max_feats=600)
    P = B.problem(0, 1, 5)
    gn, gm = m.SearchByBoW(P["kf_desc"], P["kf_angle"], P["kf_has"], {}, P["f_desc"], P["f_angle"], P["f_fv"])   # no common node
    assert gn == 0 and (gm == -1).all()
    bad = dict(P["f_fv"]); k0, k1 = sorted(bad)[:2]; bad[k1] = list(bad[k1]) + [bad[k0][0]]                       # a feature in two nodes
    with pytest.raises(PgbError):
        m.SearchByBoW(P["kf_desc"], P["kf_angle"], P["kf_has"], P["kf_fv"], P["f_desc"], P["f_angle"], bad)
    m.close()


def test_median_flow_equals_nth_element():
    """pgb_match_median_flow against the host code it replaced in optical_trajectories (std::nth_element at index n / 2 of the
    displacements of the matched keypoints, per axis; tracked = at least 20 matches): random match tables incl. duplicated
    displacements, pairs below the threshold, a pair without matches and a full table."""
    from pilotguru_b200._lib import KP_DTYPE
    from pilotguru_b200.matcher import median_flow
    rng = np.random.default_rng(77)
    cap, n_pairs = 1033, 9
    kps = np.zeros((n_pairs + 1, cap), KP_DTYPE)
    kps["x"] = rng.integers(0, 1920, kps.shape).astype(np.float32) * np.float32(1.2)     # quantised: many equal displacements
    kps["y"] = rng.random(kps.shape).astype(np.float32) * 1080
    counts = rng.integers(600, cap + 1, n_pairs + 1).astype(np.int32)
    counts[3] = cap
    match = np.full((n_pairs, cap), -1, np.int32)
    nm = np.zeros(n_pairs, np.int32)
    for p in range(n_pairs):
        want = [700, 19, 20, 0, 1000, 33, 512, 21, 5][p]
        want = min(want, counts[p + 1], counts[p])
        cur = rng.choice(counts[p + 1], want, replace=False)
        match[p, cur] = rng.choice(counts[p], want, replace=False)
        nm[p] = want
    flow, tracked = median_flow(kps, counts, match, nm)
    for p in range(n_pairs):
        t = np.nonzero(match[p, :counts[p + 1]] >= 0)[0]
        assert tracked[p] == (nm[p] >= 20 and len(t) > 0)
        if tracked[p]:
            q = match[p, t]
            fx = np.sort(kps["x"][p + 1, t] - kps["x"][p, q]); fy = np.sort(kps["y"][p + 1, t] - kps["y"][p, q])
            assert flow[p, 0] == fx[len(fx) // 2] and flow[p, 1] == fy[len(fy) // 2], p
        else:
            assert flow[p, 0] == 0 and flow[p, 1] == 0
