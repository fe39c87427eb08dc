"""Host-side mirror of ORB_SLAM2::ORBmatcher's frame-to-frame path (thirdparty/orb-slam2/include/ORBmatcher.h:41-52)
over the libpgb200 C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib, np_ptr

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # ORBmatcher.cc:38-40


class ORBmatcher:
    """``ORBmatcher(nnratio=0.6, checkOri=True)`` (ORBmatcher.cc:42)."""

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, max_feats: int = 1100, max_batch: int = 1,
                 device: int = 0, stream=None):
        self._h = None
        h = lib().pgb_matcher_create(device, nnratio, int(checkOri), max_feats, max_batch, stream)
        if not h:
            raise _lib.PgbError(-1, _lib.last_error())
        self._h = C.c_void_p(h)
        self.max_feats, self.max_batch = max_feats, max_batch

    def close(self):
        if self._h:
            lib().pgb_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray):
        """ORBmatcher::DescriptorDistance (ORBmatcher.cc:1651-1667); a, b: (32,) or (n, 32) uint8."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        check(lib().pgb_descriptor_distance(np_ptr(a), np_ptr(b), len(a), np_ptr(out), 0, None))
        return int(out[0]) if len(out) == 1 else out

    def SearchByProjection(self, cur_kps, cur_desc, q_uv, q_octave, q_angle, q_desc, q_valid, bounds, th,
                           scale_factors):
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono=True) (ORBmatcher.cc:1332-1474) on flat arrays.

        Returns (nmatches, match_of_cur) where match_of_cur[i] is the query (last-frame map point) index now held
        by current keypoint i, or -1 -- the state of CurrentFrame.mvpMapPoints after the call."""
        n_cur, n_q = len(cur_kps), len(q_octave)
        cap = max(n_cur, n_q, 1)

        def pad(a, dtype, tail=()):
            out = np.zeros((cap,) + tail, dtype)
            a = np.asarray(a)
            if len(a):
                out[:len(a)] = a
            return out
        K = pad(cur_kps, KP_DTYPE); D = pad(cur_desc, np.uint8, (32,)); UV = pad(q_uv, np.float32, (2,))
        O = pad(q_octave, np.int32); A = pad(q_angle, np.float32); QD = pad(q_desc, np.uint8, (32,))
        V = pad(q_valid, np.uint8)
        nc = np.array([n_cur], np.int32); nq = np.array([n_q], np.int32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        m = np.full(cap, -1, np.int32); nm = np.zeros(1, np.int32)
        check(lib().pgb_match_by_projection(self._h, 1, cap, np_ptr(K), np_ptr(D), np_ptr(nc), np_ptr(UV), np_ptr(O),
                                            np_ptr(A), np_ptr(QD), np_ptr(V), np_ptr(nq), bounds[0], bounds[1],
                                            bounds[2], bounds[3], th, np_ptr(sf), len(sf), np_ptr(m), np_ptr(nm), 0))
        return int(nm[0]), m[:n_cur].copy()

    def SearchForInitialization(self, kps1, desc1, kps2, desc2, prev_matched, window_size, bounds):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:407-522).
        Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
        n1, n2 = len(kps1), len(kps2)
        cap = max(n1, n2, 1)

        def pad(a, dtype, tail=()):
            out = np.zeros((cap,) + tail, dtype)
            a = np.asarray(a)
            if len(a):
                out[:len(a)] = a
            return out
        K1 = pad(kps1, KP_DTYPE); D1 = pad(desc1, np.uint8, (32,)); K2 = pad(kps2, KP_DTYPE); D2 = pad(desc2, np.uint8, (32,))
        UV = pad(prev_matched, np.float32, (2,))
        c1 = np.array([n1], np.int32); c2 = np.array([n2], np.int32)
        m12 = np.full(cap, -1, np.int32); nm = np.zeros(1, np.int32)
        check(lib().pgb_match_for_initialization(self._h, 1, cap, np_ptr(K1), np_ptr(D1), np_ptr(c1), np_ptr(K2), np_ptr(D2),
                                                 np_ptr(c2), np_ptr(UV), int(window_size), bounds[0], bounds[1], bounds[2],
                                                 bounds[3], np_ptr(m12), np_ptr(nm), 0))
        return int(nm[0]), m12[:n1].copy(), UV[:n1].copy()

    def SearchByProjectionMapPoints(self, kps, desc, has_map_point, proj_xy, track_level, view_cos, mp_desc, in_view,
                                    mp_observed, bounds, th, scale_factors):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-131) on flat arrays.
        Returns (nmatches, match_of_feature): the map point index this call assigned to each feature, or -1."""
        n, nq = len(kps), len(track_level)
        cap = max(n, nq, 1)

        def pad(a, dtype, tail=()):
            out = np.zeros((cap,) + tail, dtype)
            a = np.asarray(a)
            if len(a):
                out[:len(a)] = a
            return out
        K = pad(kps, KP_DTYPE); D = pad(desc, np.uint8, (32,)); H = pad(has_map_point, np.uint8)
        UV = pad(proj_xy, np.float32, (2,)); L = pad(track_level, np.int32); VC = pad(view_cos, np.float32)
        QD = pad(mp_desc, np.uint8, (32,)); IV = pad(in_view, np.uint8); OB = pad(mp_observed, np.uint8)
        c = np.array([n], np.int32); cq = np.array([nq], np.int32)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        mo = np.full(cap, -1, np.int32); nm = np.zeros(1, np.int32)
        check(lib().pgb_match_map_points(self._h, 1, cap, np_ptr(K), np_ptr(D), np_ptr(c), np_ptr(H), np_ptr(UV), np_ptr(L),
                                         np_ptr(VC), np_ptr(QD), np_ptr(IV), np_ptr(OB), np_ptr(cq), bounds[0], bounds[1],
                                         bounds[2], bounds[3], th, np_ptr(sf), len(sf), np_ptr(mo), np_ptr(nm), 0))
        return int(nm[0]), mo[:n].copy()

    def SearchByBoW(self, kf_desc, kf_angle, kf_has_map_point, kf_featvec, f_desc, f_angle, f_featvec):
        """SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:161-290) on flat arrays.  A feature vector
        (DBoW2::FeatureVector) is a dict {node id: [feature indices]} (iterated in ascending node id, like std::map) or an
        already flattened (node_ids, starts, indices) triple.  Returns (nmatches, match_of_feature): the keyframe feature
        index whose map point each frame feature received, or -1."""
        nk, nf = len(kf_desc), len(f_desc)
        cap = max(nk, nf, 1)

        def pad(a, dtype, tail=()):
            out = np.zeros((cap,) + tail, dtype)
            a = np.asarray(a)
            if len(a):
                out[:len(a)] = a
            return out
        kn, ks, ki = featvec_csr(kf_featvec)
        fn, fs, fi = featvec_csr(f_featvec)
        KD = pad(kf_desc, np.uint8, (32,)); KA = pad(kf_angle, np.float32); KH = pad(kf_has_map_point, np.uint8)
        FD = pad(f_desc, np.uint8, (32,)); FA = pad(f_angle, np.float32)
        ko = np.array([0, len(kn)], np.int32); fo = np.array([0, len(fn)], np.int32); c = np.array([nf], np.int32)
        mo = np.full(cap, -1, np.int32); nm = np.zeros(1, np.int32)
        check(lib().pgb_match_by_bow(self._h, 1, cap, np_ptr(KD), np_ptr(KA), np_ptr(KH), np_ptr(ko), np_ptr(kn), np_ptr(ks),
                                     np_ptr(ki), len(kn), len(ki), np_ptr(FD), np_ptr(FA), np_ptr(c), np_ptr(fo), np_ptr(fn),
                                     np_ptr(fs), np_ptr(fi), len(fn), len(fi), np_ptr(mo), np_ptr(nm), 0))
        return int(nm[0]), mo[:nf].copy()

    def match_consecutive_ptr(self, n_pairs, cap, kps_ptr, desc_ptr, counts_ptr, flow_ptr, max_x, max_y, th,
                              scale_factors, match_ptr, nmatch_ptr):
        sf = np.ascontiguousarray(scale_factors, np.float32)
        check(lib().pgb_match_consecutive(self._h, n_pairs, cap, kps_ptr, desc_ptr, counts_ptr, flow_ptr, max_x, max_y,
                                          th, np_ptr(sf), len(sf), match_ptr, nmatch_ptr))


def median_flow(kps, counts, match_of_cur, n_matches, device: int = 0):
    """pgb_match_median_flow for host arrays: kps (n_pairs + 1, cap) KP_DTYPE, counts (n_pairs + 1,), match_of_cur (n_pairs, cap),
    n_matches (n_pairs,) -> (flow (n_pairs, 2) float32, tracked (n_pairs,) bool).  The per-pair quantity optical_trajectories'
    flow-tracking loop needs: element n/2 of the sorted displacements cur - prev of the matched keypoints, per axis."""
    import torch
    dev = f"cuda:{device}"
    k = torch.from_numpy(np.ascontiguousarray(kps).view(np.uint8)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(counts, np.int32)).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(match_of_cur, np.int32)).to(dev)
    nm = torch.from_numpy(np.ascontiguousarray(n_matches, np.int32)).to(dev)
    n_pairs, cap = match_of_cur.shape
    flow = torch.zeros((n_pairs, 2), dtype=torch.float32, device=dev)
    tracked = torch.zeros(n_pairs, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(device)
    check(lib().pgb_match_median_flow(device, n_pairs, cap, k.data_ptr(), c.data_ptr(), m.data_ptr(), nm.data_ptr(), flow.data_ptr(),
                                      tracked.data_ptr(), st.cuda_stream))
    st.synchronize()
    return flow.cpu().numpy(), tracked.cpu().numpy().astype(bool)


def featvec_csr(fv):
    """DBoW2::FeatureVector as (node_ids u32 ascending, starts i32 [nodes + 1], indices u32)."""
    if isinstance(fv, dict):
        ids = sorted(fv)
        lists = [np.asarray(fv[k], np.uint32).reshape(-1) for k in ids]
        starts = np.zeros(len(ids) + 1, np.int32)
        if ids:
            starts[1:] = np.cumsum([len(x) for x in lists])
        idx = np.concatenate(lists).astype(np.uint32) if ids and starts[-1] else np.zeros(0, np.uint32)
        return np.asarray(ids, np.uint32), starts, idx
    ids, starts, idx = fv
    return np.ascontiguousarray(ids, np.uint32), np.ascontiguousarray(starts, np.int32), np.ascontiguousarray(idx, np.uint32)


def ComputeDistinctiveDescriptors(descriptor_sets):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:259-324) for a list of (n_i, 32) uint8 arrays: index of the
    descriptor with the least median distance to the rest, per map point (-1 for an empty set)."""
    sets = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in descriptor_sets]
    off = np.zeros(len(sets) + 1, np.int32)
    off[1:] = np.cumsum([len(d) for d in sets])
    flat = np.concatenate(sets) if sets and off[-1] else np.zeros((1, 32), np.uint8)
    best = np.full(max(len(sets), 1), -1, np.int32)
    check(lib().pgb_distinctive_descriptors(np_ptr(flat), np_ptr(off), len(sets), np_ptr(best), 0, None))
    return best[:len(sets)].copy()
