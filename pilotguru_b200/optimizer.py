"""Host-side mirror of ORB_SLAM2::Optimizer for the part that sits on the tracking path (Optimizer.h:43,
Optimizer.cc:239-451).  All arithmetic happens in libpgb200.so; there is no CPU fallback."""
from __future__ import annotations

import numpy as np

from ._lib import check, lib, np_ptr


def PoseOptimization(Tcw, kp_xy, kp_octave, mp_xyz, has_map_point, inv_level_sigma2, fx, fy, cx, cy, device: int = 0):
    """int Optimizer::PoseOptimization(Frame* pFrame) on flat arrays, for one frame (2-D inputs) or a batch (leading
    frame axis; every frame padded to the same number of features, `has_map_point` = 0 on the padding).

    Returns (n_inliers, Tcw, outlier): the function's return value, the pose handed to pFrame->SetPose (4x4 float32)
    and pFrame->mvbOutlier."""
    T = np.ascontiguousarray(Tcw, np.float32)
    single = T.ndim == 2
    T = T.reshape(-1, 16)
    nf = len(T)
    xy = np.ascontiguousarray(kp_xy, np.float32).reshape(nf, -1, 2)
    cap = max(xy.shape[1], 1)

    def shaped(a, dtype, tail):
        a = np.ascontiguousarray(a, dtype).reshape((nf, -1) + tail)
        if a.shape[1] == cap:
            return a
        out = np.zeros((nf, cap) + tail, dtype)
        out[:, :a.shape[1]] = a
        return out
    xy = shaped(xy, np.float32, (2,)); oc = shaped(kp_octave, np.int32, ()); X = shaped(mp_xyz, np.float32, (3,))
    hm = shaped(has_map_point, np.uint8, ())
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    counts = np.full(nf, np.asarray(kp_octave).reshape(nf, -1).shape[1], np.int32)
    To = np.zeros((nf, 16), np.float32); out = np.zeros((nf, cap), np.uint8); ni = np.zeros(nf, np.int32)
    check(lib().pgb_pose_optimization(device, nf, cap, np_ptr(T), np_ptr(xy), np_ptr(oc), np_ptr(X), np_ptr(hm), np_ptr(counts),
                                      np_ptr(s2), len(s2), fx, fy, cx, cy, np_ptr(To), np_ptr(out), np_ptr(ni), 0, None))
    n = int(counts[0])
    if single:
        return int(ni[0]), To[0].reshape(4, 4), out[0, :n].copy()
    return ni, To.reshape(nf, 4, 4), out[:, :n].copy()
