// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the body of Optimizer::PoseOptimization
// (thirdparty/orb-slam2/src/Optimizer.cc:239-451), compiled from the reference's file behind the shells of
// ref_shims/pgo_poseopt_shim.h.  Same flat arguments as the oracle's pgo_pose_optimization().
#ifndef PGO_POSEOPT_PART
#error "compile through oracle/Makefile (target _ref)"
#endif
#include <cstdint>
#include <cstring>

namespace ORB_SLAM2 { std::mutex MapPoint::mGlobalMutex; }

extern "C" int pgr_pose_optimization(const float* Tcw_in, const float* kp_xy, const int32_t* kp_octave, const float* mp_xyz,
                                     const uint8_t* has_map_point, int n, const float* inv_level_sigma2, int nlevels, float fx, float fy,
                                     float cx, float cy, float* Tcw_out, uint8_t* outlier) {
  using namespace ORB_SLAM2;
  Frame F;
  F.N = n;
  F.fx = fx; F.fy = fy; F.cx = cx; F.cy = cy;
  F.mTcw = cv::Mat(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) F.mTcw.at<float>(i, j) = Tcw_in[4 * i + j];
  F.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + nlevels);
  std::vector<MapPoint> mps(n > 0 ? n : 1);
  F.mvpMapPoints.assign(n, nullptr);
  F.mvuRight.assign(n, -1.f);
  F.mvbOutlier.assign(n, false);
  F.mvKeysUndistorted.resize(n);
  for (int i = 0; i < n; i++) {
    F.mvKeysUndistorted[i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]);
    F.mvKeysUndistorted[i].octave = kp_octave[i];
    if (has_map_point[i]) {
      mps[i].mWorldPos = cv::Mat(3, 1, CV_32F);
      for (int k = 0; k < 3; k++) mps[i].mWorldPos.at<float>(k) = mp_xyz[3 * i + k];
      F.mvpMapPoints[i] = &mps[i];
    }
  }
  const int r = Optimizer::PoseOptimization(&F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) Tcw_out[4 * i + j] = F.mTcw.at<float>(i, j);
  for (int i = 0; i < n; i++) outlier[i] = F.mvbOutlier[i] ? 1 : 0;
  return r;
}
