// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the reference's own matching functions -- the bodies of
// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono), SearchByProjection(Frame&, vector<MapPoint*>&, th),
// SearchForInitialization, SearchByBoW(KeyFrame*, Frame&, ...), ComputeThreeMaxima, DescriptorDistance
// (thirdparty/orb-slam2/src/ORBmatcher.cc) and Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid (src/Frame.cc),
// compiled from the reference's files by `make -C oracle _ref` behind the stand-in class declarations of
// oracle/ref_shims/pgo_orbslam_shim.h.  The wrapper builds Frame / MapPoint objects from flat arrays the way the oracle's
// entry points take them; tests/test_oracle_reference_pin.py compares the two.
#include <cstdint>
#include <cstring>
#include <vector>

#include "pgo_orbslam_shim.h"
#include "../include/pgb200.h"

namespace ORB_SLAM2 {
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}

using namespace ORB_SLAM2;

namespace {

void set_bounds(float minX, float maxX, float minY, float maxY) {  // Frame.cc:219-220 (first-frame initialisation)
  Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}

cv::Mat identity4() {
  cv::Mat T(4, 4, CV_32F);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T.at<float>(i, j) = i == j ? 1.f : 0.f;
  return T;
}

void fill_frame(Frame& F, const pgb_keypoint* kps, const uint8_t* desc, int n, const float* scale_factors, int nlevels) {
  F.N = n;
  F.mTcw = identity4();
  F.mvKeys.resize(n); F.mvKeysUndistorted.resize(n);
  for (int i = 0; i < n; i++) {
    cv::KeyPoint k(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
    F.mvKeys[i] = k; F.mvKeysUndistorted[i] = k;   // no distortion: mvKeysUn == mvKeys (Frame.cc:399-405 with k1 = 0)
  }
  F.mvuRight.assign(n, -1.f);                       // monocular
  F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8UC1);
  if (n > 0) memcpy(F.mDescriptors.data, desc, (size_t)n * 32);
  F.mvpMapPoints.assign(n, nullptr);
  F.mvbOutlier.assign(n, false);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.AssignFeaturesToGrid();
}

cv::Mat desc_row(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8UC1);
  memcpy(m.data, d, 32);
  return m;
}

}  // namespace

extern "C" {

int pgr_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

// SearchByProjection(CurrentFrame, LastFrame, th, bMono = true).  Every query is a map point of the last frame whose
// projection into the current frame is (u, v): both poses are the identity, fx = fy = 1, cx = cy = 0 and the point sits
// at (u, v, 1), for which the reference's float projection (ORBmatcher.cc:1365-1376) returns u and v exactly.
// match_of_cur[i2] = index of the query whose map point CurrentFrame.mvpMapPoints[i2] holds afterwards, else -1.
int pgr_search_by_projection(const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* q_uv,
                             const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc, const uint8_t* q_valid, int n_q,
                             float minX, float maxX, float minY, float maxY, float th, const float* scale_factors, int nlevels,
                             int check_ori, int32_t* match_of_cur) {
  set_bounds(minX, maxX, minY, maxY);
  Frame Cur, Last;
  fill_frame(Cur, cur_kps, cur_desc, n_cur, scale_factors, nlevels);
  std::vector<pgb_keypoint> lk(n_q > 0 ? n_q : 1);
  for (int i = 0; i < n_q; i++) { lk[i] = pgb_keypoint{q_uv[2 * i], q_uv[2 * i + 1], 31.f, q_angle[i], 0.f, q_octave[i], -1}; }
  std::vector<uint8_t> ld((size_t)(n_q > 0 ? n_q : 1) * 32);
  if (n_q > 0) memcpy(ld.data(), q_desc, (size_t)n_q * 32);
  fill_frame(Last, lk.data(), ld.data(), n_q, scale_factors, nlevels);
  std::vector<MapPoint> mps(n_q > 0 ? n_q : 1);
  for (int i = 0; i < n_q; i++) {
    mps[i].mWorldPos = cv::Mat(3, 1, CV_32F);
    mps[i].mWorldPos.at<float>(0) = q_uv[2 * i]; mps[i].mWorldPos.at<float>(1) = q_uv[2 * i + 1]; mps[i].mWorldPos.at<float>(2) = 1.f;
    mps[i].mDescriptor = desc_row(q_desc + (size_t)i * 32);
    Last.mvpMapPoints[i] = q_valid[i] ? &mps[i] : nullptr;
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int n = matcher.SearchByProjection(Cur, Last, th, true);
  for (int i = 0; i < n_cur; i++) match_of_cur[i] = Cur.mvpMapPoints[i] ? (int32_t)(Cur.mvpMapPoints[i] - mps.data()) : -1;
  return n;
}

int pgr_search_for_initialization(const pgb_keypoint* k1, const uint8_t* d1, int n1, const pgb_keypoint* k2, const uint8_t* d2, int n2,
                                  float* prev_matched, int windowSize, float minX, float maxX, float minY, float maxY, float nnratio,
                                  int check_ori, int32_t* vnMatches12) {
  set_bounds(minX, maxX, minY, maxY);
  const float sf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  Frame F1, F2;
  fill_frame(F1, k1, d1, n1, sf, 8);
  fill_frame(F2, k2, d2, n2, sf, 8);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
  for (int i = 0; i < n1; i++) { vnMatches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return n;
}

int pgr_search_map_points(const pgb_keypoint* kps, const uint8_t* desc, int n, const uint8_t* has_map_point, const float* proj_xy,
                          const int32_t* track_level, const float* view_cos, const uint8_t* mp_desc, const uint8_t* in_view,
                          const uint8_t* mp_observed, int n_mp, float minX, float maxX, float minY, float maxY, float th,
                          const float* scale_factors, int nlevels, float nnratio, int32_t* match_of_feature) {
  set_bounds(minX, maxX, minY, maxY);
  Frame F;
  fill_frame(F, kps, desc, n, scale_factors, nlevels);
  std::vector<MapPoint> held(n > 0 ? n : 1), mps(n_mp > 0 ? n_mp : 1);
  for (int i = 0; i < n; i++)
    if (has_map_point[i]) { held[i].nObs = 1; F.mvpMapPoints[i] = &held[i]; }   // a map point with observations: skipped (:85-87)
  std::vector<MapPoint*> vp(n_mp);
  for (int i = 0; i < n_mp; i++) {
    MapPoint& p = mps[i];
    p.mbTrackInView = in_view[i] != 0; p.mnTrackScaleLevel = track_level[i]; p.mTrackViewCos = view_cos[i];
    p.mTrackProjX = proj_xy[2 * i]; p.mTrackProjY = proj_xy[2 * i + 1];
    p.mDescriptor = desc_row(mp_desc + (size_t)i * 32);
    p.nObs = mp_observed[i] ? 1 : 0;
    vp[i] = &p;
  }
  ORBmatcher matcher(nnratio, true);
  const int nm = matcher.SearchByProjection(F, vp, th);
  for (int i = 0; i < n; i++) {
    MapPoint* q = F.mvpMapPoints[i];
    match_of_feature[i] = (q && q >= mps.data() && q < mps.data() + mps.size()) ? (int32_t)(q - mps.data()) : -1;
  }
  return nm;
}

int pgr_search_by_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_map_point, int kf_n, const uint32_t* kf_node_id,
                      const int32_t* kf_feat_start, const uint32_t* kf_feat_idx, int kf_nodes, const uint8_t* f_desc, const float* f_angle,
                      int f_n, const uint32_t* f_node_id, const int32_t* f_feat_start, const uint32_t* f_feat_idx, int f_nodes, float nnratio,
                      int check_ori, int32_t* match_of_feature) {
  KeyFrame KF;
  Frame F;
  std::vector<MapPoint> mps(kf_n > 0 ? kf_n : 1);
  KF.mvpMapPoints.assign(kf_n, nullptr);
  KF.mvKeysUn.resize(kf_n);
  KF.mDescriptors = cv::Mat(kf_n > 0 ? kf_n : 1, 32, CV_8UC1);
  if (kf_n > 0) memcpy(KF.mDescriptors.data, kf_desc, (size_t)kf_n * 32);
  for (int i = 0; i < kf_n; i++) {
    KF.mvKeysUn[i].angle = kf_angle[i];
    if (kf_has_map_point[i]) KF.mvpMapPoints[i] = &mps[i];
  }
  for (int k = 0; k < kf_nodes; k++)
    KF.mFeatVec[kf_node_id[k]] = std::vector<unsigned int>(kf_feat_idx + kf_feat_start[k], kf_feat_idx + kf_feat_start[k + 1]);
  F.N = f_n;
  F.mvKeys.resize(f_n);
  for (int i = 0; i < f_n; i++) F.mvKeys[i].angle = f_angle[i];
  F.mDescriptors = cv::Mat(f_n > 0 ? f_n : 1, 32, CV_8UC1);
  if (f_n > 0) memcpy(F.mDescriptors.data, f_desc, (size_t)f_n * 32);
  for (int k = 0; k < f_nodes; k++)
    F.mFeatVec[f_node_id[k]] = std::vector<unsigned int>(f_feat_idx + f_feat_start[k], f_feat_idx + f_feat_start[k + 1]);
  std::vector<MapPoint*> matches;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchByBoW(&KF, F, matches);
  for (int i = 0; i < f_n; i++) match_of_feature[i] = matches[i] ? (int32_t)(matches[i] - mps.data()) : -1;
  return n;
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:259-324) for one map point observed by n keyframes (observation i =
// row 0 of keyframe i; the keyframes sit in one array, so the std::map<KeyFrame*, size_t> iterates them in input order).
// Writes the chosen descriptor; returns 0 if the reference left mDescriptor untouched (no observations).
int pgr_distinctive_descriptor(const uint8_t* desc, int n, uint8_t* chosen32) {
  std::vector<KeyFrame> kfs(n > 0 ? n : 1);
  MapPoint mp;
  for (int i = 0; i < n; i++) {
    kfs[i].mDescriptors = cv::Mat(1, 32, CV_8UC1);
    memcpy(kfs[i].mDescriptors.data, desc + (size_t)i * 32, 32);
    mp.mObservations[&kfs[i]] = 0;
  }
  mp.ComputeDistinctiveDescriptors();
  if (mp.mDescriptor.empty()) return 0;
  memcpy(chosen32, mp.mDescriptor.data, 32);
  return 1;
}

}  // extern "C"
