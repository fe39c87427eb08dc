// Test glue for the adapters of this directory (not part of the product, not an adapter): plain-C entry points that build
// the reference's objects -- the REAL ORB_SLAM2::ORBextractor class (its own header), Frame / MapPoint stand-ins with the
// reference's member names -- run them through the adapter bodies and hand the results back as flat arrays, with the
// same signatures as the wrappers around the reference's own bodies in oracle/ref_wrap_orb.cc / ref_wrap_match.cc, so
// that tests/test_gpu_adapters.py can compare adapter vs reference source vs oracle.
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "pgo_orbslam_shim.h"
#include "pgb200.h"

namespace ORB_SLAM2 {
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}
using namespace ORB_SLAM2;

namespace {
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
    F.mvKeys[i] = k; F.mvKeysUndistorted[i] = k;
  }
  F.mvuRight.assign(n, -1.f);
  F.mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8UC1);
  if (n > 0) memcpy(F.mDescriptors.data, desc, (size_t)n * 32);
  F.mvpMapPoints.assign(n, nullptr);
  F.mvbOutlier.assign(n, false);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
}
cv::Mat desc_row(const uint8_t* d) {
  cv::Mat m(1, 32, CV_8UC1);
  memcpy(m.data, d, 32);
  return m;
}
}  // namespace

extern "C" {

void* pga_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  return new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void pga_orb_destroy(void* h) { delete static_cast<ORBextractor*>(h); }

int pga_orb_extract(void* h, const uint8_t* gray, int w, int h_px, float* kps, uint8_t* desc, int cap) {
  ORBextractor& ex = *static_cast<ORBextractor*>(h);
  cv::Mat image(h_px, w, CV_8UC1, const_cast<uint8_t*>(gray), (size_t)w);
  std::vector<cv::KeyPoint> keypoints;
  cv::Mat descriptors;
  ex(image, cv::noArray(), keypoints, descriptors);
  for (size_t i = 0; i < keypoints.size() && (int)i < cap; i++) {
    const cv::KeyPoint& k = keypoints[i];
    float* o = kps + 7 * i;
    o[0] = k.pt.x; o[1] = k.pt.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave; o[6] = (float)k.class_id;
    memcpy(desc + 32 * i, descriptors.ptr((int)i), 32);
  }
  return (int)keypoints.size();
}

void pga_orb_level(void* h, int level, uint8_t* out, int* w, int* h_px) {
  ORBextractor& ex = *static_cast<ORBextractor*>(h);
  const cv::Mat& m = ex.mvImagePyramid[level];
  *w = m.cols; *h_px = m.rows;
  if (out)
    for (int y = 0; y < m.rows; y++) memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
}

void pga_orb_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  ORBextractor& ex = *static_cast<ORBextractor*>(h);
  const int n = ex.GetLevels();
  const std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

// Same contract as oracle/ref_wrap_match.cc: pgr_search_by_projection (identity poses, map points at (u, v, 1)).
int pga_search_by_projection(const pgb_keypoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* q_uv,
                             const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc, const uint8_t* q_valid, int n_q,
                             float minX, float maxX, float minY, float maxY, float th, const float* scale_factors, int nlevels,
                             int check_ori, int32_t* match_of_cur) {
  Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
  Frame Cur, Last;
  fill_frame(Cur, cur_kps, cur_desc, n_cur, scale_factors, nlevels);
  std::vector<pgb_keypoint> lk(n_q > 0 ? n_q : 1);
  for (int i = 0; i < n_q; i++) lk[i] = pgb_keypoint{q_uv[2 * i], q_uv[2 * i + 1], 31.f, q_angle[i], 0.f, q_octave[i], -1};
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

int pga_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

}  // extern "C"
