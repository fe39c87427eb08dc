// Drop-in body for ORB_SLAM2::ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) and DescriptorDistance
// (thirdparty/orb-slam2/src/ORBmatcher.cc:1332-1474, :1651-1667) on top of libpgb200's C-ABI.  Caller unchanged:
// Tracking::TrackWithMotionModel (Tracking.cc:858-883).  The adapter flattens the Frame / MapPoint object graph into the
// arrays the C-ABI takes; the projection of the last frame's map points (:1358-1385) stays on the host because it needs
// MapPoint::GetWorldPos.
//
// In this repository it compiles against the stand-in class DECLARATIONS of oracle/ref_shims/pgo_orbslam_shim.h (the real
// Frame.h / MapPoint.h pull in the whole SLAM system, which cannot be built here); the member names are the reference's,
// so the same source compiles against the real headers.  Built by `make -C oracle _ref`, run by tests/test_gpu_adapters.py.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#ifdef PGB_ADAPTER_USE_SHIMS
#include "pgo_orbslam_shim.h"
#else
#include "ORBmatcher.h"
#endif
#include "pgb200.h"

namespace ORB_SLAM2 {

namespace {
std::mutex g_mu;
struct Entry { pgb_matcher* h = nullptr; int cap = 0; float nnratio = 0; bool checkOri = false; };
std::map<const ORBmatcher*, Entry> g_handles;  // keyed by object address: the parameters are re-checked on every call
[[noreturn]] void die(const char* what) {
  fprintf(stderr, "F ORBmatcher(pgb200): %s: %s\n", what, pgb_last_error());
  abort();
}
pgb_matcher* handle_for(const ORBmatcher* m, float nnratio, bool checkOri, int cap) {
  std::lock_guard<std::mutex> l(g_mu);
  Entry& e = g_handles[m];
  // (a matcher object that died and another one constructed at the same address look alike to this registry: the handle is
  // only reused when it was made for the same parameters)
  if (!e.h || e.cap < cap || e.nnratio != nnratio || e.checkOri != checkOri) {
    if (e.h) pgb_matcher_destroy(e.h);
    e.cap = std::max(cap, 2048); e.nnratio = nnratio; e.checkOri = checkOri;
    e.h = pgb_matcher_create(/*device*/ 0, nnratio, checkOri ? 1 : 0, e.cap, /*max_batch*/ 1, nullptr);
    if (!e.h) die("pgb_matcher_create");
  }
  return e.h;
}
}  // namespace

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  if (!bMono) die("libpgb200 covers the monocular path (mvuRight < 0)");
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);  // :1343-1344
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const int nq = LastFrame.N, nc = CurrentFrame.N, cap = std::max(std::max(nq, nc), 1);
  std::vector<float> uv(2 * (size_t)cap, 0.f), ang(cap, 0.f);
  std::vector<int32_t> oct(cap, 0);
  std::vector<uint8_t> valid(cap, 0), qd((size_t)cap * 32, 0);
  for (int i = 0; i < nq; i++) {  // :1358-1385
    MapPoint* pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    cv::Mat x3Dw = pMP->GetWorldPos();
    cv::Mat x3Dc = Rcw * x3Dw + tcw;
    const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
    const float invzc = 1.0 / x3Dc.at<float>(2);
    if (invzc < 0) continue;
    uv[2 * i] = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
    uv[2 * i + 1] = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
    oct[i] = LastFrame.mvKeys[i].octave;
    ang[i] = LastFrame.mvKeysUndistorted[i].angle;
    valid[i] = 1;
    const cv::Mat d = pMP->GetDescriptor();
    memcpy(&qd[(size_t)i * 32], d.data, 32);
  }
  // current frame: keypoints and descriptors padded to the common capacity the C-ABI's [pair][cap] layout wants
  static_assert(sizeof(cv::KeyPoint) == sizeof(pgb_keypoint), "cv::KeyPoint and pgb_keypoint share one 28-byte layout");
  std::vector<pgb_keypoint> ck(cap);
  std::vector<uint8_t> cd((size_t)cap * 32, 0);
  if (nc) memcpy((void*)ck.data(), CurrentFrame.mvKeysUndistorted.data(), (size_t)nc * sizeof(pgb_keypoint));
  for (int i = 0; i < nc; i++) memcpy(&cd[(size_t)i * 32], CurrentFrame.mDescriptors.ptr(i), 32);
  std::vector<int32_t> match(cap, -1);
  int32_t nmatches = 0;
  const int32_t ncur = nc, nqq = nq;
  pgb_matcher* h = handle_for(this, mfNNratio, mbCheckOrientation, cap);
  if (pgb_match_by_projection(h, 1, cap, ck.data(), cd.data(), &ncur, uv.data(), oct.data(), ang.data(), qd.data(), valid.data(), &nqq,
                              Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, th, CurrentFrame.mvScaleFactors.data(),
                              (int)CurrentFrame.mvScaleFactors.size(), match.data(), &nmatches, /*is_device*/ 0))
    die("pgb_match_by_projection");
  for (int i2 = 0; i2 < nc; i2++)
    if (match[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[match[i2]];
  return nmatches;
}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {  // :1651-1667; static
  int32_t d = 0;
  if (pgb_descriptor_distance(a.data, b.data, 1, &d, 0, nullptr)) die("pgb_descriptor_distance");
  return d;
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}  // :42-44

}  // namespace ORB_SLAM2
