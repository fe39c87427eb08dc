// Drop-in bodies for ORB_SLAM2::ORBextractor (thirdparty/orb-slam2/include/ORBextractor.h:45-115, replacing
// src/ORBextractor.cc:410-470 and :1042-1104) on top of libpgb200's C-ABI (include/pgb200.h).  Compiles against the
// REFERENCE'S OWN HEADER, unmodified: callers (Frame::ExtractORB, Frame.cc:251-257) do not change.
//
// The class has no member for the device handle and the header is not ours to edit, so the handle lives in a registry keyed
// by `this`; a maintainer who may touch the header adds `pgb_orb* h_` and drops the registry.  The reference's destructor is
// implicit (`~ORBextractor(){}` in the header), so handles are released at process exit.
//
// Built and exercised by `make -C oracle _ref` (-> oracle/_ref/libpgb_adapters.so) and tests/test_gpu_adapters.py: the
// reference's class, with these bodies, returns keypoints and descriptors identical to the oracle's.
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "ORBextractor.h"
#include "pgb200.h"

namespace ORB_SLAM2 {

namespace {
std::mutex g_mu;
std::map<const ORBextractor*, pgb_orb*> g_handles;
pgb_orb* handle_of(const ORBextractor* e) {
  std::lock_guard<std::mutex> l(g_mu);
  auto it = g_handles.find(e);
  return it == g_handles.end() ? nullptr : it->second;
}
[[noreturn]] void die(const char* what) {  // the reference's failure convention is a fatal glog CHECK
  fprintf(stderr, "F ORBextractor(pgb200): %s: %s\n", what, pgb_last_error());
  abort();
}
// capacity of the device buffers: the largest frame this extractor will be handed (env PGB_ADAPTER_MAX_WxH, default 1920x1080)
void max_size(int* w, int* h) {
  *w = 1920; *h = 1080;
  if (const char* e = getenv("PGB_ADAPTER_MAX_WxH")) sscanf(e, "%dx%d", w, h);
}
}  // namespace

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  int mw, mh;
  max_size(&mw, &mh);
  pgb_orb* h = pgb_orb_create(/*device*/ 0, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, mw, mh, /*max_batch*/ 1, nullptr);
  if (!h) die("pgb_orb_create");
  {
    std::lock_guard<std::mutex> l(g_mu);
    g_handles[this] = h;
  }
  // the scale tables the getters of the header return (ORBextractor.cc:415-431; the fork sizes them nlevels + 1)
  mvScaleFactor.assign(nlevels + 1, 0.f); mvInvScaleFactor.assign(nlevels + 1, 0.f);
  mvLevelSigma2.assign(nlevels + 1, 0.f); mvInvLevelSigma2.assign(nlevels + 1, 0.f);
  if (pgb_orb_scale_factors(h, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data())) die("pgb_orb_scale_factors");
  mvImagePyramid.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  std::vector<int32_t> per(nlevels);
  if (pgb_orb_features_per_level(h, per.data())) die("pgb_orb_features_per_level");
  for (int i = 0; i < nlevels; i++) mnFeaturesPerLevel[i] = per[i];
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask: ignored, as in the reference*/,
                              std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
  if (_image.empty()) return;  // ORBextractor.cc:1045
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);  // :1049
  pgb_orb* h = handle_of(this);
  if (!h) die("extractor was not constructed through the pgb200 adapter");
  const int cap = pgb_orb_max_keypoints(h);
  static_assert(sizeof(cv::KeyPoint) == sizeof(pgb_keypoint), "cv::KeyPoint and pgb_keypoint share one 28-byte layout");
  std::vector<pgb_keypoint> kps(cap);
  std::vector<uint8_t> desc((size_t)cap * 32);
  int32_t n = 0;
  if (pgb_orb_extract(h, image.data, /*where: host in, host out*/ 0, 1, image.cols, image.rows, (size_t)image.step,
                      (size_t)image.step * image.rows, kps.data(), desc.data(), &n, cap))
    die("pgb_orb_extract");
  _keypoints.resize(n);
  if (n) memcpy((void*)_keypoints.data(), kps.data(), (size_t)n * sizeof(pgb_keypoint));
  if (n == 0) { _descriptors.release(); return; }  // :1076-1077
  _descriptors.create(n, 32, CV_8U);
  cv::Mat d = _descriptors.getMat();
  for (int i = 0; i < n; i++) memcpy(d.ptr(i), desc.data() + (size_t)i * 32, 32);
  // mvImagePyramid is public (stereo matching reads it, Frame.cc:477,567-584): refreshed from the device on every call
  for (int level = 0; level < nlevels; level++) {
    int w = 0, hh = 0;
    if (pgb_orb_get_level(h, 0, level, nullptr, &w, &hh)) die("pgb_orb_get_level");
    mvImagePyramid[level].create(hh, w, CV_8UC1);
    std::vector<uint8_t> tight((size_t)w * hh);
    if (pgb_orb_get_level(h, 0, level, tight.data(), &w, &hh)) die("pgb_orb_get_level");
    for (int y = 0; y < hh; y++) memcpy(mvImagePyramid[level].ptr(y), tight.data() + (size_t)y * w, w);
  }
}

}  // namespace ORB_SLAM2
