// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the REAL ORB_SLAM2::ORBextractor
// (thirdparty/orb-slam2/src/ORBextractor.cc, compiled where it lies by `make -C oracle _ref` against the OpenCV stand-in
// of oracle/ref_shims/pgo_opencv_shim.h -- see that header for what runs from the reference's source and what is
// forwarded to the cv2-pinned primitive restatements).  tests/test_oracle_reference_pin.py runs the oracle's pipeline
// against it.
#include <cstdint>
#include <vector>

#include "ORBextractor.h"

extern "C" void pgr_set_monotonic_alloc(int on);   // ref_bump_alloc.cc

extern "C" {

void* pgr_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
  return new ORB_SLAM2::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void pgr_orb_destroy(void* h) { delete static_cast<ORB_SLAM2::ORBextractor*>(h); }

// operator()(image, noArray(), keypoints, descriptors); kps[cap][7] = x, y, size, angle, response, octave, class_id
// (octave / class_id as float-valued integers); desc[cap][32].  Returns the keypoint count (may exceed cap).
int pgr_orb_extract(void* h, const uint8_t* gray, int w, int h_px, float* kps, uint8_t* desc, int cap) {
  ORB_SLAM2::ORBextractor& ex = *static_cast<ORB_SLAM2::ORBextractor*>(h);
  cv::Mat image(h_px, w, CV_8UC1, const_cast<uint8_t*>(gray), (size_t)w);
  std::vector<cv::KeyPoint> keypoints;
  cv::Mat descriptors;
  pgr_set_monotonic_alloc(1);   // DistributeOctTree's heap-address tie-break needs increasing addresses (ref_bump_alloc.cc)
  ex(image, cv::noArray(), keypoints, descriptors);
  pgr_set_monotonic_alloc(0);
  for (size_t i = 0; i < keypoints.size() && (int)i < cap; i++) {
    const cv::KeyPoint& k = keypoints[i];
    float* o = kps + 7 * i;
    o[0] = k.pt.x; o[1] = k.pt.y; o[2] = k.size; o[3] = k.angle; o[4] = k.response; o[5] = (float)k.octave; o[6] = (float)k.class_id;
    const uint8_t* d = descriptors.ptr((int)i);
    for (int b = 0; b < 32; b++) desc[32 * i + b] = d[b];
  }
  return (int)keypoints.size();
}

// The pyramid level the extractor holds after the last call (public mvImagePyramid): copies it tight into out (if not
// null) and reports its size.
void pgr_orb_level(void* h, int level, uint8_t* out, int* w, int* h_px) {
  ORB_SLAM2::ORBextractor& ex = *static_cast<ORB_SLAM2::ORBextractor*>(h);
  const cv::Mat& m = ex.mvImagePyramid[level];
  *w = m.cols; *h_px = m.rows;
  if (out)
    for (int y = 0; y < m.rows; y++)
      for (int x = 0; x < m.cols; x++) out[(size_t)y * m.cols + x] = m.at<uint8_t>(y, x);
}

void pgr_orb_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  ORB_SLAM2::ORBextractor& ex = *static_cast<ORB_SLAM2::ORBextractor*>(h);
  const int n = ex.GetLevels();
  const std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

}  // extern "C"
