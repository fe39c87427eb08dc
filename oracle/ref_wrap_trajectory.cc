// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the reference's trajectory post-processing: SmoothHeadingDirections
// (src/slam/smoothing.cc:11-46) and ProjectDirections / Projected2DDirectionsToTurnAngles (src/slam/horizontal_flatten.cc,
// the whole file), compiled from the reference's files by `make -C oracle _ref` against the OpenCV / Eigen stand-ins
// (cv::getGaussianKernel and cv::sepFilter2D restated from OpenCV 2.4's imgproc/smooth.cpp and filter.cpp in
// ref_shims/pgo_opencv_shim.h).  tests/test_oracle_reference_pin.py compares pilotguru_b200/host/trajectory.hpp
// (through its trajectory_selftest binary) with it.
#include <cstdint>
#include <memory>
#include <vector>

#include <System.h>
#include <slam/horizontal_flatten.hpp>

namespace pilotguru {
void SmoothHeadingDirections(std::vector<ORB_SLAM2::PoseWithTimestamp>* trajectory, int sigma);
}

// poses[n][7] = tx, ty, tz, qw, qx, qy, qz; plane[6] = the 2 x 3 projection plane (first two PCA eigenvectors, as rows).
extern "C" void pgr_finish_trajectory(const double* poses, int64_t n, int sigma, const double* plane, double* out_quat_wxyz, double* out_dirs,
                                      double* out_turn) {
  std::vector<ORB_SLAM2::PoseWithTimestamp> tr((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const double* p = poses + 7 * i;
    tr[i].pose.translation = cv::Vec3d(p[0], p[1], p[2]);
    tr[i].pose.rotation = Eigen::Quaterniond(p[3], p[4], p[5], p[6]);
    tr[i].time_usec = i; tr[i].is_lost = false; tr[i].frame_id = i;
  }
  if (sigma > 0) pilotguru::SmoothHeadingDirections(&tr, sigma);   // track_image_sequence.cc:66-68
  cv::Mat projection_plane(2, 3, CV_64F);
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 3; c++) projection_plane.at<double>(r, c) = plane[3 * r + c];
  const std::unique_ptr<std::vector<cv::Mat>> dirs = pilotguru::ProjectDirections(tr, projection_plane);
  const std::vector<double> turn = pilotguru::Projected2DDirectionsToTurnAngles(*dirs);
  for (int64_t i = 0; i < n; i++) {
    const Eigen::Quaterniond& q = tr[i].pose.rotation;
    out_quat_wxyz[4 * i] = q.w(); out_quat_wxyz[4 * i + 1] = q.x(); out_quat_wxyz[4 * i + 2] = q.y(); out_quat_wxyz[4 * i + 3] = q.z();
    out_dirs[2 * i] = (*dirs)[i].at<double>(0, 0); out_dirs[2 * i + 1] = (*dirs)[i].at<double>(1, 0);
    out_turn[i] = turn[(size_t)i];
  }
}
