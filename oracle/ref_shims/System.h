// ORACLE -- TEST INFRASTRUCTURE ONLY.  Shadows thirdparty/orb-slam2/include/System.h (the whole SLAM system) for the
// trajectory post-processing code that only needs its two plain structs (System.h:46-56) and the names the real header
// chain brings into scope.  Not part of the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#include <Eigen/Geometry>
#include <glog/logging.h>
#include <opencv2/core/core.hpp>

using std::vector;
typedef int64_t int64;

namespace ORB_SLAM2 {
struct Pose {
  cv::Vec3d translation;
  Eigen::Quaterniond rotation;
};
struct PoseWithTimestamp {
  Pose pose;
  int64 time_usec;
  bool is_lost;
  int64 frame_id;
};
}  // namespace ORB_SLAM2
