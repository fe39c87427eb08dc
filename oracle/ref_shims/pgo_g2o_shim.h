// ORACLE -- TEST INFRASTRUCTURE ONLY.  Shells of the g2o classes whose per-edge arithmetic Optimizer::PoseOptimization
// relies on.  g2o (vendored by the reference under thirdparty/g2o) cannot be built here -- its core is templated over the
// real Eigen -- so oracle/Makefile (target _ref) streams the BODIES that carry arithmetic into these shells:
//   types/se3_ops.hpp:27-39                 skew
//   types/types_six_dof_expmap.cpp:37-42    project2d
//   types/se3quat.h:53-64, 104-110, 217-220, 223-257, 280-285   SE3Quat constructors, operator*, map, exp, normalizeRotation
//   types/types_six_dof_expmap.h:153-157    EdgeSE3ProjectXYZOnlyPose::computeError
//   types/types_six_dof_expmap.cpp:266-296  EdgeSE3ProjectXYZOnlyPose::linearizeOplus, cam_project
//   core/robust_kernel_impl.cpp:65-69, 78-91  RobustKernelHuber::setDelta, robustify
// The Levenberg-Marquardt driver, the quadratic-form accumulation and the dense LDLT stay restatements in
// oracle/pgo_pose.cc (they need g2o's optimizer / solver object graph).  Not part of the product.
#pragma once
#include <cmath>
#include <Eigen/Geometry>

namespace g2o {
using namespace Eigen;
typedef Matrix<double, 6, 1> Vector6d;

// ---- free functions (bodies streamed from the reference)
Matrix3d skew(const Vector3d& v);
Vector2d project2d(const Vector3d& v);

#define PGO_G2O_SE3QUAT_SHELL_BEGIN \
  class SE3Quat {                  \
   protected:                      \
    Quaterniond _r;                \
    Vector3d _t;                   \
   public:
#define PGO_G2O_SE3QUAT_SHELL_END                                    \
    const Quaterniond& rotation() const { return _r; }               \
    const Vector3d& translation() const { return _t; }               \
  };

}  // namespace g2o
