// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the g2o arithmetic that Optimizer::PoseOptimization relies on,
// compiled from the reference's vendored g2o sources (see ref_shims/pgo_g2o_shim.h for the line ranges):
// SE3Quat::exp / operator* / map, EdgeSE3ProjectXYZOnlyPose::computeError / linearizeOplus, RobustKernelHuber::robustify.
// tests/test_oracle_reference_pin.py compares the corresponding pieces of oracle/pgo_pose.cc with them.
#include <cstdint>

#ifndef PGO_G2O_PART_INCLUDED
#error "compile through oracle/Makefile (target _ref): the recipe puts the class shells, filled with the reference bodies, in front of this file"
#endif

using namespace g2o;

static SE3Quat make_pose(const double* q_wxyz_t) {
  return SE3Quat(Eigen::Quaterniond(q_wxyz_t[0], q_wxyz_t[1], q_wxyz_t[2], q_wxyz_t[3]), Eigen::Vector3d(q_wxyz_t[4], q_wxyz_t[5], q_wxyz_t[6]));
}
static void store_pose(const SE3Quat& s, double* out) {
  out[0] = s.rotation().w(); out[1] = s.rotation().x(); out[2] = s.rotation().y(); out[3] = s.rotation().z();
  out[4] = s.translation()[0]; out[5] = s.translation()[1]; out[6] = s.translation()[2];
}

extern "C" {

// VertexSE3Expmap::oplusImpl: setEstimate(SE3Quat::exp(update) * estimate())
void pgr_se3_oplus(const double* update6, const double* pose7, double* out7) {
  Vector6d u;
  for (int i = 0; i < 6; i++) u[i] = update6[i];
  store_pose(SE3Quat::exp(u) * make_pose(pose7), out7);
}

void pgr_pose_edge(const double* pose7, const double* Xw, const double* obs, double fx, double fy, double cx, double cy, double* err2,
                   double* J12) {
  VertexSE3Expmap v;
  v.setEstimate(make_pose(pose7));
  EdgeSE3ProjectXYZOnlyPose e;
  e._vertices[0] = &v;
  e._measurement = Eigen::Vector2d(obs[0], obs[1]);
  e.Xw = Eigen::Vector3d(Xw[0], Xw[1], Xw[2]);
  e.fx = fx; e.fy = fy; e.cx = cx; e.cy = cy;
  e.computeError();
  e.linearizeOplus();
  err2[0] = e._error[0]; err2[1] = e._error[1];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 6; j++) J12[6 * i + j] = e._jacobianOplusXi(i, j);
}

void pgr_huber(double delta, double e, double* rho3) {
  RobustKernelHuber k;
  k.setDelta(delta);
  Eigen::Vector3d rho;
  k.robustify(e, rho);
  rho3[0] = rho[0]; rho3[1] = rho[1]; rho3[2] = rho[2];
}

}  // extern "C"
