// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the reference's own window loop,
// ComputeAndSaveForwardVelocitiesFromImu (src/fit_motion.cc:156-293), compiled from the reference's file by
// `make -C oracle _ref` (see ref_shims/pgo_fitmotion_prelude.h): per-window LBFGS++ fit on the real
// AccelerometerCalibrator, IntegrateTrajectory, per-event averaging over the overlapping windows, timestamps,
// SmoothTimeSeries, forward-axis accumulation (KahanSum), projection and normalisation.
#include "pgo_fitmotion_prelude.h"

namespace pilotguru { PgrFitCapture g_pgr_fit_capture; }

namespace pgr_fit {
void ComputeAndSaveForwardVelocitiesFromImu(const std::vector<pilotguru::TimestampedVelocity>& gps_velocities,
                                            const std::vector<pilotguru::TimestampedRotationVelocity>& rotations,
                                            const std::vector<pilotguru::TimestampedAcceleration>& accelerations,
                                            const Eigen::Vector3d& vertical_axis, int64_t locations_batch_size,
                                            int64_t locations_shift_step, int64_t max_iters, double post_smoothing_sigma_sec,
                                            const std::string& velocities_out_json, double forward_axis_inference_min_velocity_m_s,
                                            double forward_axis_inference_min_rotation_rad, const std::string& forward_axis_out_json);
}

extern "C" int64_t pgr_fit_motion(const double* gps_v, const int64_t* gps_t, int64_t n_gps, const double* gyro_xyz, const int64_t* gyro_t,
                                  int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t, int64_t n_acc, const double* vertical_axis,
                                  int64_t batch_size, int64_t shift_step, int64_t max_iters, double sigma, double min_vel, double min_rot,
                                  int64_t* out_t_usec, double* out_smoothed, int64_t cap, double* forward_axis3) {
  std::vector<pilotguru::TimestampedVelocity> gps;
  std::vector<pilotguru::TimestampedRotationVelocity> rot;
  std::vector<pilotguru::TimestampedAcceleration> acc;
  for (int64_t i = 0; i < n_gps; i++) gps.push_back({gps_v[i], (long)gps_t[i]});
  for (int64_t i = 0; i < n_gyro; i++) rot.push_back({gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2], (long)gyro_t[i]});
  for (int64_t i = 0; i < n_acc; i++) acc.push_back({acc_xyz[3 * i], acc_xyz[3 * i + 1], acc_xyz[3 * i + 2], (long)acc_t[i]});
  pilotguru::g_pgr_fit_capture = pilotguru::PgrFitCapture();
  pgr_fit::ComputeAndSaveForwardVelocitiesFromImu(gps, rot, acc, Eigen::Vector3d(vertical_axis[0], vertical_axis[1], vertical_axis[2]),
                                                  batch_size, shift_step, max_iters, sigma, "velocities.json", min_vel, min_rot,
                                                  "forward_axis.json");
  const pilotguru::PgrFitCapture& c = pilotguru::g_pgr_fit_capture;
  for (size_t i = 0; i < c.values.size() && (int64_t)i < cap; i++) { out_t_usec[i] = c.times_usec[i]; out_smoothed[i] = c.values[i]; }
  for (int k = 0; k < 3; k++) forward_axis3[k] = c.forward_axis[k];
  return (int64_t)c.values.size();
}

// ---- src/calibration/rotation.cc:16-57 and :103-119, compiled from the reference's file (cv::PCA: see pgo_opencv_shim.h)
#include <opencv2/core/core.hpp>
namespace pilotguru {
cv::Mat GetPrincipalRotationAxes(const std::vector<TimestampedRotationVelocity>& raw_rotations, long integration_interval_usec);
std::vector<double> GetAngularVelocitiesAroundAxisDirect(const std::vector<TimestampedRotationVelocity>& raw_rotations, const cv::Vec3d& axis);
}
extern "C" void pgr_principal_rotation_axes(const double* gyro_xyz, const int64_t* gyro_t, int64_t n, int64_t interval_usec, double* axes9) {
  std::vector<pilotguru::TimestampedRotationVelocity> rot;
  for (int64_t i = 0; i < n; i++) rot.push_back({gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2], (long)gyro_t[i]});
  const cv::Mat ev = pilotguru::GetPrincipalRotationAxes(rot, (long)interval_usec);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) axes9[3 * i + j] = ev.at<double>(i, j);
}
extern "C" void pgr_angular_velocities_around_axis(const double* gyro_xyz, int64_t n, const double* axis, double* out) {
  std::vector<pilotguru::TimestampedRotationVelocity> rot;
  for (int64_t i = 0; i < n; i++) rot.push_back({gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2], 0});
  const std::vector<double> r = pilotguru::GetAngularVelocitiesAroundAxisDirect(rot, cv::Vec3d(axis[0], axis[1], axis[2]));
  for (int64_t i = 0; i < n; i++) out[i] = r[(size_t)i];
}
