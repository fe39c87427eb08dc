// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the REAL reference calibration objective:
// pilotguru::AccelerometerCalibrator (src/calibration/velocity.cc:1-256, the part fit_motion uses; the file's second class,
// FixedForwardAxisCalibrator, needs Eigen's expression templates and is not on the path) with src/geometry/geometry.cc and
// src/interpolation/align_time_series.cc, all compiled where they lie by `make -C oracle _ref` against the Eigen / glog
// stand-ins in oracle/ref_shims.  tests/test_oracle_reference_pin.py runs the oracle's restatement against it.
#include <cstdint>
#include <map>
#include <vector>

#include <calibration/velocity.hpp>

#include <LBFGS.h>  // thirdparty/LBFGS (LBFGS++, vendored by the reference), compiled against the same Eigen stand-in

namespace pilotguru {  // src/slam/smoothing.cc:56-98 (declared in include/slam/smoothing.hpp, which needs ORB-SLAM2's System.h)
std::vector<double> SmoothTimeSeries(const std::vector<double>& data_values, const std::vector<double>& data_timestamps,
                                     const std::vector<double>& target_timestamps, double sigma);
}

namespace {
struct RefCalib {
  std::vector<pilotguru::TimestampedVelocity> gps;
  std::vector<pilotguru::TimestampedRotationVelocity> rot;
  std::vector<pilotguru::TimestampedAcceleration> acc;
  pilotguru::AccelerometerCalibrator* c = nullptr;
};
}  // namespace

extern "C" {

void* pgr_calib_create(const double* gps_v, const int64_t* gps_t, int64_t n_gps, const double* gyro_xyz, const int64_t* gyro_t,
                       int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t, int64_t n_acc) {
  RefCalib* r = new RefCalib;
  for (int64_t i = 0; i < n_gps; i++) r->gps.push_back({gps_v[i], (long)gps_t[i]});
  for (int64_t i = 0; i < n_gyro; i++) r->rot.push_back({gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2], (long)gyro_t[i]});
  for (int64_t i = 0; i < n_acc; i++) r->acc.push_back({acc_xyz[3 * i], acc_xyz[3 * i + 1], acc_xyz[3 * i + 2], (long)acc_t[i]});
  r->c = new pilotguru::AccelerometerCalibrator(r->gps, r->rot, r->acc);
  return r;
}

void pgr_calib_destroy(void* h) {
  RefCalib* r = static_cast<RefCalib*>(h);
  if (!r) return;
  delete r->c;
  delete r;
}

// LossFunction::eval through the LBFGS++ functor signature, like fit_motion.cc:192-197 calls it.
double pgr_calib_eval(void* h, const double* x9, double* grad9) {
  RefCalib* r = static_cast<RefCalib*>(h);
  Eigen::VectorXd x(9), g(9);
  for (int i = 0; i < 9; i++) x[i] = x9[i];
  const double f = (*r->c)(x, g);
  for (int i = 0; i < 9; i++) grad9[i] = g[i];
  return f;
}

// IntegrateTrajectory (velocity.cc:199-256): per merged-event index, in map order.
int64_t pgr_calib_integrate(void* h, const double* x9, int64_t* idx, double* velocity_xyz, double* orientation_wxyz,
                            int64_t* duration_usec, int64_t cap) {
  RefCalib* r = static_cast<RefCalib*>(h);
  const auto out = r->c->IntegrateTrajectory(Eigen::Vector3d(x9[0], x9[1], x9[2]), Eigen::Vector3d(x9[3], x9[4], x9[5]),
                                             Eigen::Vector3d(x9[6], x9[7], x9[8]));
  int64_t k = 0;
  for (const auto& kv : out) {
    if (k < cap) {
      idx[k] = (int64_t)kv.first;
      velocity_xyz[3 * k] = kv.second.velocity.x(); velocity_xyz[3 * k + 1] = kv.second.velocity.y(); velocity_xyz[3 * k + 2] = kv.second.velocity.z();
      orientation_wxyz[4 * k] = kv.second.orientation.w(); orientation_wxyz[4 * k + 1] = kv.second.orientation.x();
      orientation_wxyz[4 * k + 2] = kv.second.orientation.y(); orientation_wxyz[4 * k + 3] = kv.second.orientation.z();
      duration_usec[k] = kv.second.duration_usec;
    }
    k++;
  }
  return k;
}

void pgr_smooth_time_series(const double* values, const double* times, int64_t n, const double* target, int64_t nt, double sigma,
                            double* out) {
  const std::vector<double> v(values, values + n), t(times, times + n), tt(target, target + nt);
  const std::vector<double> r = pilotguru::SmoothTimeSeries(v, t, tt, sigma);
  for (int64_t i = 0; i < nt; i++) out[i] = r[(size_t)i];
}

// The window fit of fit_motion.cc:166-197: LBFGSpp::LBFGSSolver<double> with epsilon 1e-5 and max_iterations, x0 = 0,
// on the REAL AccelerometerCalibrator.  Returns niter; -1 if LBFGS++ threw (line-search step out of range).
int pgr_calib_minimize(void* h, int max_iterations, double* x9, double* fx) {
  RefCalib* r = static_cast<RefCalib*>(h);
  LBFGSpp::LBFGSParam<double> params;
  params.epsilon = 1e-5;
  params.max_iterations = max_iterations;
  LBFGSpp::LBFGSSolver<double> solver(params);
  Eigen::VectorXd x = Eigen::VectorXd::Zero(9);
  double residual = 0;
  int niter;
  try {
    niter = solver.minimize(*r->c, x, residual);
  } catch (const std::exception&) {
    niter = -1;
  }
  for (int i = 0; i < 9; i++) x9[i] = x[i];
  *fx = residual;
  return niter;
}

}  // extern "C"
