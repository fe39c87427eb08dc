// fit_motion -- drop-in for the reference binary of the same name (src/fit_motion.cc): same flags, same JSON inputs
// and outputs, with every numeric stage on the B200 through libpgb200's C-ABI:
//   GetPrincipalRotationAxes            rotation.cc:16-57      -> pgb_principal_rotation_axes
//   GetAngularVelocitiesAroundAxisDirect rotation.cc:103-119    -> pgb_angular_velocities_around_axis
//   sliding-window calibration + L-BFGS  fit_motion.cc:156-221  -> pgb_imu_fit_windows_fwd (all windows in one launch set)
//   per-timestamp averaging              fit_motion.cc:250-262  -> host (speed_sum / speed_cnt from the device)
//   SmoothTimeSeries                     smoothing.cc:56-98     -> pgb_smooth_time_series
//   forward axis                         fit_motion.cc:223-248, :281-292
// There is no CPU fallback: without a B200 the first library call fails and the binary aborts like a failed CHECK.
// Extension (not in the reference): --num_gpus N shards the sliding windows over N devices.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pgb200.h"
#include "check.hpp"
#include "flags.hpp"
#include "json_lite.hpp"

using pgbhost::Table;

namespace {

std::vector<double> Interleave3(const Table& t) {
  std::vector<double> xyz(3 * t.rows());
  for (size_t i = 0; i < t.rows(); i++)
    for (int c = 0; c < 3; c++) xyz[3 * i + c] = t.real[c][i];
  return xyz;
}

struct ShardResult {
  std::vector<double> sum;
  std::vector<int32_t> cnt;
  std::vector<double> x, fx;
  std::vector<int32_t> iters;
  double fwd[3] = {0, 0, 0};
  int first = 0, count = 0;
};

}  // namespace

int main(int argc, char** argv) {
  std::string rotations_json, accelerations_json, locations_json, velocities_out_json, steering_out_json,
      forward_axis_out_json;
  int64_t locations_batch_size = 40, locations_shift_step = 5, optimization_iters = 500;
  double post_smoothing_sigma_sec = 0.003;
  int64_t principal_rotation_axis_integration_interval_usec = 500000;
  double forward_axis_inference_min_velocity_m_s = 5.0, forward_axis_inference_min_rotation_rad = 0.2;
  int64_t num_gpus = 1, device = 0;

  pgbhost::Flags flags;
  flags.String("rotations_json", &rotations_json, "JSON file with raw timestamped 3D rotations from the gyroscope");
  flags.String("accelerations_json", &accelerations_json, "JSON file with raw timestamped 3D accelerations");
  flags.String("locations_json", &locations_json, "JSON file with GPS locations and derived absolute velocities");
  flags.String("velocities_out_json", &velocities_out_json, "JSON file to write calibrated absolute velocities to");
  flags.String("steering_out_json", &steering_out_json, "JSON file to write rotations around the inferred vertical axis to");
  flags.String("forward_axis_out_json", &forward_axis_out_json, "JSON file to write the vehicle forward axis to");
  flags.Int64("locations_batch_size", &locations_batch_size, "sliding window size in GPS measurements");
  flags.Int64("locations_shift_step", &locations_shift_step, "sliding window shift in GPS measurements");
  flags.Int64("optimization_iters", &optimization_iters, "max L-BFGS iterations per calibration run");
  flags.Double("post_smoothing_sigma_sec", &post_smoothing_sigma_sec, "Gaussian kernel width of the final smoothing");
  flags.Int64("principal_rotation_axis_integration_interval_usec", &principal_rotation_axis_integration_interval_usec, "");
  flags.Double("forward_axis_inference_min_velocity_m_s", &forward_axis_inference_min_velocity_m_s, "");
  flags.Double("forward_axis_inference_min_rotation_rad", &forward_axis_inference_min_rotation_rad, "");
  flags.Int64("num_gpus", &num_gpus, "(extension) shard the sliding windows over this many B200s");
  flags.Int64("device", &device, "(extension) first CUDA device to use");
  flags.Parse(argc, argv);

  // Sanity checks (fit_motion.cc:303-313).
  PGB_CHECK(!rotations_json.empty());
  PGB_CHECK(!accelerations_json.empty());
  PGB_CHECK(!locations_json.empty());
  PGB_CHECK(optimization_iters > 0);
  PGB_CHECK(locations_batch_size > 0);
  PGB_CHECK(locations_shift_step > 0);
  PGB_CHECK(locations_batch_size >= locations_shift_step);
  PGB_CHECK(post_smoothing_sigma_sec > 0);
  PGB_CHECK(principal_rotation_axis_integration_interval_usec > 0);
  PGB_CHECK(num_gpus >= 1);

  // Read input JSONs (fit_motion.cc:315-324; only speed_m_s and time_usec of a location are used, :129-132).
  // The two IMU files of an hour-long recording are ~200 MB each: parse them side by side (a failed CHECK in a reader
  // thread aborts the process like anywhere else).
  Table gps, rot, acc;
  {
    std::thread tr([&] { rot = pgbhost::ReadTable(rotations_json, "rotations", {"x", "y", "z"}, "time_usec"); });
    std::thread ta([&] { acc = pgbhost::ReadTable(accelerations_json, "accelerations", {"x", "y", "z"}, "time_usec"); });
    gps = pgbhost::ReadTable(locations_json, "locations", {"speed_m_s"}, "time_usec");
    tr.join();
    ta.join();
  }
  const std::vector<double> gyro_xyz = Interleave3(rot), acc_xyz = Interleave3(acc);
  const int dev0 = (int)device;

  double axes[9];
  int64_t n_iv = 0;
  PGB_CALL(pgb_principal_rotation_axes(dev0, gyro_xyz.data(), rot.integer.data(), rot.rows(),
                                       principal_rotation_axis_integration_interval_usec, axes, &n_iv));
  const double vertical[3] = {axes[0], axes[1], axes[2]};
  if (flags.verbose)
    fprintf(stderr, "I principal rotation axis (%lld intervals): %.9g %.9g %.9g\n", (long long)n_iv, vertical[0], vertical[1], vertical[2]);

  std::thread steering_writer;  // formatting 1.8 M records takes longer than the whole velocity fit: do it on the side
  std::vector<double> steering;
  if (!steering_out_json.empty()) {  // ComputeAndSaveSteeringAngles, fit_motion.cc:138-154
    steering.resize(rot.rows());
    PGB_CALL(pgb_angular_velocities_around_axis(dev0, gyro_xyz.data(), rot.rows(), vertical, steering.data()));
    steering_writer = std::thread([&] {
      pgbhost::JsonWriteTimestampedRealData(rot.integer, steering, steering_out_json, "steering", "angular_velocity");
    });
  }

  if (!velocities_out_json.empty() || !forward_axis_out_json.empty()) {
    const int n_gps = (int)gps.rows();
    const int n_win = pgb_imu_num_windows(n_gps, (int)locations_shift_step);
    const int n_dev = (int)std::min<int64_t>(num_gpus, std::max(1, n_win));
    std::vector<ShardResult> shard(n_dev);
    std::vector<int64_t> merged_t;
    std::vector<std::thread> workers;
    std::vector<std::string> errors(n_dev);
    for (int d = 0; d < n_dev; d++) {
      shard[d].first = (int)((int64_t)n_win * d / n_dev);
      shard[d].count = (int)((int64_t)n_win * (d + 1) / n_dev) - shard[d].first;
      workers.emplace_back([&, d]() {
        ShardResult& r = shard[d];
        pgb_imu* imu = pgb_imu_create(dev0 + d, gyro_xyz.data(), rot.integer.data(), rot.rows(), acc_xyz.data(),
                                      acc.integer.data(), acc.rows(), nullptr);
        if (!imu) { errors[d] = pgb_last_error(); return; }
        const int64_t m = pgb_imu_merged_count(imu);
        r.sum.assign(m, 0.0);
        r.cnt.assign(m, 0);
        r.x.assign(9 * (size_t)std::max(r.count, 1), 0.0);
        r.fx.assign(std::max(r.count, 1), 0.0);
        r.iters.assign(std::max(r.count, 1), 0);
        int32_t used = 0;
        const int rc = pgb_imu_fit_windows_fwd(imu, gps.real[0].data(), gps.integer.data(), n_gps, (int)locations_batch_size,
                                               (int)locations_shift_step, (int)optimization_iters, 1e-5, r.first, r.count,
                                               r.sum.data(), r.cnt.data(), r.x.data(), r.fx.data(), r.iters.data(),
                                               forward_axis_inference_min_velocity_m_s,
                                               forward_axis_inference_min_rotation_rad, r.fwd, &used);
        if (rc) errors[d] = pgb_last_error();
        if (d == 0) {
          merged_t.resize(m);
          if (pgb_imu_merged_events(imu, merged_t.data(), nullptr, nullptr)) errors[d] = pgb_last_error();
        }
        pgb_imu_destroy(imu);
      });
    }
    for (auto& w : workers) w.join();
    for (int d = 0; d < n_dev; d++) PGB_CHECK(errors[d].empty()) << "device " << dev0 + d << ": " << errors[d];
    if (flags.verbose)
      for (int d = 0; d < n_dev; d++)
        for (int w = 0; w < shard[d].count; w++)
          fprintf(stderr, "I Sliding window optimization: %d iterations, result value: %.9g\n", shard[d].iters[w], shard[d].fx[w]);

    // Average the velocities among the sliding windows falling on every IMU measurement (fit_motion.cc:250-262);
    // shards hold disjoint window ranges in ascending order.  With ONE shard this is std::accumulate's order exactly; with
    // --num_gpus > 1 an event covered by windows of two shards is summed as (v1 + v2) + (v3 + v4) instead of
    // ((v1 + v2) + v3) + v4 -- floating-point addition is not associative, so such boundary events (and the forward-axis
    // sum below) can differ from the single-GPU / reference result in the last ulp.
    const size_t m = merged_t.size();
    std::vector<double> averaged, timestamps_sec;
    std::vector<int64_t> timestamps_usec;
    for (size_t i = 0; i < m; i++) {
      double sum = 0.0;
      int cnt = 0;
      for (int d = 0; d < n_dev; d++) {
        if (shard[d].cnt[i] == 0) continue;
        sum = cnt == 0 ? shard[d].sum[i] : sum + shard[d].sum[i];
        cnt += shard[d].cnt[i];
      }
      if (cnt == 0) continue;
      timestamps_usec.push_back(merged_t[i]);
      timestamps_sec.push_back((double)(timestamps_usec.back() - timestamps_usec.front()) * 1e-6);
      averaged.push_back(sum / cnt);
    }
    std::vector<double> smoothed(averaged.size());
    if (!averaged.empty())
      PGB_CALL(pgb_smooth_time_series(dev0, averaged.data(), timestamps_sec.data(), (int64_t)averaged.size(),
                                      timestamps_sec.data(), (int64_t)averaged.size(), post_smoothing_sigma_sec,
                                      smoothed.data()));
    if (!velocities_out_json.empty())
      pgbhost::JsonWriteTimestampedRealData(timestamps_usec, smoothed, velocities_out_json, "velocities", "speed_m_s");

    // forward_axis = total - vertical * (vertical . total); normalise with +1e-5 (fit_motion.cc:281-283)
    double f[3] = {0, 0, 0};
    for (int d = 0; d < n_dev; d++)
      for (int c = 0; c < 3; c++) f[c] += shard[d].fwd[c];
    const double dot = vertical[0] * f[0] + vertical[1] * f[1] + vertical[2] * f[2];
    for (int c = 0; c < 3; c++) f[c] -= vertical[c] * dot;
    const double nrm = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) + 1e-5;
    for (int c = 0; c < 3; c++) f[c] /= nrm;
    if (!forward_axis_out_json.empty()) {
      FILE* out = fopen(forward_axis_out_json.c_str(), "w");
      PGB_CHECK(out != nullptr) << "cannot write " << forward_axis_out_json;
      fprintf(out, "{\n  \"forward_axis\": {\n    \"x\": %s,\n    \"y\": %s,\n    \"z\": %s\n  }\n}\n",
              pgbhost::FormatDouble(f[0]).c_str(), pgbhost::FormatDouble(f[1]).c_str(), pgbhost::FormatDouble(f[2]).c_str());
      fclose(out);
    }
  }
  if (steering_writer.joinable()) steering_writer.join();
  return EXIT_SUCCESS;
}
