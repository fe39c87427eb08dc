// Host-side restatement of the trajectory post-processing that follows the tracking loop in the reference
// (fp64, glue):
//   SmoothHeadingDirections            src/slam/smoothing.cc:11-46 (cv::getGaussianKernel + cv::sepFilter2D, BORDER_REPLICATE)
//   TrajectoryToPCA + eigenvalue gate  src/slam/track_image_sequence.cc:16-30, :74-96
//   ProjectDirections                  src/slam/horizontal_flatten.cc:7-29
//   Projected2DDirectionsToTurnAngles  src/slam/horizontal_flatten.cc:44-64
//   SetPlane / SetTrajectory / dump(2) src/io/json_converters.cc:37-96
// cv::PCA is restated as in csrc/imu.cu (pinned against cv2 golden vectors through the oracle's copy).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "check.hpp"
#include "json_lite.hpp"

namespace pgbhost {

struct Pose {
  double t[3];
  double qw, qx, qy, qz;
};
struct PoseWithTimestamp {
  Pose pose;
  int64_t time_usec;
  bool is_lost;
  int64_t frame_id;
};

// cv::getGaussianKernel(n, sigma, CV_64F) for sigma > 0: exp(-(i - (n-1)/2)^2 / (2 sigma^2)), normalised to sum 1.
inline std::vector<double> GaussianKernel(int n, double sigma) {
  std::vector<double> k(n);
  const double scale2x = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) {
    const double x = i - (n - 1) * 0.5;
    k[i] = std::exp(scale2x * x * x);
    sum += k[i];
  }
  for (double& v : k) v *= 1.0 / sum;
  return k;
}

inline void SmoothHeadingDirections(std::vector<PoseWithTimestamp>* trajectory, int sigma) {
  PGB_CHECK(trajectory != nullptr);
  PGB_CHECK(sigma > 0);
  const int n = (int)trajectory->size(), ks = sigma * 4 + 1, anchor = ks / 2;
  const std::vector<double> kernel = GaussianKernel(ks, sigma);
  std::vector<double> raw(4 * (size_t)n), smooth(4 * (size_t)n);
  for (int i = 0; i < n; i++) {
    const Pose& p = (*trajectory)[i].pose;
    raw[0 * (size_t)n + i] = p.qw; raw[1 * (size_t)n + i] = p.qx; raw[2 * (size_t)n + i] = p.qy; raw[3 * (size_t)n + i] = p.qz;
  }
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < n; i++) {
      double acc = 0;
      for (int k = 0; k < ks; k++) {
        int j = i + k - anchor;
        j = j < 0 ? 0 : (j >= n ? n - 1 : j);  // BORDER_REPLICATE
        acc += raw[c * (size_t)n + j] * kernel[k];
      }
      smooth[c * (size_t)n + i] = acc;  // the 1-tap unit kernel in y is the identity
    }
  for (int i = 0; i < n; i++) {
    const double w = smooth[i], x = smooth[(size_t)n + i], y = smooth[2 * (size_t)n + i], z = smooth[3 * (size_t)n + i];
    const double norm = std::sqrt(w * w + x * x + y * y + z * z);
    Pose& p = (*trajectory)[i].pose;
    p.qw = w / norm; p.qx = x / norm; p.qy = y / norm; p.qz = z / norm;
  }
}

// cv::eigen of a symmetric 3x3 (OpenCV's Jacobi sweep), rows = eigenvectors by descending eigenvalue.
inline void Jacobi3(double A[9], double W[3], double V[9]) {
  const int n = 3;
  const double eps = 2.220446049250313e-16;
  int indR[3] = {0, 0, 0}, indC[3] = {0, 0, 0};
  for (int i = 0; i < 9; i++) V[i] = 0.0;
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  auto scan_row = [&](int k) { int m = k + 1; double mv = std::fabs(A[n * k + m]); for (int i = k + 2; i < n; i++) { double v = std::fabs(A[n * k + i]); if (mv < v) { mv = v; m = i; } } return m; };
  auto scan_col = [&](int k) { int m = 0; double mv = std::fabs(A[k]); for (int i = 1; i < k; i++) { double v = std::fabs(A[n * i + k]); if (mv < v) { mv = v; m = i; } } return m; };
  for (int k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) indR[k] = scan_row(k);
    if (k > 0) indC[k] = scan_col(k);
  }
  for (int iters = 0; iters < n * n * 30; iters++) {
    int k = 0;
    double mv = std::fabs(A[indR[0]]);
    for (int i = 1; i < n - 1; i++) { double v = std::fabs(A[n * i + indR[i]]); if (mv < v) { mv = v; k = i; } }
    int l = indR[k];
    for (int i = 1; i < n; i++) { double v = std::fabs(A[n * indC[i] + i]); if (mv < v) { mv = v; k = indC[i]; l = i; } }
    const double p = A[n * k + l];
    if (std::fabs(p) <= eps) break;
    const double y = (W[l] - W[k]) * 0.5;
    double t = std::fabs(y) + std::hypot(p, y);
    double s = std::hypot(p, t);
    const double c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[n * k + l] = 0;
    W[k] -= t; W[l] += t;
    auto rot = [&](double& v0, double& v1) { const double a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; };
    for (int i = 0; i < k; i++) rot(A[n * i + k], A[n * i + l]);
    for (int i = k + 1; i < l; i++) rot(A[n * k + i], A[n * i + l]);
    for (int i = l + 1; i < n; i++) rot(A[n * k + i], A[n * l + i]);
    for (int i = 0; i < n; i++) rot(V[n * k + i], V[n * l + i]);
    for (int j = 0; j < 2; j++) {
      const int idx = j == 0 ? k : l;
      if (idx < n - 1) indR[idx] = scan_row(idx);
      if (idx > 0) indC[idx] = scan_col(idx);
    }
  }
  for (int k = 0; k < n - 1; k++) {
    int m = k;
    for (int i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
    if (k != m) { std::swap(W[m], W[k]); for (int i = 0; i < n; i++) std::swap(V[n * m + i], V[n * k + i]); }
  }
}

// TrajectoryToPCA (track_image_sequence.cc:16-30): PCA of the translations (3 x N, DATA_AS_COL).
inline void TrajectoryPCA(const std::vector<PoseWithTimestamp>& tr, double eigvec[9], double eigval[3]) {
  const size_t n = tr.size();
  double mean[3] = {0, 0, 0};
  for (const auto& p : tr) for (int c = 0; c < 3; c++) mean[c] += p.pose.t[c];
  for (int c = 0; c < 3; c++) mean[c] /= (double)n;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const auto& p : tr) {
    double d[3];
    for (int c = 0; c < 3; c++) d[c] = p.pose.t[c] - mean[c];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cov[3 * a + b] += d[a] * d[b];
  }
  for (int i = 0; i < 9; i++) cov[i] /= (double)n;
  Jacobi3(cov, eigval, eigvec);
}

// Eigen::Quaterniond::_transformVector
inline void TransformVector(const Pose& p, const double v[3], double out[3]) {
  const double qv[3] = {p.qx, p.qy, p.qz};
  double uv[3] = {qv[1] * v[2] - qv[2] * v[1], qv[2] * v[0] - qv[0] * v[2], qv[0] * v[1] - qv[1] * v[0]};
  for (double& u : uv) u = u + u;
  const double c[3] = {qv[1] * uv[2] - qv[2] * uv[1], qv[2] * uv[0] - qv[0] * uv[2], qv[0] * uv[1] - qv[1] * uv[0]};
  for (int i = 0; i < 3; i++) out[i] = v[i] + p.qw * uv[i] + c[i];
}

// ProjectDirections: plane (2x3, rows = first two PCA eigenvectors) times the camera's optical axis in the world frame.
inline std::vector<std::array<double, 2>> ProjectDirections(const std::vector<PoseWithTimestamp>& tr, const double plane[6]) {
  std::vector<std::array<double, 2>> out;
  const double z_axis[3] = {0, 0, 1};
  for (const auto& p : tr) {
    double d[3];
    TransformVector(p.pose, z_axis, d);
    out.push_back({plane[0] * d[0] + plane[1] * d[1] + plane[2] * d[2], plane[3] * d[0] + plane[4] * d[1] + plane[5] * d[2]});
  }
  return out;
}

inline std::vector<double> Projected2DDirectionsToTurnAngles(const std::vector<std::array<double, 2>>& dirs) {
  std::vector<double> turn(dirs.size(), 0.0);
  for (size_t i = 1; i < dirs.size(); i++) {
    const double px = dirs[i - 1][0], py = dirs[i - 1][1], cx = dirs[i][0], cy = dirs[i][1];
    const double rotation_cos = (px * cx + py * cy) / std::sqrt(px * px + py * py) / std::sqrt(cx * cx + cy * cy);
    const double cross_z = px * cy - py * cx;
    turn[i] = std::acos(rotation_cos) * (cross_z > 0 ? 1.0 : -1.0);
  }
  return turn;
}

// SetPlane + SetTrajectory + dump(2): keys alphabetical at every level, as nlohmann's std::map orders them.
inline void WriteTrajectoryJson(const std::string& path, const std::vector<PoseWithTimestamp>& tr, const double plane[6],
                                const std::vector<std::array<double, 2>>& dirs, const std::vector<double>& turn,
                                int64_t frame_id_offset) {
  PGB_CHECK(tr.size() == dirs.size());   // CHECK_EQ, json_converters.cc:63-68
  PGB_CHECK(tr.size() == turn.size());
  FILE* f = fopen(path.c_str(), "w");
  PGB_CHECK(f != nullptr) << "cannot write " << path;
  auto D = [](double v) { return FormatDouble(v); };
  fprintf(f, "{\n  \"plane\": [\n    [\n      %s,\n      %s,\n      %s\n    ],\n    [\n      %s,\n      %s,\n      %s\n    ]\n  ],\n",
          D(plane[0]).c_str(), D(plane[1]).c_str(), D(plane[2]).c_str(), D(plane[3]).c_str(), D(plane[4]).c_str(), D(plane[5]).c_str());
  fprintf(f, "  \"trajectory\": [\n");
  for (size_t i = 0; i < tr.size(); i++) {
    const PoseWithTimestamp& p = tr[i];
    std::string av = "0";  // point_json[kAngularVelocity] = 0 (an integer) for the first point
    if (i > 0) {
      const double rotation_time_sec = (double)(p.time_usec - tr[i - 1].time_usec) * 1e-6;
      av = D(turn[i] / (rotation_time_sec + 1e-10));
    }
    fprintf(f,
            "    {\n      \"angular_velocity\": %s,\n      \"frame_id\": %lld,\n      \"is_lost\": %s,\n"
            "      \"planar_direction\": [\n        %s,\n        %s\n      ],\n      \"pose\": {\n        \"rotation\": {\n"
            "          \"w\": %s,\n          \"x\": %s,\n          \"y\": %s,\n          \"z\": %s\n        },\n"
            "        \"translation\": [\n          %s,\n          %s,\n          %s\n        ]\n      },\n      \"time_usec\": %lld\n    }%s\n",
            av.c_str(), (long long)(p.frame_id - frame_id_offset), p.is_lost ? "true" : "false", D(dirs[i][0]).c_str(),
            D(dirs[i][1]).c_str(), D(p.pose.qw).c_str(), D(p.pose.qx).c_str(), D(p.pose.qy).c_str(), D(p.pose.qz).c_str(),
            D(p.pose.t[0]).c_str(), D(p.pose.t[1]).c_str(), D(p.pose.t[2]).c_str(), (long long)p.time_usec,
            i + 1 < tr.size() ? "," : "");
  }
  fprintf(f, "  ]\n}\n");
  fclose(f);
}

// The tail of TrackImageSequence (track_image_sequence.cc:63-109) for a finished segment; returns false (and writes
// nothing) for an empty trajectory or when the 3rd eigenvalue is too large.
inline bool FinishTrajectory(std::vector<PoseWithTimestamp> trajectory, int rotation_smooth_sigma, int64_t frame_id_offset,
                             const std::string& out_file, bool verbose) {
  if (trajectory.empty()) return false;
  if (rotation_smooth_sigma > 0) SmoothHeadingDirections(&trajectory, rotation_smooth_sigma);
  double eigvec[9], eigval[3];
  TrajectoryPCA(trajectory, eigvec, eigval);
  if (verbose)
    fprintf(stderr, "I PCA eigenvalues: %.9g %.9g %.9g\n", eigval[0], eigval[1], eigval[2]);
  if (eigval[2] > eigval[1] * 1e-2) {
    fprintf(stderr, "W 3rd eigenvalue was too large, dropping the trajectory. Relative magnitude wrt the 2nd eigenvalue: %g\n",
            eigval[2] / eigval[1]);
    return false;
  }
  const auto dirs = ProjectDirections(trajectory, eigvec);  // rows 0..1 = the plane
  const auto turn = Projected2DDirectionsToTurnAngles(dirs);
  WriteTrajectoryJson(out_file, trajectory, eigvec, dirs, turn, frame_id_offset);
  return true;
}

}  // namespace pgbhost
