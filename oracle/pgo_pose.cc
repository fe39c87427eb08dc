// ORACLE -- TEST INFRASTRUCTURE ONLY (see pgo_orb.cc header for the rules).
// CPU restatement of Optimizer::PoseOptimization (thirdparty/orb-slam2/src/Optimizer.cc:239-451), monocular edges
// only (mvuRight < 0), together with the parts of the vendored g2o it drives -- restated, not compiled (g2o needs Eigen,
// which this image does not have):
//   thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:60-186   LM step, lambda init (tau 1e-5), scale
//   thirdparty/g2o/g2o/core/sparse_optimizer.cpp:100-114,206-267,354-419  robust chi2, level-0 active edges, optimize()
//   thirdparty/g2o/g2o/core/block_solver.hpp:502-604                      buildSystem, setLambda / restoreDiagonal
//   thirdparty/g2o/g2o/core/base_unary_edge.hpp:43-72                     b -= rho' J^T W e, H += J^T (rho' W) J
//   thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:65-91                  Huber
//   thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-111               Eigen::LDLT of the dense 6x6
//   thirdparty/g2o/g2o/types/types_six_dof_expmap.{h:59-77,143-171; cpp:266-296}, types/se3quat.h (exp, operator*)
//   thirdparty/orb-slam2/src/Converter.cc:39-55                           float cv::Mat <-> SE3Quat
// Eigen's Quaterniond(Matrix3d), quaternion product, _transformVector, toRotationMatrix and LDLT (diagonal pivoting)
// are restated from their published algorithms.  PARITY UNPINNED against the real g2o/Eigen as a whole: the reference
// holds no test or golden vector for this function.  Pins: (a) the per-edge arithmetic -- SE3Quat::exp * estimate, the
// reprojection error and its 2x6 Jacobian, the Huber kernel -- is bit-identical to the g2o sources compiled in place
// (oracle/_ref, tests/test_oracle_reference_pin.py), and lm_solve() gives bit-identical poses to g2o's own
// OptimizationAlgorithmLevenberg::solve compiled from source and run on these primitives, and pgo_pose_optimization() gives
// identical inlier counts, outlier flags and pose bit patterns to the body of Optimizer::PoseOptimization compiled from
// Optimizer.cc:239-451; the quadratic form, the robust-chi2 sum and the 6x6 LDLT stay restatements (they need g2o's
// optimizer graph on the real Eigen), held by (b) recovery of the true pose on synthetic scenes and (c)
// an independent scipy least-squares fit on the same inlier set (tests/test_oracle_pose.py).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pgo.h"

namespace {

struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

void normalize_rotation(Quat& q) {  // SE3Quat::normalizeRotation
  if (q.w < 0) { q.x *= -1; q.y *= -1; q.z *= -1; q.w *= -1; }
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

Quat quat_from_matrix(const double m[3][3]) {  // Eigen quaternionbase_assign_impl<Matrix3>
  Quat q;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[2][1] - m[1][2]) * t;
    q.y = (m[0][2] - m[2][0]) * t;
    q.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[k][j] - m[j][k]) * t;
    v[j] = (m[j][i] + m[i][j]) * t;
    v[k] = (m[k][i] + m[i][k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}

Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

void quat_rotate(const Quat& q, const double v[3], double out[3]) {  // Eigen _transformVector
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

void quat_to_matrix(const Quat& q, double R[3][3]) {  // Eigen toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}

void mat3_mul(const double A[3][3], const double B[3][3], double C[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}

SE3 se3_exp(const double u[6]) {  // SE3Quat::exp: u[0..2] = omega, u[3..5] = upsilon
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
  double O2[3][3], R[3][3], V[3][3];
  mat3_mul(O, O, O2);
  if (theta < 0.00001) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) V[i][j] = R[i][j] = (i == j ? 1.0 : 0.0) + O[i][j] + O2[i][j];
  } else {
    const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
    const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + a * O[i][j] + b * O2[i][j];
        V[i][j] = (i == j ? 1.0 : 0.0) + b * O[i][j] + c * O2[i][j];
      }
  }
  SE3 s;
  s.r = quat_from_matrix(R);
  for (int i = 0; i < 3; i++) s.t[i] = V[i][0] * up[0] + V[i][1] * up[1] + V[i][2] * up[2];
  normalize_rotation(s.r);  // SE3Quat(const Quaterniond&, const Vector3d&)
  return s;
}

SE3 se3_mul(const SE3& a, const SE3& b) {  // SE3Quat::operator*
  SE3 r = a;
  double rt[3];
  quat_rotate(a.r, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] += rt[i];
  r.r = quat_mul(a.r, b.r);
  normalize_rotation(r.r);
  return r;
}

SE3 se3_from_cv(const float* T) {  // Converter::toSE3Quat, T row-major 4x4 float
  double R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T[4 * i + j];
  SE3 s;
  s.r = quat_from_matrix(R);
  normalize_rotation(s.r);
  for (int i = 0; i < 3; i++) s.t[i] = T[4 * i + 3];
  return s;
}

void se3_to_cv(const SE3& s, float* T) {  // SE3Quat::to_homogeneous_matrix + Converter::toCvMat
  double R[3][3];
  quat_to_matrix(s.r, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[i][j];
    T[4 * i + 3] = (float)s.t[i];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}

// Eigen::LDLT<MatrixXd>::compute + isPositive + solve for a 6x6 (unblocked, diagonal pivoting on the largest |a_kk|).
bool ldlt6_solve(const double H[6][6], const double b[6], double x[6]) {
  double A[6][6];
  int tr[6];
  memcpy(A, H, sizeof A);
  int sign = 0;  // ZeroSign; 1 PositiveSemiDef, -1 NegativeSemiDef, 2 Indefinite
  for (int k = 0; k < 6; k++) {
    int p = k;
    double big = std::fabs(A[k][k]);
    for (int i = k + 1; i < 6; i++)
      if (std::fabs(A[i][i]) > big) { big = std::fabs(A[i][i]); p = i; }
    tr[k] = p;
    if (p != k) {  // symmetric row/column swap on the lower triangle
      for (int j = 0; j < k; j++) std::swap(A[k][j], A[p][j]);
      for (int i = p + 1; i < 6; i++) std::swap(A[i][k], A[i][p]);
      std::swap(A[k][k], A[p][p]);
      for (int i = k + 1; i < p; i++) std::swap(A[i][k], A[p][i]);
    }
    double temp[6];
    for (int j = 0; j < k; j++) temp[j] = A[j][j] * A[k][j];
    for (int j = 0; j < k; j++) A[k][k] -= A[k][j] * temp[j];
    for (int i = k + 1; i < 6; i++)
      for (int j = 0; j < k; j++) A[i][k] -= A[i][j] * temp[j];
    const double akk = A[k][k];
    if (std::fabs(akk) > 0)
      for (int i = k + 1; i < 6; i++) A[i][k] /= akk;
    if (sign == 1) { if (akk < 0) sign = 2; }
    else if (sign == -1) { if (akk > 0) sign = 2; }
    else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
  }
  if (!(sign == 1 || sign == 0)) return false;  // isPositive()
  double y[6];
  memcpy(y, b, sizeof y);
  for (int k = 0; k < 6; k++) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i][j] * y[j];
  for (int i = 0; i < 6; i++) y[i] = std::fabs(A[i][i]) > DBL_MIN ? y[i] / A[i][i] : 0.0;
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) y[i] -= A[j][i] * y[j];
  for (int k = 5; k >= 0; k--) std::swap(y[k], y[tr[k]]);
  memcpy(x, y, sizeof y);
  return true;
}

struct Edge {
  double obs[2], Xw[3], info;  // information = Identity * invSigma2
  double err[2];               // _error of the last computeError()
  int level;                   // 0 active, 1 outlier
  bool robust;
  int idx;                     // feature index
};

struct Problem {
  std::vector<Edge> edges;
  double fx, fy, cx, cy, delta, dsqr;
  SE3 est;
  // LM state (OptimizationAlgorithmLevenberg members)
  double lambda = -1, ni = 2;
  int nBad = 0;
  double x[6] = {0, 0, 0, 0, 0, 0};

  void compute_error(Edge& e) const {  // EdgeSE3ProjectXYZOnlyPose::computeError
    double p[3], r[3];
    quat_rotate(est.r, e.Xw, r);
    for (int i = 0; i < 3; i++) p[i] = r[i] + est.t[i];  // SE3Quat::map
    const double px = p[0] / p[2], py = p[1] / p[2];     // project2d
    e.err[0] = e.obs[0] - (px * fx + cx);
    e.err[1] = e.obs[1] - (py * fy + cy);
  }
  static double chi2(const Edge& e) { return e.err[0] * (e.info * e.err[0]) + e.err[1] * (e.info * e.err[1]); }
  void robustify(double e, double rho[3]) const {
    if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else { const double s = std::sqrt(e); rho[0] = 2 * s * delta - dsqr; rho[1] = delta / s; rho[2] = -0.5 * rho[1] / e; }
  }
  void compute_active_errors() {
    for (Edge& e : edges) if (e.level == 0) compute_error(e);
  }
  double active_robust_chi2() const {
    double chi = 0.0, rho[3];
    for (const Edge& e : edges) {
      if (e.level != 0) continue;
      if (e.robust) { robustify(chi2(e), rho); chi += rho[0]; } else chi += chi2(e);
    }
    return chi;
  }
  // EdgeSE3ProjectXYZOnlyPose::linearizeOplus (types_six_dof_expmap.cpp:266-288) at the camera-frame point p
  void jacobian(const double p[3], double J[2][6]) const {
    const double x = p[0], y = p[1], invz = 1.0 / p[2], invz_2 = invz * invz;
    J[0][0] = x * y * invz_2 * fx; J[0][1] = -(1 + (x * x * invz_2)) * fx; J[0][2] = y * invz * fx;
    J[0][3] = -invz * fx; J[0][4] = 0; J[0][5] = x * invz_2 * fx;
    J[1][0] = (1 + y * y * invz_2) * fy; J[1][1] = -x * y * invz_2 * fy; J[1][2] = -x * invz * fy;
    J[1][3] = 0; J[1][4] = -invz * fy; J[1][5] = y * invz_2 * fy;
  }
  void build_system(double H[6][6], double b[6]) const {
    memset(H, 0, 36 * sizeof(double));
    memset(b, 0, 6 * sizeof(double));
    for (const Edge& e : edges) {
      if (e.level != 0) continue;
      double p[3], r[3];
      quat_rotate(est.r, e.Xw, r);
      for (int i = 0; i < 3; i++) p[i] = r[i] + est.t[i];
      double J[2][6];
      jacobian(p, J);
      double w = 1.0;
      if (e.robust) { double rho[3]; robustify(chi2(e), rho); w = rho[1]; }
      for (int i = 0; i < 6; i++) {
        b[i] -= w * (J[0][i] * (e.info * e.err[0]) + J[1][i] * (e.info * e.err[1]));
        for (int j = 0; j < 6; j++) H[i][j] += J[0][i] * (w * e.info) * J[0][j] + J[1][i] * (w * e.info) * J[1][j];
      }
    }
  }
  // OptimizationAlgorithmLevenberg::solve; returns true for OK, false for Terminate
  bool lm_solve(int iteration) {
    compute_active_errors();
    double currentChi = active_robust_chi2(), tempChi = currentChi;
    const double iniChi = currentChi;
    double H[6][6], b[6];
    build_system(H, b);
    if (iteration == 0) {
      double maxDiagonal = 0.;
      for (int j = 0; j < 6; j++) maxDiagonal = std::max(std::fabs(H[j][j]), maxDiagonal);
      lambda = 1e-5 * maxDiagonal;
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      const SE3 backup = est;
      double Hl[6][6];
      memcpy(Hl, H, sizeof Hl);
      for (int j = 0; j < 6; j++) Hl[j][j] += lambda;
      const bool ok2 = ldlt6_solve(Hl, b, x);
      est = se3_mul(se3_exp(x), est);  // VertexSE3Expmap::oplusImpl
      compute_active_errors();
      tempChi = active_robust_chi2();
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0.;
      for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        const double scaleFactor = std::max(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        est = backup;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) return false;
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
    if (nBad >= 3) return false;
    return true;
  }
  void optimize(int iterations) {
    bool any = false;
    for (const Edge& e : edges) any |= e.level == 0;
    if (!any) return;  // "0 vertices to optimize"
    for (int i = 0; i < iterations; i++)
      if (!lm_solve(i)) break;
  }
};

}  // namespace

extern "C" {

// One frame.  kp_xy[n][2] = mvKeysUn[i].pt, kp_octave[n], mp_xyz[n][3] = GetWorldPos() (float), has_map_point[n].
// Tcw_in / Tcw_out: 4x4 row-major float (cv::Mat CV_32F).  outlier[n] = mvbOutlier (0 where there is no map point).
// round_outliers (optional, [4][n]): mvbOutlier after each of the 4 rounds (rows of rounds not run are left alone).
// Returns nInitialCorrespondences - nBad.
int pgo_pose_optimization(const float* Tcw_in, const float* kp_xy, const int32_t* kp_octave, const float* mp_xyz,
                          const uint8_t* has_map_point, int n, const float* inv_level_sigma2, float fx, float fy, float cx,
                          float cy, float* Tcw_out, uint8_t* outlier, uint8_t* round_outliers) {
  Problem P;
  P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
  const float deltaMono = (float)std::sqrt(5.991);
  P.delta = deltaMono;
  P.dsqr = P.delta * P.delta;
  int nInitialCorrespondences = 0;
  for (int i = 0; i < n; i++) {
    outlier[i] = 0;
    if (!has_map_point[i]) continue;
    nInitialCorrespondences++;
    Edge e;
    e.obs[0] = kp_xy[2 * i]; e.obs[1] = kp_xy[2 * i + 1];
    e.info = inv_level_sigma2[kp_octave[i]];
    for (int k = 0; k < 3; k++) e.Xw[k] = mp_xyz[3 * i + k];
    e.err[0] = e.err[1] = 0;
    e.level = 0; e.robust = true; e.idx = i;
    P.edges.push_back(e);
  }
  memcpy(Tcw_out, Tcw_in, 16 * sizeof(float));
  if (nInitialCorrespondences < 3) return 0;
  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f};
  const int its[4] = {10, 10, 10, 10};
  int nBad = 0;
  for (size_t it = 0; it < 4; it++) {
    P.est = se3_from_cv(Tcw_in);
    P.optimize(its[it]);
    nBad = 0;
    for (Edge& e : P.edges) {
      if (outlier[e.idx]) P.compute_error(e);
      const float chi2 = (float)Problem::chi2(e);
      if (chi2 > chi2Mono[it]) { outlier[e.idx] = 1; e.level = 1; nBad++; }
      else { outlier[e.idx] = 0; e.level = 0; }
      if (it == 2) e.robust = false;
    }
    if (round_outliers) memcpy(round_outliers + it * n, outlier, n);
    if (P.edges.size() < 10) break;
  }
  se3_to_cv(P.est, Tcw_out);
  return nInitialCorrespondences - nBad;
}

// ---- the per-edge pieces, exposed so that tests/test_oracle_reference_pin.py can hold them to the g2o sources compiled in
// place (oracle/_ref).  pose7 = quaternion w, x, y, z + translation.
void pgo_pose_se3_oplus(const double* update6, const double* pose7, double* out7) {   // VertexSE3Expmap::oplusImpl
  SE3 est;
  est.r = Quat{pose7[1], pose7[2], pose7[3], pose7[0]};
  normalize_rotation(est.r);                                                            // SE3Quat(q, t) normalises
  for (int i = 0; i < 3; i++) est.t[i] = pose7[4 + i];
  const SE3 r = se3_mul(se3_exp(update6), est);
  out7[0] = r.r.w; out7[1] = r.r.x; out7[2] = r.r.y; out7[3] = r.r.z;
  for (int i = 0; i < 3; i++) out7[4 + i] = r.t[i];
}

void pgo_pose_edge(const double* pose7, const double* Xw, const double* obs, double fx, double fy, double cx, double cy, double* err2,
                   double* J12) {
  Problem P;
  P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy; P.delta = 1; P.dsqr = 1;
  P.est.r = Quat{pose7[1], pose7[2], pose7[3], pose7[0]};
  normalize_rotation(P.est.r);
  for (int i = 0; i < 3; i++) P.est.t[i] = pose7[4 + i];
  Edge e;
  e.obs[0] = obs[0]; e.obs[1] = obs[1]; e.info = 1; e.level = 0; e.robust = false; e.idx = 0;
  for (int i = 0; i < 3; i++) e.Xw[i] = Xw[i];
  P.compute_error(e);
  err2[0] = e.err[0]; err2[1] = e.err[1];
  double r[3], p[3], J[2][6];
  quat_rotate(P.est.r, e.Xw, r);
  for (int i = 0; i < 3; i++) p[i] = r[i] + P.est.t[i];
  P.jacobian(p, J);
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 6; j++) J12[6 * i + j] = J[i][j];
}

void pgo_pose_huber(double delta, double e, double* rho3) {
  Problem P;
  P.delta = delta; P.dsqr = delta * delta;
  P.robustify(e, rho3);
}

// ---- a Problem as a handle, one primitive per call: lets the g2o Levenberg-Marquardt driver compiled from source
// (oracle/_ref, ref_wrap_g2o_lm.cc) run on exactly the arithmetic lm_solve() uses, so that a difference between the two
// can only come from the driver's control flow.
void* pgo_pose_problem_create(const float* Tcw, const float* kp_xy, const int32_t* kp_octave, const float* mp_xyz,
                              const uint8_t* has_map_point, int n, const float* inv_level_sigma2, float fx, float fy, float cx,
                              float cy, int robust) {
  Problem* P = new Problem;
  P->fx = fx; P->fy = fy; P->cx = cx; P->cy = cy;
  P->delta = (float)std::sqrt(5.991);
  P->dsqr = P->delta * P->delta;
  for (int i = 0; i < n; i++) {
    if (!has_map_point[i]) continue;
    Edge e;
    e.obs[0] = kp_xy[2 * i]; e.obs[1] = kp_xy[2 * i + 1];
    e.info = inv_level_sigma2[kp_octave[i]];
    for (int k = 0; k < 3; k++) e.Xw[k] = mp_xyz[3 * i + k];
    e.err[0] = e.err[1] = 0;
    e.level = 0; e.robust = robust != 0; e.idx = i;
    P->edges.push_back(e);
  }
  P->est = se3_from_cv(Tcw);
  return P;
}
void pgo_pose_problem_destroy(void* h) { delete static_cast<Problem*>(h); }
void pgo_pose_problem_reset(void* h, const float* Tcw) {
  Problem* P = static_cast<Problem*>(h);
  P->est = se3_from_cv(Tcw);
  P->lambda = -1; P->ni = 2; P->nBad = 0;
  for (int i = 0; i < 6; i++) P->x[i] = 0;
}
void pgo_pose_problem_get_estimate(void* h, double* pose7) {
  const Problem* P = static_cast<Problem*>(h);
  pose7[0] = P->est.r.w; pose7[1] = P->est.r.x; pose7[2] = P->est.r.y; pose7[3] = P->est.r.z;
  for (int i = 0; i < 3; i++) pose7[4 + i] = P->est.t[i];
}
void pgo_pose_problem_set_estimate(void* h, const double* pose7) {
  Problem* P = static_cast<Problem*>(h);
  P->est.r = Quat{pose7[1], pose7[2], pose7[3], pose7[0]};
  for (int i = 0; i < 3; i++) P->est.t[i] = pose7[4 + i];
}
void pgo_pose_problem_compute_active_errors(void* h) { static_cast<Problem*>(h)->compute_active_errors(); }
double pgo_pose_problem_active_robust_chi2(void* h) { return static_cast<Problem*>(h)->active_robust_chi2(); }
void pgo_pose_problem_build_system(void* h, double* H36, double* b6) {
  double H[6][6];
  static_cast<Problem*>(h)->build_system(H, b6);
  memcpy(H36, H, sizeof H);
}
int pgo_pose_ldlt6_solve(const double* H36, const double* b6, double* x6) {
  double H[6][6];
  memcpy(H, H36, sizeof H);
  return ldlt6_solve(H, b6, x6) ? 1 : 0;
}
void pgo_pose_problem_oplus(void* h, const double* x6) {   // VertexSE3Expmap::oplusImpl on the current estimate
  Problem* P = static_cast<Problem*>(h);
  P->est = se3_mul(se3_exp(x6), P->est);
}
void pgo_pose_problem_optimize(void* h, int iterations) { static_cast<Problem*>(h)->optimize(iterations); }   // the restated driver

// ---- more handle primitives, for the outer four-round loop of Optimizer.cc:239-451 compiled from source (ref_wrap_poseopt.cc)
void* pgo_pose_problem_create_raw(int n, const double* obs2, const double* Xw3, const double* info, double fx, double fy, double cx,
                                  double cy, double delta) {
  Problem* P = new Problem;
  P->fx = fx; P->fy = fy; P->cx = cx; P->cy = cy; P->delta = delta; P->dsqr = delta * delta;
  for (int i = 0; i < n; i++) {
    Edge e;
    e.obs[0] = obs2[2 * i]; e.obs[1] = obs2[2 * i + 1];
    for (int k = 0; k < 3; k++) e.Xw[k] = Xw3[3 * i + k];
    e.info = info[i];
    e.err[0] = e.err[1] = 0; e.level = 0; e.robust = true; e.idx = i;
    P->edges.push_back(e);
  }
  P->est.r = Quat{0, 0, 0, 1};
  P->est.t[0] = P->est.t[1] = P->est.t[2] = 0;
  return P;
}
void pgo_pose_problem_edge_set_level(void* h, int k, int level) { static_cast<Problem*>(h)->edges[k].level = level; }
void pgo_pose_problem_edge_set_robust(void* h, int k, int robust) { static_cast<Problem*>(h)->edges[k].robust = robust != 0; }
void pgo_pose_problem_edge_compute_error(void* h, int k) { Problem* P = static_cast<Problem*>(h); P->compute_error(P->edges[k]); }
double pgo_pose_problem_edge_chi2(void* h, int k) { return Problem::chi2(static_cast<Problem*>(h)->edges[k]); }
int pgo_pose_problem_num_active(void* h) {
  int n = 0;
  for (const Edge& e : static_cast<Problem*>(h)->edges) n += e.level == 0;
  return n;
}
void pgo_pose_T_to_pose7(const float* T, double* pose7) {   // Converter::toSE3Quat
  const SE3 s = se3_from_cv(T);
  pose7[0] = s.r.w; pose7[1] = s.r.x; pose7[2] = s.r.y; pose7[3] = s.r.z;
  for (int i = 0; i < 3; i++) pose7[4 + i] = s.t[i];
}
void pgo_pose_pose7_to_T(const double* pose7, float* T) {   // Converter::toCvMat(SE3Quat)
  SE3 s;
  s.r = Quat{pose7[1], pose7[2], pose7[3], pose7[0]};
  for (int i = 0; i < 3; i++) s.t[i] = pose7[4 + i];
  se3_to_cv(s, T);
}

}  // extern "C"
