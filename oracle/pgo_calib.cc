// ORACLE -- TEST INFRASTRUCTURE ONLY (see pgo_orb.cc header for the rules).
// CPU restatement of the reference's IMU+GPS calibration path, literal and sequential:
//   src/interpolation/align_time_series.cc  MergeTimeSeries :29-113, GetEffectiveTimeStamp :115-128,
//                                           MakeInterpolationIntervals :155-196
//   src/geometry/geometry.cc                RotationMotionToQuaternion :6-22, IntegrateMotion :24-53
//   src/calibration/velocity.cc             ctor :29-39, eval :41-180, IntegrateTrajectory :199-256
//   thirdparty/LBFGS/LBFGS.h :79-182, LBFGS/LineSearch.h :41-111, LBFGS/Param.h :162-175
//   src/slam/smoothing.cc                   NormalCdf :49-53, SmoothTimeSeries :56-98
//   src/fit_motion.cc                       ComputeAndSaveForwardVelocitiesFromImu :156-293
//   src/calibration/rotation.cc             GetPrincipalRotationAxes :16-57, GetAngularVelocitiesAroundAxisDirect :103-119
// Eigen (un-vendored) is restated from its standard formulas; libm sin/cos/sqrt/erf are used
// as the reference does.  PINNED against the reference's OWN SOURCES where they compile: align_time_series.cc,
// geometry.cc, velocity.cc:1-256, smoothing.cc:48-end and the vendored LBFGS++ are built where they lie into
// oracle/_ref (make _ref; Eigen / glog stand-ins in ref_shims/) and tests/test_oracle_reference_pin.py runs this file
// against them: indices and interval tables equal, loss equal to the bit, gradient within 1 ulp, SmoothTimeSeries
// bit-exact, L-BFGS iterates equal to 1e-14 through 20 iterations.  PARITY UNPINNED against a reference BINARY (no
// Eigen in this image, no fixtures): the real Eigen's association of reductions is version dependent.  Further pins:
// the doc-comment example of align_time_series.hpp:17-26, an independent numpy restatement in
// tests/test_oracle_calib.py and finite-difference / invariance properties.
//
// The "_core" entry points evaluate the SAME quantities through include/pgb200_imu_core.h (the arithmetic
// contract the CUDA kernels are compiled from) on the host, so GPU results can be compared bit for bit; the
// literal/core agreement (<= 1e-10 relative per evaluation) is itself a test.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "../include/pgb200_imu_core.h"
#include "pgo.h"

namespace {

struct Vec3 { double x, y, z; };
struct Quat { double w, x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(Vec3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline Quat qmul(Quat a, Quat b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Vec3 transform_vector(Quat q, Vec3 v) {  // Eigen::QuaternionBase::_transformVector
  Vec3 qv{q.x, q.y, q.z};
  Vec3 uv = cross(qv, v);
  uv = uv + uv;
  Vec3 c = cross(qv, uv);
  return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
struct Mat3 { double m[9]; };
inline Mat3 to_rotation_matrix(Quat q) {  // Eigen::QuaternionBase::toRotationMatrix
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r.m[0] = 1.0 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1.0 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1.0 - (txx + tyy);
  return r;
}

// ------------------------------------------------------------------ align_time_series.cc
typedef std::vector<int64_t> Times;

std::vector<std::vector<size_t>> merge_time_series(const std::vector<const Times*>& in) {
  std::vector<std::vector<size_t>> result;
  for (const Times* c : in) {
    if (!c || c->empty()) throw std::invalid_argument("empty component");
    for (size_t i = 0; i + 1 < c->size(); ++i)
      if (!((*c)[i] < (*c)[i + 1])) throw std::invalid_argument("timestamps not increasing");
  }
  std::vector<int64_t> start_times, end_times;
  for (const Times* c : in) { start_times.push_back(c->front()); end_times.push_back(c->back()); }
  const int64_t start_time = *std::max_element(start_times.begin(), start_times.end());
  const int64_t end_time = *std::min_element(end_times.begin(), end_times.end());
  if (end_time < start_time) return {};
  std::vector<size_t> cur;
  for (const Times* c : in) {
    const size_t idx = std::lower_bound(c->begin(), c->end(), start_time) - c->begin();
    if ((*c)[idx] > start_time) cur.push_back(idx - 1); else cur.push_back(idx);
  }
  while (true) {
    result.push_back(cur);
    std::vector<int64_t> next_times;
    for (size_t i = 0; i < cur.size(); ++i) {
      const size_t next = cur[i] + 1;
      if (next >= in[i]->size()) return result;
      next_times.push_back((*in[i])[next]);
    }
    const int64_t next_time = *std::min_element(next_times.begin(), next_times.end());
    for (size_t i = 0; i < cur.size(); ++i)
      if ((*in[i])[cur[i] + 1] == next_time) cur[i] = cur[i] + 1;
  }
}

int64_t effective_time(const std::vector<const Times*>& comp, const std::vector<size_t>& ev) {
  int64_t r = std::numeric_limits<int64_t>::min();
  for (size_t i = 0; i < comp.size(); ++i) r = std::max(r, (*comp[i])[ev[i]]);
  return r;
}

struct Interval { size_t ref_idx, interp_idx; int64_t start, end; };

std::vector<std::vector<Interval>> make_intervals(const Times& ref, const Times& interp) {
  std::vector<std::vector<Interval>> result;
  int64_t latest = std::min(interp.front(), ref.front());
  for (size_t r = 0, k = 0; r < ref.size(); ++r) {
    const int64_t ref_ts = ref[r];
    std::vector<Interval> iv;
    while (k < interp.size() && interp[k] <= ref_ts) {
      const int64_t ts = interp[k];
      if (ts > latest && k > 0 && r > 0) iv.push_back({r, k, latest, ts});
      latest = ts;
      ++k;
    }
    if (k > 0 && r > 0 && k < interp.size() && ref_ts > latest) iv.push_back({r, k, latest, ref_ts});
    latest = ref_ts;
    result.push_back(iv);
  }
  return result;
}

// ------------------------------------------------------------------ geometry.cc
Quat rotation_motion_to_quaternion(double rx, double ry, double rz, double duration_sec) {
  const double rate = std::sqrt(rx * rx + ry * ry + rz * rz);
  const double half_theta = rate * duration_sec * 0.5;
  const double k = std::sin(half_theta) / (rate + 1e-30);
  return {std::cos(half_theta), rx * k, ry * k, rz * k};
}

struct Outcome { Quat orientation; Vec3 velocity; int64_t duration_usec; };

Outcome integrate_motion(Quat start_q, Vec3 start_v, Quat raw_rot, Vec3 raw_acc, Vec3 g, Vec3 h, int64_t dur_usec) {
  const double dt = (double)dur_usec * 1e-6;
  const Vec3 local = raw_acc + h;
  const Vec3 rotated = transform_vector(start_q, local);
  const Vec3 global = rotated + g;
  const Vec3 v = start_v + global * dt;
  return {qmul(start_q, raw_rot), v, dur_usec};
}

}  // namespace

// ------------------------------------------------------------------ calibrator
struct pgo_calib {
  std::vector<double> gps_v; Times gps_t;
  std::vector<double> gyro, acc; Times gyro_t, acc_t;
  std::vector<std::vector<size_t>> merged;
  Times merged_t;
  std::vector<std::vector<Interval>> intervals;
  // core-path products
  std::vector<pgbimu::WinRec> rec;
  int64_t total_usec = 0;
  bool core_ready = false;
};

namespace {

void build_core(pgo_calib* c) {
  using namespace pgbimu;
  c->rec.clear();
  WinState ws;
  win_init(&ws);
  c->total_usec = 0;
  for (size_t r = 1; r < c->intervals.size(); ++r) {
    GpsLocal gl;
    SweepState ss;
    sweep_init(&ss, &gl);
    for (const Interval& iv : c->intervals[r]) {
      const auto& ev = c->merged.at(iv.interp_idx);
      ImuStep st;
      st.wx = c->gyro[3 * ev[0]]; st.wy = c->gyro[3 * ev[0] + 1]; st.wz = c->gyro[3 * ev[0] + 2];
      st.ax = c->acc[3 * ev[1]]; st.ay = c->acc[3 * ev[1] + 1]; st.az = c->acc[3 * ev[1] + 2];
      st.dur_usec = iv.end - iv.start;
      const double dt = sweep_step(&ss, st);
      sweep_accumulate(ss, dt, &gl);
    }
    sweep_finish(ss, &gl);
    WinRec wr;
    win_chain(&ws, gl, c->gps_v[r], &wr);
    c->rec.push_back(wr);
    c->total_usec += gl.dur;
  }
  c->core_ready = true;
}

double eval_literal(const pgo_calib* c, const double* in, double* gradient) {
  for (int i = 0; i < 9; i++) gradient[i] = 0.0;
  const Vec3 g{in[0], in[1], in[2]}, h{in[3], in[4], in[5]}, v0{in[6], in[7], in[8]};
  double result = 0;
  Quat q{1.0, 0.0, 0.0, 0.0};
  Vec3 v = v0;
  double W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int64_t total_time_usec = 0;
  std::vector<Outcome> outcomes;
  for (const std::vector<Interval>& intervals : c->intervals) {
    Vec3 travel{0, 0, 0};
    double reference_distance = 0;
    outcomes.clear();
    for (const Interval& iv : intervals) {
      const auto& ev = c->merged.at(iv.interp_idx);
      const double* w = &c->gyro[3 * ev.at(0)];
      const double* a = &c->acc[3 * ev.at(1)];
      const double dsec = (double)(iv.end - iv.start) * 1e-6;
      const Quat raw = rotation_motion_to_quaternion(w[0], w[1], w[2], dsec);
      const Outcome o = integrate_motion(q, v, raw, Vec3{a[0], a[1], a[2]}, g, h, iv.end - iv.start);
      outcomes.push_back(o);
      q = o.orientation;
      v = o.velocity;
      travel = travel + o.velocity * dsec;
      reference_distance += dsec * c->gps_v.at(iv.ref_idx);
    }
    const double tn = norm(travel);
    const double diff = tn - reference_distance;
    result += diff * diff;
    const double k = 2.0 * diff;
    const Vec3 dL{k * travel.x / (tn + 1e-5), k * travel.y / (tn + 1e-5), k * travel.z / (tn + 1e-5)};
    for (const Outcome& o : outcomes) {
      const double isec = (double)o.duration_usec * 1e-6;
      total_time_usec += o.duration_usec;
      const double tsec = (double)total_time_usec * 1e-6;
      gradient[0] += tsec * isec * dL.x;
      gradient[1] += tsec * isec * dL.y;
      gradient[2] += tsec * isec * dL.z;
      const Mat3 R = to_rotation_matrix(o.orientation);
      for (int i = 0; i < 9; i++) W[i] += R.m[i] * isec;
      // interval_sec * W^T * dL  (Eigen evaluates (isec * W^T) * dL; scalar placement differs by <= 1 ulp)
      gradient[3] += isec * (W[0] * dL.x + W[3] * dL.y + W[6] * dL.z);
      gradient[4] += isec * (W[1] * dL.x + W[4] * dL.y + W[7] * dL.z);
      gradient[5] += isec * (W[2] * dL.x + W[5] * dL.y + W[8] * dL.z);
      gradient[6] += isec * dL.x;
      gradient[7] += isec * dL.y;
      gradient[8] += isec * dL.z;
    }
  }
  const double total = (double)total_time_usec * 1e-6;
  result /= total;
  for (int i = 0; i < 9; i++) gradient[i] /= total;
  return result;
}

// LBFGSSolver<double>::minimize with LBFGSParam defaults except epsilon / max_iterations (fit_motion.cc:167-169).
// Written against LBFGS.h / LineSearch.h independently of the contract header's driver.
template <typename F>
int lbfgs_literal(F f, std::vector<double>& x, double& fx, double epsilon, int max_iterations, int* n_eval) {
  const int n = (int)x.size(), m = 6, max_linesearch = 20;
  const double ftol = 1e-4, min_step = 1e-20, max_step = 1e20;
  std::vector<std::vector<double>> s(m, std::vector<double>(n)), y(m, std::vector<double>(n));
  std::vector<double> ys(m), alpha(m), xp(n), grad(n), gradp(n), drt(n);
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) {
    double r = 0;
    for (int i = 0; i < n; i++) r += a[i] * b[i];
    return r;
  };
  int evals = 0;
  fx = f(x, grad); evals++;
  double xnorm = std::sqrt(dot(x, x)), gnorm = std::sqrt(dot(grad, grad));
  if (gnorm <= epsilon * std::max(xnorm, 1.0)) { if (n_eval) *n_eval = evals; return 1; }
  for (int i = 0; i < n; i++) drt[i] = -grad[i];
  double step = 1.0 / std::sqrt(dot(drt, drt));
  int k = 1, end = 0;
  for (;;) {
    xp = x; gradp = grad;
    {  // LineSearch::Backtracking (Armijo)
      const double dec = 0.5;
      const double fx_init = fx, dg_init = dot(grad, drt), dg_test = ftol * dg_init;
      for (int iter = 0; iter < max_linesearch; iter++) {
        for (int i = 0; i < n; i++) x[i] = xp[i] + step * drt[i];
        fx = f(x, grad); evals++;
        double width;
        if (fx > fx_init + step * dg_test) width = dec; else break;
        if (step < min_step) throw std::runtime_error("the line search step became smaller than the minimum value allowed");
        if (step > max_step) throw std::runtime_error("the line search step became larger than the maximum value allowed");
        step *= width;
      }
    }
    xnorm = std::sqrt(dot(x, x)); gnorm = std::sqrt(dot(grad, grad));
    if (gnorm <= epsilon * std::max(xnorm, 1.0)) break;
    if (max_iterations != 0 && k >= max_iterations) break;
    for (int i = 0; i < n; i++) { s[end][i] = x[i] - xp[i]; y[end][i] = grad[i] - gradp[i]; }
    const double ysv = dot(y[end], s[end]), yy = dot(y[end], y[end]);
    ys[end] = ysv;
    for (int i = 0; i < n; i++) drt[i] = -grad[i];
    const int bound = std::min(m, k);
    end = (end + 1) % m;
    int j = end;
    for (int i = 0; i < bound; i++) {
      j = (j + m - 1) % m;
      alpha[j] = dot(s[j], drt) / ys[j];
      for (int q = 0; q < n; q++) drt[q] -= alpha[j] * y[j][q];
    }
    for (int q = 0; q < n; q++) drt[q] *= (ysv / yy);
    for (int i = 0; i < bound; i++) {
      const double beta = dot(y[j], drt) / ys[j];
      for (int q = 0; q < n; q++) drt[q] += (alpha[j] - beta) * s[j][q];
      j = (j + 1) % m;
    }
    step = 1.0;
    k++;
  }
  if (n_eval) *n_eval = evals;
  return k;
}

struct TrajPoint { Quat q; Vec3 v; int64_t dur; };

std::map<size_t, TrajPoint> integrate_literal(const pgo_calib* c, const double* x) {
  std::map<size_t, TrajPoint> result;
  const Vec3 g{x[0], x[1], x[2]}, h{x[3], x[4], x[5]};
  Quat q{1.0, 0.0, 0.0, 0.0};
  Vec3 v{x[6], x[7], x[8]};
  for (const auto& intervals : c->intervals)
    for (const Interval& iv : intervals) {
      const auto& ev = c->merged.at(iv.interp_idx);
      const double* w = &c->gyro[3 * ev.at(0)];
      const double* a = &c->acc[3 * ev.at(1)];
      const Quat raw = rotation_motion_to_quaternion(w[0], w[1], w[2], (double)(iv.end - iv.start) * 1e-6);
      const Outcome o = integrate_motion(q, v, raw, Vec3{a[0], a[1], a[2]}, g, h, iv.end - iv.start);
      q = o.orientation; v = o.velocity;
      auto it = result.find(iv.interp_idx);
      if (it == result.end()) result.insert({iv.interp_idx, TrajPoint{o.orientation, o.velocity, o.duration_usec}});
      else { it->second.q = o.orientation; it->second.v = o.velocity; it->second.dur += o.duration_usec; }
    }
  return result;
}

// Same trajectory through the contract header (what the CUDA K10 kernel computes).
std::map<size_t, TrajPoint> integrate_core(pgo_calib* c, const double* x) {
  using namespace pgbimu;
  if (!c->core_ready) build_core(c);
  std::map<size_t, TrajPoint> result;
  const V3 g = v3(x[0], x[1], x[2]), h = v3(x[3], x[4], x[5]), v0 = v3(x[6], x[7], x[8]);
  for (size_t r = 1; r < c->intervals.size(); ++r) {
    const WinRec& wr = c->rec[r - 1];
    const V3 Vr = add(add(v0, wr.Sa), add(mv(wr.SE, h), scale(g, wr.St)));
    // orientation at the start of the GPS interval: chain of the previous intervals' products
    GpsLocal gl;
    SweepState ss;
    sweep_init(&ss, &gl);
    for (const Interval& iv : c->intervals[r]) {
      const auto& ev = c->merged.at(iv.interp_idx);
      ImuStep st;
      st.wx = c->gyro[3 * ev[0]]; st.wy = c->gyro[3 * ev[0] + 1]; st.wz = c->gyro[3 * ev[0] + 2];
      st.ax = c->acc[3 * ev[1]]; st.ay = c->acc[3 * ev[1] + 1]; st.az = c->acc[3 * ev[1] + 2];
      st.dur_usec = iv.end - iv.start;
      sweep_step(&ss, st);
      const V3 loc = add(ss.pa, mv(ss.pR, h));
      const V3 v = add(add(Vr, mv(wr.RQ, loc)), scale(g, (double)ss.tau * 1e-6));
      TrajPoint tp;
      const Q4 qq = pgbimu::qmul(wr.Q, ss.l);  // orientation after the step, as k_imu_speeds forms it
      tp.q = {qq.w, qq.x, qq.y, qq.z};
      tp.v = {v.x, v.y, v.z};
      tp.dur = st.dur_usec;
      auto it = result.find(iv.interp_idx);
      if (it == result.end()) result.insert({iv.interp_idx, tp});
      else { it->second.q = tp.q; it->second.v = tp.v; it->second.dur += tp.dur; }
    }
  }
  return result;
}

double normal_cdf(double x, double mean, double sigma) {
  static const double sqrt_2 = std::sqrt(2.0);
  return 0.5 * (1.0 + std::erf((x - mean) / (sqrt_2 * sigma)));
}

}  // namespace

extern "C" {

pgo_calib* pgo_calib_create(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz,
                            const int64_t* gyro_t, int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t,
                            int64_t n_acc) {
  try {
    pgo_calib* c = new pgo_calib;
    c->gps_v.assign(gps_v, gps_v + n_gps); c->gps_t.assign(gps_t, gps_t + n_gps);
    c->gyro.assign(gyro_xyz, gyro_xyz + 3 * n_gyro); c->gyro_t.assign(gyro_t, gyro_t + n_gyro);
    c->acc.assign(acc_xyz, acc_xyz + 3 * n_acc); c->acc_t.assign(acc_t, acc_t + n_acc);
    std::vector<const Times*> comp{&c->gyro_t, &c->acc_t};
    c->merged = merge_time_series(comp);
    for (const auto& ev : c->merged) c->merged_t.push_back(effective_time(comp, ev));
    if (c->merged.empty() || c->gps_t.empty()) { delete c; return nullptr; }
    for (size_t i = 0; i + 1 < c->gps_t.size(); ++i)
      if (!(c->gps_t[i] < c->gps_t[i + 1])) { delete c; return nullptr; }
    c->intervals = make_intervals(c->gps_t, c->merged_t);
    return c;
  } catch (const std::exception&) {
    return nullptr;
  }
}
void pgo_calib_destroy(pgo_calib* c) { delete c; }

int64_t pgo_calib_merged_count(const pgo_calib* c) { return (int64_t)c->merged.size(); }
void pgo_calib_merged_events(const pgo_calib* c, int64_t* t, int64_t* gi, int64_t* ai) {
  for (size_t i = 0; i < c->merged.size(); i++) {
    if (t) t[i] = c->merged_t[i];
    if (gi) gi[i] = (int64_t)c->merged[i][0];
    if (ai) ai[i] = (int64_t)c->merged[i][1];
  }
}
int64_t pgo_calib_num_intervals(const pgo_calib* c) {
  int64_t n = 0;
  for (const auto& v : c->intervals) n += (int64_t)v.size();
  return n;
}
void pgo_calib_intervals(const pgo_calib* c, int64_t* ref_idx, int64_t* merged_idx, int64_t* start, int64_t* end) {
  size_t k = 0;
  for (const auto& v : c->intervals)
    for (const Interval& iv : v) {
      ref_idx[k] = (int64_t)iv.ref_idx; merged_idx[k] = (int64_t)iv.interp_idx; start[k] = iv.start; end[k] = iv.end;
      k++;
    }
}

double pgo_calib_eval(const pgo_calib* c, const double* x, double* grad) { return eval_literal(c, x, grad); }
double pgo_calib_eval_core(pgo_calib* c, const double* x, double* grad) {
  if (!c->core_ready) build_core(c);
  return pgbimu::imu_eval(c->rec.data(), (int)c->rec.size(), c->total_usec, x, grad);
}

int pgo_calib_minimize(pgo_calib* c, double* x, double* fx, int max_iterations, double epsilon, int* n_eval) {
  std::vector<double> xv(x, x + 9);
  int it;
  try {
    it = lbfgs_literal([&](const std::vector<double>& p, std::vector<double>& g) { return eval_literal(c, p.data(), g.data()); },
                       xv, *fx, epsilon, max_iterations, n_eval);
  } catch (const std::exception&) {
    return -4;
  }
  for (int i = 0; i < 9; i++) x[i] = xv[i];
  return it;
}

// literal L-BFGS driver over the core evaluation: isolates driver differences from evaluation differences
int pgo_calib_minimize_literal_driver_core_eval(pgo_calib* c, double* x, double* fx, int max_iterations, double epsilon,
                                                int* n_eval) {
  if (!c->core_ready) build_core(c);
  std::vector<double> xv(x, x + 9);
  int it;
  try {
    it = lbfgs_literal([&](const std::vector<double>& p, std::vector<double>& g) {
      return pgbimu::imu_eval(c->rec.data(), (int)c->rec.size(), c->total_usec, p.data(), g.data()); },
                       xv, *fx, epsilon, max_iterations, n_eval);
  } catch (const std::exception&) {
    return -4;
  }
  for (int i = 0; i < 9; i++) x[i] = xv[i];
  return it;
}

int pgo_calib_minimize_core(pgo_calib* c, double* x, double* fx, int max_iterations, double epsilon, int* n_eval) {
  if (!c->core_ready) build_core(c);
  struct F {
    pgo_calib* c;
    double operator()(const double* p, double* g) {
      return pgbimu::imu_eval(c->rec.data(), (int)c->rec.size(), c->total_usec, p, g);
    }
  } f{c};
  pgbimu::LbfgsParam P = pgbimu::lbfgs_default();
  P.epsilon = epsilon; P.max_iterations = max_iterations;
  double ws[pgbimu::PGB_LBFGS_WS_DOUBLES];
  return pgbimu::lbfgs_minimize9(f, x, fx, P, n_eval, ws);
}

static int64_t dump_traj(const std::map<size_t, TrajPoint>& tr, int64_t cap, int64_t* idx, double* speed, double* quat,
                         double* vel, int64_t* dur) {
  if ((int64_t)tr.size() > cap) return -1;
  int64_t k = 0;
  for (const auto& p : tr) {
    idx[k] = (int64_t)p.first;
    speed[k] = norm(p.second.v);
    if (quat) { quat[4 * k] = p.second.q.w; quat[4 * k + 1] = p.second.q.x; quat[4 * k + 2] = p.second.q.y; quat[4 * k + 3] = p.second.q.z; }
    if (vel) { vel[3 * k] = p.second.v.x; vel[3 * k + 1] = p.second.v.y; vel[3 * k + 2] = p.second.v.z; }
    if (dur) dur[k] = p.second.dur;
    k++;
  }
  return k;
}
int64_t pgo_calib_integrate(const pgo_calib* c, const double* x, int64_t cap, int64_t* idx, double* speed, double* quat,
                            double* vel, int64_t* dur) {
  return dump_traj(integrate_literal(c, x), cap, idx, speed, quat, vel, dur);
}
int64_t pgo_calib_integrate_core(pgo_calib* c, const double* x, int64_t cap, int64_t* idx, double* speed, double* vel,
                                 int64_t* dur) {
  auto tr = integrate_core(c, x);
  if ((int64_t)tr.size() > cap) return -1;
  int64_t k = 0;
  for (const auto& p : tr) {
    idx[k] = (int64_t)p.first;
    speed[k] = pgbimu::norm3(pgbimu::v3(p.second.v.x, p.second.v.y, p.second.v.z));
    if (vel) { vel[3 * k] = p.second.v.x; vel[3 * k + 1] = p.second.v.y; vel[3 * k + 2] = p.second.v.z; }
    if (dur) dur[k] = p.second.dur;
    k++;
  }
  return k;
}

void pgo_smooth_time_series(const double* values, const double* times, int64_t n, const double* target, int64_t nt,
                            double sigma, double* out) {
  size_t left = 0, right = 0;
  for (int64_t ti = 0; ti < nt; ++ti) {
    const double t = target[ti];
    while (left + 1 < (size_t)n && (t - times[left + 1]) > 3 * sigma) ++left;
    while (right + 1 < (size_t)n && (times[right] - t) < 3 * sigma) ++right;
    double prev = 0, acc = 0;
    for (size_t i = left; i < right; ++i) {
      const double mid = (times[i] + times[i + 1]) / 2.0;
      const double cdf = normal_cdf(mid, t, sigma);
      acc += values[i] * (cdf - prev);
      prev = cdf;
    }
    acc += values[right] * (1.0 - prev);
    out[ti] = acc;
  }
}

// ComputeAndSaveForwardVelocitiesFromImu's window loop (fit_motion.cc:156-273).  mode 0 = literal evaluation +
// literal driver; mode 1 = contract-header evaluation + contract driver (what the CUDA path must equal bit for bit).
// Outputs: merged indices covered (ascending), their timestamps, averaged speeds, smoothed speeds; per-window x.
// Returns the number of covered merged events, or <0 on error.  nthreads > 1 runs windows concurrently (results
// are accumulated in window order afterwards, so they do not depend on nthreads).
int64_t pgo_fit_motion(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz, const int64_t* gyro_t,
                       int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t, int64_t n_acc, int batch_size,
                       int shift_step, int max_iters, double sigma, int mode, int64_t cap, int64_t* out_idx,
                       int64_t* out_t_usec, double* out_avg, double* out_smoothed, double* x_out, int32_t* iters_out,
                       double* fx_out, int64_t* n_evals_total) {
  std::map<size_t, std::vector<double>> integrated;
  int64_t evals = 0;
  int w = 0;
  for (size_t start = 0; start < (size_t)n_gps; start += shift_step, ++w) {
    const size_t end = std::min(start + (size_t)batch_size, (size_t)n_gps);
    pgo_calib* c = pgo_calib_create(gps_v + start, gps_t + start, (int)(end - start), gyro_xyz, gyro_t, n_gyro, acc_xyz,
                                    acc_t, n_acc);
    if (!c) return -2;
    double x[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, fx = 0;
    int ne = 0, it = 0;
    if (pgo_calib_num_intervals(c) > 0) {
      it = mode == 0 ? pgo_calib_minimize(c, x, &fx, max_iters, 1e-5, &ne) : pgo_calib_minimize_core(c, x, &fx, max_iters, 1e-5, &ne);
      if (it < 0) { pgo_calib_destroy(c); return -4; }
    }
    evals += ne;
    if (x_out) for (int i = 0; i < 9; i++) x_out[9 * w + i] = x[i];
    if (iters_out) iters_out[w] = it;
    if (fx_out) fx_out[w] = fx;
    if (mode == 0) {
      for (const auto& p : integrate_literal(c, x)) integrated[p.first].push_back(norm(p.second.v));
    } else {
      for (const auto& p : integrate_core(c, x))
        integrated[p.first].push_back(pgbimu::norm3(pgbimu::v3(p.second.v.x, p.second.v.y, p.second.v.z)));
    }
    pgo_calib_destroy(c);
  }
  if (n_evals_total) *n_evals_total = evals;
  if ((int64_t)integrated.size() > cap) return -1;
  // merged timestamps of the whole recording (the reference builds one more calibrator for this, :247-248)
  pgo_calib* all = pgo_calib_create(gps_v, gps_t, n_gps, gyro_xyz, gyro_t, n_gyro, acc_xyz, acc_t, n_acc);
  if (!all) return -2;
  std::vector<double> avg, ts_sec;
  std::vector<int64_t> ts_usec;
  int64_t k = 0;
  for (auto& p : integrated) {
    ts_usec.push_back(all->merged_t.at(p.first));
    ts_sec.push_back((double)(ts_usec.back() - ts_usec.front()) * 1e-6);
    const double sum = std::accumulate(p.second.begin(), p.second.end(), 0.0);
    avg.push_back(sum / p.second.size());
    out_idx[k] = (int64_t)p.first;
    out_t_usec[k] = ts_usec.back();
    out_avg[k] = avg.back();
    k++;
  }
  pgo_calib_destroy(all);
  if (k > 0) pgo_smooth_time_series(avg.data(), ts_sec.data(), k, ts_sec.data(), k, sigma, out_smoothed);
  return k;
}

// total_velocity_local of fit_motion.cc:172-173,223-248 for given per-window solutions x_all[w][9]: Kahan sum
// (include/math/math.hpp:8-27) of conj(orientation)*velocity over the trajectory points with |v| >= min_vel of the
// windows whose largest rotation acos(min |q.w|) reaches min_rot.  mode as in pgo_fit_motion.
int pgo_forward_axis_sum(const double* gps_v, const int64_t* gps_t, int n_gps, const double* gyro_xyz,
                         const int64_t* gyro_t, int64_t n_gyro, const double* acc_xyz, const int64_t* acc_t,
                         int64_t n_acc, int batch_size, int shift_step, const double* x_all, int mode, double min_vel,
                         double min_rot, double* sum_out, int32_t* windows_used) {
  Vec3 sum{0, 0, 0}, rem{0, 0, 0};
  int w = 0, used = 0;
  for (size_t start = 0; start < (size_t)n_gps; start += shift_step, ++w) {
    const size_t end = std::min(start + (size_t)batch_size, (size_t)n_gps);
    pgo_calib* c = pgo_calib_create(gps_v + start, gps_t + start, (int)(end - start), gyro_xyz, gyro_t, n_gyro, acc_xyz,
                                    acc_t, n_acc);
    if (!c) return -2;
    const auto traj = mode == 0 ? integrate_literal(c, x_all + 9 * w) : integrate_core(c, x_all + 9 * w);
    pgo_calib_destroy(c);
    double min_cos = 1.0;
    for (const auto& p : traj) min_cos = std::min(min_cos, std::abs(p.second.q.w));
    if (!(std::acos(min_cos) >= min_rot)) continue;
    used++;
    for (const auto& p : traj) {
      if (norm(p.second.v) >= min_vel) {
        const Quat qc{p.second.q.w, -p.second.q.x, -p.second.q.y, -p.second.q.z};
        const Vec3 v = transform_vector(qc, p.second.v);
        const Vec3 prop = v + rem;                       // KahanSum<Eigen::Vector3d>::add
        const Vec3 upd = sum + prop;
        const Vec3 act{upd.x - sum.x, upd.y - sum.y, upd.z - sum.z};
        rem = Vec3{prop.x - act.x, prop.y - act.y, prop.z - act.z};
        sum = upd;
      }
    }
  }
  sum_out[0] = sum.x; sum_out[1] = sum.y; sum_out[2] = sum.z;
  if (windows_used) *windows_used = used;
  return 0;
}

// cv::eigen for a symmetric 3x3 (OpenCV core/src/lapack.cpp JacobiImpl_, un-vendored; restated from the published
// algorithm): rows of V = eigenvectors by descending eigenvalue, signs as the sweep leaves them.
static void jacobi_sym(double* A, double* W, double* V, int n) {
  const double eps = std::numeric_limits<double>::epsilon();
  std::vector<int> indR(n, 0), indC(n, 0);
  for (int i = 0; i < n * n; i++) V[i] = 0.0;
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  auto scan_row = [&](int k) { int m = k + 1; double mv = std::abs(A[n * k + m]); for (int i = k + 2; i < n; i++) { double v = std::abs(A[n * k + i]); if (mv < v) { mv = v; m = i; } } return m; };
  auto scan_col = [&](int k) { int m = 0; double mv = std::abs(A[k]); for (int i = 1; i < k; i++) { double v = std::abs(A[n * i + k]); if (mv < v) { mv = v; m = i; } } return m; };
  for (int k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) indR[k] = scan_row(k);
    if (k > 0) indC[k] = scan_col(k);
  }
  if (n > 1)
    for (int iters = 0; iters < n * n * 30; iters++) {
      int k = 0;
      double mv = std::abs(A[indR[0]]);
      for (int i = 1; i < n - 1; i++) { double v = std::abs(A[n * i + indR[i]]); if (mv < v) { mv = v; k = i; } }
      int l = indR[k];
      for (int i = 1; i < n; i++) { double v = std::abs(A[n * indC[i] + i]); if (mv < v) { mv = v; k = indC[i]; l = i; } }
      const double p = A[n * k + l];
      if (std::abs(p) <= eps) break;
      const double y = (W[l] - W[k]) * 0.5;
      double t = std::abs(y) + std::hypot(p, y);
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

// cv::PCA(data, noArray(), CV_PCA_DATA_AS_ROW) for an n x 3 matrix: eigenvectors (3x3 row-major) and eigenvalues.
void pgo_pca3(const double* rows, int64_t n, double* eigvec9, double* eigval3, double* mean3) {
  double mean[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; i++) for (int c = 0; c < 3; c++) mean[c] += rows[3 * i + c];
  for (int c = 0; c < 3; c++) mean[c] /= (double)n;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; i++) {
    double d[3];
    for (int c = 0; c < 3; c++) d[c] = rows[3 * i + c] - mean[c];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cov[3 * a + b] += d[a] * d[b];
  }
  for (int i = 0; i < 9; i++) cov[i] /= (double)n;
  double W[3];
  jacobi_sym(cov, W, eigvec9, 3);
  if (eigval3) for (int c = 0; c < 3; c++) eigval3[c] = W[c];
  if (mean3) for (int c = 0; c < 3; c++) mean3[c] = mean[c];
}

// GetPrincipalRotationAxes (rotation.cc:16-57), literal: returns the number of integration intervals (rows of the PCA
// input, optionally copied to rows_out[cap][3]) or -1 when fewer than 3 (CHECK_GE).
int64_t pgo_principal_rotation_axes(const double* gyro_xyz, const int64_t* gyro_t, int64_t n, int64_t interval_usec,
                                    double* axes9, double* rows_out, int64_t cap) {
  std::vector<double> rows;
  Quat cur{1, 0, 0, 0};
  int64_t cur_usec = 0;
  for (int64_t i = 1; i < n; i++) {
    const int64_t dur = gyro_t[i] - gyro_t[i - 1];
    cur_usec += dur;
    const Quat rq = rotation_motion_to_quaternion(gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2], (double)dur * 1e-6);
    cur = qmul(cur, rq);
    if (cur_usec >= interval_usec) {
      rows.push_back(cur.x); rows.push_back(cur.y); rows.push_back(cur.z);
      cur = Quat{1, 0, 0, 0};
      cur_usec = 0;
    }
  }
  const int64_t m = (int64_t)rows.size() / 3;
  if (m < 3) return -1;
  if (rows_out) for (int64_t i = 0; i < std::min(m, cap) * 3; i++) rows_out[i] = rows[i];
  pgo_pca3(rows.data(), m, axes9, nullptr, nullptr);
  return m;
}

// GetAngularVelocitiesAroundAxisDirect (rotation.cc:103-119)
void pgo_angular_velocities_around_axis(const double* gyro_xyz, int64_t n, const double* axis, double* out) {
  const double nrm = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);  // cv::norm(axis, NORM_L2)
  for (int64_t i = 0; i < n; i++)
    out[i] = (gyro_xyz[3 * i] * axis[0] + gyro_xyz[3 * i + 1] * axis[1] + gyro_xyz[3 * i + 2] * axis[2]) / nrm;
}

// annotate_frames.cc:59-72 over TimeSeries<double>::TimeAveragedValue / MostRecentPreviousValue / LinearInterpolate
// (include/interpolation/time_series.hpp:103-225), literal including the forward scans from the hint.  Returns 0,
// or -1 where the reference would CHECK-fail.  out/valid have n_frames-1 entries.
int pgo_time_averaged_values(const double* values, const int64_t* times, int64_t n, const int64_t* ft, int64_t n_frames,
                             double* out, uint8_t* valid) {
  if (n <= 0) return -1;
  auto interval_sec = [](int64_t a, int64_t b) { return (double)(b - a) * 1e-6; };
  struct R { double value; bool ok; int64_t end; };
  auto most_recent = [&](int64_t q, int64_t hint, bool* fatal) -> R {
    if (hint >= n) { *fatal = true; return {0, false, 0}; }
    if (times[0] > q) return {std::numeric_limits<double>::quiet_NaN(), false, 0};
    if (!(times[hint] <= q)) { *fatal = true; return {0, false, 0}; }
    int64_t next = hint;
    while (next < n && times[next] <= q) ++next;
    return {values[next - 1], true, next - 1};
  };
  auto lerp = [&](int64_t l, int64_t r, int64_t q, bool* fatal) -> double {
    if (!(l < r) || !(r < n) || !(times[l] <= q) || !(q <= times[r])) { *fatal = true; return 0; }
    const double lt = interval_sec(times[l], q), rt = interval_sec(q, times[r]), tt = interval_sec(times[l], times[r]);
    return (lt / tt) * values[r] + (rt / tt) * values[l];
  };
  R ann{0, false, 0};
  for (int64_t i = 1; i < n_frames; i++) {
    const int64_t start = ft[i - 1], end = ft[i];
    bool fatal = false;
    if (!(end > start)) return -1;
    if (start < times[0] || end > times[n - 1]) {
      ann = {std::numeric_limits<double>::quiet_NaN(), false, 0};
    } else {
      const R se = most_recent(start, ann.end, &fatal);
      if (fatal || !se.ok) return -1;
      const R ee = most_recent(end, se.end, &fatal);
      if (fatal || !ee.ok) return -1;
      double total = 0;
      for (int64_t k = se.end + 1; k < ee.end; ++k) total += (interval_sec(times[k], times[k + 1]) * 0.5 * (values[k] + values[k + 1]));
      const double lv = lerp(se.end, se.end + 1, start, &fatal), rv = lerp(ee.end, ee.end + 1, end, &fatal);
      if (fatal) return -1;
      if (se.end == ee.end) {
        total += (lv + rv) * 0.5 * interval_sec(start, end);
      } else {
        total += (lv + values[se.end + 1]) * 0.5 * interval_sec(start, times[se.end + 1]);
        total += (values[ee.end] + rv) * 0.5 * interval_sec(times[ee.end], end);
      }
      ann = {total / interval_sec(start, end), true, ee.end};
    }
    out[i - 1] = ann.value;
    valid[i - 1] = ann.ok ? 1 : 0;
  }
  return 0;
}

}  // extern "C"

extern "C" void pgo_det_sincos(double x, double* s, double* c) { pgbimu::det_sincos(x, s, c); }
