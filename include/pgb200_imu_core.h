/*
 * pgb200_imu_core.h -- the fp64 arithmetic contract of the IMU+GPS calibration path (K9/K10), shared by the
 * sm_100a kernels and by any host code that must reproduce them bit for bit.
 *
 * Everything here is plain +,-,*,/ and sqrt in a fixed order.  It must be compiled with floating-point
 * contraction OFF (nvcc -fmad=false, gcc -ffp-contract=off): IEEE-754 then makes host and device results
 * identical, which is what the 1e-6 velocity parity of a 500-iteration L-BFGS run on an ill-conditioned valley
 * needs.
 *
 * Reference formulas (file:line in waiwnf/pilotguru):
 *   RotationMotionToQuaternion / IntegrateMotion      src/geometry/geometry.cc:6-22, :24-53
 *   AccelerometerCalibrator::eval (loss + 9-gradient)  src/calibration/velocity.cc:41-180
 *   IntegrateTrajectory                                src/calibration/velocity.cc:199-256
 *   LBFGSSolver::minimize, LineSearch::Backtracking    thirdparty/LBFGS/LBFGS.h:79-182, LBFGS/LineSearch.h:41-111
 *   Eigen quaternion product / _transformVector / toRotationMatrix (un-vendored Eigen 3)
 *
 * Formulation.  The orientation never depends on the unknowns x = (g, h, v0), so the reference's per-interval
 * recursion is linear in x.  One GPS interval r (the IMU sub-intervals between two GPS fixes) is swept ONCE, in
 * the frame of the orientation at its start, into a GpsLocal record; a window (<= 40 GPS fixes) chains its
 * records with the running orientation into per-interval coefficient records WinRec, after which
 *     D_j = T_j*v0 + a_j + B_j*h + g*c_j          (integrated travel over GPS interval j)
 * and the loss / gradient of velocity.cc:41-180 cost O(#GPS) per evaluation.
 */
#ifndef PGB200_IMU_CORE_H_
#define PGB200_IMU_CORE_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define PGB_HD __host__ __device__ __forceinline__
#else
#define PGB_HD inline
#endif

#if defined(__FP_FAST_FMA) && !defined(__CUDACC__) && !defined(PGB_ALLOW_FMA_MACRO)
/* __FP_FAST_FMA only says the target has FMA; contraction must still be disabled by -ffp-contract=off. */
#endif

namespace pgbimu {

struct V3 { double x, y, z; };
struct Q4 { double w, x, y, z; };
struct M3 { double m[9]; };  /* row-major */

PGB_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PGB_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PGB_HD V3 scale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
PGB_HD M3 m3zero() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0.0; return r; }
PGB_HD M3 madd(M3 a, const M3& b) { for (int i = 0; i < 9; i++) a.m[i] = a.m[i] + b.m[i]; return a; }
PGB_HD M3 mscale(M3 a, double s) { for (int i = 0; i < 9; i++) a.m[i] = a.m[i] * s; return a; }
PGB_HD V3 mv(const M3& a, V3 v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
PGB_HD V3 mtv(const M3& a, V3 v) { /* transpose(a) * v */
  return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
PGB_HD M3 mm(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}

/* Eigen::Quaterniond product a*b */
PGB_HD Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
/* Eigen::Quaterniond::toRotationMatrix */
PGB_HD M3 qmat(Q4 q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0] = 1.0 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
  r.m[3] = txy + twz; r.m[4] = 1.0 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1.0 - (txx + tyy);
  return r;
}
/* Eigen::Quaterniond::_transformVector: v + w*uv + vec x uv, uv = 2 * (vec x v) */
PGB_HD V3 qrot(Q4 q, V3 v) {
  V3 uv = v3(q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x);
  uv = add(uv, uv);
  const V3 c = v3(q.y * uv.z - q.z * uv.y, q.z * uv.x - q.x * uv.z, q.x * uv.y - q.y * uv.x);
  return v3(v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z);
}

/* Deterministic sin/cos: Cody-Waite reduction by pi/2 (three-part constant, exact products for |k| < 2^20) and
 * the fdlibm kernel polynomials.  Absolute error < 2e-16 on the reduced range; identical on host and device. */
PGB_HD void det_sincos(double x, double* s, double* c) {
  const double invpio2 = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00, p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21;
  const double p3t = 8.47842766036889956997e-32;
  const double t = x * invpio2;
  const long long k = (long long)(t + (t < 0.0 ? -0.5 : 0.5));
  const double fk = (double)k;
  double r = x - fk * p1;
  r = r - fk * p2;
  r = r - fk * p3;
  r = r - fk * p3t;
  const double z = r * r;
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  const double sn = r + (z * r) * (S1 + z * ps);
  const double pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double cs = 1.0 - (0.5 * z - z * pc);
  switch ((int)(k & 3)) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}

#if defined(__CUDA_ARCH__)
PGB_HD double det_sqrt(double v) { return __dsqrt_rn(v); }
#else
}  /* namespace */
#include <math.h>
namespace pgbimu {
PGB_HD double det_sqrt(double v) { return sqrt(v); }
#endif

PGB_HD double norm3(V3 v) { return det_sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

/* RotationMotionToQuaternion (geometry.cc:6-22) */
PGB_HD Q4 rotation_motion_to_quaternion(double rx, double ry, double rz, double duration_sec) {
  const double rate = det_sqrt(rx * rx + ry * ry + rz * rz);
  const double half_theta = rate * duration_sec * 0.5;
  double sn, cs;
  det_sincos(half_theta, &sn, &cs);
  const double k = sn / (rate + 1e-30);
  Q4 q;
  q.w = cs; q.x = rx * k; q.y = ry * k; q.z = rz * k;
  return q;
}

/* ---------------------------------------------------------------- per-GPS-interval sweep (K9, x-independent) */
struct GpsLocal {
  double T;        /* sum of dt_k (seconds) */
  double ct;       /* sum dt_k * pt_k, pt_k = elapsed seconds since the start of this GPS interval (from usec) */
  double et;       /* pt_n */
  long long dur;   /* usec */
  Q4 P;            /* product of the dq_k */
  V3 ca; M3 CR;    /* sum dt_k*pa_k, sum dt_k*pR_k ; pa_k = sum_{j<=k} R(l_{j-1}) a_j dt_j, pR_k = sum R(l_{j-1}) dt_j */
  V3 ea; M3 ER;    /* pa_n, pR_n */
  M3 CRa, ERa;     /* sum dt_k*pRa_k, pRa_n ; pRa_k = sum_{j<=k} R(l_j) dt_j */
};

/* One IMU sub-interval of a GPS interval: the merged-event samples that end it (velocity.cc:78-84) and its span. */
struct ImuStep {
  double wx, wy, wz;  /* gyro rate rad/s */
  double ax, ay, az;  /* raw acceleration */
  long long dur_usec;
};

struct SweepState {
  Q4 l;            /* orientation relative to the start of the GPS interval */
  V3 pa; M3 pR, pRa;
  long long tau;   /* usec since the start of the GPS interval */
};

PGB_HD void sweep_init(SweepState* s, GpsLocal* g) {
  s->l.w = 1.0; s->l.x = 0.0; s->l.y = 0.0; s->l.z = 0.0;
  s->pa = v3(0, 0, 0); s->pR = m3zero(); s->pRa = m3zero(); s->tau = 0;
  g->T = 0.0; g->ct = 0.0; g->et = 0.0; g->dur = 0;
  g->ca = v3(0, 0, 0); g->CR = m3zero(); g->CRa = m3zero();
}

/* Advances the local state by one sub-interval; returns its dt.  After the call s holds pa_k, pR_k, pRa_k, l_k. */
PGB_HD double sweep_step(SweepState* s, const ImuStep& st) {
  const double dt = (double)st.dur_usec * 1e-6;
  const M3 Rb = qmat(s->l);
  const V3 ra = mv(Rb, v3(st.ax, st.ay, st.az));
  s->pa = add(s->pa, scale(ra, dt));
  s->pR = madd(s->pR, mscale(Rb, dt));
  s->l = qmul(s->l, rotation_motion_to_quaternion(st.wx, st.wy, st.wz, dt));
  s->pRa = madd(s->pRa, mscale(qmat(s->l), dt));
  s->tau += st.dur_usec;
  return dt;
}

PGB_HD void sweep_accumulate(const SweepState& s, double dt, GpsLocal* g) {
  const double pt = (double)s.tau * 1e-6;
  g->T = g->T + dt;
  g->ct = g->ct + dt * pt;
  g->ca = add(g->ca, scale(s.pa, dt));
  g->CR = madd(g->CR, mscale(s.pR, dt));
  g->CRa = madd(g->CRa, mscale(s.pRa, dt));
}

PGB_HD void sweep_finish(const SweepState& s, GpsLocal* g) {
  g->et = (double)s.tau * 1e-6;
  g->dur = s.tau;
  g->P = s.l;
  g->ea = s.pa; g->ER = s.pR; g->ERa = s.pRa;
}

/* ---------------------------------------------------------------- per-window chaining */
struct WinRec {     /* one GPS interval of a window, in the window's fixed frame */
  double T, c, dref, T2;   /* dref = gps speed * T ; T2 = sum_k tau_k[s] * dt_k with tau from the window start */
  V3 a; M3 B;              /* D = T*v0 + a + B*h + g*c */
  M3 MW;                   /* sum_k dt_k * W_k  (grad_h += MW^T * dL) */
  M3 RQ;                   /* rotation at the start of the interval (K10) */
  Q4 Q;                    /* the same as a quaternion (K10 orientation output) */
  V3 Sa; M3 SE; double St; /* velocity at the start of the interval: v0 + Sa + SE*h + g*St (K10) */
};

struct WinState {
  Q4 Q; M3 W; long long tau;
  V3 Sa; M3 SE; double St;
};

PGB_HD void win_init(WinState* w) {
  w->Q.w = 1.0; w->Q.x = 0.0; w->Q.y = 0.0; w->Q.z = 0.0;
  w->W = m3zero(); w->tau = 0; w->Sa = v3(0, 0, 0); w->SE = m3zero(); w->St = 0.0;
}

PGB_HD void win_chain(WinState* w, const GpsLocal& g, double gps_speed, WinRec* r) {
  const M3 RQ = qmat(w->Q);
  r->RQ = RQ;
  r->Q = w->Q;
  r->T = g.T;
  r->dref = gps_speed * g.T;
  r->T2 = ((double)w->tau * 1e-6) * g.T + g.ct;
  r->Sa = w->Sa; r->SE = w->SE; r->St = w->St;
  r->a = add(scale(w->Sa, g.T), mv(RQ, g.ca));
  r->B = madd(mscale(w->SE, g.T), mm(RQ, g.CR));
  r->c = g.T * w->St + g.ct;
  r->MW = madd(mscale(w->W, g.T), mm(RQ, g.CRa));
  /* advance */
  w->Sa = add(w->Sa, mv(RQ, g.ea));
  w->SE = madd(w->SE, mm(RQ, g.ER));
  w->St = w->St + g.et;
  w->W = madd(w->W, mm(RQ, g.ERa));
  w->Q = qmul(w->Q, g.P);
  w->tau += g.dur;
}

/* AccelerometerCalibrator::eval (velocity.cc:41-180) on a chained window. x = (g, h, v0).
 *
 * Summation shape (part of the contract).  The reference adds the per-GPS-interval terms of the loss and of the
 * gradient sequentially; here the sum over the window's n records has the shape of ONE WARP:
 *   lane l (0..31) adds the terms of records l, l+32, l+64, ... in ascending order onto +0.0   (imu_eval_lane);
 *   then five butterfly steps k = 16, 8, 4, 2, 1: every lane replaces its partial P_l by P_l + P_(l xor k);
 *   the result is lane 0's value (IEEE addition is commutative, so all 32 lanes end with identical bits).
 * The device runs one warp per window and does the butterfly with __shfl_xor_sync (csrc/imu.cu: k_imu_solve); the host
 * emulation below walks the same 32 partials.  Per evaluation this differs from the sequential order by rounding only
 * (<= 1e-15 relative on loss and gradient, asserted against the literal restatement in tests/test_oracle_calib.py). */
enum { PGB_EVAL_LANES = 32 };

/* Partial sums of lane `lane`: acc[0] = sum e^2, acc[1..3] = grad_g, acc[4..6] = grad_h, acc[7..9] = grad_v0 terms. */
PGB_HD void imu_eval_lane(const WinRec* rec, int n, int lane, const double* x, double* acc) {
  const V3 g = v3(x[0], x[1], x[2]), h = v3(x[3], x[4], x[5]), v0 = v3(x[6], x[7], x[8]);
  double loss = 0.0;
  V3 gg = v3(0, 0, 0), gh = v3(0, 0, 0), gv = v3(0, 0, 0);
  for (int j = lane; j < n; j += PGB_EVAL_LANES) {
    const WinRec& r = rec[j];
    V3 D = add(add(scale(v0, r.T), r.a), add(mv(r.B, h), scale(g, r.c)));
    const double dn = norm3(D);
    const double e = dn - r.dref;
    loss = loss + e * e;
    const double k = 2.0 * e / (dn + 1e-5);
    const V3 dL = scale(D, k);
    gg = add(gg, scale(dL, r.T2));
    gh = add(gh, mtv(r.MW, dL));
    gv = add(gv, scale(dL, r.T));
  }
  acc[0] = loss;
  acc[1] = gg.x; acc[2] = gg.y; acc[3] = gg.z;
  acc[4] = gh.x; acc[5] = gh.y; acc[6] = gh.z;
  acc[7] = gv.x; acc[8] = gv.y; acc[9] = gv.z;
}

/* Normalisation by the window's total time (velocity.cc:176-179); returns the loss. */
PGB_HD double imu_eval_finish(const double* acc, long long total_usec, double* grad) {
  const double total = (double)total_usec * 1e-6;
  for (int i = 0; i < 9; i++) grad[i] = acc[1 + i] / total;
  return acc[0] / total;
}

/* The whole evaluation by ONE thread (host side of the contract, and the single-thread device paths): emulates the 32
 * lanes and the butterfly. */
PGB_HD double imu_eval(const WinRec* rec, int n, long long total_usec, const double* x, double* grad) {
  double P[PGB_EVAL_LANES][10], T[PGB_EVAL_LANES][10];
  for (int l = 0; l < PGB_EVAL_LANES; l++) imu_eval_lane(rec, n, l, x, P[l]);
  for (int k = PGB_EVAL_LANES / 2; k >= 1; k >>= 1) {
    for (int l = 0; l < PGB_EVAL_LANES; l++)
      for (int i = 0; i < 10; i++) T[l][i] = P[l][i] + P[l ^ k][i];
    for (int l = 0; l < PGB_EVAL_LANES; l++)
      for (int i = 0; i < 10; i++) P[l][i] = T[l][i];
  }
  return imu_eval_finish(P[0], total_usec, grad);
}

/* ---------------------------------------------------------------- L-BFGS (LBFGS.h:79-182, LineSearch.h:41-111) */
struct LbfgsParam {
  int m;               /* fixed at 6 by the storage below */
  double epsilon;
  int max_iterations;
  int max_linesearch;
  double min_step, max_step, ftol;
};
PGB_HD LbfgsParam lbfgs_default() {
  LbfgsParam p;
  p.m = 6; p.epsilon = 1e-5; p.max_iterations = 0; p.max_linesearch = 20;
  p.min_step = 1e-20; p.max_step = 1e+20; p.ftol = 1e-4;
  return p;
}

PGB_HD double dot9(const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < 9; i++) s = s + a[i] * b[i];
  return s;
}

/* Returns the iteration count (as LBFGSSolver::minimize) or a negative status: -4 when the line-search step
 * leaves [min_step, max_step] (the reference throws std::runtime_error there). n is fixed at 9, m at 6. */
/* ws: storage for the m = 6 correction pairs, 2 * 6 * 9 doubles (the device keeps it in shared memory, one block per
 * window: every lane of the window's warp runs this function redundantly on identical values). */
enum { PGB_LBFGS_WS_DOUBLES = 2 * 6 * 9 };
template <typename F>
PGB_HD int lbfgs_minimize9(F& f, double* x, double* fx_out, const LbfgsParam& P, int* n_eval, double* ws) {
  const int n = 9, m = 6;
  double (*S)[9] = reinterpret_cast<double (*)[9]>(ws);
  double (*Y)[9] = reinterpret_cast<double (*)[9]>(ws + 6 * 9);
  double ys_h[6], alpha[6];
  double xp[9], grad[9], gradp[9], drt[9];
  int evals = 0;
  double fx = f(x, grad);
  evals++;
  double xnorm = det_sqrt(dot9(x, x)), gnorm = det_sqrt(dot9(grad, grad));
  if (gnorm <= P.epsilon * (xnorm > 1.0 ? xnorm : 1.0)) { *fx_out = fx; if (n_eval) *n_eval = evals; return 1; }
  for (int i = 0; i < n; i++) drt[i] = -grad[i];
  double step = 1.0 / det_sqrt(dot9(drt, drt));
  int k = 1, end = 0;
  for (;;) {
    for (int i = 0; i < n; i++) { xp[i] = x[i]; gradp[i] = grad[i]; }
    /* LineSearch::Backtracking, Armijo */
    {
      const double fx_init = fx;
      const double dg_init = dot9(grad, drt);
      const double dg_test = P.ftol * dg_init;
      for (int iter = 0; iter < P.max_linesearch; iter++) {
        for (int i = 0; i < n; i++) x[i] = xp[i] + step * drt[i];
        fx = f(x, grad);
        evals++;
        if (!(fx > fx_init + step * dg_test)) break;
        if (step < P.min_step || step > P.max_step) { *fx_out = fx; if (n_eval) *n_eval = evals; return -4; }
        step = step * 0.5;
      }
    }
    xnorm = det_sqrt(dot9(x, x));
    gnorm = det_sqrt(dot9(grad, grad));
    if (gnorm <= P.epsilon * (xnorm > 1.0 ? xnorm : 1.0)) break;
    if (P.max_iterations != 0 && k >= P.max_iterations) break;
    double* sv = S[end];
    double* yv = Y[end];
    for (int i = 0; i < n; i++) { sv[i] = x[i] - xp[i]; yv[i] = grad[i] - gradp[i]; }
    const double ys = dot9(yv, sv), yy = dot9(yv, yv);
    ys_h[end] = ys;
    for (int i = 0; i < n; i++) drt[i] = -grad[i];
    const int bound = m < k ? m : k;
    end = (end + 1) % m;
    int j = end;
    for (int i = 0; i < bound; i++) {
      j = (j + m - 1) % m;
      alpha[j] = dot9(S[j], drt) / ys_h[j];
      for (int q = 0; q < n; q++) drt[q] = drt[q] - alpha[j] * Y[j][q];
    }
    const double sc = ys / yy;
    for (int q = 0; q < n; q++) drt[q] = drt[q] * sc;
    for (int i = 0; i < bound; i++) {
      const double beta = dot9(Y[j], drt) / ys_h[j];
      const double co = alpha[j] - beta;
      for (int q = 0; q < n; q++) drt[q] = drt[q] + co * S[j][q];
      j = (j + 1) % m;
    }
    step = 1.0;
    k++;
  }
  *fx_out = fx;
  if (n_eval) *n_eval = evals;
  return k;
}

}  /* namespace pgbimu */
#endif /* PGB200_IMU_CORE_H_ */
