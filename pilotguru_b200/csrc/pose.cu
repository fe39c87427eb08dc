// Pose-only bundle adjustment for a batch of frames (sm_100a): Optimizer::PoseOptimization
// (thirdparty/orb-slam2/src/Optimizer.cc:239-451), monocular edges, with the g2o machinery it drives restated as one
// kernel: Levenberg-Marquardt (thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:60-186), the active-edge /
// robust-chi2 bookkeeping of SparseOptimizer (sparse_optimizer.cpp:100-114,206-267,354-419), the unary-edge quadratic
// form with the Huber kernel (base_unary_edge.hpp:43-72, robust_kernel_impl.cpp:65-91), EdgeSE3ProjectXYZOnlyPose
// (types_six_dof_expmap.cpp:266-296), SE3Quat::exp / operator* (types/se3quat.h) and the dense LDLT of the 6x6 system
// (solvers/linear_solver_dense.h:65-111).
//
// One CTA per frame.  A frame's edges (<= cap, one per feature that holds a map point) are spread over the threads; each
// pass over the edges ends in a fixed-shape fp64 reduction (per-thread partial sums in edge order, xor-shuffle tree,
// four warp partials added in order) so results do not depend on scheduling.  The 6x6 solve and the SE(3) update are
// done by thread 0; every thread then replays the (scalar) LM control flow on the same broadcast values.  The only
// per-edge state g2o keeps between passes is _error (it is what chi2() reads when the four rounds classify inliers,
// including after a rejected trial step) -- it lives in shared memory next to the edge's level.
// fp64 throughout, like g2o; -fmad=false keeps every per-edge value identical to an unfused CPU evaluation, only the
// summation order of the reductions differs.
#include <cuda_runtime.h>

#include <cfloat>

#include "common.cuh"

namespace pgb {

namespace {

constexpr int kPoThreads = 128;
constexpr int kPoWarps = kPoThreads / 32;
constexpr int kPoSums = 28;  // 21 upper-triangle entries of H, 6 of b, 1 chi2

struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

struct PoseArgs {
  int cap, nlevels;
  double fx, fy, cx, cy, delta, dsqr;
  float invSigma2[16];
  const float* TcwIn;      // [frame][16]
  const float* kpXY;       // [frame][cap][2]
  const int* kpOctave;     // [frame][cap]
  const float* mpXYZ;      // [frame][cap][3]
  const uint8_t* hasMp;    // [frame][cap]
  const int* counts;       // [frame]
  float* TcwOut;           // [frame][16]
  uint8_t* outlier;        // [frame][cap]
  int* nInliers;           // [frame]
  int* err;
};

__device__ __forceinline__ void normalize_rotation(Quat& q) {
  if (q.w < 0) { q.x *= -1; q.y *= -1; q.z *= -1; q.w *= -1; }
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

__device__ __noinline__ Quat quat_from_matrix(const double m[3][3]) {
  Quat q;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0) {
    t = sqrt(t + 1.0);
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
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
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

__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

__device__ __forceinline__ void quat_rotate(const Quat& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}

__device__ __noinline__ void se3_exp_mul(const double u[6], const SE3& est, SE3& out) {  // SE3Quat::exp(u) * est
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
  double O2[3][3], R[3][3], V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  if (theta < 0.00001) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) V[i][j] = R[i][j] = (i == j ? 1.0 : 0.0) + O[i][j] + O2[i][j];
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
    const double c = (theta - sin(theta)) / pow(theta, 3.0);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        R[i][j] = (i == j ? 1.0 : 0.0) + a * O[i][j] + b * O2[i][j];
        V[i][j] = (i == j ? 1.0 : 0.0) + b * O[i][j] + c * O2[i][j];
      }
  }
  SE3 e;
  e.r = quat_from_matrix(R);
  for (int i = 0; i < 3; i++) e.t[i] = V[i][0] * up[0] + V[i][1] * up[1] + V[i][2] * up[2];
  normalize_rotation(e.r);
  double rt[3];
  quat_rotate(e.r, est.t, rt);
  for (int i = 0; i < 3; i++) out.t[i] = e.t[i] + rt[i];
  out.r = quat_mul(e.r, est.r);
  normalize_rotation(out.r);
}

// Eigen::LDLT of a 6x6 (unblocked, pivoting on the largest remaining |diagonal|), isPositive(), solve.
__device__ __noinline__ bool ldlt6_solve(double A[6][6], const double b[6], double x[6]) {
  int tr[6];
  int sign = 0;
  _Pragma("unroll 1") for (int k = 0; k < 6; k++) {
    int p = k;
    double big = fabs(A[k][k]);
    _Pragma("unroll 1") for (int i = k + 1; i < 6; i++)
      if (fabs(A[i][i]) > big) { big = fabs(A[i][i]); p = i; }
    tr[k] = p;
    if (p != k) {
      double t;
      _Pragma("unroll 1") for (int j = 0; j < k; j++) { t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
      _Pragma("unroll 1") for (int i = p + 1; i < 6; i++) { t = A[i][k]; A[i][k] = A[i][p]; A[i][p] = t; }
      t = A[k][k]; A[k][k] = A[p][p]; A[p][p] = t;
      _Pragma("unroll 1") for (int i = k + 1; i < p; i++) { t = A[i][k]; A[i][k] = A[p][i]; A[p][i] = t; }
    }
    double temp[6];
    _Pragma("unroll 1") for (int j = 0; j < k; j++) temp[j] = A[j][j] * A[k][j];
    _Pragma("unroll 1") for (int j = 0; j < k; j++) A[k][k] -= A[k][j] * temp[j];
    _Pragma("unroll 1") for (int i = k + 1; i < 6; i++)
      _Pragma("unroll 1") for (int j = 0; j < k; j++) A[i][k] -= A[i][j] * temp[j];
    const double akk = A[k][k];
    if (fabs(akk) > 0)
      _Pragma("unroll 1") for (int i = k + 1; i < 6; i++) A[i][k] /= akk;
    if (sign == 1) { if (akk < 0) sign = 2; }
    else if (sign == -1) { if (akk > 0) sign = 2; }
    else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
  }
  if (!(sign == 1 || sign == 0)) return false;
  double y[6];
  _Pragma("unroll 1") for (int i = 0; i < 6; i++) y[i] = b[i];
  _Pragma("unroll 1") for (int k = 0; k < 6; k++) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  _Pragma("unroll 1") for (int i = 0; i < 6; i++)
    _Pragma("unroll 1") for (int j = 0; j < i; j++) y[i] -= A[i][j] * y[j];
  _Pragma("unroll 1") for (int i = 0; i < 6; i++) y[i] = fabs(A[i][i]) > DBL_MIN ? y[i] / A[i][i] : 0.0;
  _Pragma("unroll 1") for (int i = 5; i >= 0; i--)
    _Pragma("unroll 1") for (int j = i + 1; j < 6; j++) y[i] -= A[j][i] * y[j];
  _Pragma("unroll 1") for (int k = 5; k >= 0; k--) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  _Pragma("unroll 1") for (int i = 0; i < 6; i++) x[i] = y[i];
  return true;
}

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sRed, double* sOut) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double a = v[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) sRed[warp * K + k] = a;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double a = sRed[threadIdx.x];
#pragma unroll
    for (int w = 1; w < kPoWarps; w++) a += sRed[w * K + threadIdx.x];
    sOut[threadIdx.x] = a;
  }
  __syncthreads();
}

struct EdgeIn { double ox, oy, X[3], info; };

__global__ void __launch_bounds__(kPoThreads) k_pose_optimization(PoseArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double sRed[kPoWarps * kPoSums];
  __shared__ double sOut[kPoSums];
  __shared__ SE3 sTrial;
  __shared__ double sX[6];
  __shared__ int sOk, sCount[kPoWarps];
  const int p = blockIdx.x, tid = threadIdx.x, cap = A.cap;
  const int n = min(max(A.counts[p], 0), cap);
  double2* sErr = reinterpret_cast<double2*>(smem);                 // [cap] _error of the edge's last computeError()
  uint8_t* sLevel = reinterpret_cast<uint8_t*>(sErr + cap);         // [cap] 0 active, 1 outlier, 2 no edge
  const float* kpXY = A.kpXY + (size_t)p * cap * 2;
  const float* mpXYZ = A.mpXYZ + (size_t)p * cap * 3;
  const int* kpOct = A.kpOctave + (size_t)p * cap;
  const double fx = A.fx, fy = A.fy, cx = A.cx, cy = A.cy, delta = A.delta, dsqr = A.dsqr;

  auto block_count = [&](int mine) -> int {   // sum of an int over the CTA
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    __syncthreads();
    if ((tid & 31) == 0) sCount[tid >> 5] = mine;
    __syncthreads();
    int s = 0;
    for (int w = 0; w < kPoWarps; w++) s += sCount[w];
    return s;
  };
  auto load_edge = [&](int i) -> EdgeIn {
    EdgeIn e;
    e.ox = kpXY[2 * i]; e.oy = kpXY[2 * i + 1];
    e.X[0] = mpXYZ[3 * i]; e.X[1] = mpXYZ[3 * i + 1]; e.X[2] = mpXYZ[3 * i + 2];
    e.info = A.invSigma2[min(max(kpOct[i], 0), 15)];
    return e;
  };

  if (tid < 6) sX[tid] = 0.0;
  int mine = 0;
  bool bad = false;
  for (int i = tid; i < cap; i += kPoThreads) {
    const bool has = i < n && A.hasMp[(size_t)p * cap + i] != 0;
    if (has && (kpOct[i] < 0 || kpOct[i] >= A.nlevels)) bad = true;
    sLevel[i] = has ? 0 : 2;
    sErr[i] = make_double2(0.0, 0.0);
    A.outlier[(size_t)p * cap + i] = 0;
    mine += has;
  }
  if (bad) atomicOr(A.err, 1);
  const int nInitial = block_count(mine);
  const float* Tin = A.TcwIn + (size_t)p * 16;
  float* Tout = A.TcwOut + (size_t)p * 16;
  if (nInitial < 3) {  // Optimizer.cc:363-364
    if (tid < 16) Tout[tid] = Tin[tid];
    if (tid == 0) A.nInliers[p] = 0;
    return;
  }

  SE3 est0;  // Converter::toSE3Quat(pFrame->mTcw)
  {
    double R[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i][j] = Tin[4 * i + j];
    est0.r = quat_from_matrix(R);
    normalize_rotation(est0.r);
    for (int i = 0; i < 3; i++) est0.t[i] = Tin[4 * i + 3];
  }
  SE3 est = est0;
  bool robust = true;
  int nBadEdges = 0;

  // error of edge e at pose s; chi2 = e^T (info * e)
  auto edge_error = [&](const EdgeIn& e, const SE3& s, double& ex, double& ey, double pc[3]) {
    double r[3];
    quat_rotate(s.r, e.X, r);
    pc[0] = r[0] + s.t[0]; pc[1] = r[1] + s.t[1]; pc[2] = r[2] + s.t[2];
    const double px = pc[0] / pc[2], py = pc[1] / pc[2];
    ex = e.ox - (px * fx + cx);
    ey = e.oy - (py * fy + cy);
  };
  auto robust_rho0 = [&](double c) -> double { return c <= dsqr ? c : 2 * sqrt(c) * delta - dsqr; };

  for (int it = 0; it < 4; it++) {
    est = est0;
    int act = 0;
    for (int i = tid; i < cap; i += kPoThreads) act += sLevel[i] == 0;
    const bool anyActive = block_count(act) > 0;   // else g2o: "0 vertices to optimize", optimize() returns at once
    double lambda = -1., ni = 2.;
    int lmBad = 0;
    for (int iter = 0; anyActive && iter < 10; iter++) {
      // computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate
      double acc[kPoSums];
#pragma unroll
      for (int k = 0; k < kPoSums; k++) acc[k] = 0.0;
      for (int i = tid; i < cap; i += kPoThreads) {
        if (sLevel[i] != 0) continue;
        const EdgeIn e = load_edge(i);
        double ex, ey, pc[3];
        edge_error(e, est, ex, ey, pc);
        sErr[i] = make_double2(ex, ey);
        const double c = ex * (e.info * ex) + ey * (e.info * ey);
        double w = 1.0;
        if (robust) {
          acc[27] += robust_rho0(c);
          if (c > dsqr) w = delta / sqrt(c);
        } else {
          acc[27] += c;
        }
        const double x = pc[0], y = pc[1], invz = 1.0 / pc[2], invz_2 = invz * invz;
        double J0[6], J1[6];
        J0[0] = x * y * invz_2 * fx; J0[1] = -(1 + (x * x * invz_2)) * fx; J0[2] = y * invz * fx;
        J0[3] = -invz * fx; J0[4] = 0; J0[5] = x * invz_2 * fx;
        J1[0] = (1 + y * y * invz_2) * fy; J1[1] = -x * y * invz_2 * fy; J1[2] = -x * invz * fy;
        J1[3] = 0; J1[4] = -invz * fy; J1[5] = y * invz_2 * fy;
        const double wi = w * e.info, iex = e.info * ex, iey = e.info * ey;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
          for (int b = a; b < 6; b++) acc[k++] += J0[a] * wi * J0[b] + J1[a] * wi * J1[b];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] -= w * (J0[a] * iex + J1[a] * iey);
      }
      block_sum<kPoSums>(acc, sRed, sOut);
      double H[6][6], b[6];
      {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
          for (int c = a; c < 6; c++) { H[a][c] = sOut[k]; H[c][a] = sOut[k]; k++; }
#pragma unroll
        for (int a = 0; a < 6; a++) b[a] = sOut[21 + a];
      }
      double currentChi = sOut[27], tempChi = currentChi;
      const double iniChi = currentChi;
      if (iter == 0) {
        double maxDiagonal = 0.;
        for (int j = 0; j < 6; j++) maxDiagonal = fmax(fabs(H[j][j]), maxDiagonal);
        lambda = 1e-5 * maxDiagonal;
        ni = 2;
        lmBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        __syncthreads();  // everyone is done with sOut / sTrial of the previous trial
        if (tid == 0) {
          double Hl[6][6], x[6];
          for (int a = 0; a < 6; a++)
            for (int c = 0; c < 6; c++) Hl[a][c] = H[a][c] + (a == c ? lambda : 0.0);
          for (int a = 0; a < 6; a++) x[a] = sX[a];
          const bool ok2 = ldlt6_solve(Hl, b, x);   // on failure x keeps the previous solution, like g2o's buffer
          for (int a = 0; a < 6; a++) sX[a] = x[a];
          sOk = ok2;
          SE3 tr;
          se3_exp_mul(x, est, tr);
          sTrial = tr;
        }
        __syncthreads();
        const SE3 trial = sTrial;
        double chi[1] = {0.0};
        for (int i = tid; i < cap; i += kPoThreads) {
          if (sLevel[i] != 0) continue;
          const EdgeIn e = load_edge(i);
          double ex, ey, pc[3];
          edge_error(e, trial, ex, ey, pc);
          sErr[i] = make_double2(ex, ey);
          const double c = ex * (e.info * ex) + ey * (e.info * ey);
          chi[0] += robust ? robust_rho0(c) : c;
        }
        block_sum<1>(chi, sRed, sOut);
        tempChi = sOut[0];
        if (!sOk) tempChi = DBL_MAX;
        rho = currentChi - tempChi;
        double scale = 0.;
        for (int j = 0; j < 6; j++) scale += sX[j] * (lambda * sX[j] + b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - pow((2 * rho - 1), 3.0);
          alpha = fmin(alpha, 2. / 3.);
          const double scaleFactor = fmax(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
          est = trial;
        } else {
          lambda *= ni;
          ni *= 2;
        }
        qmax++;
      } while (rho < 0 && qmax < 10);
      if (qmax == 10 || rho == 0) break;
      if ((iniChi - currentChi) * 1e3 < iniChi) lmBad++; else lmBad = 0;
      if (lmBad >= 3) break;
    }
    __syncthreads();

    // classify (Optimizer.cc:381-407): outliers are re-evaluated at the new pose, inliers keep the _error of the last pass
    int nb = 0;
    for (int i = tid; i < cap; i += kPoThreads) {
      const int lv = sLevel[i];
      if (lv == 2) continue;
      const EdgeIn e = load_edge(i);
      if (lv == 1) {
        double ex, ey, pc[3];
        edge_error(e, est, ex, ey, pc);
        sErr[i] = make_double2(ex, ey);
      }
      const double2 er = sErr[i];
      const float chi2 = (float)(er.x * (e.info * er.x) + er.y * (e.info * er.y));
      if (chi2 > 5.991f) { sLevel[i] = 1; nb++; } else sLevel[i] = 0;
    }
    nBadEdges = block_count(nb);
    if (it == 2) robust = false;
    if (nInitial < 10) break;  // optimizer.edges().size() < 10
  }

  for (int i = tid; i < cap; i += kPoThreads) A.outlier[(size_t)p * cap + i] = sLevel[i] == 1;
  if (tid == 0) {
    const Quat q = est.r;  // toRotationMatrix + float conversion (Converter::toCvMat)
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Tout[0] = (float)(1 - (tyy + tzz)); Tout[1] = (float)(txy - twz); Tout[2] = (float)(txz + twy); Tout[3] = (float)est.t[0];
    Tout[4] = (float)(txy + twz); Tout[5] = (float)(1 - (txx + tzz)); Tout[6] = (float)(tyz - twx); Tout[7] = (float)est.t[1];
    Tout[8] = (float)(txz - twy); Tout[9] = (float)(tyz + twx); Tout[10] = (float)(1 - (txx + tyy)); Tout[11] = (float)est.t[2];
    Tout[12] = 0.f; Tout[13] = 0.f; Tout[14] = 0.f; Tout[15] = 1.f;
    A.nInliers[p] = nInitial - nBadEdges;
  }
}

template <typename T>
int stage_pose_in(TempBuf<T>& d, const T*& ptr, size_t n, bool is_device, cudaStream_t s) {
  if (is_device) return PGB_OK;
  if (d.alloc(n)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpyAsync(d.p, ptr, n * sizeof(T), cudaMemcpyHostToDevice, s));
  ptr = d.p;
  return PGB_OK;
}

}  // namespace

}  // namespace pgb

using namespace pgb;

extern "C" int pgb_pose_optimization(int device, int n_frames, int cap, const float* Tcw_in, const float* kp_xy,
                                     const int32_t* kp_octave, const float* mp_xyz, const uint8_t* has_map_point,
                                     const int32_t* counts, const float* inv_level_sigma2, int nlevels, float fx, float fy,
                                     float cx, float cy, float* Tcw_out, uint8_t* outlier, int32_t* n_inliers,
                                     int is_device, void* stream) {
  if (n_frames < 0 || cap <= 0 || nlevels <= 0 || nlevels > 16 || !inv_level_sigma2)
    return fail(PGB_ERR_INVALID, "pgb_pose_optimization: invalid argument");
  if (n_frames == 0) return PGB_OK;
  if (!Tcw_in || !kp_xy || !kp_octave || !mp_xyz || !has_map_point || !counts || !Tcw_out || !outlier || !n_inliers)
    return fail(PGB_ERR_INVALID, "pgb_pose_optimization: null buffer");
  const size_t smem = (size_t)cap * 17 + 16;
  if (smem > 200 * 1024) return fail(PGB_ERR_CAPACITY, "cap %d needs %zu B of shared memory", cap, smem);
  if (use_device(device)) return PGB_ERR_CUDA;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)n_frames * cap;
  TempScope scope(device, s);
  if (scope.rc) return PGB_ERR_CUDA;
  TempBuf<float> dT, dXY, dX, dTo;
  TempBuf<int> dOct, dCnt, dNi, dErr;
  TempBuf<uint8_t> dHas, dOut;
  int rc = stage_pose_in(dT, Tcw_in, (size_t)n_frames * 16, is_device, s) | stage_pose_in(dXY, kp_xy, n * 2, is_device, s) |
           stage_pose_in(dOct, kp_octave, n, is_device, s) | stage_pose_in(dX, mp_xyz, n * 3, is_device, s) |
           stage_pose_in(dHas, has_map_point, n, is_device, s) | stage_pose_in(dCnt, counts, n_frames, is_device, s);
  if (rc || dErr.alloc(1)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemsetAsync(dErr.p, 0, sizeof(int), s));
  float* to = Tcw_out;
  uint8_t* out = outlier;
  int* ni = n_inliers;
  if (!is_device) {
    if (dTo.alloc((size_t)n_frames * 16) || dOut.alloc(n) || dNi.alloc(n_frames)) return PGB_ERR_CUDA;
    to = dTo.p; out = dOut.p; ni = dNi.p;
  }
  PoseArgs A;
  memset(&A, 0, sizeof A);
  A.cap = cap; A.nlevels = nlevels;
  A.fx = fx; A.fy = fy; A.cx = cx; A.cy = cy;
  const float deltaMono = (float)sqrt(5.991);  // Optimizer.cc:272: const float deltaMono = sqrt(5.991)
  A.delta = deltaMono;
  A.dsqr = A.delta * A.delta;
  for (int i = 0; i < nlevels; i++) A.invSigma2[i] = inv_level_sigma2[i];
  A.TcwIn = Tcw_in; A.kpXY = kp_xy; A.kpOctave = kp_octave; A.mpXYZ = mp_xyz; A.hasMp = has_map_point; A.counts = counts;
  A.TcwOut = to; A.outlier = out; A.nInliers = ni; A.err = dErr.p;
  static DynSmemLimit lim;
  if (int rc = lim.ensure(k_pose_optimization, smem)) return rc;
  k_pose_optimization<<<n_frames, kPoThreads, smem, s>>>(A);
  PGB_CHECK_LAUNCH();
  int e = 0;
  PGB_CUDA(cudaMemcpyAsync(&e, dErr.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (!is_device) {
    PGB_CUDA(cudaMemcpyAsync(Tcw_out, dTo.p, (size_t)n_frames * 16 * sizeof(float), cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaMemcpyAsync(outlier, dOut.p, n, cudaMemcpyDeviceToHost, s));
    PGB_CUDA(cudaMemcpyAsync(n_inliers, dNi.p, n_frames * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  PGB_CUDA(cudaStreamSynchronize(s));
  if (e) return fail(PGB_ERR_INVALID, "pgb_pose_optimization: keypoint octave outside [0, nlevels)");
  return PGB_OK;
}
