// K9/K10: IMU+GPS calibration on the device, and its C-ABI ("IMU + GPS calibration" section of pgb200.h).
//
// Reference (file:line in waiwnf/pilotguru): AccelerometerCalibrator src/calibration/velocity.cc:29-256,
// MergeTimeSeries / MakeInterpolationIntervals src/interpolation/align_time_series.cc:29-196, the window loop of
// src/fit_motion.cc:156-273, LBFGS++ thirdparty/LBFGS/LBFGS.h:79-182, SmoothTimeSeries src/slam/smoothing.cc:49-98.
//
// Device pipeline (all fp64, arithmetic from include/pgb200_imu_core.h, compiled with -fmad=false):
//   k_imu_sweep    one thread per GPS interval: the rotation sweep over its IMU sub-intervals -> GpsLocal
//   k_imu_chain    one thread per window: chains <= batch_size-1 GpsLocal records -> WinRec coefficients
//   k_imu_solve    one thread per window: the whole L-BFGS (<= 500 iterations) on the 9 unknowns
//   k_imu_speeds   one thread per (window, GPS interval): |v| at every IMU sub-interval with the fitted x (K10)
//   k_imu_average  one thread per merged IMU event: sum over the covering windows, in window order
// Host side: merged-event table and interpolation intervals are built once per recording / GPS series in O(N)
// (the reference rebuilds them per window).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/pgb200_imu_core.h"
#include "common.cuh"

namespace pgb {
using namespace pgbimu;

struct WinDesc {
  int s, e;          // GPS index range [s, e) of the window; its GPS intervals are refs s+1 .. e-1
  long long spOff;   // offset of the window's first sub-interval in the speeds array
};

__global__ void k_imu_sweep(int nRef, const int* __restrict__ ioff, const int* __restrict__ ivM,
                            const long long* __restrict__ ivDur, const int* __restrict__ mG, const int* __restrict__ mA,
                            const double* __restrict__ gyro, const double* __restrict__ acc, GpsLocal* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nRef) return;
  GpsLocal gl;
  SweepState ss;
  sweep_init(&ss, &gl);
  for (int k = ioff[r]; k < ioff[r + 1]; k++) {
    const int m = ivM[k];
    const int gi = mG[m], ai = mA[m];
    ImuStep st;
    st.wx = gyro[3 * (size_t)gi]; st.wy = gyro[3 * (size_t)gi + 1]; st.wz = gyro[3 * (size_t)gi + 2];
    st.ax = acc[3 * (size_t)ai]; st.ay = acc[3 * (size_t)ai + 1]; st.az = acc[3 * (size_t)ai + 2];
    st.dur_usec = ivDur[k];
    const double dt = sweep_step(&ss, st);
    sweep_accumulate(ss, dt, &gl);
  }
  sweep_finish(ss, &gl);
  out[r] = gl;
}

__global__ void k_imu_chain(int nWin, const WinDesc* __restrict__ win, const GpsLocal* __restrict__ loc,
                            const double* __restrict__ gpsV, int recStride, WinRec* __restrict__ rec,
                            long long* __restrict__ totalUsec) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nWin) return;
  WinState ws;
  win_init(&ws);
  long long tot = 0;
  int j = 0;
  for (int r = win[w].s + 1; r < win[w].e; r++, j++) {
    const GpsLocal gl = loc[r];
    WinRec wr;
    win_chain(&ws, gl, gpsV[r], &wr);
    rec[(size_t)w * recStride + j] = wr;
    tot += gl.dur;
  }
  totalUsec[w] = tot;
}

// One warp per window: lane l holds the partial sums of records l, l + 32, ...; the butterfly is the contract's
// (pgb200_imu_core.h: imu_eval).  Every lane ends with identical bits, so every lane can run the L-BFGS bookkeeping
// redundantly -- no broadcasts, no divergence.
struct WarpEval {
  const WinRec* rec;
  int n, lane;
  long long total;
  __device__ double operator()(const double* x, double* g) const {
    double acc[10];
    imu_eval_lane(rec, n, lane, x, acc);
#pragma unroll
    for (int k = PGB_EVAL_LANES / 2; k >= 1; k >>= 1) {
#pragma unroll
      for (int i = 0; i < 10; i++) acc[i] = acc[i] + __shfl_xor_sync(0xffffffffu, acc[i], k);
    }
    return imu_eval_finish(acc, total, g);
  }
};

__global__ void __launch_bounds__(32) k_imu_eval(const WinRec* __restrict__ rec, int n, long long total,
                                                 const double* __restrict__ x, double* __restrict__ out10) {
  double xx[9], g[9];
  for (int i = 0; i < 9; i++) xx[i] = x[i];
  WarpEval f{rec, n, (int)threadIdx.x, total};
  const double loss = f(xx, g);
  if (threadIdx.x == 0) {
    out10[0] = loss;
    for (int i = 0; i < 9; i++) out10[1 + i] = g[i];
  }
}

// The whole L-BFGS of a window (<= 500 iterations, ~1.2 evaluations each) by one warp; kSolveWarps windows per CTA.  The
// round-1 kernel ran one THREAD per window (720 threads on 148 SMs for BASELINE configs[3]): 60 of the fit's 80 ms
// (profiles/r02_calibration_phases.txt); per evaluation it walked the window's 39 records sequentially.
constexpr int kSolveWarps = 4;
__global__ void __launch_bounds__(kSolveWarps * 32) k_imu_solve(int nWin, const WinDesc* __restrict__ win, int recStride,
                                                               const WinRec* __restrict__ rec,
                                                               const long long* __restrict__ totalUsec, int maxIter,
                                                               double epsilon, int useX0, double* __restrict__ xOut,
                                                               double* __restrict__ fxOut, int* __restrict__ itOut,
                                                               int* __restrict__ evalOut) {
  __shared__ double sWs[kSolveWarps][PGB_LBFGS_WS_DOUBLES];  // the correction pairs S, Y of each window
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * kSolveWarps + warp;
  if (w >= nWin) return;  // warp-uniform
  double x[9];
  for (int i = 0; i < 9; i++) x[i] = useX0 ? xOut[(size_t)w * 9 + i] : 0.0;
  const int n = win[w].e - win[w].s - 1;
  double fx = 0.0;
  int it = 0, ne = 0;
  if (n > 0 && totalUsec[w] > 0) {
    WarpEval f{rec + (size_t)w * recStride, n, lane, totalUsec[w]};
    LbfgsParam P = lbfgs_default();
    P.epsilon = epsilon;
    P.max_iterations = maxIter;
    it = lbfgs_minimize9(f, x, &fx, P, &ne, sWs[warp]);
  }
  if (lane == 0) {
    for (int i = 0; i < 9; i++) xOut[(size_t)w * 9 + i] = x[i];
    fxOut[w] = fx;
    itOut[w] = it;
    if (evalOut) evalOut[w] = ne;
  }
}

// K10: per sub-interval speed (and optionally orientation / velocity) with the fitted parameters.
__global__ void k_imu_speeds(int nWin, int maxRefs, const WinDesc* __restrict__ win, int recStride,
                             const WinRec* __restrict__ rec, const double* __restrict__ xAll,
                             const int* __restrict__ ioff, const int* __restrict__ ivM,
                             const long long* __restrict__ ivDur, const int* __restrict__ mG, const int* __restrict__ mA,
                             const double* __restrict__ gyro, const double* __restrict__ acc,
                             double* __restrict__ speeds, double* __restrict__ quat, double* __restrict__ vel,
                             double minVel, double* __restrict__ fwdPart) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nWin * maxRefs) return;
  const int w = t / maxRefs, j = t - w * maxRefs;
  const int r = win[w].s + 1 + j;
  if (fwdPart) { fwdPart[4 * (size_t)t] = 0.0; fwdPart[4 * (size_t)t + 1] = 0.0; fwdPart[4 * (size_t)t + 2] = 0.0; fwdPart[4 * (size_t)t + 3] = 1.0; }
  if (r >= win[w].e) return;
  V3 fsum = v3(0.0, 0.0, 0.0);
  double minW = 1.0;
  const int kEnd = ioff[win[w].e];  // one past the window's last sub-interval
  const double* x = xAll + (size_t)w * 9;
  const V3 g = v3(x[0], x[1], x[2]), h = v3(x[3], x[4], x[5]), v0 = v3(x[6], x[7], x[8]);
  const WinRec wr = rec[(size_t)w * recStride + j];
  const V3 Vr = add(add(v0, wr.Sa), add(mv(wr.SE, h), scale(g, wr.St)));
  GpsLocal gl;
  SweepState ss;
  sweep_init(&ss, &gl);
  const long long base = win[w].spOff - ioff[win[w].s + 1];
  for (int k = ioff[r]; k < ioff[r + 1]; k++) {
    const int m = ivM[k];
    const int gi = mG[m], ai = mA[m];
    ImuStep st;
    st.wx = gyro[3 * (size_t)gi]; st.wy = gyro[3 * (size_t)gi + 1]; st.wz = gyro[3 * (size_t)gi + 2];
    st.ax = acc[3 * (size_t)ai]; st.ay = acc[3 * (size_t)ai + 1]; st.az = acc[3 * (size_t)ai + 2];
    st.dur_usec = ivDur[k];
    sweep_step(&ss, st);
    const V3 loc = add(ss.pa, mv(ss.pR, h));
    const V3 v = add(add(Vr, mv(wr.RQ, loc)), scale(g, (double)ss.tau * 1e-6));
    speeds[base + k] = norm3(v);
    if (vel) { vel[3 * (base + k)] = v.x; vel[3 * (base + k) + 1] = v.y; vel[3 * (base + k) + 2] = v.z; }
    if (quat) {
      const Q4 q = qmul(wr.Q, ss.l);
      quat[4 * (base + k)] = q.w; quat[4 * (base + k) + 1] = q.x; quat[4 * (base + k) + 2] = q.y; quat[4 * (base + k) + 3] = q.z;
    }
    // forward-axis evidence (fit_motion.cc:223-248): one trajectory point per merged event -- the LAST sub-interval
    // carrying it inside the window (velocity.cc:236-250)
    if (fwdPart && (k + 1 >= kEnd || ivM[k + 1] != m)) {
      const Q4 q = qmul(wr.Q, ss.l);
      minW = fmin(minW, fabs(q.w));
      if (norm3(v) >= minVel) {
        Q4 qc; qc.w = q.w; qc.x = -q.x; qc.y = -q.y; qc.z = -q.z;
        fsum = add(fsum, qrot(qc, v));
      }
    }
  }
  if (fwdPart) { fwdPart[4 * (size_t)t] = fsum.x; fwdPart[4 * (size_t)t + 1] = fsum.y; fwdPart[4 * (size_t)t + 2] = fsum.z; fwdPart[4 * (size_t)t + 3] = minW; }
}

// MakeInterpolationIntervals (align_time_series.cc:155-196) on the device: the host only finds, per GPS fix r, the first
// merged IMU event after it (hiIdx[r], 3600 bisections for an hour of GPS) and the prefix sums ioff; one thread per
// GPS interval then writes its sub-intervals: every merged event k in (gps[r-1], gps[r]] ends a sub-interval that
// starts at the later of the previous event and gps[r-1], and the stretch from the last such event to gps[r] is a
// partial sub-interval carrying the NEXT event's samples (:187-193).  (Round 1 built the three arrays in a host loop
// and uploaded 29 MB per fit: 8 of the fit's 80 ms.)
__global__ void k_imu_build_intervals(int nRef, long long nEvents, const long long* __restrict__ gpsT,
                                      const long long* __restrict__ evT, const int* __restrict__ hiIdx,
                                      const int* __restrict__ ioff, int* __restrict__ ivM, int* __restrict__ ivRef,
                                      long long* __restrict__ ivDur) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < 1 || r >= nRef) return;
  const int lo = hiIdx[r - 1], hi = hiIdx[r];
  int out = ioff[r];
  long long latest = gpsT[r - 1];
  for (int k = lo; k < hi; k++) {
    const long long ts = evT[k];
    if (k > 0) { ivM[out] = k; ivRef[out] = r; ivDur[out] = ts - latest; out++; }
    latest = ts;
  }
  if (hi > 0 && hi < nEvents && gpsT[r] > latest) { ivM[out] = hi; ivRef[out] = r; ivDur[out] = gpsT[r] - latest; }
}

// Per merged event: first / last sub-interval carrying it (the run is contiguous in the interval list), -1 = none.
__global__ void k_imu_event_runs(int kLo, int kHi, int mLo, const int* __restrict__ ivM, int* __restrict__ firstIv,
                                 int* __restrict__ lastIv) {
  const int k = kLo + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kHi) return;
  const int m = ivM[k];
  if (k == kLo || ivM[k - 1] != m) firstIv[m - mLo] = k;
  if (k + 1 == kHi || ivM[k + 1] != m) lastIv[m - mLo] = k;
}

// GetPrincipalRotationAxes (rotation.cc:16-57): one thread per integration interval multiplies the rotation
// quaternions of its gyro samples (the interval boundaries are an integer prefix computed on the host).
__global__ void k_rot_intervals(int nIv, const int* __restrict__ ivStart, const double* __restrict__ gyro,
                                const long long* __restrict__ gyroT, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nIv) return;
  Q4 q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
  for (int k = ivStart[i]; k < ivStart[i + 1]; k++) {
    const double dt = (double)(gyroT[k] - gyroT[k - 1]) * 1e-6;
    q = qmul(q, rotation_motion_to_quaternion(gyro[3 * (size_t)k], gyro[3 * (size_t)k + 1], gyro[3 * (size_t)k + 2], dt));
  }
  out[3 * (size_t)i] = q.x; out[3 * (size_t)i + 1] = q.y; out[3 * (size_t)i + 2] = q.z;
}

// GetAngularVelocitiesAroundAxisDirect (rotation.cc:103-119)
__global__ void k_axis_project(long long n, const double* __restrict__ gyro, double ax, double ay, double az, double norm,
                               double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = ((gyro[3 * i] * ax + gyro[3 * i + 1] * ay) + gyro[3 * i + 2] * az) / norm;
}

// Per merged event: the sub-intervals carrying it are a contiguous run [ka, kb] of the global interval list.
// A window contributes the value of the LAST of them it contains (IntegrateTrajectory overwrites the partial result
// of an event split by a GPS boundary, velocity.cc:236-250); windows are added in ascending order
// (std::accumulate over push_back order, fit_motion.cc:218-221, :256-258).
__global__ void k_imu_average(int mLo, int mCount, const int* __restrict__ firstIv, const int* __restrict__ lastIv,
                              const int* __restrict__ ivRef, int nWin, int firstWin, int batch, int step, int nGps,
                              const WinDesc* __restrict__ win, const int* __restrict__ ioff,
                              const double* __restrict__ speeds, double* __restrict__ sum, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mCount) return;
  const int ka = firstIv[i], kb = lastIv[i];
  double s = 0.0;
  int c = 0;
  if (ka >= 0) {
    const int ra = ivRef[ka], rb = ivRef[kb];
    int wlo = (ra - batch) / step - 1, whi = rb / step + 1;
    wlo = max(wlo, firstWin);
    whi = min(whi, firstWin + nWin - 1);
    for (int wg = wlo; wg <= whi; wg++) {
      const WinDesc wd = win[wg - firstWin];
      const int lo = ioff[wd.s + 1], hi = ioff[wd.e];  // the window's sub-intervals [lo, hi)
      const int k = min(kb, hi - 1);
      if (hi > lo && k >= max(ka, lo)) {
        s = s + speeds[wd.spOff + (k - lo)];
        c++;
      }
    }
  }
  sum[mLo + i] = s;
  cnt[mLo + i] = c;
}

// SmoothTimeSeries (smoothing.cc:56-98): one thread per target; the reference's moving window bounds are
// monotone in the target time, so each thread finds its own by bisection over the same predicates.
__global__ void k_smooth(const double* __restrict__ v, const double* __restrict__ t, long long n,
                         const double* __restrict__ tt, long long nt, double sigma, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const double target = tt[i];
  // left = number of leading indices j (j+1 < n) with target - t[j+1] > 3 sigma  (monotone in j)
  long long lo = 0, hi = n - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if ((target - t[mid + 1]) > 3 * sigma) lo = mid + 1; else hi = mid;
  }
  const long long left = lo;
  // right = first index j with !(t[j] - target < 3 sigma) or n-1
  lo = 0; hi = n - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if ((t[mid] - target) < 3 * sigma) lo = mid + 1; else hi = mid;
  }
  const long long right = lo;
  const double sqrt2 = sqrt(2.0);
  double prev = 0.0, acc = 0.0;
  for (long long k = left; k < right; k++) {
    const double midp = (t[k] + t[k + 1]) / 2.0;
    const double cdf = 0.5 * (1.0 + erf((midp - target) / (sqrt2 * sigma)));
    acc = acc + v[k] * (cdf - prev);
    prev = cdf;
  }
  acc = acc + v[right] * (1.0 - prev);
  out[i] = acc;
}

// TimeSeries<double>::TimeAveragedValue (include/interpolation/time_series.hpp:129-189) for every frame interval
// (t[i-1], t[i]] of annotate_frames.cc:59-72: one thread per frame; the reference's forward linear scans
// (MostRecentPreviousValue, :103-125) become bisections for the same index (last event with time <= query).
// status: 0 = not covered by the series (is_valid = false), 1 = valid, 2 = the reference would CHECK-fail
// (LinearInterpolate needs an event after the query's end: :210-213).
__device__ __forceinline__ long long last_not_after(const long long* __restrict__ t, long long n, long long q) {
  long long lo = 0, hi = n;  // first index with t > q
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (t[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}
__device__ __forceinline__ double interval_sec(long long a, long long b) { return (double)(b - a) * 1e-6; }
__device__ __forceinline__ double lerp_at(const double* __restrict__ v, const long long* __restrict__ t, long long l,
                                          long long r, long long q) {
  const double left = interval_sec(t[l], q), right = interval_sec(q, t[r]), total = interval_sec(t[l], t[r]);
  return (left / total) * v[r] + (right / total) * v[l];
}
__global__ void k_time_average(const double* __restrict__ v, const long long* __restrict__ t, long long n,
                               const long long* __restrict__ ft, long long nFrames, double* __restrict__ out,
                               unsigned char* __restrict__ status) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= nFrames) return;
  const long long start = ft[i - 1], end = ft[i];
  if (start < t[0] || end > t[n - 1]) { out[i - 1] = nan(""); status[i - 1] = 0; return; }
  const long long si = last_not_after(t, n, start), ei = last_not_after(t, n, end);
  if (ei + 1 >= n || si + 1 >= n) { out[i - 1] = nan(""); status[i - 1] = 2; return; }
  double total = 0.0;
  for (long long k = si + 1; k < ei; k++) total += (interval_sec(t[k], t[k + 1]) * 0.5 * (v[k] + v[k + 1]));
  const double lv = lerp_at(v, t, si, si + 1, start), rv = lerp_at(v, t, ei, ei + 1, end);
  if (si == ei) {
    total += (lv + rv) * 0.5 * interval_sec(start, end);
  } else {
    total += (lv + v[si + 1]) * 0.5 * interval_sec(start, t[si + 1]);
    total += (v[ei] + rv) * 0.5 * interval_sec(t[ei], end);
  }
  out[i - 1] = total / interval_sec(start, end);
  status[i - 1] = 1;
}

}  // namespace pgb

using namespace pgb;

struct pgb_imu {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  // recording
  std::vector<int64_t> gyroT, accT, mergedT;
  std::vector<int> mG, mA;
  DevBuf<double> dGyro, dAcc;
  DevBuf<int> dMG, dMA;
  // prepared GPS series
  std::vector<int64_t> gpsT;
  std::vector<int> ioff, hiIdx;  // per GPS fix: offset of its first sub-interval; first merged event after it
  DevBuf<long long> dMergedT, dGpsT;
  DevBuf<int> dHiIdx;
  std::vector<WinDesc> win;
  int firstWin = 0, batch = 0, step = 0, maxRefs = 0;
  long long spTotal = 0;
  DevBuf<int> dIoff, dIvM, dIvRef, dFirstIv, dLastIv, dIt, dNe, dCnt;
  DevBuf<long long> dIvDur, dTotal;
  DevBuf<double> dGpsV, dX, dFx, dSpeeds, dQuat, dVel, dSum, dOut10, dFwd;
  DevBuf<GpsLocal> dLoc;
  DevBuf<WinRec> dRec;
  DevBuf<WinDesc> dWin;
  std::vector<long long> hTotal;
  bool windowReady = false;
  // CUDA events around the kernels of the last fit (pgb_imu_last_kernel_ms): 0/1 sweep, 2/3 solve, 4/5 speeds
  cudaEvent_t ev[6] = {};
  bool evValid = false;
};

namespace {

// MergeTimeSeries for the two IMU components (align_time_series.cc:29-113): every emitted event holds, per
// component, the index of its latest sample at or before the event time.
int merge_two(const std::vector<int64_t>& a, const std::vector<int64_t>& b, std::vector<int>& ia, std::vector<int>& ib,
              std::vector<int64_t>& t) {
  ia.clear(); ib.clear(); t.clear();
  const int64_t start = std::max(a.front(), b.front()), end = std::min(a.back(), b.back());
  if (end < start) return 0;
  auto first = [&](const std::vector<int64_t>& c) {
    size_t i = std::lower_bound(c.begin(), c.end(), start) - c.begin();
    return c[i] > start ? i - 1 : i;
  };
  size_t i = first(a), j = first(b);
  for (;;) {
    ia.push_back((int)i); ib.push_back((int)j);
    t.push_back(std::max(a[i], b[j]));
    if (i + 1 >= a.size() || j + 1 >= b.size()) break;
    const int64_t next = std::min(a[i + 1], b[j + 1]);
    const bool adva = a[i + 1] == next, advb = b[j + 1] == next;
    if (adva) i++;
    if (advb) j++;
  }
  return (int)t.size();
}

int check_increasing(const int64_t* t, size_t n, const char* what) {
  for (size_t i = 0; i + 1 < n; i++)
    if (!(t[i] < t[i + 1])) return fail(PGB_ERR_INVALID, "%s timestamps must be strictly increasing (index %zu)", what, i);
  return PGB_OK;
}

// Host part of MakeInterpolationIntervals (align_time_series.cc:155-196): per GPS fix r the index of the first merged
// event after it and the number of sub-intervals of GPS interval r (k_imu_build_intervals writes them).
void count_intervals(const std::vector<int64_t>& ref, const std::vector<int64_t>& interp, std::vector<int>& ioff,
                     std::vector<int>& hiIdx) {
  const size_t R = ref.size(), N = interp.size();
  ioff.assign(R + 1, 0); hiIdx.assign(R, 0);
  for (size_t r = 0; r < R; r++) hiIdx[r] = (int)(std::upper_bound(interp.begin(), interp.end(), ref[r]) - interp.begin());
  int total = 0;
  for (size_t r = 0; r < R; r++) {
    ioff[r] = total;
    if (r == 0) continue;
    const int lo = hiIdx[r - 1], hi = hiIdx[r];
    int n = hi - lo - (lo == 0 && hi > 0 ? 1 : 0);  // event 0 never ends a sub-interval (k > 0, :176)
    const int64_t latest = hi > lo ? interp[hi - 1] : ref[r - 1];
    if (hi > 0 && (size_t)hi < N && ref[r] > latest) n++;
    total += n;
  }
  ioff[R] = total;
}

// PGB_IMU_TIMING=1: wall time of every phase of a fit (stream synchronised at the marks) on stderr -- the breakdown
// profiles/r02_calibration_phases.txt was taken with it.
struct PhaseTimer {
  bool on;
  cudaStream_t s;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(cudaStream_t st) : on(getenv("PGB_IMU_TIMING") != nullptr), s(st), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(s);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[pgb_imu] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

template <typename T>
int upload(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t s) {
  if (d.n < h.size() || !d.p) { if (d.alloc(h.size())) return PGB_ERR_CUDA; }
  if (!h.empty()) PGB_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return PGB_OK;
}
template <typename T>
int ensure(DevBuf<T>& d, size_t n) {
  if (d.n < n || !d.p) return d.alloc(n);
  return PGB_OK;
}

// Builds intervals for the GPS series, the window list, and runs sweep + chain.
int prepare(pgb_imu* o, const double* gps_v, const int64_t* gps_t, int n_gps, int batch, int step, int first_window,
            int n_windows) {
  NvtxRange range("pgb:imu:intervals+sweep+chain");
  if (n_gps <= 0 || !gps_v || !gps_t) return fail(PGB_ERR_INVALID, "empty GPS series");
  int rc = check_increasing(gps_t, n_gps, "GPS");
  if (rc) return rc;
  o->windowReady = false;
  PhaseTimer pt(o->stream);
  o->gpsT.assign(gps_t, gps_t + n_gps);
  count_intervals(o->gpsT, o->mergedT, o->ioff, o->hiIdx);
  pt.mark("host: interval counts");
  const int allWin = (n_gps + step - 1) / step;
  if (first_window < 0 || first_window > allWin) return fail(PGB_ERR_INVALID, "first_window out of range");
  if (n_windows < 0 || first_window + n_windows > allWin) n_windows = allWin - first_window;
  o->win.clear();
  o->firstWin = first_window; o->batch = batch; o->step = step;
  long long sp = 0;
  int maxRefs = 1;
  for (int w = first_window; w < first_window + n_windows; w++) {
    WinDesc d;
    d.s = w * step;
    d.e = std::min(d.s + batch, n_gps);
    d.spOff = sp;
    sp += o->ioff[d.e] - o->ioff[std::min(d.s + 1, n_gps)];
    maxRefs = std::max(maxRefs, d.e - d.s - 1);
    o->win.push_back(d);
  }
  o->spTotal = sp;
  o->maxRefs = maxRefs;
  cudaStream_t s = o->stream;
  std::vector<double> gv(gps_v, gps_v + n_gps);
  std::vector<long long> gt(gps_t, gps_t + n_gps);
  const size_t nIv = (size_t)o->ioff[n_gps];
  if (upload(o->dIoff, o->ioff, s) || upload(o->dHiIdx, o->hiIdx, s) || upload(o->dGpsT, gt, s) || upload(o->dGpsV, gv, s) ||
      upload(o->dWin, o->win, s) || ensure(o->dIvM, nIv + 1) || ensure(o->dIvRef, nIv + 1) || ensure(o->dIvDur, nIv + 1))
    return PGB_ERR_CUDA;
  k_imu_build_intervals<<<(n_gps + 63) / 64, 64, 0, s>>>(n_gps, (long long)o->mergedT.size(), o->dGpsT.p, o->dMergedT.p, o->dHiIdx.p,
                                                        o->dIoff.p, o->dIvM.p, o->dIvRef.p, o->dIvDur.p);
  PGB_CHECK_LAUNCH();
  const size_t nW = o->win.size();
  if (ensure(o->dLoc, n_gps) || ensure(o->dRec, nW * maxRefs) || ensure(o->dTotal, nW) || ensure(o->dX, nW * 9) ||
      ensure(o->dFx, nW) || ensure(o->dIt, nW) || ensure(o->dNe, nW) || ensure(o->dOut10, 16))
    return PGB_ERR_CUDA;
  if (nW == 0) { PGB_CUDA(cudaStreamSynchronize(s)); return PGB_OK; }
  pt.mark("uploads + k_imu_build_intervals");
  o->evValid = false;
  for (auto& e : o->ev)
    if (!e) PGB_CUDA(cudaEventCreate(&e));
  PGB_CUDA(cudaEventRecord(o->ev[0], s));
  k_imu_sweep<<<(n_gps + 63) / 64, 64, 0, s>>>(n_gps, o->dIoff.p, o->dIvM.p, o->dIvDur.p, o->dMG.p, o->dMA.p,
                                                o->dGyro.p, o->dAcc.p, o->dLoc.p);
  PGB_CHECK_LAUNCH();
  PGB_CUDA(cudaEventRecord(o->ev[1], s));
  pt.mark("k_imu_sweep");
  k_imu_chain<<<((int)nW + 63) / 64, 64, 0, s>>>((int)nW, o->dWin.p, o->dLoc.p, o->dGpsV.p, maxRefs, o->dRec.p, o->dTotal.p);
  PGB_CHECK_LAUNCH();
  o->hTotal.resize(nW);
  PGB_CUDA(cudaMemcpyAsync(o->hTotal.data(), o->dTotal.p, nW * sizeof(long long), cudaMemcpyDeviceToHost, s));
  PGB_CUDA(cudaStreamSynchronize(s));  // gv, gt and the host vectors were sources of async copies
  pt.mark("k_imu_chain");
  o->windowReady = true;
  return PGB_OK;
}

int solve(pgb_imu* o, int maxIter, double eps, int useX0) {
  const int nW = (int)o->win.size();
  if (nW == 0) return PGB_OK;
  NvtxRange range("pgb:imu:lbfgs_solve");
  k_imu_solve<<<(nW + kSolveWarps - 1) / kSolveWarps, kSolveWarps * 32, 0, o->stream>>>(nW, o->dWin.p, o->maxRefs, o->dRec.p, o->dTotal.p, maxIter, eps, useX0,
                                                    o->dX.p, o->dFx.p, o->dIt.p, o->dNe.p);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

int speeds(pgb_imu* o, bool full, bool fwd = false, double minVel = 0.0) {
  const int nW = (int)o->win.size();
  if (nW == 0 || o->spTotal == 0) return PGB_OK;
  NvtxRange range("pgb:imu:speeds");
  if (ensure(o->dSpeeds, o->spTotal)) return PGB_ERR_CUDA;
  if (fwd && ensure(o->dFwd, 4 * (size_t)nW * o->maxRefs)) return PGB_ERR_CUDA;
  if (full && (ensure(o->dQuat, 4 * o->spTotal) || ensure(o->dVel, 3 * o->spTotal))) return PGB_ERR_CUDA;
  const int n = nW * o->maxRefs;
  k_imu_speeds<<<(n + 63) / 64, 64, 0, o->stream>>>(nW, o->maxRefs, o->dWin.p, o->maxRefs, o->dRec.p, o->dX.p, o->dIoff.p,
                                                    o->dIvM.p, o->dIvDur.p, o->dMG.p, o->dMA.p, o->dGyro.p, o->dAcc.p,
                                                    o->dSpeeds.p, full ? o->dQuat.p : nullptr, full ? o->dVel.p : nullptr,
                                                    minVel, fwd ? o->dFwd.p : nullptr);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

}  // namespace

extern "C" {

pgb_imu* pgb_imu_create(int device, const double* gyro_xyz, const int64_t* gyro_t, size_t n_gyro, const double* acc_xyz,
                        const int64_t* acc_t, size_t n_acc, void* stream) {
  if (!gyro_xyz || !gyro_t || !acc_xyz || !acc_t || n_gyro == 0 || n_acc == 0) {  // CHECK(!component->empty())
    fail(PGB_ERR_INVALID, "pgb_imu_create: empty sensor series");
    return nullptr;
  }
  if (check_increasing(gyro_t, n_gyro, "gyro") || check_increasing(acc_t, n_acc, "accelerometer")) return nullptr;
  if (use_device(device)) return nullptr;
  pgb_imu* o = new pgb_imu;
  o->device = device;
  o->gyroT.assign(gyro_t, gyro_t + n_gyro);
  o->accT.assign(acc_t, acc_t + n_acc);
  merge_two(o->gyroT, o->accT, o->mG, o->mA, o->mergedT);
  if (o->mergedT.empty()) {
    fail(PGB_ERR_INVALID, "gyro and accelerometer series do not overlap in time");
    delete o;
    return nullptr;
  }
  if (stream) o->stream = (cudaStream_t)stream;
  else {
    if (cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking) != cudaSuccess) { fail(PGB_ERR_CUDA, "cudaStreamCreate failed"); delete o; return nullptr; }
    o->ownStream = true;
  }
  auto bad = [&]() -> pgb_imu* { pgb_imu_destroy(o); return nullptr; };
  if (o->dGyro.alloc(3 * n_gyro) || o->dAcc.alloc(3 * n_acc)) return bad();
  if (cudaMemcpyAsync(o->dGyro.p, gyro_xyz, 3 * n_gyro * sizeof(double), cudaMemcpyHostToDevice, o->stream) != cudaSuccess ||
      cudaMemcpyAsync(o->dAcc.p, acc_xyz, 3 * n_acc * sizeof(double), cudaMemcpyHostToDevice, o->stream) != cudaSuccess) {
    fail(PGB_ERR_CUDA, "H2D copy of the sensor series failed");
    return bad();
  }
  {
    std::vector<long long> mt(o->mergedT.begin(), o->mergedT.end());
    if (upload(o->dMG, o->mG, o->stream) || upload(o->dMA, o->mA, o->stream) || upload(o->dMergedT, mt, o->stream)) return bad();
    if (cudaStreamSynchronize(o->stream) != cudaSuccess) { fail(PGB_ERR_CUDA, "stream sync failed"); return bad(); }
  }
  if (cudaStreamSynchronize(o->stream) != cudaSuccess) { fail(PGB_ERR_CUDA, "stream sync failed"); return bad(); }
  return o;
}

int pgb_imu_last_kernel_ms(pgb_imu* o, float* sweep_ms, float* solve_ms, float* speeds_ms, int64_t* n_intervals) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  if (!o->evValid) return fail(PGB_ERR_INVALID, "pgb_imu_last_kernel_ms: no completed pgb_imu_fit_windows call");
  PGB_CUDA(cudaSetDevice(o->device));
  PGB_CUDA(cudaEventSynchronize(o->ev[5]));
  float a = 0, b = 0, c = 0;
  PGB_CUDA(cudaEventElapsedTime(&a, o->ev[0], o->ev[1]));
  PGB_CUDA(cudaEventElapsedTime(&b, o->ev[2], o->ev[3]));
  PGB_CUDA(cudaEventElapsedTime(&c, o->ev[4], o->ev[5]));
  if (sweep_ms) *sweep_ms = a;
  if (solve_ms) *solve_ms = b;
  if (speeds_ms) *speeds_ms = c;
  if (n_intervals) *n_intervals = o->ioff.empty() ? 0 : (int64_t)o->ioff.back();
  return PGB_OK;
}

void pgb_imu_destroy(pgb_imu* o) {
  if (!o) return;
  cudaSetDevice(o->device);
  if (o->stream) cudaStreamSynchronize(o->stream);
  for (auto& e : o->ev)
    if (e) cudaEventDestroy(e);
  if (o->ownStream && o->stream) cudaStreamDestroy(o->stream);
  delete o;
}

int64_t pgb_imu_merged_count(const pgb_imu* o) { return o ? (int64_t)o->mergedT.size() : PGB_ERR_INVALID; }

int pgb_imu_merged_events(const pgb_imu* o, int64_t* t_usec, int64_t* gyro_idx, int64_t* acc_idx) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  for (size_t i = 0; i < o->mergedT.size(); i++) {
    if (t_usec) t_usec[i] = o->mergedT[i];
    if (gyro_idx) gyro_idx[i] = o->mG[i];
    if (acc_idx) acc_idx[i] = o->mA[i];
  }
  return PGB_OK;
}

int pgb_imu_set_window(pgb_imu* o, const double* gps_v, const int64_t* gps_t, int n_gps) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  PGB_CUDA(cudaSetDevice(o->device));
  return prepare(o, gps_v, gps_t, n_gps, n_gps, n_gps, 0, 1);
}

int64_t pgb_imu_window_intervals(const pgb_imu* o) {
  if (!o || !o->windowReady || o->win.empty()) return 0;
  return (int64_t)o->spTotal;
}

int pgb_imu_eval(pgb_imu* o, const double x[9], double* loss, double grad[9]) {
  if (!o || !x || !loss || !grad) return fail(PGB_ERR_INVALID, "null argument");  // CHECK_NOTNULL(gradient)
  if (!o->windowReady || o->win.size() != 1) return fail(PGB_ERR_INVALID, "pgb_imu_eval: call pgb_imu_set_window first");
  PGB_CUDA(cudaSetDevice(o->device));
  const int n = o->win[0].e - o->win[0].s - 1;
  PGB_CUDA(cudaMemcpyAsync(o->dX.p, x, 9 * sizeof(double), cudaMemcpyHostToDevice, o->stream));
  k_imu_eval<<<1, 32, 0, o->stream>>>(o->dRec.p, n, o->hTotal[0], o->dX.p, o->dOut10.p);
  PGB_CHECK_LAUNCH();
  double out[10];
  PGB_CUDA(cudaMemcpyAsync(out, o->dOut10.p, sizeof out, cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  *loss = out[0];
  for (int i = 0; i < 9; i++) grad[i] = out[1 + i];
  return PGB_OK;
}

int pgb_imu_minimize(pgb_imu* o, double x[9], double* fx, int* n_iter, int max_iterations, double epsilon) {
  if (!o || !x || !fx) return fail(PGB_ERR_INVALID, "null argument");
  if (!o->windowReady || o->win.size() != 1) return fail(PGB_ERR_INVALID, "pgb_imu_minimize: call pgb_imu_set_window first");
  if (max_iterations < 0 || !(epsilon > 0)) return fail(PGB_ERR_INVALID, "bad L-BFGS parameters");  // LBFGSParam::check_param
  PGB_CUDA(cudaSetDevice(o->device));
  PGB_CUDA(cudaMemcpyAsync(o->dX.p, x, 9 * sizeof(double), cudaMemcpyHostToDevice, o->stream));
  int rc = solve(o, max_iterations, epsilon, 1);
  if (rc) return rc;
  int it = 0;
  PGB_CUDA(cudaMemcpyAsync(x, o->dX.p, 9 * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaMemcpyAsync(fx, o->dFx.p, sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaMemcpyAsync(&it, o->dIt.p, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  if (it < 0) return fail(PGB_ERR_NUMERIC, "the line search step left [min_step, max_step]");
  if (n_iter) *n_iter = it;
  return PGB_OK;
}

int pgb_imu_integrate(pgb_imu* o, const double x[9], int64_t cap, int64_t* merged_idx, double* speed, double* quat_wxyz,
                      double* vel_xyz, int64_t* duration_usec, int64_t* n_out) {
  if (!o || !x || !n_out) return fail(PGB_ERR_INVALID, "null argument");
  if (!o->windowReady || o->win.size() != 1) return fail(PGB_ERR_INVALID, "pgb_imu_integrate: call pgb_imu_set_window first");
  PGB_CUDA(cudaSetDevice(o->device));
  PGB_CUDA(cudaMemcpyAsync(o->dX.p, x, 9 * sizeof(double), cudaMemcpyHostToDevice, o->stream));
  int rc = speeds(o, true);
  if (rc) return rc;
  const long long n = o->spTotal;
  std::vector<double> sp(n), q(4 * n), v(3 * n);
  if (n) {
    PGB_CUDA(cudaMemcpyAsync(sp.data(), o->dSpeeds.p, n * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    PGB_CUDA(cudaMemcpyAsync(q.data(), o->dQuat.p, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    PGB_CUDA(cudaMemcpyAsync(v.data(), o->dVel.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  }
  // the window's sub-interval table (built on the device): merged event and span of each
  const int lo = o->ioff[std::min(o->win[0].s + 1, (int)o->gpsT.size())];
  std::vector<int> ivM(n);
  std::vector<long long> ivDur(n);
  if (n) {
    PGB_CUDA(cudaMemcpyAsync(ivM.data(), o->dIvM.p + lo, n * sizeof(int), cudaMemcpyDeviceToHost, o->stream));
    PGB_CUDA(cudaMemcpyAsync(ivDur.data(), o->dIvDur.p + lo, n * sizeof(long long), cudaMemcpyDeviceToHost, o->stream));
  }
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  // fold sub-intervals of the same merged event: later overwrites, durations add (velocity.cc:236-250)
  int64_t k = 0;
  for (long long i = 0; i < n; i++) {
    const int m = ivM[i];
    const bool cont = i > 0 && ivM[i - 1] == m;
    if (!cont) {
      if (k >= cap) return fail(PGB_ERR_CAPACITY, "trajectory has more than %lld points", (long long)cap);
      k++;
      if (duration_usec) duration_usec[k - 1] = 0;
    }
    if (merged_idx) merged_idx[k - 1] = m;
    if (speed) speed[k - 1] = sp[i];
    if (quat_wxyz) for (int c = 0; c < 4; c++) quat_wxyz[4 * (k - 1) + c] = q[4 * i + c];
    if (vel_xyz) for (int c = 0; c < 3; c++) vel_xyz[3 * (k - 1) + c] = v[3 * i + c];
    if (duration_usec) duration_usec[k - 1] += ivDur[i];
  }
  *n_out = k;
  return PGB_OK;
}

int pgb_imu_num_windows(int n_gps, int shift_step) {
  if (n_gps <= 0 || shift_step <= 0) return 0;
  return (n_gps + shift_step - 1) / shift_step;
}

int pgb_imu_fit_windows(pgb_imu* o, const double* gps_v, const int64_t* gps_t, int n_gps, int batch_size, int shift_step,
                        int max_iterations, double epsilon, int first_window, int n_windows, double* speed_sum,
                        int32_t* speed_cnt, double* x_out, double* fx_out, int32_t* iters_out) {
  return pgb_imu_fit_windows_fwd(o, gps_v, gps_t, n_gps, batch_size, shift_step, max_iterations, epsilon, first_window,
                                 n_windows, speed_sum, speed_cnt, x_out, fx_out, iters_out, 0.0, 0.0, nullptr, nullptr);
}

int pgb_imu_fit_windows_fwd(pgb_imu* o, const double* gps_v, const int64_t* gps_t, int n_gps, int batch_size,
                            int shift_step, int max_iterations, double epsilon, int first_window, int n_windows,
                            double* speed_sum, int32_t* speed_cnt, double* x_out, double* fx_out, int32_t* iters_out,
                            double fwd_min_velocity, double fwd_min_rotation_rad, double* fwd_sum_xyz,
                            int32_t* fwd_windows_used) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  // the CHECKs of fit_motion.cc:304-309
  if (batch_size <= 0 || shift_step <= 0 || batch_size < shift_step || max_iterations <= 0 || !(epsilon > 0))
    return fail(PGB_ERR_INVALID, "pgb_imu_fit_windows: invalid window/optimiser parameters");
  PGB_CUDA(cudaSetDevice(o->device));
  int rc = prepare(o, gps_v, gps_t, n_gps, batch_size, shift_step, first_window, n_windows);
  if (rc) return rc;
  const int nW = (int)o->win.size();
  const size_t M = o->mergedT.size();
  // events outside the shard's range get 0 / 0; the range itself is overwritten by the device results below
  auto zero_outside = [&](size_t lo, size_t hi) {
    if (speed_sum) { memset(speed_sum, 0, lo * sizeof(double)); memset(speed_sum + hi, 0, (M - hi) * sizeof(double)); }
    if (speed_cnt) { memset(speed_cnt, 0, lo * sizeof(int32_t)); memset(speed_cnt + hi, 0, (M - hi) * sizeof(int32_t)); }
  };
  if (nW == 0) { zero_outside(0, 0); return PGB_OK; }
  PhaseTimer pt(o->stream);
  PGB_CUDA(cudaEventRecord(o->ev[2], o->stream));
  rc = solve(o, max_iterations, epsilon, 0);
  if (rc) return rc;
  PGB_CUDA(cudaEventRecord(o->ev[3], o->stream));
  pt.mark("k_imu_solve");
  PGB_CUDA(cudaEventRecord(o->ev[4], o->stream));
  rc = speeds(o, false, fwd_sum_xyz != nullptr, fwd_min_velocity);
  if (rc) return rc;
  PGB_CUDA(cudaEventRecord(o->ev[5], o->stream));
  o->evValid = true;
  pt.mark("k_imu_speeds");
  cudaStream_t s = o->stream;
  std::vector<int> its(nW);
  std::vector<double> fwd;
  if (fwd_sum_xyz && o->spTotal > 0) {
    fwd.resize(4 * (size_t)nW * o->maxRefs);
    PGB_CUDA(cudaMemcpyAsync(fwd.data(), o->dFwd.p, fwd.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
  }
  PGB_CUDA(cudaMemcpyAsync(its.data(), o->dIt.p, nW * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (x_out) PGB_CUDA(cudaMemcpyAsync(x_out, o->dX.p, (size_t)nW * 9 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (fx_out) PGB_CUDA(cudaMemcpyAsync(fx_out, o->dFx.p, (size_t)nW * sizeof(double), cudaMemcpyDeviceToHost, s));
  size_t covLo = 0, covHi = 0;  // merged-event range [covLo, covHi) written from the device
  if ((speed_sum || speed_cnt) && o->spTotal > 0) {
    // merged-event range touched by the shard, and each event's run of sub-intervals (device: k_imu_event_runs)
    const int kLo = o->ioff[std::min(o->win.front().s + 1, n_gps)], kHi = o->ioff[o->win.back().e];
    if (kHi > kLo) {
      int mEnds[2] = {0, 0};
      PGB_CUDA(cudaMemcpyAsync(&mEnds[0], o->dIvM.p + kLo, sizeof(int), cudaMemcpyDeviceToHost, s));
      PGB_CUDA(cudaMemcpyAsync(&mEnds[1], o->dIvM.p + kHi - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
      PGB_CUDA(cudaStreamSynchronize(s));
      const int mLo = mEnds[0], mHi = mEnds[1];
      const int mc = mHi - mLo + 1;
      if (ensure(o->dFirstIv, mc) || ensure(o->dLastIv, mc) || ensure(o->dSum, M) || ensure(o->dCnt, M)) return PGB_ERR_CUDA;
      PGB_CUDA(cudaMemsetAsync(o->dFirstIv.p, 0xff, (size_t)mc * sizeof(int), s));
      PGB_CUDA(cudaMemsetAsync(o->dLastIv.p, 0xff, (size_t)mc * sizeof(int), s));
      k_imu_event_runs<<<(kHi - kLo + 255) / 256, 256, 0, s>>>(kLo, kHi, mLo, o->dIvM.p, o->dFirstIv.p, o->dLastIv.p);
      PGB_CHECK_LAUNCH();
      pt.mark("k_imu_event_runs");
      covLo = (size_t)mLo; covHi = (size_t)mLo + mc;
      k_imu_average<<<(mc + 255) / 256, 256, 0, s>>>(mLo, mc, o->dFirstIv.p, o->dLastIv.p, o->dIvRef.p, nW, o->firstWin,
                                                     batch_size, shift_step, n_gps, o->dWin.p, o->dIoff.p, o->dSpeeds.p,
                                                     o->dSum.p, o->dCnt.p);
      PGB_CHECK_LAUNCH();
      if (speed_sum) PGB_CUDA(cudaMemcpyAsync(speed_sum + mLo, o->dSum.p + mLo, (size_t)mc * sizeof(double), cudaMemcpyDeviceToHost, s));
      if (speed_cnt) PGB_CUDA(cudaMemcpyAsync(speed_cnt + mLo, o->dCnt.p + mLo, (size_t)mc * sizeof(int), cudaMemcpyDeviceToHost, s));
      zero_outside(covLo, covHi);  // host work while the copies are in flight
      PGB_CUDA(cudaStreamSynchronize(s));
      pt.mark("k_imu_average + D2H");
    }
  }
  if (covHi == covLo) zero_outside(0, 0);
  PGB_CUDA(cudaStreamSynchronize(s));
  if (fwd_sum_xyz) {
    // total_velocity_local (fit_motion.cc:172-173, :232-248): Kahan sum (math.hpp:8-27) over the windows whose largest
    // rotation reaches the threshold; per (window, GPS interval) partial sums come from k_imu_speeds.
    double sum[3] = {0, 0, 0}, rem[3] = {0, 0, 0};
    int used = 0;
    for (int w = 0; w < nW && !fwd.empty(); w++) {
      const int nr = o->win[w].e - o->win[w].s - 1;
      double minCos = 1.0;
      for (int j = 0; j < nr; j++) minCos = std::min(minCos, fwd[4 * ((size_t)w * o->maxRefs + j) + 3]);
      if (!(std::acos(minCos) >= fwd_min_rotation_rad)) continue;
      used++;
      for (int j = 0; j < nr; j++)
        for (int c = 0; c < 3; c++) {
          const double v = fwd[4 * ((size_t)w * o->maxRefs + j) + c];
          const double prop = v + rem[c], upd = sum[c] + prop, act = upd - sum[c];
          rem[c] = prop - act;
          sum[c] = upd;
        }
    }
    for (int c = 0; c < 3; c++) fwd_sum_xyz[c] = sum[c];
    if (fwd_windows_used) *fwd_windows_used = used;
  }
  for (int w = 0; w < nW; w++) {
    if (iters_out) iters_out[w] = its[w];
    if (its[w] < 0) return fail(PGB_ERR_NUMERIC, "window %d: the line search step left [min_step, max_step]", o->firstWin + w);
  }
  return PGB_OK;
}

// Symmetric 3x3 eigen-decomposition as cv::PCA performs it (cv::eigen -> the cyclic Jacobi sweep of OpenCV's
// core/src/lapack.cpp JacobiImpl_, un-vendored: restated from the published algorithm so that the eigenvector SIGNS
// follow OpenCV's; rows sorted by descending eigenvalue).
static void jacobi3(double A[9], double W[3], double V[9]) {
  const int n = 3;
  const double eps = 2.220446049250313e-16;
  for (int i = 0; i < 9; i++) V[i] = 0.0;
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  int indR[3] = {0, 0, 0}, indC[3] = {0, 0, 0};
  for (int k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) {
      int m = k + 1;
      double mv = std::fabs(A[n * k + m]);
      for (int i = k + 2; i < n; i++) { const double val = std::fabs(A[n * k + i]); if (mv < val) { mv = val; m = i; } }
      indR[k] = m;
    }
    if (k > 0) {
      int m = 0;
      double mv = std::fabs(A[k]);
      for (int i = 1; i < k; i++) { const double val = std::fabs(A[n * i + k]); if (mv < val) { mv = val; m = i; } }
      indC[k] = m;
    }
  }
  for (int iters = 0, maxIters = n * n * 30; iters < maxIters; iters++) {
    int k = 0;
    double mv = std::fabs(A[indR[0]]);
    for (int i = 1; i < n - 1; i++) { const double val = std::fabs(A[n * i + indR[i]]); if (mv < val) { mv = val; k = i; } }
    int l = indR[k];
    for (int i = 1; i < n; i++) { const double val = std::fabs(A[n * indC[i] + i]); if (mv < val) { mv = val; k = indC[i]; l = i; } }
    const double p = A[n * k + l];
    if (std::fabs(p) <= eps) break;
    const double y = (W[l] - W[k]) * 0.5;
    double t = std::fabs(y) + std::hypot(p, y);
    double s = std::hypot(p, t);
    const double c = t / s;
    s = p / s;
    t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[n * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    auto rot = [&](double& v0, double& v1) { const double a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; };
    for (int i = 0; i < k; i++) rot(A[n * i + k], A[n * i + l]);
    for (int i = k + 1; i < l; i++) rot(A[n * k + i], A[n * i + l]);
    for (int i = l + 1; i < n; i++) rot(A[n * k + i], A[n * l + i]);
    for (int i = 0; i < n; i++) rot(V[n * k + i], V[n * l + i]);
    for (int j = 0; j < 2; j++) {
      const int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        int m = idx + 1;
        double mv2 = std::fabs(A[n * idx + m]);
        for (int i = idx + 2; i < n; i++) { const double val = std::fabs(A[n * idx + i]); if (mv2 < val) { mv2 = val; m = i; } }
        indR[idx] = m;
      }
      if (idx > 0) {
        int m = 0;
        double mv2 = std::fabs(A[idx]);
        for (int i = 1; i < idx; i++) { const double val = std::fabs(A[n * i + idx]); if (mv2 < val) { mv2 = val; m = i; } }
        indC[idx] = m;
      }
    }
  }
  for (int k = 0; k < n - 1; k++) {
    int m = k;
    for (int i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
    if (k != m) {
      std::swap(W[m], W[k]);
      for (int i = 0; i < n; i++) std::swap(V[n * m + i], V[n * k + i]);
    }
  }
}

int pgb_principal_rotation_axes(int device, const double* gyro_xyz, const int64_t* gyro_t, size_t n,
                                int64_t integration_interval_usec, double axes_out[9], int64_t* n_intervals) {
  if (!gyro_xyz || !gyro_t || !axes_out) return fail(PGB_ERR_INVALID, "null argument");
  if (integration_interval_usec <= 0) return fail(PGB_ERR_INVALID, "integration interval must be positive");  // CHECK_GT
  // interval boundaries: the integer running sum of rotation.cc:24-43
  std::vector<int> start;
  start.push_back(1);
  int64_t cur = 0;
  for (size_t k = 1; k < n; k++) {
    cur += gyro_t[k] - gyro_t[k - 1];
    if (cur >= integration_interval_usec) { start.push_back((int)k + 1); cur = 0; }
  }
  const int nIv = (int)start.size() - 1;
  if (n_intervals) *n_intervals = nIv;
  if (nIv < 3) return fail(PGB_ERR_INVALID, "fewer than 3 rotation integration intervals (%d)", nIv);  // CHECK_GE(.., 3)
  if (use_device(device)) return PGB_ERR_CUDA;
  DevBuf<double> dG, dOut;
  DevBuf<long long> dT;
  DevBuf<int> dS;
  if (dG.alloc(3 * n) || dT.alloc(n) || dS.alloc(start.size()) || dOut.alloc(3 * (size_t)nIv)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpy(dG.p, gyro_xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dT.p, gyro_t, n * sizeof(int64_t), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dS.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
  k_rot_intervals<<<(nIv + 63) / 64, 64>>>(nIv, dS.p, dG.p, dT.p, dOut.p);
  PGB_CHECK_LAUNCH();
  std::vector<double> rows(3 * (size_t)nIv);
  PGB_CUDA(cudaMemcpy(rows.data(), dOut.p, rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
  // cv::PCA(data, noArray(), DATA_AS_ROW): mean, covariance scaled by 1/nsamples, eigenvectors as rows
  double mean[3] = {0, 0, 0};
  for (int i = 0; i < nIv; i++) for (int c = 0; c < 3; c++) mean[c] += rows[3 * (size_t)i + c];
  for (int c = 0; c < 3; c++) mean[c] /= nIv;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nIv; i++) {
    double d[3];
    for (int c = 0; c < 3; c++) d[c] = rows[3 * (size_t)i + c] - mean[c];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) cov[3 * a + b] += d[a] * d[b];
  }
  for (int i = 0; i < 9; i++) cov[i] /= nIv;
  double W[3];
  jacobi3(cov, W, axes_out);
  return PGB_OK;
}

int pgb_angular_velocities_around_axis(int device, const double* gyro_xyz, size_t n, const double axis[3], double* out) {
  if (!gyro_xyz || !axis || !out) return fail(PGB_ERR_INVALID, "null argument");
  const double norm = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  if (!(norm > 1.0 - 1e-2) || !(norm < 1.0 + 1e-2)) return fail(PGB_ERR_INVALID, "axis is not normalised");  // rotation.cc:108-109
  if (n == 0) return PGB_OK;
  if (use_device(device)) return PGB_ERR_CUDA;
  DevBuf<double> dG, dOut;
  if (dG.alloc(3 * n) || dOut.alloc(n)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpy(dG.p, gyro_xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  k_axis_project<<<(unsigned)((n + 255) / 256), 256>>>((long long)n, dG.p, axis[0], axis[1], axis[2], norm, dOut.p);
  PGB_CHECK_LAUNCH();
  PGB_CUDA(cudaMemcpy(out, dOut.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return PGB_OK;
}

int pgb_time_averaged_values(int device, const double* values, const int64_t* times_usec, int64_t n,
                             const int64_t* frame_times_usec, int64_t n_frames, double* out_values, uint8_t* out_valid) {
  if (n <= 0 || !values || !times_usec) return fail(PGB_ERR_INVALID, "empty time series");  // CHECK_LT(hint, size)
  if (n_frames < 0 || (n_frames > 0 && (!frame_times_usec || (n_frames > 1 && (!out_values || !out_valid)))))
    return fail(PGB_ERR_INVALID, "pgb_time_averaged_values: invalid argument");
  if (n_frames < 2) return PGB_OK;
  for (int64_t i = 1; i < n_frames; i++)
    if (!(frame_times_usec[i] > frame_times_usec[i - 1]))
      return fail(PGB_ERR_INVALID, "frame timestamps must be strictly increasing (frame %lld)", (long long)i);  // CHECK_GT(end, start)
  for (int64_t i = 1; i < n; i++)
    if (times_usec[i] < times_usec[i - 1]) return fail(PGB_ERR_INVALID, "series timestamps must be non-decreasing");
  if (use_device(device)) return PGB_ERR_CUDA;
  DevBuf<double> dv, dout;
  DevBuf<long long> dt, dft;
  DevBuf<unsigned char> dst;
  if (dv.alloc(n) || dt.alloc(n) || dft.alloc(n_frames) || dout.alloc(n_frames - 1) || dst.alloc(n_frames - 1)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpy(dv.p, values, n * sizeof(double), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dt.p, times_usec, n * sizeof(int64_t), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dft.p, frame_times_usec, n_frames * sizeof(int64_t), cudaMemcpyHostToDevice));
  k_time_average<<<(unsigned)((n_frames - 1 + 127) / 128), 128>>>(dv.p, dt.p, n, dft.p, n_frames, dout.p, dst.p);
  PGB_CHECK_LAUNCH();
  PGB_CUDA(cudaMemcpy(out_values, dout.p, (n_frames - 1) * sizeof(double), cudaMemcpyDeviceToHost));
  PGB_CUDA(cudaMemcpy(out_valid, dst.p, n_frames - 1, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i + 1 < n_frames; i++)
    if (out_valid[i] == 2)
      return fail(PGB_ERR_INVALID, "frame %lld ends at or after the last event of the series: the reference's LinearInterpolate CHECK fails "
                  "(time_series.hpp:210-213)", (long long)(i + 1));
  return PGB_OK;
}

int pgb_smooth_time_series(int device, const double* values, const double* times, int64_t n, const double* target_times,
                           int64_t n_target, double sigma, double* out) {
  if (!(sigma > 0)) return fail(PGB_ERR_INVALID, "sigma must be positive");  // CHECK_GT(sigma, 0)
  if (n < 0 || n_target < 0 || (n_target > 0 && (!values || !times || !target_times || !out || n == 0)))
    return fail(PGB_ERR_INVALID, "pgb_smooth_time_series: invalid argument");
  if (n_target == 0) return PGB_OK;
  // the reference's window bounds only ever move forward (smoothing.cc:70-78): targets and data must be ascending
  for (int64_t i = 0; i + 1 < n_target; i++)
    if (target_times[i + 1] < target_times[i]) return fail(PGB_ERR_INVALID, "target timestamps must be non-decreasing");
  for (int64_t i = 0; i + 1 < n; i++)
    if (times[i + 1] < times[i]) return fail(PGB_ERR_INVALID, "data timestamps must be non-decreasing");
  if (use_device(device)) return PGB_ERR_CUDA;
  DevBuf<double> dv, dt, dtt, dout;
  if (dv.alloc(n) || dt.alloc(n) || dtt.alloc(n_target) || dout.alloc(n_target)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemcpy(dv.p, values, n * sizeof(double), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dt.p, times, n * sizeof(double), cudaMemcpyHostToDevice));
  PGB_CUDA(cudaMemcpy(dtt.p, target_times, n_target * sizeof(double), cudaMemcpyHostToDevice));
  k_smooth<<<(unsigned)((n_target + 255) / 256), 256>>>(dv.p, dt.p, n, dtt.p, n_target, sigma, dout.p);
  PGB_CHECK_LAUNCH();
  PGB_CUDA(cudaMemcpy(out, dout.p, n_target * sizeof(double), cudaMemcpyDeviceToHost));
  return PGB_OK;
}

}  // extern "C"
