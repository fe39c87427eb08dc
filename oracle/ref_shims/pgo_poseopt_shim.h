// ORACLE -- TEST INFRASTRUCTURE ONLY.  Shells of the g2o / ORB-SLAM2 classes that the body of Optimizer::PoseOptimization
// (thirdparty/orb-slam2/src/Optimizer.cc:239-451) names, so that this body can be compiled from the reference's file
// (oracle/Makefile, target _ref).  The shells keep a mirrored oracle Problem (oracle/pgo.h: pgo_pose_problem_*): an edge's
// setLevel / setRobustKernel / computeError / chi2 act on its mirror, SparseOptimizer::optimize() runs g2o's own
// Levenberg-Marquardt driver compiled from source (pgr_lm_optimize, ref_wrap_g2o_lm.cc) -- or returns at once when no edge
// is at level 0, like initializeOptimization() + optimize() do.  What this pins is the function's own logic: which
// observations become edges, the four rounds restarting from mTcw, the chi2 classification with the re-evaluation of
// outliers, levels, dropping the robust kernel after the third round, the < 3 and < 10 exits, the returned inlier count.
// Not part of the product.
#pragma once
#include <cmath>
#include <cstddef>
#include <mutex>
#include <vector>

#include <Eigen/Geometry>

#include "pgo_opencv_shim.h"
#include "../pgo.h"

extern "C" void pgr_lm_optimize(void* problem, int iterations);   // g2o's LM driver from source (ref_wrap_g2o_lm.cc)

using namespace std;

namespace g2o {

struct SE3Quat { double pose7[7]; };   // w, x, y, z, t -- the arithmetic lives in the mirrored Problem

struct OptimizableGraph {
  struct Vertex { virtual ~Vertex() {} };
};
class SparseOptimizer;

class VertexSE3Expmap : public OptimizableGraph::Vertex {
 public:
  SparseOptimizer* owner = nullptr;
  SE3Quat pending;
  void setEstimate(const SE3Quat& e);
  void setId(int) {}
  void setFixed(bool) {}
  SE3Quat estimate() const;
};

class RobustKernelHuber {
 public:
  double delta = 0;
  void setDelta(double d) { delta = d; }
};

struct PoseEdgeBase {
  SparseOptimizer* owner = nullptr;
  int index = -1;
  RobustKernelHuber* kernel = nullptr;
  void setVertex(int, OptimizableGraph::Vertex*) {}
  void setRobustKernel(RobustKernelHuber* k);
  void setLevel(int l);
  void computeError();
  double chi2() const;
};
class EdgeSE3ProjectXYZOnlyPose : public PoseEdgeBase {
 public:
  Eigen::Matrix<double, 2, 1> measurement;
  double info = 0;
  Eigen::Vector3d Xw;
  double fx = 0, fy = 0, cx = 0, cy = 0;
  void setMeasurement(const Eigen::Matrix<double, 2, 1>& m) { measurement = m; }
  void setInformation(const Eigen::Matrix2d& I) { info = I(0, 0); }
};
class EdgeStereoSE3ProjectXYZOnlyPose : public PoseEdgeBase {   // compiled, never instantiated: the path is monocular
 public:
  Eigen::Vector3d Xw;
  double fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
  void setMeasurement(const Eigen::Matrix<double, 3, 1>&) {}
  void setInformation(const Eigen::Matrix3d&) {}
};

struct BlockSolver_6_3 {
  typedef int PoseMatrixType;
  struct LinearSolverType {};
  explicit BlockSolver_6_3(LinearSolverType*) {}
};
template <typename M> struct LinearSolverDense : BlockSolver_6_3::LinearSolverType {};
struct OptimizationAlgorithmLevenberg {
  explicit OptimizationAlgorithmLevenberg(BlockSolver_6_3*) {}
};

class SparseOptimizer {
 public:
  VertexSE3Expmap* v = nullptr;
  std::vector<EdgeSE3ProjectXYZOnlyPose*> mono;
  void* problem = nullptr;
  struct EdgeSet { size_t n = 0; size_t size() const { return n; } } edgeSet;
  ~SparseOptimizer() { if (problem) pgo_pose_problem_destroy(problem); }
  void setAlgorithm(OptimizationAlgorithmLevenberg*) {}
  void addVertex(VertexSE3Expmap* vv) { v = vv; vv->owner = this; }
  OptimizableGraph::Vertex* vertex(int) { return v; }
  void addEdge(EdgeSE3ProjectXYZOnlyPose* e) { e->owner = this; e->index = (int)mono.size(); mono.push_back(e); edgeSet.n++; }
  void addEdge(EdgeStereoSE3ProjectXYZOnlyPose*) { edgeSet.n++; }
  const EdgeSet& edges() const { return edgeSet; }
  void ensure() {   // the mirrored Problem, built when the graph is first used
    if (problem || mono.empty()) return;
    std::vector<double> obs, X, info;
    for (const EdgeSE3ProjectXYZOnlyPose* e : mono) {
      obs.push_back(e->measurement[0]); obs.push_back(e->measurement[1]);
      for (int k = 0; k < 3; k++) X.push_back(e->Xw[k]);
      info.push_back(e->info);
    }
    const EdgeSE3ProjectXYZOnlyPose* e0 = mono[0];
    problem = pgo_pose_problem_create_raw((int)mono.size(), obs.data(), X.data(), info.data(), e0->fx, e0->fy, e0->cx, e0->cy, e0->kernel->delta);
    pgo_pose_problem_set_estimate(problem, v->pending.pose7);
  }
  bool initializeOptimization(int /*level*/) { ensure(); return true; }
  int optimize(int iterations) {
    ensure();
    if (!problem || pgo_pose_problem_num_active(problem) == 0) return -1;   // "0 vertices to optimize"
    pgr_lm_optimize(problem, iterations);
    return iterations;
  }
};

inline void VertexSE3Expmap::setEstimate(const SE3Quat& e) {
  pending = e;
  if (owner && owner->problem) pgo_pose_problem_set_estimate(owner->problem, e.pose7);
}
inline SE3Quat VertexSE3Expmap::estimate() const {
  SE3Quat s = pending;
  if (owner && owner->problem) pgo_pose_problem_get_estimate(owner->problem, s.pose7);
  return s;
}
inline void PoseEdgeBase::setRobustKernel(RobustKernelHuber* k) {
  if (k) kernel = k;
  if (owner) { owner->ensure(); if (owner->problem) pgo_pose_problem_edge_set_robust(owner->problem, index, k != nullptr); }
}
inline void PoseEdgeBase::setLevel(int l) { owner->ensure(); pgo_pose_problem_edge_set_level(owner->problem, index, l); }
inline void PoseEdgeBase::computeError() { owner->ensure(); pgo_pose_problem_edge_compute_error(owner->problem, index); }
inline double PoseEdgeBase::chi2() const { owner->ensure(); return pgo_pose_problem_edge_chi2(owner->problem, index); }

}  // namespace g2o

namespace ORB_SLAM2 {

class MapPoint {
 public:
  static std::mutex mGlobalMutex;
  cv::Mat mWorldPos;
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
};

class Frame {
 public:
  int N = 0;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<float> mvuRight;
  std::vector<bool> mvbOutlier;
  std::vector<cv::KeyPoint> mvKeysUndistorted;
  std::vector<float> mvInvLevelSigma2;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
  cv::Mat mTcw;
  void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }
};

class Converter {
 public:
  static g2o::SE3Quat toSE3Quat(const cv::Mat& cvT) {   // Converter.cc:39-49 (the arithmetic: oracle se3_from_cv)
    float T[16];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) T[4 * i + j] = cvT.at<float>(i, j);
    g2o::SE3Quat s;
    pgo_pose_T_to_pose7(T, s.pose7);
    return s;
  }
  static cv::Mat toCvMat(const g2o::SE3Quat& SE3) {     // Converter.cc:51-55, 65-73
    float T[16];
    pgo_pose_pose7_to_T(SE3.pose7, T);
    cv::Mat m(4, 4, CV_32F);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) m.at<float>(i, j) = T[4 * i + j];
    return m;
  }
};

class Optimizer {
 public:
  static int PoseOptimization(Frame* pFrame);
};

}  // namespace ORB_SLAM2
