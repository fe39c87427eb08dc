// ORACLE -- TEST INFRASTRUCTURE ONLY.  Shell of g2o::OptimizationAlgorithmLevenberg (members as in
// core/optimization_algorithm_levenberg.h) and of the optimizer / solver interfaces its solve() calls, forwarding to the
// oracle's primitives (oracle/pgo.h: pgo_pose_problem_*).  The bodies of solve(), computeLambdaInit() and computeScale()
// come from the reference's file (oracle/Makefile, target _ref).  Not part of the product.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <limits>
#include <vector>

#include "../pgo.h"

#define g2o_isfinite(x) std::isfinite(x)

namespace g2o {
using std::cerr;
using std::endl;

struct G2OBatchStatistics {
  double timeResiduals, timeQuadraticForm, timeLinearSolution, timeUpdate;
  int levenbergIterations;
  static G2OBatchStatistics* globalStats() { return nullptr; }
};
inline double get_monotonic_time() { return 0.; }
template <typename T> struct Property {
  T v;
  T value() const { return v; }
};

struct VertexShell {   // the single 6-dof pose vertex: hessian(j, j) is what computeLambdaInit() reads
  const double* H;
  int dimension() const { return 6; }
  double hessian(int i, int j) const { return H[6 * i + j]; }
};
struct OptimizableGraph { typedef VertexShell Vertex; };

class SparseOptimizerShell {
 public:
  void* P;
  std::vector<double> stack;           // push() / pop() / discardTop(): the estimate stack of the vertex
  VertexShell vertex;
  std::vector<VertexShell*> mapping;
  explicit SparseOptimizerShell(void* problem) : P(problem) { vertex.H = nullptr; mapping.push_back(&vertex); }
  void computeActiveErrors() { pgo_pose_problem_compute_active_errors(P); }
  double activeRobustChi2() const { return pgo_pose_problem_active_robust_chi2(P); }
  void push() { double e[7]; pgo_pose_problem_get_estimate(P, e); stack.insert(stack.end(), e, e + 7); }
  void pop() { pgo_pose_problem_set_estimate(P, &stack[stack.size() - 7]); stack.resize(stack.size() - 7); }
  void discardTop() { stack.resize(stack.size() - 7); }
  void update(const double* x) { pgo_pose_problem_oplus(P, x); }
  bool terminate() { return false; }
  const std::vector<VertexShell*>& indexMapping() const { return mapping; }
};

class SolverShell {   // BlockSolver_6_3 + LinearSolverDense on one 6x6 block
 public:
  SparseOptimizerShell* opt;
  double H[36], b_[6], x_[6], diag[6];
  explicit SolverShell(SparseOptimizerShell* o) : opt(o) { memset(H, 0, sizeof H); memset(b_, 0, sizeof b_); memset(x_, 0, sizeof x_); opt->vertex.H = H; }
  SparseOptimizerShell* optimizer() const { return opt; }
  bool buildStructure() { return true; }
  bool buildSystem() { pgo_pose_problem_build_system(opt->P, H, b_); return true; }
  bool setLambda(double lambda, bool backup) {
    for (int i = 0; i < 6; i++) { if (backup) diag[i] = H[7 * i]; H[7 * i] += lambda; }
    return true;
  }
  void restoreDiagonal() { for (int i = 0; i < 6; i++) H[7 * i] = diag[i]; }
  bool solve() { return pgo_pose_ldlt6_solve(H, b_, x_) != 0; }   // on failure x keeps its previous content, like LinearSolverDense
  double* x() { return x_; }
  double* b() { return b_; }
  size_t vectorSize() const { return 6; }
  bool schur() { return false; }
};

class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
};

class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  SolverResult solve(int iteration, bool online = false);
  double computeLambdaInit() const;
  double computeScale() const;
  SparseOptimizerShell* _optimizer;
  SolverShell* _solver;
  Property<int>* _maxTrialsAfterFailure;
  Property<double>* _userLambdaInit;
  double _currentLambda, _tau, _goodStepUpperScale, _goodStepLowerScale, _ni;
  int _levenbergIterations, _nBad;
};

}  // namespace g2o
