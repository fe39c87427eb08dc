// ORACLE -- TEST INFRASTRUCTURE ONLY.  The Levenberg-Marquardt driver of the reference's vendored g2o --
// OptimizationAlgorithmLevenberg::solve / computeLambdaInit / computeScale
// (thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-164, 166-180, 182-189), streamed from the reference's
// file by `make -C oracle _ref` into the class shell below -- running on an optimizer / solver pair that forwards every
// primitive (active errors, robust chi2, quadratic form, 6x6 LDLT, SE3 update, estimate stack) to the SAME functions
// oracle/pgo_pose.cc uses.  Whatever differs between this and the oracle's lm_solve() is therefore the driver's control
// flow: lambda initialisation and update, rho / scale, the trial loop, the termination tests (incl. the "_nBad" criterion
// ORB-SLAM2 added).  SparseOptimizer::optimize()'s loop (sparse_optimizer.cpp:376-414: call solve(i) until it stops
// returning OK) is three lines in pgr_lm_optimize().
#ifndef PGO_G2O_LM_PART
#error "compile through oracle/Makefile (target _ref)"
#endif

extern "C" void pgr_lm_optimize(void* problem, int iterations) {
  using namespace g2o;
  SparseOptimizerShell opt(problem);
  SolverShell solver(&opt);
  Property<int> maxTrials{10};
  Property<double> userLambdaInit{0.};
  OptimizationAlgorithmLevenberg lm;          // constructor values: optimization_algorithm_levenberg.cpp:43-55
  lm._optimizer = &opt; lm._solver = &solver;
  lm._currentLambda = -1.; lm._tau = 1e-5; lm._goodStepUpperScale = 2. / 3.; lm._goodStepLowerScale = 1. / 3.;
  lm._userLambdaInit = &userLambdaInit; lm._maxTrialsAfterFailure = &maxTrials;
  lm._ni = 2.; lm._levenbergIterations = 0; lm._nBad = 0;
  bool ok = true;
  for (int i = 0; i < iterations && ok; i++) ok = lm.solve(i, false) == OptimizationAlgorithm::OK;
}
