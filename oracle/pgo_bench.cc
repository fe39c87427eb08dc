// ORACLE -- TEST INFRASTRUCTURE ONLY.  Timed CPU baseline driver used by bench.py's cpu_baseline leg and by
// `bench.py --impl reference`: runs the oracle's extract + consecutive-match over n frames on `nthreads` host threads,
// as a frame-parallel CPU deployment of the reference would do.  Phase 1: the threads pull frames off a shared counter
// and extract them (one oracle extractor per thread); phase 2: they pull the n-1 frame pairs (f-1, f) and match them --
// EVERY pair is matched (the round-1 version gave each thread a contiguous block and skipped the match of the block's
// first frame).  Returns elapsed wall seconds over both phases; writes the total number of keypoints and matches and,
// when the arrays are given, the per-frame counts (matches_per_frame[0] = -1: frame 0 has no predecessor).
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

#include "pgo.h"

extern "C" double pgo_bench_extract_match(const uint8_t* frames, int n, int w, int h, const float* flow_xy,
                                          int nfeatures, float scale, int nlevels, int iniTh, int minTh, float th,
                                          int nthreads, int64_t* total_kps, int64_t* total_matches,
                                          int32_t* kps_per_frame, int32_t* matches_per_frame) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  const int cap = nfeatures + 4 * nlevels + 64;
  std::vector<pgb_keypoint> K((size_t)n * cap);
  std::vector<uint8_t> D((size_t)n * cap * 32);
  std::vector<int32_t> N(n, -1), M(n, -1);
  std::vector<float> sf(nlevels);
  std::atomic<int> next{0}, nextPair{1}, arrived{0};
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tix) {
    pgo_orb* o = pgo_orb_create(nfeatures, scale, nlevels, iniTh, minTh);
    if (tix == 0) pgo_orb_tables(o, sf.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
    for (int f; (f = next.fetch_add(1)) < n;)
      N[f] = pgo_orb_extract(o, frames + (size_t)f * w * h, w, h, (size_t)w, K.data() + (size_t)f * cap,
                             D.data() + (size_t)f * cap * 32, cap);
    pgo_orb_destroy(o);
    arrived.fetch_add(1);
    while (arrived.load() < nthreads) std::this_thread::yield();  // every frame's features are in place
    std::vector<int32_t> match(cap);
    for (int f; (f = nextPair.fetch_add(1)) < n;) {
      if (N[f - 1] < 0 || N[f] < 0) continue;
      M[f] = pgo_match_consecutive(K.data() + (size_t)(f - 1) * cap, D.data() + (size_t)(f - 1) * cap * 32, N[f - 1],
                                   K.data() + (size_t)f * cap, D.data() + (size_t)f * cap * 32, N[f], flow_xy[2 * f],
                                   flow_xy[2 * f + 1], (float)w, (float)h, th, sf.data(), nlevels, match.data());
    }
  };
  std::vector<std::thread> ts;
  for (int t = 1; t < nthreads; t++) ts.emplace_back(work, t);
  work(0);
  for (auto& t : ts) t.join();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int64_t a = 0, b = 0;
  for (int f = 0; f < n; f++) {
    if (N[f] > 0) a += N[f];
    if (M[f] > 0) b += M[f];
    if (kps_per_frame) kps_per_frame[f] = N[f];
    if (matches_per_frame) matches_per_frame[f] = M[f];
  }
  if (total_kps) *total_kps = a;
  if (total_matches) *total_matches = b;
  return el;
}
