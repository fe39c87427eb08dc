// ORACLE -- TEST INFRASTRUCTURE ONLY.  Timed CPU baseline driver used by bench.py's cpu_baseline leg and by
// `bench.py --impl reference`: runs the oracle's extract + consecutive-match over n frames on `nthreads` host threads
// (frames are split into contiguous chunks, one oracle extractor per thread, as a frame-parallel CPU deployment of
// the reference would do).  Returns elapsed wall seconds; writes the total number of keypoints and matches.
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

#include "pgo.h"

extern "C" double pgo_bench_extract_match(const uint8_t* frames, int n, int w, int h, const float* flow_xy,
                                          int nfeatures, float scale, int nlevels, int iniTh, int minTh, float th,
                                          int nthreads, int64_t* total_kps, int64_t* total_matches) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n;
  std::vector<int64_t> kp(nthreads, 0), mt(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int tix) {
    const int a = (int)((int64_t)n * tix / nthreads), b = (int)((int64_t)n * (tix + 1) / nthreads);
    pgo_orb* o = pgo_orb_create(nfeatures, scale, nlevels, iniTh, minTh);
    std::vector<float> sf(nlevels);
    pgo_orb_tables(o, sf.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
    const int cap = nfeatures + 4 * nlevels + 64;
    std::vector<pgb_keypoint> k0(cap), k1(cap);
    std::vector<uint8_t> d0((size_t)cap * 32), d1((size_t)cap * 32);
    std::vector<int32_t> match(cap);
    int nPrev = -1;
    for (int f = a; f < b; f++) {
      const int nc = pgo_orb_extract(o, frames + (size_t)f * w * h, w, h, (size_t)w, k1.data(), d1.data(), cap);
      if (nc < 0) break;
      kp[tix] += nc;
      if (nPrev >= 0)
        mt[tix] += pgo_match_consecutive(k0.data(), d0.data(), nPrev, k1.data(), d1.data(), nc, flow_xy[2 * f],
                                         flow_xy[2 * f + 1], (float)w, (float)h, th, sf.data(), nlevels, match.data());
      k0.swap(k1); d0.swap(d1); nPrev = nc;
    }
    pgo_orb_destroy(o);
  };
  std::vector<std::thread> ts;
  for (int t = 1; t < nthreads; t++) ts.emplace_back(work, t);
  work(0);
  for (auto& t : ts) t.join();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  int64_t a = 0, b = 0;
  for (int t = 0; t < nthreads; t++) { a += kp[t]; b += mt[t]; }
  if (total_kps) *total_kps = a;
  if (total_matches) *total_matches = b;
  return el;
}
