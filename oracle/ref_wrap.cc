// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the REAL reference implementation of the time-series alignment
// (src/interpolation/align_time_series.cc, compiled where it lies by `make -C oracle _ref`; its only external
// dependency, glog's CHECK macros, comes from oracle/ref_shims).  tests/test_oracle_reference_pin.py runs the oracle's
// restatement (oracle/pgo_calib.cc) against it.  The other files of the path need Eigen / OpenCV and cannot be built here.
#include <cstddef>
#include <cstdint>
#include <vector>

#include <interpolation/align_time_series.hpp>

extern "C" {

// pilotguru::MergedTimeSeries({&a, &b}) (align_time_series.cc:115-143): events as (index into a, index into b,
// effective time).  Returns the number of events; fills at most cap.
int64_t pgr_merge_two(const int64_t* a, int64_t na, const int64_t* b, int64_t nb, int64_t* idx_a, int64_t* idx_b,
                      int64_t* t_usec, int64_t cap) {
  const std::vector<long> va(a, a + na), vb(b, b + nb);
  const pilotguru::MergedTimeSeries m({&va, &vb});
  const auto& ev = m.MergedEvents();
  for (size_t i = 0; i < ev.size() && (int64_t)i < cap; i++) {
    idx_a[i] = (int64_t)ev[i][0];
    idx_b[i] = (int64_t)ev[i][1];
    t_usec[i] = m.MergedEventTimeUsec(i);
  }
  return (int64_t)ev.size();
}

// pilotguru::MakeInterpolationIntervals(reference, interpolation) (align_time_series.cc:155-196), flattened in order:
// reference_end_time_index, interpolation_end_time_index, start_usec, end_usec; per_ref[r] = intervals of reference r.
int64_t pgr_make_interpolation_intervals(const int64_t* ref, int64_t nr, const int64_t* interp, int64_t ni, int64_t* ref_idx,
                                         int64_t* interp_idx, int64_t* start, int64_t* end, int64_t cap, int64_t* per_ref) {
  const std::vector<long> vr(ref, ref + nr), vi(interp, interp + ni);
  const auto all = pilotguru::MakeInterpolationIntervals(vr, vi);
  int64_t k = 0;
  for (size_t r = 0; r < all.size(); r++) {
    if (per_ref) per_ref[r] = (int64_t)all[r].size();
    for (const auto& iv : all[r]) {
      if (k < cap) {
        ref_idx[k] = (int64_t)iv.reference_end_time_index;
        interp_idx[k] = (int64_t)iv.interpolation_end_time_index;
        start[k] = iv.start_usec;
        end[k] = iv.end_usec;
      }
      k++;
    }
  }
  return k;
}

int64_t pgr_num_reference_rows(const int64_t* ref, int64_t nr, const int64_t* interp, int64_t ni) {
  const std::vector<long> vr(ref, ref + nr), vi(interp, interp + ni);
  return (int64_t)pilotguru::MakeInterpolationIntervals(vr, vi).size();
}

}  // extern "C"
