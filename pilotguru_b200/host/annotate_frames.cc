// annotate_frames -- drop-in for the reference binary (src/annotate_frames.cc): labels every video frame with the
// time-weighted average of a JSON time series over the interval since the previous frame (the
// step that turns fit_motion's output into per-frame velocity / steering labels).  Same flags, same JSON in/out.
//   RealTimeSeries(in_json, root, value)      include/interpolation/time_series.hpp:243-263 -> ReadTable
//   GaussianSmooth(sigma)                     time_series.hpp:91-100                         -> pgb_smooth_time_series
//   TimeAveragedValue per frame               time_series.hpp:129-189, annotate_frames.cc:59-72 -> pgb_time_averaged_values
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pgb200.h"
#include "check.hpp"
#include "flags.hpp"
#include "json_lite.hpp"

int main(int argc, char** argv) {
  std::string frames_json, in_json, json_root_element_name, json_value_name, out_json;
  double smoothing_sigma = -1.0;
  int64_t device = 0;
  pgbhost::Flags flags;
  flags.String("frames_json", &frames_json, "JSON file with video frames timestamps.");
  flags.String("in_json", &in_json, "JSON file with the source time series data.");
  flags.String("json_root_element_name", &json_root_element_name, "Root JSON element, pointing to the time series list.");
  flags.String("json_value_name", &json_value_name, "Value element name in the time series JSON file.");
  flags.String("out_json", &out_json, "Filename to write to.");
  flags.Double("smoothing_sigma", &smoothing_sigma, "If positive, Gaussian smoothing sigma (seconds) applied first.");
  flags.Int64("device", &device, "(extension) CUDA device");
  flags.Parse(argc, argv);

  PGB_CHECK(!frames_json.empty());
  const pgbhost::Table frames = pgbhost::ReadTable(frames_json, "frames", {"frame_id"}, "time_usec");
  PGB_CHECK(!in_json.empty());
  PGB_CHECK(!json_root_element_name.empty());
  PGB_CHECK(!json_value_name.empty());
  const pgbhost::Table series = pgbhost::ReadTable(in_json, json_root_element_name, {json_value_name}, "time_usec");
  std::vector<double> values = series.real[0];
  const int64_t n = (int64_t)series.rows();
  if (smoothing_sigma > 0) {  // TimeSeries::GaussianSmooth
    std::vector<double> ts(n), sm(n);
    for (int64_t i = 0; i < n; i++) ts[i] = (double)(series.integer[i] - series.integer[0]) * 1e-6;
    PGB_CALL(pgb_smooth_time_series((int)device, values.data(), ts.data(), n, ts.data(), n, smoothing_sigma, sm.data()));
    values.swap(sm);
  }
  const int64_t nf = (int64_t)frames.rows();
  std::vector<double> out(nf > 1 ? nf - 1 : 0);
  std::vector<uint8_t> valid(nf > 1 ? nf - 1 : 0);
  PGB_CALL(pgb_time_averaged_values((int)device, values.data(), series.integer.data(), n, frames.integer.data(), nf,
                                    out.data(), valid.data()));
  FILE* f = fopen(out_json.c_str(), "w");
  PGB_CHECK(f != nullptr) << "cannot write " << out_json;
  const bool value_first = json_value_name < std::string("frame_id");
  bool any = false;
  for (int64_t i = 1; i < nf; i++) {
    if (!valid[i - 1]) continue;
    fprintf(f, any ? ",\n" : "{\n  \"%s\": [\n", json_root_element_name.c_str());
    any = true;
    const std::string v = pgbhost::FormatDouble(out[i - 1]);
    const long long id = (long long)frames.real[0][i];
    if (value_first) fprintf(f, "    {\n      \"%s\": %s,\n      \"frame_id\": %lld\n    }", json_value_name.c_str(), v.c_str(), id);
    else fprintf(f, "    {\n      \"frame_id\": %lld,\n      \"%s\": %s\n    }", id, json_value_name.c_str(), v.c_str());
  }
  if (any) fprintf(f, "\n  ]\n}\n");
  else fprintf(f, "{\n  \"%s\": null\n}\n", json_root_element_name.c_str());  // out_json[root] = {} stays null
  fclose(f);
  return EXIT_SUCCESS;
}
