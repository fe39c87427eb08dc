// ORACLE -- TEST INFRASTRUCTURE ONLY.  C wrapper around the reference's per-frame annotation: the REAL
// include/interpolation/time_series.hpp (TimeSeries<T>::MostRecentPreviousValue / TimeAveragedValue / LinearInterpolate /
// GaussianSmooth, RealTimeSeries, the header's SmoothTimeSeries template), included as is, and the frame loop of
// src/annotate_frames.cc:56-69, streamed from the reference's file into pgr_annotate_loop() below by `make -C oracle _ref`.
// The JSON documents the code reads are built in memory (ref_shims/json.hpp, ref_shims/io/json_converters.hpp).
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <glog/logging.h>
#include <interpolation/time_series.hpp>

namespace {
std::unique_ptr<nlohmann::json> g_next_document;
}
namespace pilotguru {
std::unique_ptr<nlohmann::json> ReadJsonFile(const std::string& /*filename*/) { return std::move(g_next_document); }
}

void pgr_annotate_loop(pilotguru::RealTimeSeries* in_series, const nlohmann::json& frames, nlohmann::json& out_list,
                       const std::string& FLAGS_json_value_name);   // body: annotate_frames.cc:56-69 (ref_annotate_part)

extern "C" int64_t pgr_annotate_frames(const double* values, const int64_t* times_usec, int64_t n, const int64_t* frame_times_usec,
                                       int64_t n_frames, double smoothing_sigma, int64_t* out_frame_id, double* out_value, int64_t cap) {
  g_next_document.reset(new nlohmann::json);
  nlohmann::json& series = (*g_next_document)["series"];
  for (int64_t i = 0; i < n; i++) {
    nlohmann::json e;
    e["time_usec"] = nlohmann::json::integer((long)times_usec[i]);
    e["v"] = nlohmann::json::real(values[i]);
    series.push_back(e);
  }
  std::unique_ptr<pilotguru::RealTimeSeries> in_series(new pilotguru::RealTimeSeries("in.json", "series", "v"));
  if (smoothing_sigma > 0) in_series->GaussianSmooth(smoothing_sigma);   // annotate_frames.cc:49-51
  nlohmann::json frames, out_list;
  for (int64_t f = 0; f < n_frames; f++) {
    nlohmann::json e;
    e["time_usec"] = nlohmann::json::integer((long)frame_times_usec[f]);
    e["frame_id"] = nlohmann::json::integer((long)f);
    frames.push_back(e);
  }
  pgr_annotate_loop(in_series.get(), frames, out_list, "v");
  int64_t k = 0;
  for (const nlohmann::json& e : out_list) {
    if (k < cap) { out_frame_id[k] = (long)e["frame_id"]; out_value[k] = (double)e["v"]; }
    k++;
  }
  return k;
}
