// ORACLE -- TEST INFRASTRUCTURE ONLY.  What src/fit_motion.cc:156-293 (ComputeAndSaveForwardVelocitiesFromImu, the window
// loop of the fit_motion binary) needs around it when that line range is compiled from the reference's file (oracle/Makefile,
// target _ref): the reference's own headers where they build (velocity.hpp, geometry.hpp, math.hpp, the vendored LBFGS.h,
// all against the Eigen / glog stand-ins), a declaration of SmoothTimeSeries (its header needs ORB-SLAM2), and CAPTURING
// stand-ins for the JSON surface (src/io/json_converters.cc is written on nlohmann/json, absent here): what the function
// would write to its two output files is kept in memory for the wrapper to hand back.  Not part of the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include <Eigen/Geometry>
#include <LBFGS.h>
#include <glog/logging.h>

#include <calibration/velocity.hpp>
#include <geometry/geometry.hpp>
#include <math/math.hpp>

#include <json.hpp>

namespace pilotguru {
std::vector<double> SmoothTimeSeries(const std::vector<double>& data_values, const std::vector<double>& data_timestamps,
                                     const std::vector<double>& target_timestamps, double sigma);
// include/io/json_converters.hpp:10-35
const char kVelocities[] = "velocities";
const char kSpeedMS[] = "speed_m_s";
const char kForwardAxis[] = "forward_axis";
const char kX[] = "x";
const char kY[] = "y";
const char kZ[] = "z";

struct PgrFitCapture {
  std::vector<long> times_usec;
  std::vector<double> values;
  double forward_axis[3] = {0, 0, 0};
  bool have_axis = false;
};
extern PgrFitCapture g_pgr_fit_capture;

inline void JsonWriteTimestampedRealData(const std::vector<long>& times_usec, const std::vector<double>& values,
                                         const std::string& /*filename*/, const std::string& /*root*/, const std::string& /*name*/) {
  g_pgr_fit_capture.times_usec = times_usec;
  g_pgr_fit_capture.values = values;
}
inline void WriteJsonFile(nlohmann::json& root, const std::string& /*filename*/) {
  nlohmann::json& a = root[kForwardAxis];
  g_pgr_fit_capture.forward_axis[0] = a[kX].num;
  g_pgr_fit_capture.forward_axis[1] = a[kY].num;
  g_pgr_fit_capture.forward_axis[2] = a[kZ].num;
  g_pgr_fit_capture.have_axis = true;
}
}  // namespace pilotguru
