// ORACLE -- TEST INFRASTRUCTURE ONLY.  Shadows the reference's include/io/json_converters.hpp (which needs ORB-SLAM2's
// System.h, OpenCV and nlohmann/json) for the one header that includes it on the path, interpolation/time_series.hpp:
// the field-name constants that header uses (json_converters.hpp:10-35) and ReadJsonFile, which here hands back a
// document the wrapper built in memory.  Not part of the product.
#pragma once
#include <cmath>    // the real header chain brings <cmath> in; time_series.hpp calls sqrt / erf unqualified
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include <json.hpp>

using std::vector;   // the real header chain brings this in; time_series.hpp writes `vector<T>` unqualified

namespace pilotguru {
const char kTimeUsec[] = "time_usec";
const char kFrameId[] = "frame_id";
const char kFrames[] = "frames";
std::unique_ptr<nlohmann::json> ReadJsonFile(const std::string& filename);
}  // namespace pilotguru
