// optical_trajectories -- the reference CLI (src/optical_trajectories.cc:36-62) with its per-frame hot path on the
// B200: ORB extraction of every frame (ORBextractor::operator(), Frame.cc:251-257) and Hamming projection matching
// against the previous frame (ORBmatcher::SearchByProjection, Tracking.cc:860-883) through libpgb200's C-ABI, and
// the reference's trajectory post-processing + JSON writer (track_image_sequence.cc:63-109) restated in trajectory.hpp.
//
// Scope: the SLAM back end (map initialisation, pose optimisation, local BA, loop
// closing, DBoW2 relocalisation) and video decoding are out of scope.  So this binary runs in FLOW-TRACKING mode:
//   * --in_video takes raw frames:  raw:<path>:<width>x<height> (8-bit gray, frame i at byte i*width*height) or
//     raw24:<path>:<width>x<height> (interleaved 24-bit colour, the format the reference's reader hands to the tracker;
//     channel order from Camera.RGB).  Flips and the colour conversion run on the device (pgb_frames_to_gray).
//   * the tracked quantity is the dominant image translation between consecutive frames (median displacement of the
//     matched keypoints); the camera is modelled as translating in its x-z plane by minus that flow, heading along
//     its motion.  Poses are therefore in pixel units, not metres -- monocular SLAM scale is arbitrary as well.
//   * tracking is "lost" (segment closed, new trajectory-<k>.json started, as the reference's outer loop does) when
//     fewer than 20 matches survive the 2*th retry.
// --vocabulary_file and --camera_settings keep the reference's CHECKs; the settings file supplies ORBextractor.* and
// Camera.fps (Tracking.cc:52-135).  --visualize and --output_per_segment_videos are accepted and ignored.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../include/pgb200.h"
#include "check.hpp"
#include "flags.hpp"
#include "trajectory.hpp"

namespace {

// "Key: value" lines of an OpenCV FileStorage YAML (%YAML:1.0), enough for the settings keys the tracker reads.
std::map<std::string, double> ReadSettings(const std::string& path) {
  std::ifstream f(path);
  PGB_CHECK(f.good()) << "cannot open camera settings " << path;
  std::map<std::string, double> kv;
  std::string line;
  while (std::getline(f, line)) {
    const size_t c = line.find(':');
    if (c == std::string::npos || line[0] == '%' || line[0] == '#') continue;
    std::string k = line.substr(0, c), v = line.substr(c + 1);
    k.erase(0, k.find_first_not_of(" \t")); k.erase(k.find_last_not_of(" \t") + 1);
    char* end = nullptr;
    const double d = strtod(v.c_str(), &end);
    if (end != v.c_str()) kv[k] = d;
  }
  return kv;
}

struct FrameFeats {
  std::vector<pgb_keypoint> k;
  std::vector<uint8_t> d;
};

}  // namespace

int main(int argc, char** argv) {
  std::string vocabulary_file, camera_settings, out_dir, in_video;
  bool visualize = true, vertical_flip = false, horizontal_flip = false, output_per_segment_videos = false;
  int64_t rotation_smooth_sigma = -1, device = 0, batch = 32;
  pgbhost::Flags flags;
  flags.String("vocabulary_file", &vocabulary_file, "ORB vocabulary file.");
  flags.String("camera_settings", &camera_settings, ".yml file with the camera calibration and ORB parameters.");
  flags.String("out_dir", &out_dir, "Directory to write trajectory-<segment>.json files to.");
  flags.String("in_video", &in_video, "Input frames: raw:<path>:<width>x<height> (8-bit gray).");
  flags.Bool("visualize", &visualize, "accepted, ignored");
  flags.Bool("vertical_flip", &vertical_flip, "Whether to flip input frames vertically.");
  flags.Bool("horizontal_flip", &horizontal_flip, "Whether to flip input frames horizontally.");
  flags.Bool("output_per_segment_videos", &output_per_segment_videos, "accepted, ignored");
  flags.Int64("rotation_smooth_sigma", &rotation_smooth_sigma, "Gaussian sigma (frames) for smoothing rotations; <0: none");
  flags.Int64("device", &device, "(extension) CUDA device");
  flags.Int64("batch", &batch, "(extension) frames per extraction batch");
  flags.Parse(argc, argv);
  PGB_CHECK(!vocabulary_file.empty());
  PGB_CHECK(!camera_settings.empty());
  PGB_CHECK(!in_video.empty());
  PGB_CHECK(batch >= 2);

  int width = 0, height = 0, channels = 1;
  char path[4096];
  if (sscanf(in_video.c_str(), "raw24:%4095[^:]:%dx%d", path, &width, &height) == 3) channels = 3;
  else
    PGB_CHECK(sscanf(in_video.c_str(), "raw:%4095[^:]:%dx%d", path, &width, &height) == 3)
        << "--in_video must be raw:<path>:<width>x<height> or raw24:<path>:<width>x<height> (video decoding is out of "
           "scope, see the header comment)";
  PGB_CHECK(width > 0 && height > 0);
  const auto cfg = ReadSettings(camera_settings);
  auto get = [&](const char* k, double dflt) { auto it = cfg.find(k); return it == cfg.end() ? dflt : it->second; };
  const int nfeatures = (int)get("ORBextractor.nFeatures", 1000), nlevels = (int)get("ORBextractor.nLevels", 8);
  const int ini_th = (int)get("ORBextractor.iniThFAST", 20), min_th = (int)get("ORBextractor.minThFAST", 7);
  const float scale_factor = (float)get("ORBextractor.scaleFactor", 1.2);
  double fps = get("Camera.fps", 30.0);
  if (fps == 0) fps = 30;  // Tracking.cc:86-88

  FILE* in = fopen(path, "rb");
  PGB_CHECK(in != nullptr) << "cannot open " << path;
  const size_t frame_bytes = (size_t)width * height, in_bytes = frame_bytes * channels;
  const int rgb_order = (int)get("Camera.RGB", 1.0);  // Tracking.cc:90-96

  const int B = (int)batch;
  pgb_orb* orb = pgb_orb_create((int)device, nfeatures, scale_factor, nlevels, ini_th, min_th, width, height, B, nullptr);
  PGB_CHECK(orb != nullptr) << pgb_last_error();
  const int cap = pgb_orb_max_keypoints(orb);
  pgb_matcher* matcher = pgb_matcher_create((int)device, 0.9f, 1, cap, B, nullptr);  // ORBmatcher(0.9, true), Tracking.cc:860
  PGB_CHECK(matcher != nullptr) << pgb_last_error();
  std::vector<float> sf(nlevels), inv(nlevels), s2(nlevels), is2(nlevels);
  PGB_CALL(pgb_orb_scale_factors(orb, sf.data(), inv.data(), s2.data(), is2.data()));

  std::vector<uint8_t> frames(frame_bytes * B), raw(in_bytes * B);
  std::vector<pgb_keypoint> kps((size_t)B * cap);
  std::vector<uint8_t> desc((size_t)B * cap * 32);
  std::vector<int32_t> counts(B);
  // matcher staging (pair p: current = frame p of the batch, queries = its predecessor)
  std::vector<pgb_keypoint> curK((size_t)B * cap);
  std::vector<uint8_t> curD((size_t)B * cap * 32), qD((size_t)B * cap * 32), qValid((size_t)B * cap);
  std::vector<float> qUV((size_t)B * cap * 2), qAng((size_t)B * cap);
  std::vector<int32_t> qOct((size_t)B * cap), curN(B), qN(B), matchOf((size_t)B * cap), nMatch(B);

  FrameFeats prev;  // last frame of the previous batch (empty before the first frame / after a lost segment)
  bool have_prev = false;
  double pos[3] = {0, 0, 0}, heading = 0.0, vflow[2] = {0, 0};
  std::vector<pgbhost::PoseWithTimestamp> trajectory;
  int segment_id = 0;
  int64_t frame_id = 0, total_matches = 0, total_kps = 0;
  auto close_segment = [&]() {
    if (trajectory.empty()) return;
    char name[64];
    snprintf(name, sizeof name, "/trajectory-%d.json", segment_id);
    const bool ok = pgbhost::FinishTrajectory(trajectory, (int)rotation_smooth_sigma, 0, out_dir + name, flags.verbose);
    if (flags.verbose) fprintf(stderr, "I segment %d: %zu poses%s\n", segment_id, trajectory.size(), ok ? "" : " (dropped)");
    trajectory.clear();
    segment_id++;
    pos[0] = pos[1] = pos[2] = 0; heading = 0; vflow[0] = vflow[1] = 0;
  };

  for (;;) {
    const size_t got = fread(raw.data(), in_bytes, B, in);
    if (got == 0) break;
    const int n = (int)got;
    // cv::flip (image_sequence_reader.cc:163-175) and cvtColor (Tracking.cc:243-258; OpenCV 2.4 fixed point) on the device
    PGB_CALL(pgb_frames_to_gray((int)device, raw.data(), 0, n, width, height, channels, rgb_order, (size_t)width * channels, in_bytes,
                                vertical_flip ? 1 : 0, horizontal_flip ? 1 : 0, 0, frames.data(), 0, width, frame_bytes, nullptr));
    PGB_CALL(pgb_orb_extract(orb, frames.data(), 0, n, width, height, width, frame_bytes, kps.data(), desc.data(), counts.data(), cap));
    // queries of pair p = keypoints of frame p-1 projected with the motion guess (TrackWithMotionModel's role); the
    // guess is zero motion: the th=15 / 30 px windows (times the octave scale) cover ordinary inter-frame flow
    int first = have_prev ? 0 : 1;
    for (int p = first; p < n; p++) {
      const pgb_keypoint* pk = p == 0 ? prev.k.data() : kps.data() + (size_t)(p - 1) * cap;
      const uint8_t* pd = p == 0 ? prev.d.data() : desc.data() + (size_t)(p - 1) * cap * 32;
      const int np = p == 0 ? (int)prev.k.size() : counts[p - 1];
      qN[p] = np; curN[p] = counts[p];
      for (int i = 0; i < np; i++) {
        const size_t o = (size_t)p * cap + i;
        qUV[2 * o] = pk[i].x + (float)vflow[0]; qUV[2 * o + 1] = pk[i].y + (float)vflow[1];
        qOct[o] = pk[i].octave; qAng[o] = pk[i].angle; qValid[o] = 1;
      }
      std::copy(pd, pd + (size_t)np * 32, qD.begin() + (size_t)p * cap * 32);
      std::copy(kps.begin() + (size_t)p * cap, kps.begin() + (size_t)p * cap + counts[p], curK.begin() + (size_t)p * cap);
      std::copy(desc.begin() + (size_t)p * cap * 32, desc.begin() + ((size_t)p * cap + counts[p]) * 32, curD.begin() + (size_t)p * cap * 32);
    }
    const int np = n - first;
    if (np > 0) {
      auto run = [&](float th, int p0, int cnt) {
        const size_t o = (size_t)p0 * cap;
        PGB_CALL(pgb_match_by_projection(matcher, cnt, cap, curK.data() + o, curD.data() + o * 32, curN.data() + p0, qUV.data() + 2 * o,
                                         qOct.data() + o, qAng.data() + o, qD.data() + o * 32, qValid.data() + o, qN.data() + p0, 0.f,
                                         (float)width, 0.f, (float)height, th, sf.data(), nlevels, matchOf.data() + o,
                                         nMatch.data() + p0, 0));
      };
      run(15.f, first, np);                                        // th = 15 for monocular, Tracking.cc:871-876
      for (int p = first; p < n; p++)
        if (nMatch[p] < 20) run(30.f, p, 1);                       // Tracking.cc:876-883
    }
    for (int p = 0; p < n; p++, frame_id++) {
      total_kps += counts[p];
      bool tracked = true;
      double dx = 0, dy = 0;
      if (p >= first) {
        const pgb_keypoint* pk = p == 0 ? prev.k.data() : kps.data() + (size_t)(p - 1) * cap;
        std::vector<float> fx, fy;
        for (int t = 0; t < counts[p]; t++) {
          const int q = matchOf[(size_t)p * cap + t];
          if (q < 0) continue;
          fx.push_back(kps[(size_t)p * cap + t].x - pk[q].x);
          fy.push_back(kps[(size_t)p * cap + t].y - pk[q].y);
        }
        total_matches += nMatch[p];
        tracked = nMatch[p] >= 20 && !fx.empty();
        if (tracked) {
          std::nth_element(fx.begin(), fx.begin() + fx.size() / 2, fx.end());
          std::nth_element(fy.begin(), fy.begin() + fy.size() / 2, fy.end());
          dx = fx[fx.size() / 2]; dy = fy[fy.size() / 2];
        }
      }
      if (!tracked) {  // LOST: close the segment; this frame starts the next one
        close_segment();
      }
      pos[0] -= dx; pos[2] -= dy;
      if (dx != 0 || dy != 0) heading = std::atan2(-dx, -dy);
      pgbhost::PoseWithTimestamp pw;
      pw.pose.t[0] = pos[0]; pw.pose.t[1] = pos[1]; pw.pose.t[2] = pos[2];
      pw.pose.qw = std::cos(heading * 0.5); pw.pose.qx = 0; pw.pose.qy = std::sin(heading * 0.5); pw.pose.qz = 0;
      pw.time_usec = (int64_t)std::llround((double)frame_id * 1e6 / fps);
      pw.is_lost = false;
      pw.frame_id = frame_id;
      trajectory.push_back(pw);
    }
    prev.k.assign(kps.begin() + (size_t)(n - 1) * cap, kps.begin() + (size_t)(n - 1) * cap + counts[n - 1]);
    prev.d.assign(desc.begin() + (size_t)(n - 1) * cap * 32, desc.begin() + ((size_t)(n - 1) * cap + counts[n - 1]) * 32);
    have_prev = true;
  }
  fclose(in);
  close_segment();
  if (flags.verbose)
    fprintf(stderr, "I %lld frames, %.1f keypoints/frame, %.1f matches/frame, %d segment(s)\n", (long long)frame_id,
            frame_id ? (double)total_kps / frame_id : 0.0, frame_id > 1 ? (double)total_matches / (frame_id - 1) : 0.0, segment_id);
  pgb_matcher_destroy(matcher);
  pgb_orb_destroy(orb);
  return EXIT_SUCCESS;
}
