// optical_trajectories -- the reference CLI (src/optical_trajectories.cc:36-62) with its per-frame hot path on B200s:
// ORB extraction of every frame (ORBextractor::operator(), Frame.cc:251-257) and Hamming projection matching against the
// previous frame (ORBmatcher::SearchByProjection with the th -> 2*th retry, Tracking.cc:860-883) through libpgb200's
// C-ABI, frames and features resident on the device, and the reference's trajectory post-processing + JSON writer
// (track_image_sequence.cc:63-109) restated in trajectory.hpp.
//
// --num_gpus N (extension; BASELINE configs[2]: 10 000 frames over 8 B200s): the frame sequence is cut into N contiguous
// blocks, one host thread + one GPU per block.  Extraction is independent per frame; the only dependency that crosses a
// block boundary is the match of a block's FIRST frame against the previous block's LAST frame: every rank packs that
// frame's features into a record, ONE NCCL all-gather (pgb_allgather_feats) hands each rank its left neighbour's, and
// the N-1 boundary pairs are matched.  The result is the same trajectory, byte for byte, as --num_gpus 1.
//
// Scope: the SLAM back end (map initialisation, local BA, loop closing, DBoW2
// relocalisation) and video decoding are out of scope.  So this binary runs in FLOW-TRACKING mode:
//   * --in_video takes a Motion-JPEG AVI file (demuxed and decoded on the device through pgb_video_*: the frames the
//     reference gets from VideoImageSequenceSource; the frame rate comes from the container), or raw frames:
//     raw:<path>:<width>x<height> (8-bit gray, frame i at byte i*width*height),
//     raw24:<path>:<width>x<height> (interleaved 24-bit colour, the format the reference's reader hands to the tracker;
//     channel order from Camera_RGB; flips and the colour conversion run on the device, pgb_frames_to_gray), or
//     synth:<canvas path>:<canvas width>x<canvas height>:<frames>:<width>x<height> -- the SURVEY 8(d) synthetic sequence
//     rendered on the device from a canvas file (pgb_synth_frames; stands where a hardware decoder would);
//   * the tracked quantity is the dominant image translation between consecutive frames (median displacement of the
//     matched keypoints); the camera is modelled as translating in its x-z plane by minus that flow, heading along
//     its motion.  Poses are therefore in pixel units, not metres -- monocular SLAM scale is arbitrary as well.
//   * tracking is "lost" (segment closed, new trajectory-<k>.json started, as the reference's outer loop does) when
//     fewer than 20 matches survive the 2*th retry.
// --vocabulary_file and --camera_settings keep the reference's CHECKs.  The settings file is the fork's format
// (written by src/calibrate.cc:504-544, read by Tracking.cc:52-135): underscore keys -- Camera_fps, Camera_RGB,
// ORBextractor_nFeatures / _scaleFactor / _nLevels / _iniThFAST / _minThFAST.  A file with none of the ORBextractor_*
// keys is rejected (upstream ORB-SLAM2's dotted keys are accepted with a warning).  --visualize and
// --output_per_segment_videos are accepted and ignored.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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

struct Source {
  enum Kind { kRawGray, kRaw24, kSynth, kAvi } kind = kRawGray;
  double fps = 0;  // kAvi: the container's frame rate (timestamps = pts * time_base, image_sequence_reader.cc:153-155)
  std::string path;
  int width = 0, height = 0, channels = 1;
  int canvasW = 0, canvasH = 0;
  int64_t frames = 0;
};

Source ParseSource(const std::string& spec) {
  Source s;
  char path[4096];
  long long n = 0;
  if (spec.size() > 4 && spec.compare(0, 4, "raw:") != 0 && spec.compare(0, 6, "raw24:") != 0 && spec.compare(0, 6, "synth:") != 0) {
    // a video FILE, as the reference takes it (VideoImageSequenceSource, image_sequence_reader.cc:74-120): demuxed and
    // decoded through libpgb200 (Motion-JPEG AVI; any other container / codec is refused there with its name)
    pgb_video* v = pgb_video_open(0, spec.c_str());
    PGB_CHECK(v != nullptr) << pgb_last_error()
                            << " (--in_video takes a Motion-JPEG .avi file, raw:<path>:<w>x<h>, raw24:<path>:<w>x<h> or "
                               "synth:<canvas>:<cw>x<ch>:<frames>:<w>x<h>)";
    int rot = 0;
    PGB_CALL(pgb_video_info(v, &s.width, &s.height, &s.frames, &s.fps, &rot));
    pgb_video_close(v);
    s.kind = Source::kAvi; s.channels = 3; s.path = spec;
    return s;
  }
  if (sscanf(spec.c_str(), "synth:%4095[^:]:%dx%d:%lld:%dx%d", path, &s.canvasW, &s.canvasH, &n, &s.width, &s.height) == 6) {
    s.kind = Source::kSynth; s.frames = n;
  } else if (sscanf(spec.c_str(), "raw24:%4095[^:]:%dx%d", path, &s.width, &s.height) == 3) {
    s.kind = Source::kRaw24; s.channels = 3;
  } else {
    PGB_CHECK(sscanf(spec.c_str(), "raw:%4095[^:]:%dx%d", path, &s.width, &s.height) == 3)
        << "--in_video must be a Motion-JPEG .avi file, raw:<path>:<w>x<h>, raw24:<path>:<w>x<h> or "
           "synth:<canvas>:<cw>x<ch>:<frames>:<w>x<h> (see the header comment)";
  }
  s.path = path;
  PGB_CHECK(s.width > 0 && s.height > 0);
  if (s.kind != Source::kSynth) {
    FILE* f = fopen(path, "rb");
    PGB_CHECK(f != nullptr) << "cannot open " << path;
    fseeko(f, 0, SEEK_END);
    s.frames = (int64_t)(ftello(f) / ((off_t)s.width * s.height * s.channels));
    fclose(f);
  }
  return s;
}

struct Config {
  int nfeatures, nlevels, ini_th, min_th, rgb_order;
  float scale_factor;
  double fps;
  bool vertical_flip, horizontal_flip;
  int batch;
};

// What the trajectory builder needs from one frame.
struct FrameResult {
  int32_t n_kps = 0, n_matches = -1;  // n_matches = -1: the frame has no predecessor (first frame of the sequence)
  float dx = 0, dy = 0;               // median displacement of the matched keypoints (valid when n_matches >= 20)
  bool tracked = false;
};

// One rank: frames [t0, t1) on `device`.  Features live in a device region of batch+1 slots (slot 0 = predecessor).
struct Rank {
  int device = 0, rank = 0;
  int64_t t0 = 0, t1 = 0;
  pgb_comm* comm = nullptr;
  const Source* src = nullptr;
  const Config* cfg = nullptr;
  FrameResult* out = nullptr;  // out[t - 0], the whole sequence's array
  double seconds = 0;

  void Run() {
    const int B = cfg->batch, W = src->width, H = src->height;
    const size_t frameBytes = (size_t)W * H, inBytes = frameBytes * src->channels;
    pgb_orb* orb = pgb_orb_create(device, cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th, cfg->min_th, W, H, B, nullptr);
    PGB_CHECK(orb != nullptr) << pgb_last_error();
    void* st = pgb_orb_stream(orb);
    const int cap = pgb_orb_max_keypoints(orb);
    pgb_matcher* matcher = pgb_matcher_create(device, 0.9f, 1, cap, B, st);  // ORBmatcher(0.9, true), Tracking.cc:860
    PGB_CHECK(matcher != nullptr) << pgb_last_error();
    std::vector<float> sf(cfg->nlevels), inv(cfg->nlevels), s2(cfg->nlevels), is2(cfg->nlevels);
    PGB_CALL(pgb_orb_scale_factors(orb, sf.data(), inv.data(), s2.data(), is2.data()));

    auto dalloc = [&](size_t bytes) { void* p = pgb_device_malloc(device, bytes); PGB_CHECK(p != nullptr) << pgb_last_error(); return p; };
    pgb_keypoint* dK = (pgb_keypoint*)dalloc((size_t)(B + 1) * cap * sizeof(pgb_keypoint));
    uint8_t* dD = (uint8_t*)dalloc((size_t)(B + 1) * cap * 32);
    int32_t* dN = (int32_t*)dalloc((size_t)(B + 1) * sizeof(int32_t));
    int32_t* dMatch = (int32_t*)dalloc((size_t)B * cap * sizeof(int32_t));
    int32_t* dNm = (int32_t*)dalloc((size_t)B * sizeof(int32_t));
    float* dFlow = (float*)dalloc((size_t)B * 2 * sizeof(float));  // zero motion guess (TrackWithMotionModel's role)
    uint8_t* dGray = (uint8_t*)dalloc(frameBytes * B);
    const size_t recBytes = pgb_frame_record_bytes(cap);
    uint8_t* dFirstRec = (uint8_t*)dalloc(recBytes);   // this block's first frame, kept for the boundary pair
    uint8_t* dLastRec = (uint8_t*)dalloc(recBytes);
    uint8_t* dAllRec = comm ? (uint8_t*)dalloc(recBytes * pgb_comm_size(comm)) : nullptr;
    uint8_t* dCanvas = nullptr;
    uint8_t* hRaw = nullptr;
    uint8_t* dRgb = nullptr;
    pgb_video* video = nullptr;
    int fd = -1;
    if (src->kind == Source::kAvi) {
      video = pgb_video_open(device, src->path.c_str());  // one demuxer per rank: Motion-JPEG frames decode independently
      PGB_CHECK(video != nullptr) << pgb_last_error();
      dRgb = (uint8_t*)dalloc(inBytes * B);
    } else if (src->kind == Source::kSynth) {
      std::vector<uint8_t> canvas((size_t)src->canvasW * src->canvasH);
      FILE* f = fopen(src->path.c_str(), "rb");
      PGB_CHECK(f != nullptr) << "cannot open canvas " << src->path;
      PGB_CHECK(fread(canvas.data(), 1, canvas.size(), f) == canvas.size()) << "canvas file is smaller than " << src->canvasW << "x" << src->canvasH;
      fclose(f);
      dCanvas = (uint8_t*)dalloc(canvas.size());
      PGB_CALL(pgb_memcpy_async(device, dCanvas, canvas.data(), canvas.size(), 0, st));
      PGB_CALL(pgb_stream_synchronize(device, st));
    } else {
      hRaw = (uint8_t*)pgb_host_malloc_pinned(inBytes * B);
      PGB_CHECK(hRaw != nullptr) << pgb_last_error();
      fd = open(src->path.c_str(), O_RDONLY);
      PGB_CHECK(fd >= 0) << "cannot open " << src->path;
    }
    // What the trajectory builder needs from a pair -- the median displacement of its matched keypoints -- is computed on
    // the device (pgb_match_median_flow); a batch sends back 16 bytes per frame instead of every keypoint and match.  Two
    // sets of pinned result buffers and an event each: batch k + 1 is issued before the host looks at batch k's results,
    // so the stream never drains between batches.
    float* dFlowOut = (float*)dalloc((size_t)B * 2 * sizeof(float));
    int32_t* dTracked = (int32_t*)dalloc((size_t)B * sizeof(int32_t));
    struct HostSet {
      int32_t* N; int32_t* Nm; float* Flow; int32_t* Tracked; void* ev;
      int n = 0, first = 0; int64_t b0 = 0; bool pending = false;
    } hs[2];
    for (HostSet& h : hs) {
      h.N = (int32_t*)pgb_host_malloc_pinned((size_t)(B + 1) * sizeof(int32_t));
      h.Nm = (int32_t*)pgb_host_malloc_pinned((size_t)B * sizeof(int32_t));
      h.Flow = (float*)pgb_host_malloc_pinned((size_t)B * 2 * sizeof(float));
      h.Tracked = (int32_t*)pgb_host_malloc_pinned((size_t)B * sizeof(int32_t));
      h.ev = pgb_event_create(device);
      PGB_CHECK(h.N && h.Nm && h.Flow && h.Tracked && h.ev) << pgb_last_error();
    }
    auto finish_set = [&](HostSet& h) {
      PGB_CALL(pgb_event_synchronize(device, h.ev));
      for (int i = 0; i < h.n; i++) {  // frame i of the batch sits in slot i + 1; its pair (slot i, slot i + 1) is matched pair i - first
        FrameResult& r = out[h.b0 + i];
        r.n_kps = h.N[i + 1];
        if (i >= h.first) {
          const int p = i - h.first;
          r.n_matches = h.Nm[p];
          r.tracked = h.Tracked[p] != 0;
          if (r.tracked) { r.dx = h.Flow[2 * p]; r.dy = h.Flow[2 * p + 1]; }
        }
      }
      h.pending = false;
    };

    const auto wall0 = std::chrono::steady_clock::now();
    bool have_prev = false;
    int k = 0;
    for (int64_t b0 = t0; b0 < t1; b0 += B, k++) {
      const int n = (int)std::min<int64_t>(B, t1 - b0);
      HostSet& h = hs[k & 1];
      if (h.pending) finish_set(h);  // the batch before last: its results have long arrived
      // host-resident sources refill one pinned input buffer: the previous batch's upload must have left it
      if (hRaw && hs[(k + 1) & 1].pending) PGB_CALL(pgb_event_synchronize(device, hs[(k + 1) & 1].ev));
      // ---- frames -> gray on the device -> features in slots 1..n
      if (src->kind == Source::kAvi) {
        // decode (nvJPEG) -> RGB24 on the device -> cv::flip + cvtColor (pgb_frames_to_gray: the decoder hands out RGB order,
        // image_sequence_reader.cc:157-170) -> extractor; nothing crosses PCIe but the compressed frames
        PGB_CALL(pgb_video_read_rgb(video, b0, n, dRgb, (size_t)W * 3, inBytes, nullptr, st));
        PGB_CALL(pgb_frames_to_gray(device, dRgb, 1, n, W, H, 3, 1, (size_t)W * 3, inBytes, cfg->vertical_flip ? 1 : 0,
                                    cfg->horizontal_flip ? 1 : 0, 0, dGray, 1, W, frameBytes, st));
        PGB_CALL(pgb_orb_extract(orb, dGray, PGB_IN_DEVICE | PGB_OUT_DEVICE, n, W, H, W, frameBytes, dK + cap, dD + (size_t)cap * 32, dN + 1, cap));
      } else if (src->kind == Source::kSynth) {
        PGB_CALL(pgb_synth_frames(device, dCanvas, src->canvasW, src->canvasH, (int)b0, n, W, H, dGray, st));
        PGB_CALL(pgb_orb_extract(orb, dGray, PGB_IN_DEVICE | PGB_OUT_DEVICE, n, W, H, W, frameBytes, dK + cap, dD + (size_t)cap * 32, dN + 1, cap));
      } else {
        const ssize_t want = (ssize_t)(inBytes * n);
        PGB_CHECK(pread(fd, hRaw, want, (off_t)(b0 * (int64_t)inBytes)) == want) << "short read from " << src->path;
        if (src->kind == Source::kRaw24 || cfg->vertical_flip || cfg->horizontal_flip) {
          // cv::flip (image_sequence_reader.cc:163-175) and cvtColor (Tracking.cc:243-258; OpenCV 2.4 fixed point) on the device
          PGB_CALL(pgb_frames_to_gray(device, hRaw, 0, n, W, H, src->channels, cfg->rgb_order, (size_t)W * src->channels, inBytes,
                                      cfg->vertical_flip ? 1 : 0, cfg->horizontal_flip ? 1 : 0, 0, dGray, 1, W, frameBytes, st));
          PGB_CALL(pgb_orb_extract(orb, dGray, PGB_IN_DEVICE | PGB_OUT_DEVICE, n, W, H, W, frameBytes, dK + cap, dD + (size_t)cap * 32, dN + 1, cap));
        } else {
          PGB_CALL(pgb_orb_extract(orb, hRaw, PGB_OUT_DEVICE, n, W, H, W, frameBytes, dK + cap, dD + (size_t)cap * 32, dN + 1, cap));
        }
      }
      if (b0 == t0 && rank > 0) PGB_CALL(pgb_frame_record_pack(dK, dD, dN, 1, cap, dFirstRec, st));
      // ---- pairs (slot p, slot p + 1): all n when slot 0 holds the predecessor, else n - 1 starting at slot 1
      const int first = have_prev ? 0 : 1, np = n - first;
      if (np > 0)
        PGB_CALL(pgb_match_consecutive(matcher, np, cap, dK + (size_t)first * cap, dD + (size_t)first * cap * 32, dN + first, dFlow,
                                       (float)W, (float)H, 15.f, sf.data(), cfg->nlevels, dMatch, dNm));
      if (np > 0)
        PGB_CALL(pgb_match_median_flow(device, np, cap, dK + (size_t)first * cap, dN + first, dMatch, dNm, dFlowOut, dTracked, st));
      PGB_CALL(pgb_memcpy_async(device, h.N, dN, (size_t)(n + 1) * sizeof(int32_t), 1, st));
      if (np > 0) {
        PGB_CALL(pgb_memcpy_async(device, h.Nm, dNm, (size_t)np * sizeof(int32_t), 1, st));
        PGB_CALL(pgb_memcpy_async(device, h.Flow, dFlowOut, (size_t)np * 2 * sizeof(float), 1, st));
        PGB_CALL(pgb_memcpy_async(device, h.Tracked, dTracked, (size_t)np * sizeof(int32_t), 1, st));
      }
      // slot 0 <- this batch's last frame, for the next batch (stream-ordered after the copies above)
      PGB_CALL(pgb_frame_record_pack(dK, dD, dN, n, cap, dLastRec, st));
      PGB_CALL(pgb_frame_record_unpack(dLastRec, dK, dD, dN, 0, cap, st));
      PGB_CALL(pgb_event_record(device, h.ev, st));
      h.n = n; h.first = first; h.b0 = b0; h.pending = true;
      have_prev = true;
    }
    for (int q = 0; q < 2; q++)
      if (hs[(k + q) & 1].pending) finish_set(hs[(k + q) & 1]);  // the older of the two first
    PGB_CALL(pgb_orb_check(orb));  // synchronises the stream, surfaces device-side capacity flags of every batch
    int32_t* hN = hs[0].N; int32_t* hNm = hs[0].Nm; float* hFlow = hs[0].Flow; int32_t* hTracked = hs[0].Tracked;
    // ---- block boundary: ONE all-gather of every rank's last-frame record; rank r > 0 matches its first frame against
    //      the left neighbour's last frame
    if (comm) {
      PGB_CALL(pgb_allgather_feats(comm, dLastRec, dAllRec, recBytes, st));
      if (rank > 0 && t1 > t0) {
        // (every rank owns frames: main() refuses --num_gpus beyond what the frame count can use)
        PGB_CALL(pgb_frame_record_unpack(dAllRec + (size_t)(rank - 1) * recBytes, dK, dD, dN, 0, cap, st));
        PGB_CALL(pgb_frame_record_unpack(dFirstRec, dK, dD, dN, 1, cap, st));
        PGB_CALL(pgb_match_consecutive(matcher, 1, cap, dK, dD, dN, dFlow, (float)W, (float)H, 15.f, sf.data(), cfg->nlevels, dMatch, dNm));
        PGB_CALL(pgb_match_median_flow(device, 1, cap, dK, dN, dMatch, dNm, dFlowOut, dTracked, st));
        PGB_CALL(pgb_memcpy_async(device, hN, dN, 2 * sizeof(int32_t), 1, st));
        PGB_CALL(pgb_memcpy_async(device, hNm, dNm, sizeof(int32_t), 1, st));
        PGB_CALL(pgb_memcpy_async(device, hFlow, dFlowOut, 2 * sizeof(float), 1, st));
        PGB_CALL(pgb_memcpy_async(device, hTracked, dTracked, sizeof(int32_t), 1, st));
        PGB_CALL(pgb_orb_check(orb));
        FrameResult& r = out[t0];
        r.n_matches = hNm[0];
        r.tracked = hTracked[0] != 0;
        if (r.tracked) { r.dx = hFlow[0]; r.dy = hFlow[1]; }
      } else {
        PGB_CALL(pgb_stream_synchronize(device, st));
      }
    }
    seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    if (fd >= 0) close(fd);
    pgb_video_close(video);
    pgb_device_free(device, dRgb);
    pgb_host_free_pinned(hRaw);
    for (HostSet& h : hs) { pgb_host_free_pinned(h.N); pgb_host_free_pinned(h.Nm); pgb_host_free_pinned(h.Flow); pgb_host_free_pinned(h.Tracked); pgb_event_destroy(device, h.ev); }
    for (void* p : {(void*)dK, (void*)dD, (void*)dN, (void*)dMatch, (void*)dNm, (void*)dFlow, (void*)dGray, (void*)dFirstRec, (void*)dLastRec,
                    (void*)dAllRec, (void*)dCanvas, (void*)dFlowOut, (void*)dTracked})
      pgb_device_free(device, p);
    pgb_matcher_destroy(matcher);
    pgb_orb_destroy(orb);
  }
};

}  // namespace

int main(int argc, char** argv) {
  std::string vocabulary_file, camera_settings, out_dir, in_video;
  bool visualize = true, vertical_flip = false, horizontal_flip = false, output_per_segment_videos = false;
  int64_t rotation_smooth_sigma = -1, device = 0, batch = 64, num_gpus = 1;
  pgbhost::Flags flags;
  flags.String("vocabulary_file", &vocabulary_file, "ORB vocabulary file.");
  flags.String("camera_settings", &camera_settings, ".yml file with the camera calibration and ORB parameters.");
  flags.String("out_dir", &out_dir, "Directory to write trajectory-<segment>.json files to.");
  flags.String("in_video", &in_video, "Input frames: raw:<path>:<w>x<h>, raw24:<path>:<w>x<h> or synth:<canvas>:<cw>x<ch>:<frames>:<w>x<h>.");
  flags.Bool("visualize", &visualize, "accepted, ignored");
  flags.Bool("vertical_flip", &vertical_flip, "Whether to flip input frames vertically.");
  flags.Bool("horizontal_flip", &horizontal_flip, "Whether to flip input frames horizontally.");
  flags.Bool("output_per_segment_videos", &output_per_segment_videos, "accepted, ignored");
  flags.Int64("rotation_smooth_sigma", &rotation_smooth_sigma, "Gaussian sigma (frames) for smoothing rotations; <0: none");
  flags.Int64("device", &device, "(extension) first CUDA device");
  flags.Int64("num_gpus", &num_gpus, "(extension) GPUs to shard the frames over (devices device .. device+num_gpus-1)");
  flags.Int64("batch", &batch, "(extension) frames per extraction batch");
  flags.Parse(argc, argv);
  PGB_CHECK(!vocabulary_file.empty());
  PGB_CHECK(!camera_settings.empty());
  PGB_CHECK(!in_video.empty());
  PGB_CHECK(batch >= 2);
  PGB_CHECK(num_gpus >= 1);

  const Source src = ParseSource(in_video);
  const auto kv = ReadSettings(camera_settings);
  // The fork's keys (Tracking.cc:80,102,131-135; written by src/calibrate.cc:504-544).  Upstream ORB-SLAM2 spells them
  // with a dot: accepted, with a warning, so that stock example files still work.
  bool dotted = false;
  auto get = [&](const char* stem, const char* field, double dflt) {
    auto it = kv.find(std::string(stem) + "_" + field);
    if (it != kv.end()) return it->second;
    it = kv.find(std::string(stem) + "." + field);
    if (it != kv.end()) { dotted = true; return it->second; }
    return dflt;
  };
  const bool any_orb = kv.count("ORBextractor_nFeatures") || kv.count("ORBextractor_scaleFactor") || kv.count("ORBextractor_nLevels") ||
                       kv.count("ORBextractor_iniThFAST") || kv.count("ORBextractor_minThFAST") || kv.count("ORBextractor.nFeatures") ||
                       kv.count("ORBextractor.scaleFactor") || kv.count("ORBextractor.nLevels") || kv.count("ORBextractor.iniThFAST") ||
                       kv.count("ORBextractor.minThFAST");
  PGB_CHECK(any_orb) << "the settings file " << camera_settings << " has none of the ORBextractor_* keys (nFeatures, scaleFactor, nLevels, "
                        "iniThFAST, minThFAST) that Tracking.cc:131-135 reads; refusing to run on defaults";
  Config cfg;
  cfg.nfeatures = (int)get("ORBextractor", "nFeatures", 1000);
  cfg.nlevels = (int)get("ORBextractor", "nLevels", 8);
  cfg.ini_th = (int)get("ORBextractor", "iniThFAST", 20);
  cfg.min_th = (int)get("ORBextractor", "minThFAST", 7);
  cfg.scale_factor = (float)get("ORBextractor", "scaleFactor", 1.2);
  cfg.fps = get("Camera", "fps", 30.0);
  if (cfg.fps == 0) cfg.fps = 30;  // Tracking.cc:103-104
  cfg.rgb_order = (int)get("Camera", "RGB", 1.0);  // Tracking.cc:114-120
  cfg.vertical_flip = vertical_flip; cfg.horizontal_flip = horizontal_flip;
  cfg.batch = (int)batch;
  if (dotted) fprintf(stderr, "W settings file uses upstream ORB-SLAM2's dotted keys (Camera.fps ...); this fork writes Camera_fps ...\n");

  const int N = (int)num_gpus;
  PGB_CHECK(pgb_device_count() >= (int)device + N) << "--num_gpus " << N << " from device " << device << ": only " << pgb_device_count() << " CUDA device(s)";
  std::vector<FrameResult> results((size_t)std::max<int64_t>(src.frames, 1));
  std::vector<pgb_comm*> comms(N, nullptr);
  if (N > 1) {
    std::vector<int> devs(N);
    for (int i = 0; i < N; i++) devs[i] = (int)device + i;
    PGB_CALL(pgb_comm_create_all(N, devs.data(), comms.data()));
  }
  std::vector<Rank> ranks(N);
  const int64_t per = (src.frames + N - 1) / N;  // contiguous blocks (pilotguru_b200/dist.py: shard_range)
  for (int r = 0; r < N; r++) {
    ranks[r].device = (int)device + r; ranks[r].rank = r; ranks[r].comm = comms[r];
    ranks[r].t0 = std::min<int64_t>(r * per, src.frames); ranks[r].t1 = std::min<int64_t>(ranks[r].t0 + per, src.frames);
    ranks[r].src = &src; ranks[r].cfg = &cfg; ranks[r].out = results.data();
  }
  if (N > 1) PGB_CHECK(ranks[N - 1].t1 > ranks[N - 1].t0) << "--num_gpus " << N << " exceeds what " << src.frames << " frames can use";
  const auto wall0 = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (int r = 1; r < N; r++) th.emplace_back([&ranks, r] { ranks[r].Run(); });
    ranks[0].Run();
    for (auto& t : th) t.join();
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
  for (pgb_comm* c : comms) pgb_comm_destroy(c);
  const auto post0 = std::chrono::steady_clock::now();

  // ---- the sequential part: accumulate the per-pair flows into poses, cut segments where tracking was lost
  double pos[3] = {0, 0, 0}, heading = 0.0;
  std::vector<pgbhost::PoseWithTimestamp> trajectory;
  int segment_id = 0;
  int64_t total_matches = 0, total_kps = 0;
  auto close_segment = [&]() {
    if (trajectory.empty()) return;
    char name[64];
    snprintf(name, sizeof name, "/trajectory-%d.json", segment_id);
    const bool ok = pgbhost::FinishTrajectory(trajectory, (int)rotation_smooth_sigma, 0, out_dir + name, flags.verbose);
    if (flags.verbose) fprintf(stderr, "I segment %d: %zu poses%s\n", segment_id, trajectory.size(), ok ? "" : " (dropped)");
    trajectory.clear();
    segment_id++;
    pos[0] = pos[1] = pos[2] = 0; heading = 0;
  };
  for (int64_t t = 0; t < src.frames; t++) {
    const FrameResult& r = results[t];
    total_kps += r.n_kps;
    double dx = 0, dy = 0;
    if (r.n_matches >= 0) {
      total_matches += r.n_matches;
      if (r.tracked) { dx = r.dx; dy = r.dy; }
      else close_segment();  // LOST: this frame starts the next segment
    }
    pos[0] -= dx; pos[2] -= dy;
    if (dx != 0 || dy != 0) heading = std::atan2(-dx, -dy);
    pgbhost::PoseWithTimestamp pw;
    pw.pose.t[0] = pos[0]; pw.pose.t[1] = pos[1]; pw.pose.t[2] = pos[2];
    pw.pose.qw = std::cos(heading * 0.5); pw.pose.qx = 0; pw.pose.qy = std::sin(heading * 0.5); pw.pose.qz = 0;
    // a video file carries its own time base (timestamp = pts * time_base, image_sequence_reader.cc:153-155); raw frames
    // are stamped with the settings file's Camera_fps
    pw.time_usec = (int64_t)std::llround((double)t * 1e6 / (src.kind == Source::kAvi ? src.fps : cfg.fps));
    pw.is_lost = false;
    pw.frame_id = t;
    trajectory.push_back(pw);
  }
  close_segment();
  if (flags.verbose) {
    fprintf(stderr, "I %lld frames, %.1f keypoints/frame, %.1f matches/frame, %d segment(s)\n", (long long)src.frames,
            src.frames ? (double)total_kps / src.frames : 0.0, src.frames > 1 ? (double)total_matches / (src.frames - 1) : 0.0, segment_id);
    fprintf(stderr, "I extract+match: %.3f s on %d GPU(s) = %.0f frames/s (per-rank seconds:", wall, N, src.frames / std::max(wall, 1e-9));
    for (int r = 0; r < N; r++) fprintf(stderr, " %.3f", ranks[r].seconds);
    fprintf(stderr, "); totals: %lld keypoints, %lld matches\n", (long long)total_kps, (long long)total_matches);
    fprintf(stderr, "I trajectory post-processing + JSON: %.3f s\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - post0).count());
  }
  return EXIT_SUCCESS;
}
