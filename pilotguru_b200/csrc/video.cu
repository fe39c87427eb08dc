// Frame feed, decode part: what VideoImageSequenceSource does
// (src/io/image_sequence_reader.cc:74-208: open the container, find the video stream, decode a frame, convert it to
// RGB24, stamp it with best-effort-pts * time_base) for Motion-JPEG AVI files, with the decoded RGB frames appearing
// in DEVICE memory where pgb_frames_to_gray_rotated / pgb_orb_extract pick them up.
//
// Scope.  The reference decodes through libavformat / libavcodec (any codec).  This image has neither libav nor the
// NVDEC headers; what it has is nvJPEG (CUDA toolkit library, like cuBLAS: not on the hot path, which starts at the gray
// frame).  Motion-JPEG is the one codec whose frames are independent JPEG images, so:
//   * the container side (RIFF 'AVI ' + OpenDML 'AVIX' segments, 'hdrl' / 'strl' headers, the 'movi' chunk walk incl.
//     'rec ' lists, stream selection = first 'vids' stream as VideoStreamIndexOrDie does, frame timestamps
//     = frame_index * dwScale / dwRate, i.e. pts * time_base of an AVI stream) is restated here in plain C++;
//   * every frame is decoded by nvjpegDecode to interleaved RGB (the RGB24 raw_frame_image_ of the reference) on the
//     caller's stream;
//   * any other codec is refused with its FOURCC in the message (no silent fallback).
// nvJPEG is bound at run time (dlopen), like NCCL in comm.cu: libpgb200.so has no link-time dependency on it, opening
// and indexing a file needs no GPU, and the library is only loaded when a frame is decoded.
// Decoder parity: JPEG decoders are allowed to differ in the last bit of the IDCT and in chroma upsampling, so a decoded
// frame is not bit-defined by the reference either (libavcodec's mjpeg decoder + swscale vs libjpeg-turbo vs nvJPEG);
// tests/test_gpu_video.py holds the GRAY frames (what the extractor sees) within 3 grey levels of cv2's decode of the
// same file and the container walk identical to an independent Python walker and to cv2.VideoCapture's frame count / fps.
#include <dlfcn.h>
#include <fcntl.h>
#include <nvjpeg.h>
#include <strings.h>
#include <sys/stat.h>
#include <unistd.h>

#include <mutex>
#include <vector>

#include "../../include/pgb200_jpeg_std_dht.h"
#include "common.cuh"

namespace {

// The typical Huffman tables of ITU-T T.81 Annex K.3 as DHT segments: Motion-JPEG frames in AVI files may omit theirs
// ("AVI1" frames of capture hardware; libavcodec's mjpeg decoder falls back to these tables as well).
const unsigned char kStdDht[PGB200_JPEG_STD_DHT_BYTES] = PGB200_JPEG_STD_DHT_INIT;

// Offset of the first SOS marker of a JPEG image and whether a DHT segment precedes it; false if the header is malformed.
bool jpeg_header_scan(const uint8_t* j, size_t n, size_t* sosAt, bool* hasDht) {
  if (n < 4 || j[0] != 0xFF || j[1] != 0xD8) return false;
  size_t i = 2;
  *hasDht = false;
  while (i + 4 <= n) {
    if (j[i] != 0xFF) return false;
    const uint8_t m = j[i + 1];
    if (m == 0xFF) { i++; continue; }                                   // fill byte
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { i += 2; continue; }  // markers without a length
    if (m == 0xDA) { *sosAt = i; return true; }
    if (m == 0xC4) *hasDht = true;
    const size_t len = ((size_t)j[i + 2] << 8) | j[i + 3];
    if (len < 2) return false;
    i += 2 + len;
  }
  return false;
}

struct NvjpegApi {
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  void* handle = nullptr;
  std::string error;
};

NvjpegApi* nvjpeg_api() {
  static NvjpegApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
                           "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12", "/usr/local/cuda-12.9/targets/x86_64-linux/lib/libnvjpeg.so.12"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load libnvjpeg.so.12: ") + dlerror(); return; }
    auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("libnvjpeg lacks ") + s; return p; };
    api.CreateSimple = (decltype(api.CreateSimple))sym("nvjpegCreateSimple");
    api.Destroy = (decltype(api.Destroy))sym("nvjpegDestroy");
    api.JpegStateCreate = (decltype(api.JpegStateCreate))sym("nvjpegJpegStateCreate");
    api.JpegStateDestroy = (decltype(api.JpegStateDestroy))sym("nvjpegJpegStateDestroy");
    api.GetImageInfo = (decltype(api.GetImageInfo))sym("nvjpegGetImageInfo");
    api.Decode = (decltype(api.Decode))sym("nvjpegDecode");
  });
  return &api;
}

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
bool is4(const uint8_t* p, const char* cc) { return memcmp(p, cc, 4) == 0; }

}  // namespace

struct pgb_video {
  int device = 0;
  int fd = -1;
  std::string path;
  uint64_t fileSize = 0;
  int width = 0, height = 0;
  int stream = -1;            // index of the first 'vids' stream (VideoStreamIndexOrDie, image_sequence_reader.cc:63-71)
  uint32_t scale = 0, rate = 0;
  char fourcc[5] = {0, 0, 0, 0, 0};
  struct Frame { uint64_t off; uint32_t size; };
  std::vector<Frame> frames;
  std::vector<uint8_t> bits;  // bitstreams of the frames of the decode call in flight (alive until the next call)
  nvjpegHandle_t nj = nullptr;
  nvjpegJpegState_t st = nullptr;
};

namespace {

bool pread_all(int fd, void* dst, size_t n, uint64_t off) {
  uint8_t* p = (uint8_t*)dst;
  while (n > 0) {
    const ssize_t r = pread(fd, p, n, (off_t)off);
    if (r <= 0) return false;
    p += r; n -= (size_t)r; off += (uint64_t)r;
  }
  return true;
}

// Walks the chunks of [off, end): headers fill the stream description, 'movi' data chunks of the selected stream are
// appended to the frame index.  Returns false (g_last_error set) on a malformed file.
bool walk(pgb_video* v, uint64_t off, uint64_t end, int depth, int* nextStream, bool inMovi) {
  uint8_t h[12];
  while (off + 8 <= end) {
    if (!pread_all(v->fd, h, 8, off)) { pgb::fail(PGB_ERR_INVALID, "%s: short read at offset %llu", v->path.c_str(), (unsigned long long)off); return false; }
    uint64_t size = rd32(h + 4);
    const uint64_t body = off + 8;
    if (body + size > end) {
      // a recording that was cut short (the sizes in the list headers still describe the full file): lists are walked
      // as far as they go, an incomplete frame is dropped, an incomplete header is an error
      if (is4(h, "LIST") || is4(h, "RIFF")) size = end - body;
      else if (inMovi) return true;
      else {
        pgb::fail(PGB_ERR_INVALID, "%s: chunk '%.4s' at %llu runs past the end of the file", v->path.c_str(), (const char*)h, (unsigned long long)off);
        return false;
      }
    }
    if (is4(h, "LIST") || is4(h, "RIFF")) {
      if (size < 4 || !pread_all(v->fd, h + 8, 4, body)) { pgb::fail(PGB_ERR_INVALID, "%s: bad list chunk", v->path.c_str()); return false; }
      if (depth > 8) { pgb::fail(PGB_ERR_INVALID, "%s: lists nested too deep", v->path.c_str()); return false; }
      const bool movi = is4(h + 8, "movi") || (inMovi && is4(h + 8, "rec "));
      if (is4(h + 8, "hdrl") || is4(h + 8, "strl") || movi || is4(h + 8, "AVI ") || is4(h + 8, "AVIX")) {
        if (!walk(v, body + 4, body + size, depth + 1, nextStream, movi)) return false;
      }
    } else if (is4(h, "strh")) {
      uint8_t s[56] = {0};
      if (size < 36 || !pread_all(v->fd, s, std::min<uint64_t>(size, sizeof s), body)) { pgb::fail(PGB_ERR_INVALID, "%s: bad strh", v->path.c_str()); return false; }
      const int idx = (*nextStream)++;
      if (v->stream < 0 && is4(s, "vids")) {
        v->stream = idx;
        memcpy(v->fourcc, s + 4, 4);
        v->scale = rd32(s + 20);
        v->rate = rd32(s + 24);
      }
    } else if (is4(h, "strf")) {
      // the strf that follows the selected stream's strh: BITMAPINFOHEADER
      if (v->stream >= 0 && v->stream == *nextStream - 1 && v->width == 0) {
        uint8_t s[40] = {0};
        if (size < 20 || !pread_all(v->fd, s, std::min<uint64_t>(size, sizeof s), body)) { pgb::fail(PGB_ERR_INVALID, "%s: bad strf", v->path.c_str()); return false; }
        v->width = (int)rd32(s + 4);
        const int32_t hgt = (int32_t)rd32(s + 8);
        v->height = hgt < 0 ? -hgt : hgt;
        if (rd32(s + 16) != 0) memcpy(v->fourcc, s + 16, 4);  // biCompression names the codec (strh.fccHandler may be 0)
      }
    } else if (inMovi && v->stream >= 0 && h[0] == (uint8_t)('0' + v->stream / 10) && h[1] == (uint8_t)('0' + v->stream % 10) &&
               ((h[2] == 'd' && (h[3] == 'c' || h[3] == 'b')))) {
      if (size > 0) v->frames.push_back({body, (uint32_t)size});  // (an empty chunk is a dropped frame: the decoder outputs nothing for it)
    }
    off = body + size + (size & 1);
  }
  return true;
}

}  // namespace

using namespace pgb;

extern "C" pgb_video* pgb_video_open(int device, const char* path) {
  if (!path) { fail(PGB_ERR_INVALID, "pgb_video_open: null path"); return nullptr; }
  pgb_video* v = new pgb_video();
  v->device = device;
  v->path = path;
  auto bail = [&](void) -> pgb_video* { if (v->fd >= 0) close(v->fd); delete v; return nullptr; };
  v->fd = open(path, O_RDONLY);
  if (v->fd < 0) { fail(PGB_ERR_INVALID, "cannot open %s", path); return bail(); }  // CHECK_EQ(avformat_open_input(...), 0) (:79-80)
  struct stat sb;
  if (fstat(v->fd, &sb) != 0) { fail(PGB_ERR_INVALID, "cannot stat %s", path); return bail(); }
  v->fileSize = (uint64_t)sb.st_size;
  uint8_t h[12];
  if (v->fileSize < 12 || !pread_all(v->fd, h, 12, 0) || !is4(h, "RIFF") || !is4(h + 8, "AVI ")) {
    fail(PGB_ERR_INVALID, "%s is not a RIFF AVI file (the only container this build demuxes; raw frames go through raw: / raw24:)", path);
    return bail();
  }
  int nextStream = 0;
  if (!walk(v, 0, v->fileSize, 0, &nextStream, false)) return bail();
  if (v->stream < 0) { fail(PGB_ERR_INVALID, "%s: inspected all the streams, but no video stream found", path); return bail(); }  // :69
  bool mjpeg = false;
  for (const char* cc : {"MJPG", "JPEG", "AVRN", "AVDJ", "dmb1"})  // the FOURCCs libavformat's riff.c maps to AV_CODEC_ID_MJPEG (baseline ones)
    mjpeg = mjpeg || !strncasecmp(v->fourcc, cc, 4);
  if (!mjpeg) {
    fail(PGB_ERR_INVALID, "%s: video codec '%.4s' is not decoded by this build (Motion-JPEG only: the image has nvJPEG but no NVDEC binding)", path,
         v->fourcc);
    return bail();
  }
  if (v->width <= 0 || v->height <= 0 || v->rate == 0 || v->scale == 0) {
    fail(PGB_ERR_INVALID, "%s: incomplete stream header (%dx%d, scale %u, rate %u)", path, v->width, v->height, v->scale, v->rate);
    return bail();
  }
  return v;
}

extern "C" int pgb_video_info(pgb_video* v, int* width, int* height, int64_t* n_frames, double* fps, int* rotate_degrees) {
  if (!v) return fail(PGB_ERR_INVALID, "pgb_video_info: null handle");
  if (width) *width = v->width;
  if (height) *height = v->height;
  if (n_frames) *n_frames = (int64_t)v->frames.size();
  if (fps) *fps = (double)v->rate / (double)v->scale;
  if (rotate_degrees) *rotate_degrees = 0;  // the `rotate` entry is MOV / MP4 stream metadata (:113-118); AVI has none
  return PGB_OK;
}

extern "C" int pgb_video_frame_span(pgb_video* v, int64_t frame, uint64_t* offset, uint32_t* size) {
  if (!v || frame < 0 || frame >= (int64_t)v->frames.size()) return fail(PGB_ERR_INVALID, "pgb_video_frame_span: frame out of range");
  if (offset) *offset = v->frames[(size_t)frame].off;
  if (size) *size = v->frames[(size_t)frame].size;
  return PGB_OK;
}

extern "C" int pgb_video_read_rgb(pgb_video* v, int64_t first_frame, int n_frames, uint8_t* rgb_dev, size_t pitch, size_t frame_stride,
                                  double* timestamps_sec, void* stream) {
  if (!v || n_frames < 0 || first_frame < 0 || first_frame + n_frames > (int64_t)v->frames.size())
    return fail(PGB_ERR_INVALID, "pgb_video_read_rgb: frames [%lld, %lld) out of range (the file has %zu)", (long long)first_frame,
                (long long)(first_frame + n_frames), v ? v->frames.size() : (size_t)0);
  if (n_frames == 0) return PGB_OK;
  if (!rgb_dev || pitch < (size_t)v->width * 3 || frame_stride < pitch * v->height)
    return fail(PGB_ERR_INVALID, "pgb_video_read_rgb: null buffer or pitch/stride smaller than the frame");
  if (use_device(v->device)) return PGB_ERR_CUDA;
  NvjpegApi* nj = nvjpeg_api();
  if (!nj->error.empty()) return fail(PGB_ERR_CUDA, "%s", nj->error.c_str());
  cudaStream_t s = (cudaStream_t)stream;
  if (!v->nj) {
    if (nj->CreateSimple(&v->nj) != NVJPEG_STATUS_SUCCESS) return fail(PGB_ERR_CUDA, "nvjpegCreateSimple failed");
    if (nj->JpegStateCreate(v->nj, &v->st) != NVJPEG_STATUS_SUCCESS) return fail(PGB_ERR_CUDA, "nvjpegJpegStateCreate failed");
  }
  // the previous call's bitstreams may still be in use by its decode: drain the stream before they are replaced
  PGB_CUDA(cudaStreamSynchronize(s));
  size_t total = 0;
  for (int i = 0; i < n_frames; i++) total += v->frames[(size_t)(first_frame + i)].size + sizeof kStdDht;  // room for inserted tables
  v->bits.resize(total);
  size_t at = 0;
  NvtxRange range("pgb:video:mjpeg_decode");
  for (int i = 0; i < n_frames; i++) {
    const pgb_video::Frame& f = v->frames[(size_t)(first_frame + i)];
    uint8_t* b = v->bits.data() + at;
    if (!pread_all(v->fd, b, f.size, f.off)) return fail(PGB_ERR_INVALID, "%s: short read of frame %lld", v->path.c_str(), (long long)(first_frame + i));
    size_t fsize = f.size, sosAt = 0;
    bool hasDht = false;
    if (jpeg_header_scan(b, fsize, &sosAt, &hasDht) && !hasDht) {  // a frame without Huffman tables: the standard ones go in front of its scan
      memmove(b + sosAt + sizeof kStdDht, b + sosAt, fsize - sosAt);
      memcpy(b + sosAt, kStdDht, sizeof kStdDht);
      fsize += sizeof kStdDht;
    }
    at += f.size + sizeof kStdDht;
    int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    nvjpegStatus_t r = nj->GetImageInfo(v->nj, b, fsize, &comps, &sub, ws, hs);
    if (r != NVJPEG_STATUS_SUCCESS)
      return fail(PGB_ERR_INVALID, "%s: frame %lld is not a JPEG image nvJPEG can parse (status %d)",
                  v->path.c_str(), (long long)(first_frame + i), (int)r);
    if (ws[0] != v->width || hs[0] != v->height)
      return fail(PGB_ERR_INVALID, "%s: frame %lld is %dx%d, the stream header says %dx%d", v->path.c_str(), (long long)(first_frame + i), ws[0], hs[0],
                  v->width, v->height);
    nvjpegImage_t out;
    memset(&out, 0, sizeof out);
    out.channel[0] = rgb_dev + (size_t)i * frame_stride;
    out.pitch[0] = pitch;
    r = nj->Decode(v->nj, v->st, b, fsize, NVJPEG_OUTPUT_RGBI, &out, s);
    if (r != NVJPEG_STATUS_SUCCESS) return fail(PGB_ERR_CUDA, "%s: nvjpegDecode failed on frame %lld (status %d)", v->path.c_str(), (long long)(first_frame + i), (int)r);
    // av_frame_get_best_effort_timestamp * av_q2d(time_base) (:153-155): an AVI video stream's time base is dwScale / dwRate
    // and the pts of frame k is k
    if (timestamps_sec) timestamps_sec[i] = (double)(first_frame + i) * (double)v->scale / (double)v->rate;
  }
  return PGB_OK;
}

extern "C" void pgb_video_close(pgb_video* v) {
  if (!v) return;
  if (v->nj) {  // (the library was loaded by the decode call that created these)
    NvjpegApi* nj = nvjpeg_api();
    if (v->st && nj->JpegStateDestroy) nj->JpegStateDestroy(v->st);
    if (nj->Destroy) nj->Destroy(v->nj);
  }
  if (v->fd >= 0) close(v->fd);
  delete v;
}
