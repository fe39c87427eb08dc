// Frame feed: what happens to a decoded frame between the video reader and ORBextractor --
// cv::flip of the RGB24 frame (src/io/image_sequence_reader.cc:163-175, flags --vertical_flip / --horizontal_flip) and
// cvtColor to gray in Tracking::GrabImageMonocular (thirdparty/orb-slam2/src/Tracking.cc:243-258: RGB/BGR/RGBA/BGRA by
// Camera.RGB) -- as one pass over the pixels, so frames can go to the extractor without leaving the device.
// cvtColor(8U) is fixed point; the reference pins OpenCV 2.4.x, whose RGB2Gray is (R*4899 + G*9617 + B*1868 + 8192) >> 14
// (formula 0, un-vendored, restated from the published source); OpenCV >= 3 uses (R*9798 + G*19235 + B*3735 + 16384) >> 15
// (formula 1, pinned bit-exact against cv2 4.13 golden vectors).  They differ on ~0.3 % of random pixels by one level.
// Video decoding itself (libav in the reference) stays outside the library.
#include <cuda_runtime.h>

#include "common.cuh"

namespace pgb {

// The same with the container's `rotate` metadata applied first (VideoImageSequenceSource::fetchNext,
// src/io/image_sequence_reader.cc:186-207: 90 -> cv::flip(raw.t(), out, 0), 180 -> cv::flip(raw, out, -1),
// 270 -> cv::flip(raw.t(), out, 1)), then the CLI's flips, then cvtColor.  (w, h) are the OUTPUT dimensions; thread =
// four output pixels of a row.  For 90 / 270 an output row walks down a source column: a 32 x 32 tile goes through shared
// memory so that both the source reads and the destination writes are row-contiguous.
template <int kCh>
__global__ void __launch_bounds__(256) k_to_gray_rot(const uint8_t* __restrict__ src, size_t srcPitch, size_t srcStride,
                                                     uint8_t* __restrict__ dst, size_t dstPitch, size_t dstStride, int w, int h,
                                                     int rgbOrder, int vflip, int hflip, int formula, int rot) {
  __shared__ uint8_t tile[32][33];
  const int f = blockIdx.z;
  const int cr = formula ? 9798 : 4899, cg = formula ? 19235 : 9617, cb = formula ? 3735 : 1868;
  const int rnd = formula ? 16384 : 8192, sh = formula ? 15 : 14;
  const int sw = (rot == 90 || rot == 270) ? h : w, shh = (rot == 90 || rot == 270) ? w : h;  // source dimensions
  // output (y, x) -> rotated image (yr, xr) through the flips -> source (sy, sx)
  auto src_of = [&](int y, int x, int* sy, int* sx) {
    const int yr = vflip ? h - 1 - y : y, xr = hflip ? w - 1 - x : x;
    if (rot == 0) { *sy = yr; *sx = xr; }
    else if (rot == 90) { *sy = xr; *sx = sw - 1 - yr; }       // out(y, x) = raw(x, w_src - 1 - y)
    else if (rot == 180) { *sy = shh - 1 - yr; *sx = sw - 1 - xr; }
    else { *sy = shh - 1 - xr; *sx = yr; }                      // 270: out(y, x) = raw(h_src - 1 - x, y)
  };
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const bool transposed = rot == 90 || rot == 270;
  // load phase: thread (tx, ty + 8k) reads the source pixel that is contiguous along tx
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int a = ty + 8 * k;
    // not transposed: (y, x) = (y0 + a, x0 + tx); transposed: the source row index follows the output x, so swap roles
    const int y = transposed ? y0 + tx : y0 + a, x = transposed ? x0 + a : x0 + tx;
    if (y < h && x < w) {
      int sy, sx;
      src_of(y, x, &sy, &sx);
      const uint8_t* p = src + (size_t)f * srcStride + (size_t)sy * srcPitch + (size_t)sx * kCh;
      int v;
      if (kCh == 1) v = p[0];
      else {
        const int c0 = p[0], c1 = p[1], c2 = p[2];
        const int r = rgbOrder ? c0 : c2, b = rgbOrder ? c2 : c0;
        v = (r * cr + c1 * cg + b * cb + rnd) >> sh;
      }
      if (transposed) tile[tx][a] = (uint8_t)v; else tile[a][tx] = (uint8_t)v;   // tile[y - y0][x - x0]
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int y = y0 + ty + 8 * k, x = x0 + tx;
    if (y < h && x < w) dst[(size_t)f * dstStride + (size_t)y * dstPitch + x] = tile[ty + 8 * k][tx];
  }
}

template <int kCh>
__global__ void __launch_bounds__(256) k_to_gray(const uint8_t* __restrict__ src, size_t srcPitch, size_t srcStride,
                                                 uint8_t* __restrict__ dst, size_t dstPitch, size_t dstStride, int w, int h,
                                                 int rgbOrder, int vflip, int hflip, int formula) {
  const int xq = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (xq * 4 >= w) return;
  const uint8_t* srow = src + (size_t)f * srcStride + (size_t)(vflip ? h - 1 - y : y) * srcPitch;
  const int cr = formula ? 9798 : 4899, cg = formula ? 19235 : 9617, cb = formula ? 3735 : 1868;
  const int rnd = formula ? 16384 : 8192, sh = formula ? 15 : 14;
  uint32_t packed = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = xq * 4 + i;
    if (x < w) {
      const uint8_t* p = srow + (size_t)(hflip ? w - 1 - x : x) * kCh;
      int v;
      if (kCh == 1) {
        v = p[0];
      } else {
        const int c0 = p[0], c1 = p[1], c2 = p[2];
        const int r = rgbOrder ? c0 : c2, b = rgbOrder ? c2 : c0;
        v = (r * cr + c1 * cg + b * cb + rnd) >> sh;
      }
      packed |= (uint32_t)v << (8 * i);
    }
  }
  uint8_t* drow = dst + (size_t)f * dstStride + (size_t)y * dstPitch;
  if (xq * 4 + 3 < w && (dstPitch & 3) == 0 && ((size_t)dst & 3) == 0 && (dstStride & 3) == 0) {
    *reinterpret_cast<uint32_t*>(drow + xq * 4) = packed;
  } else {
    for (int i = 0; i < 4 && xq * 4 + i < w; i++) drow[xq * 4 + i] = (uint8_t)(packed >> (8 * i));
  }
}

}  // namespace pgb

using namespace pgb;

extern "C" int pgb_frames_to_gray(int device, const uint8_t* src, int src_is_device, int n_frames, int width, int height,
                                  int channels, int rgb_order, size_t src_pitch, size_t src_frame_stride, int vertical_flip,
                                  int horizontal_flip, int formula, uint8_t* dst_gray, int dst_is_device, size_t dst_pitch,
                                  size_t dst_frame_stride, void* stream) {
  return pgb_frames_to_gray_rotated(device, src, src_is_device, n_frames, width, height, channels, rgb_order, src_pitch,
                                    src_frame_stride, 0, vertical_flip, horizontal_flip, formula, dst_gray, dst_is_device,
                                    dst_pitch, dst_frame_stride, stream);
}

extern "C" int pgb_frames_to_gray_rotated(int device, const uint8_t* src, int src_is_device, int n_frames, int src_width,
                                          int src_height, int channels, int rgb_order, size_t src_pitch, size_t src_frame_stride,
                                          int rotate_degrees, int vertical_flip, int horizontal_flip, int formula,
                                          uint8_t* dst_gray, int dst_is_device, size_t dst_pitch, size_t dst_frame_stride,
                                          void* stream) {
  if (n_frames < 0 || src_width < 0 || src_height < 0 || (channels != 1 && channels != 3 && channels != 4) || (formula != 0 && formula != 1))
    return fail(PGB_ERR_INVALID, "pgb_frames_to_gray: invalid argument");
  rotate_degrees %= 360;  // image_sequence_reader.cc:118
  if (rotate_degrees != 0 && rotate_degrees != 90 && rotate_degrees != 180 && rotate_degrees != 270)
    return fail(PGB_ERR_INVALID, "Unsupported rotation angle in video metadata: %d. Only multiples of 90 degrees rotations are supported.",
                rotate_degrees);  // the reference's LOG(FATAL) (:203-206)
  const bool transposed = rotate_degrees == 90 || rotate_degrees == 270;
  const int width = transposed ? src_height : src_width, height = transposed ? src_width : src_height;  // output dimensions
  if (n_frames == 0 || width == 0 || height == 0) return PGB_OK;
  if (!src || !dst_gray || src_pitch < (size_t)src_width * channels || dst_pitch < (size_t)width ||
      src_frame_stride < src_pitch * src_height || dst_frame_stride < dst_pitch * height)
    return fail(PGB_ERR_INVALID, "pgb_frames_to_gray: null buffer or pitch/stride smaller than the image");
  if (use_device(device)) return PGB_ERR_CUDA;
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf<uint8_t> dIn, dOut;
  const uint8_t* in = src;
  uint8_t* out = dst_gray;
  if (!src_is_device) {
    if (dIn.alloc(src_frame_stride * n_frames)) return PGB_ERR_CUDA;
    PGB_CUDA(cudaMemcpyAsync(dIn.p, src, src_frame_stride * (n_frames - 1) + src_pitch * src_height, cudaMemcpyHostToDevice, s));
    in = dIn.p;
  }
  if (!dst_is_device) {
    if (dOut.alloc(dst_frame_stride * n_frames)) return PGB_ERR_CUDA;
    out = dOut.p;
  }
  if (rotate_degrees != 0) {
    dim3 gridR((width + 31) / 32, (height + 31) / 32, n_frames);
    if (channels == 1) k_to_gray_rot<1><<<gridR, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula, rotate_degrees);
    else if (channels == 3) k_to_gray_rot<3><<<gridR, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula, rotate_degrees);
    else k_to_gray_rot<4><<<gridR, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula, rotate_degrees);
  } else {
  dim3 grid(((width + 3) / 4 + 255) / 256, height, n_frames);
  if (channels == 1) k_to_gray<1><<<grid, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula);
  else if (channels == 3) k_to_gray<3><<<grid, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula);
  else k_to_gray<4><<<grid, 256, 0, s>>>(in, src_pitch, src_frame_stride, out, dst_pitch, dst_frame_stride, width, height, rgb_order, vertical_flip, horizontal_flip, formula);
  }
  PGB_CHECK_LAUNCH();
  if (!dst_is_device) {
    PGB_CUDA(cudaMemcpy2DAsync(dst_gray, dst_pitch, dOut.p, dst_pitch, width, (size_t)height, cudaMemcpyDeviceToHost, s));
    for (int f = 1; f < n_frames; f++)
      PGB_CUDA(cudaMemcpy2DAsync(dst_gray + f * dst_frame_stride, dst_pitch, dOut.p + f * dst_frame_stride, dst_pitch, width,
                                 (size_t)height, cudaMemcpyDeviceToHost, s));
  }
  if (!src_is_device || !dst_is_device) PGB_CUDA(cudaStreamSynchronize(s));
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------ synthetic frame source
// The SURVEY.md 8(d) frame generator rendered on the device, for the long BASELINE configs (10 000 and 54 000 frames: 21 and
// 112 GB as raw files): frame t = w x h crop of a canvas at the ping-pong origin of pilotguru_b200/synth.py, plus per-frame
// noise.  It stands where a hardware video decoder would stand: frames appear in device memory.  The noise is a counter
// hash (sum of four uniform bytes, sigma 1 like the numpy generator's N(0, 1)), so the frames are deterministic functions
// of (t, x, y) but NOT bit-identical to synth.frame(t): parity tests use the numpy frames, timing runs use these.
namespace pgb {
__device__ __forceinline__ uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__global__ void __launch_bounds__(256) k_synth_frames(const uint8_t* __restrict__ canvas, int cw, int ch, int firstT, int w, int h,
                                                      uint8_t* __restrict__ out) {
  const int xq = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (xq * 4 >= w) return;
  const int t = firstT + f;
  const int px = cw - w - 32, py = ch - h - 32;  // synth.frame_origin: 16 + pp(2t, px), 16 + pp(t, py)
  auto pp = [](int a, int p) { const int m = a % (2 * p); return p - abs(m - p); };
  const int x0 = 16 + pp(2 * t, px), y0 = 16 + pp(t, py);
  uint32_t packed = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = xq * 4 + i;
    if (x < w) {
      const uint32_t r = hash3((uint32_t)t, (uint32_t)y, (uint32_t)x);
      const int s4 = (int)(r & 0xff) + (int)((r >> 8) & 0xff) + (int)((r >> 16) & 0xff) + (int)(r >> 24);
      const float v = (float)canvas[(size_t)(y0 + y) * cw + x0 + x] + (float)(s4 - 510) * (1.0f / 147.8f);
      packed |= (uint32_t)min(max(__float2int_rn(v), 0), 255) << (8 * i);
    }
  }
  uint8_t* drow = out + (size_t)f * w * h + (size_t)y * w;
  if ((w & 3) == 0) *reinterpret_cast<uint32_t*>(drow + xq * 4) = packed;
  else for (int i = 0; i < 4 && xq * 4 + i < w; i++) drow[xq * 4 + i] = (uint8_t)(packed >> (8 * i));
}
}  // namespace pgb

extern "C" int pgb_synth_frames(int device, const uint8_t* canvas_dev, int canvas_w, int canvas_h, int first_t, int n_frames,
                                int width, int height, uint8_t* out_dev, void* stream) {
  if (!canvas_dev || !out_dev || n_frames < 0 || width <= 0 || height <= 0 || canvas_w < width + 33 || canvas_h < height + 33 || first_t < 0)
    return fail(PGB_ERR_INVALID, "pgb_synth_frames: invalid argument (the canvas must exceed the frame by at least 33 px)");
  if (n_frames == 0) return PGB_OK;
  if (use_device(device)) return PGB_ERR_CUDA;
  dim3 grid(((width + 3) / 4 + 255) / 256, height, n_frames);
  k_synth_frames<<<grid, 256, 0, (cudaStream_t)stream>>>(canvas_dev, canvas_w, canvas_h, first_t, width, height, out_dev);
  PGB_CHECK_LAUNCH();
  return PGB_OK;
}

// ------------------------------------------------------------------------------------------------ device memory for hosts
// Host programs that keep frames and features on the device between calls (optical_trajectories) link only libpgb200.so; the
// CUDA runtime inside it is private, so the few memory operations such a host needs are part of the C-ABI.
extern "C" void* pgb_device_malloc(int device, size_t bytes) {
  if (use_device(device)) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { fail(PGB_ERR_CUDA, "cudaMalloc(%zu) failed", bytes); return nullptr; }
  cudaMemset(p, 0, bytes);
  return p;
}
extern "C" void pgb_device_free(int device, void* p) {
  if (!p) return;
  cudaSetDevice(device);
  cudaFree(p);
}
extern "C" void* pgb_host_malloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { fail(PGB_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes); return nullptr; }
  return p;
}
extern "C" void pgb_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }
extern "C" int pgb_memcpy_async(int device, void* dst, const void* src, size_t bytes, int kind, void* stream) {
  if (bytes == 0) return PGB_OK;
  if (!dst || !src || kind < 0 || kind > 2) return fail(PGB_ERR_INVALID, "pgb_memcpy_async: invalid argument");
  PGB_CUDA(cudaSetDevice(device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  PGB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream));
  return PGB_OK;
}
extern "C" int pgb_stream_synchronize(int device, void* stream) {
  PGB_CUDA(cudaSetDevice(device));
  PGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return PGB_OK;
}
// Completion markers for callers that keep several batches in flight on one stream (optical_trajectories): an event is
// recorded behind a batch's device-to-host copies and waited for when the host needs that batch's results.
extern "C" void* pgb_event_create(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { fail(PGB_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { fail(PGB_ERR_CUDA, "cudaEventCreate failed"); return nullptr; }
  return e;
}
extern "C" int pgb_event_record(int device, void* event, void* stream) {
  if (!event) return fail(PGB_ERR_INVALID, "pgb_event_record: null event");
  PGB_CUDA(cudaSetDevice(device));
  PGB_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
  return PGB_OK;
}
extern "C" int pgb_event_synchronize(int device, void* event) {
  if (!event) return fail(PGB_ERR_INVALID, "pgb_event_synchronize: null event");
  PGB_CUDA(cudaSetDevice(device));
  PGB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return PGB_OK;
}
extern "C" void pgb_event_destroy(int device, void* event) {
  if (!event) return;
  cudaSetDevice(device);
  cudaEventDestroy((cudaEvent_t)event);
}
extern "C" int pgb_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
