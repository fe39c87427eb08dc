// Device-side geometry and kernel declarations for the ORB extraction path (K1..K7 of SURVEY.md section 2.1).
// Data layout in HBM (per extractor handle, B = max_batch):
//   pyr    [B][frameStride]  u8   level l of frame f at f*frameStride + lv[l].off, rows lv[l].pitch apart
//   score  same layout as pyr (FAST score map, 0 = not a corner at minTh / outside the tested region)
//   slots  [B][slotsPerFrame] u32  per FAST cell: up to slotCap packed candidates x:12|y:12|score:8
//   cellCnt[B][totalCells]   i32
//   cand   [B][candPerFrame] u64  per level: candidates in reference order; hi32 = octree node id | quadrant<<30
//   staged [B][kpCapInternal]     per level: octree survivors in list order; lvlCnt[B][nlevels]
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)

#include <cstdint>

#include "../../include/pgb200.h"

namespace pgb {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;        // EDGE_THRESHOLD (ORBextractor.cc:74)
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD - 3 (ORBextractor.cc:773)
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE

// pyramid kernel: destination tile of k_pyramid_tiled and the capacity of its source staging buffers
// (the PGB_* macros exist for A/B builds: tools/build_variants.sh)
#ifndef PGB_PY_H
#define PGB_PY_H 32
#endif
constexpr int kPyW = 256, kPyH = PGB_PY_H;
constexpr int kPySrcPitch = 416, kPySrcRows = kPyH * 5 / 4 + 4;  // staged source footprint of a destination tile
inline bool pyramid_tile_fits(int sw, int sh, int dw, int dh) {
  const double rx = (double)sw / dw, ry = (double)sh / dh;
  return rx * kPyW + 34 <= kPySrcPitch && ry * kPyH + 3 <= kPySrcRows;
}

// FAST score kernel: one CTA per 256x32 tile (4 warps x 8-row bands), TMA-staged with a halo
constexpr int kF2W = 256, kF2H = 32, kF2Threads = 128;
constexpr int kF2InWords = kF2W / 4 + 8;  // 72 words per row: 16-byte halo left and right (a TMA box must start
                                          // on a 16-byte boundary in the innermost dimension; measured: tools/probe)
constexpr int kF2InRows = kF2H + 6;       // 70
constexpr int kF2InBytes = kF2InWords * 4 * kF2InRows;  // 20160 = TMA transaction size

// fused FAST + cell NMS kernel (fast_cells.cu): one CTA per row of up to fcKc FAST cells; the TMA box of a tile is
// 72 words x (8 * bands + 6) rows, the score tile (8 * bands + 2) rows of 272 bytes, one candidate queue per band
constexpr int kFcThreads = 128, kFcWarps = 4;
#ifndef PGB_FC_INW
#define PGB_FC_INW 72
#endif
#ifndef PGB_FC_QCAP
#define PGB_FC_QCAP 384
#endif
constexpr int kFcInWords = PGB_FC_INW;  // staged input rows: >= 72 words (<= 18 bytes of alignment slack + 249 px + 3-px halos)
constexpr int kFcTilePitch = 272;    // score tile row: 16 pad bytes + 256 px
constexpr int kFcQueueCap = PGB_FC_QCAP;  // candidates per band queue (u32 flag + u8 code each; >= 256: a refilled queue holds one row)
constexpr int kFcQueueBytes = kFcQueueCap * 5;
constexpr int kFcListCap = 128;     // NMS survivors per tile kept in the list (natural images: ~20); more -> bitmap emission
constexpr int kFcMaxFrame = 249;     // widest tested tile inside the 256-px lane frame (first pixel at offset 0..7)
#ifndef PGB_FC_OCC
#define PGB_FC_OCC 9
#endif
constexpr int kFcOccA = PGB_FC_OCC;  // resident CTAs per SM the class-A instantiation is compiled for
struct FcSmem {
  int tile, queue, misc, total;  // byte offsets inside the dynamic shared memory (input stage at 0)
};
__host__ __device__ inline FcSmem fc_smem_layout(int nb) {
  FcSmem s;
  s.tile = (kFcInWords * 4 * (8 * nb + 6) + 127) & ~127;
  s.queue = s.tile + (8 * nb + 2) * kFcTilePitch;
  s.misc = s.queue + nb * kFcQueueBytes;
  s.total = s.misc + 64;  // mbarrier (16 B) + per-band corner counts (8 ints) + the tile's survivor count
  return s;
}
// two-tier kernel (k_fast_cells2<nb>, one warp per band): smaller queues (the iniThFAST pass has about half the
// candidates), the survivor list in its own region (the input stage stays live for the minThFAST pass)
#ifndef PGB_FC2_QCAP
#define PGB_FC2_QCAP 256
#endif
constexpr int kFc2QueueCap = PGB_FC2_QCAP;  // per band, general path (>= 256: a refilled queue holds one row); the pooled queue of the
                                   // fast path holds kFc2QueueCap * nb * 3 / 2 16-bit entries in the same bytes
constexpr int kFc2TilePitch = 264;  // score tile row: 8 pad bytes + 256 px
__host__ __device__ constexpr FcSmem fc2_smem_layout(int nb) {
  FcSmem s{};
  s.tile = kFcInWords * 4 * (8 * nb + 6);  // (a multiple of 16)
  s.queue = s.tile + (8 * nb + 2) * kFc2TilePitch;
  s.misc = s.queue + nb * kFc2QueueCap * 3;  // general path, per band: 16-bit queue entries + the 256-byte survivor bitmap
  s.total = s.misc + kFcListCap * 4 + 128;   // survivor list, mbarrier (16 B), corner counts (8 ints), survivor count, cell mask,
                                             // pooled queue counts (2 ints), bit -> (x, row) table (64 B)
  return s;
}

struct LevelGeo {
  int w, h, pitch;
  unsigned long long off;
  int maxBX, maxBY;             // w-16, h-16
  int nCols, nRows, wCell, hCell;
  int cellBase, slotCap;
  unsigned long long slotBase;  // in u32 units inside a frame's slot block
  int quota, nIni;
  float hX;
  int candCap;
  unsigned long long candBase;  // in u64 units inside a frame's cand block
  int nodeCap, kpBase;
  int tile2Base, tiles2X, tiles2Y;
  int fcKc, fcNb, fcClassB;     // fused kernel: cells per tile, bands per tile; class 0 = two-tier kernel with 4 bands (cells <= 32 x 32),
                                // 2 = two-tier with 5 bands (<= 32 x 40), 1 = generic single-pass instantiation (anything larger)
  unsigned fcRecip;             // ceil(65536 / wCell): (x * fcRecip) >> 16 == x / wCell for x < 1024
  float scale;
  int patchSize;
};

struct OrbGeo {
  int nlevels, iniTh, minTh;
  unsigned one;      // = 1, opaque to the compiler (FAST v3 issues its additions as IMAD on the idle FMA pipe)
  unsigned absMask;  // FAST prefilter: 0x80 - (minTh + 1) in every byte
  unsigned absMaskIni;  // the same for iniTh (two-tier kernel)
  int totalCells, totalTiles2, kpCapInternal, maxNodeCap;
  int fcTilesA, fcTilesB, fcNbB;  // fused kernel: tiles of class 0 / class 1 levels, bands per class-1 tile
  int fcTilesA5;                  // tiles of class 2 levels
  unsigned long long frameStride, slotsPerFrame, candPerFrame;
  // Level 0 in place: when the caller's frames are device-resident and 16-byte aligned, level 0 is read where it lies
  // (ext0 + f * ext0Stride, rows ext0Pitch apart) instead of being copied into the pyramid buffer.
  const uint8_t* ext0;
  unsigned long long ext0Stride;
  int ext0Pitch;
  LevelGeo lv[kMaxLevels];
};

#ifdef __CUDACC__
// Base address and row pitch of level `level` of frame f (f relative to the pointers the kernel was given).
__device__ __forceinline__ const uint8_t* level_base(const OrbGeo& g, const uint8_t* pyr, int level, int f, int* pitch) {
  if (level == 0 && g.ext0) {
    *pitch = g.ext0Pitch;
    return g.ext0 + (size_t)f * g.ext0Stride;
  }
  *pitch = g.lv[level].pitch;
  return pyr + (size_t)f * g.frameStride + g.lv[level].off;
}
#endif

struct alignas(64) TmapPack {  // per level: u32 views of the pyramid (load) and of the score map (store)
  CUtensorMap in[kMaxLevels];
  CUtensorMap out[kMaxLevels];
};

struct alignas(64) TmapIn {  // fused kernel: per level the u32 view of the pyramid with that level's box height
  CUtensorMap in[kMaxLevels];
};

struct ResizeTab {  // one entry per destination column / row
  short s, a0, a1, pad;
};

struct StagedKp {
  int x, y, score, level;
};

enum OrbErr { kErrCandOverflow = 1, kErrNodeOverflow = 2, kErrOutCap = 4, kErrCellChunks = 8 };

void launch_pyramid_level(const OrbGeo& g, const TmapIn& tm, int level, int frame0, int nFrames, uint8_t* pyr,
                          const ResizeTab* xtab, const ResizeTab* ytab, const int2* tileX, const int2* tileY,
                          cudaStream_t st);
int configure_fast_score();
int launch_fast_score(const OrbGeo& g, const TmapPack& tm, const int4* tileTab, int frame0, int nFrames,
                         cudaStream_t st);
int launch_fast_cells(const OrbGeo& g, const TmapIn& tm, const int4* tileTabA, const int4* tileTabB, const int4* tileTabA5,
                      int nFrames, uint32_t* slots, int* cellCnt, int* err, cudaStream_t st, int frame0,
                      cudaStream_t side = nullptr, cudaEvent_t evFork = nullptr, cudaEvent_t evJoin = nullptr);
void launch_cells(const OrbGeo& g, int nFrames, const int* cellTab, const uint8_t* score, uint32_t* slots, int* cellCnt,
                  int* err, cudaStream_t st);
void launch_octree(const OrbGeo& g, int nFrames, const uint32_t* slots, const int* cellCnt, unsigned long long* cand,
                   StagedKp* staged, int* lvlCnt, int* err, cudaStream_t st);
void launch_orient_desc(const OrbGeo& g, int nFrames, const uint8_t* pyr, const StagedKp* staged, const int* lvlCnt,
                        pgb_keypoint* kps, uint8_t* desc, int* counts, int cap, int* err, cudaStream_t st);
void launch_blur_level(const OrbGeo& g, int level, int frame, const uint8_t* pyr, uint8_t* out, cudaStream_t st);

}  // namespace pgb
