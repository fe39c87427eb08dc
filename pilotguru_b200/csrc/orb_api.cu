// C-ABI of the ORB extractor (include/pgb200.h, "ORB extractor" section): handle, geometry, orchestration.
// Host-side restatement of the reference's constructor tables (ORBextractor.cc:410-470) and of the per-level
// grid arithmetic (ORBextractor.cc:769-787, :542-545), which must use the same float expressions.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "orb_kernels.cuh"

namespace pgb {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int use_device(int device) {
  // The capability check runs once per device: cudaGetDeviceProperties costs milliseconds (it queries the driver for
  // every field), and entry points such as pgb_synth_frames / pgb_frames_to_gray come through here once per batch.
  static std::atomic<unsigned long long> verified{0};
  if (device >= 0 && device < 64 && (verified.load(std::memory_order_relaxed) >> device & 1ull)) {
    PGB_CUDA(cudaSetDevice(device));
    return PGB_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(PGB_ERR_CUDA, "no CUDA device available (%s); libpgb200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(PGB_ERR_INVALID, "device %d out of range (have %d)", device, n);
  PGB_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  PGB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  PGB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return fail(PGB_ERR_CUDA, "device %d is sm_%d%d; libpgb200 is built for sm_100a only", device, major, minor);
  if (device < 64) verified.fetch_or(1ull << device, std::memory_order_relaxed);
  return PGB_OK;
}

int keep_pool_memory(int device) {
  static std::atomic<unsigned> done{0};
  if (device >= 0 && device < 32 && (done.load() >> device & 1u)) return PGB_OK;
  cudaMemPool_t pool;
  PGB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;
  PGB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (device >= 0 && device < 32) done.fetch_or(1u << device);
  return PGB_OK;
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }

}  // namespace pgb

using namespace pgb;

struct pgb_orb {
  int device = 0;
  int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
  float scaleFactor = 0;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nPerLevel;
  int maxW = 0, maxH = 0, maxBatch = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;

  OrbGeo capGeo{};  // geometry at (maxW, maxH): sizes the buffers
  OrbGeo geo{};     // geometry of the last call
  int curW = 0, curH = 0, curFrames = 0;

  DevBuf<uint8_t> pyr, score;
  DevBuf<uint32_t> slots;
  DevBuf<int> cellCnt, lvlCnt, err, counts;
  DevBuf<unsigned long long> cand;
  DevBuf<StagedKp> staged;
  DevBuf<ResizeTab> xtab, ytab;
  std::vector<int> xtabOff, ytabOff, tileXOff, tileYOff;
  DevBuf<int2> tileX, tileY;
  DevBuf<pgb_keypoint> kps;
  DevBuf<uint8_t> desc;
  int outCap = 0;
  DevBuf<uint8_t> tmpLevel;
  TmapPack tmaps{};
  TmapPack tmapsCur{};  // = tmaps, with in[0] re-encoded on the caller's buffer while level 0 is read in place
  DevBuf<int4> tileTab;
  TmapIn tmapsFc{};     // fused kernel: per level, box height of the level's class
  TmapIn tmapsFcCur{};  // = tmapsFc, with in[0] on the caller's buffer while level 0 is read in place
  TmapIn tmapsPy{}, tmapsPyCur{};  // pyramid kernel: source-footprint boxes (kPySrcPitch x kPySrcRows) of every level
  DevBuf<int4> fcTabA, fcTabB, fcTabA5;  // fused kernel tile tables: level, first tested x, first tested y, cell row | first cell column << 16
  bool unfused = false;  // PGB_UNFUSED=1: the round-1 pair k_fast_score -> k_cells instead of k_fast_cells (A/B, stage debugging)
  bool scoreValid = false;  // the score map of the resident batch has been produced (only the unfused path writes it)
  DevBuf<int> cellTab;  // per FAST cell: level | grid row << 8 | grid column << 20
  cudaStream_t copyStream = nullptr;
  cudaStream_t auxStream[2] = {nullptr, nullptr};
  cudaEvent_t evAux[2] = {nullptr, nullptr};
  cudaEvent_t evDone = nullptr;
  // fused FAST kernel: the small launches (5-band and generic tiles) run on a side stream next to the main one
  cudaStream_t fcSide = nullptr;
  cudaEvent_t evFcFork = nullptr, evFcJoin = nullptr;
  // octree kernel (one long-lived, latency-bound CTA per level and frame): on a HIGH-PRIORITY side stream, so that with
  // several handles in flight its CTAs are placed ahead of another handle's throughput kernels instead of behind them
  cudaStream_t octSide = nullptr;
  cudaEvent_t evOctFork = nullptr, evOctJoin = nullptr;
  cudaEvent_t evChunk[16] = {};
  int h2dChunk = 16;  // largest H2D/compute pipeline chunk in frames (PGB_H2D_CHUNK)
  int h2dMinChunk = 4;  // smallest chunk of the ramp-down at the end of a batch (PGB_H2D_MIN_CHUNK)
  int resChunk = 0;     // chunk of a device-resident batch, alternated over the three compute streams (PGB_RES_CHUNK; 0 = off, the default: see run_stages_resident)
  int numSMs = 148;
};

namespace {

constexpr int kMaxChunkEvents = 16;

int build_geo(const pgb_orb* o, int w, int h, OrbGeo* g) {
  memset(g, 0, sizeof *g);
  g->nlevels = o->nlevels;
  g->iniTh = o->iniTh;
  g->minTh = o->minTh;
  {
    // prefilter constant K = 0x80 - (minTh + 1) per byte: x + K has its msb set iff x > minTh, for x < 0x80 (x >= 0x80
    // is caught by "| x"); thresholds >= 0x7f degrade to the weaker but still necessary test x >= 0x80
    g->one = 1u;
    g->absMask = (uint32_t)std::max(0, 0x80 - (o->minTh + 1)) * 0x01010101u;
    g->absMaskIni = (uint32_t)std::max(0, 0x80 - (o->iniTh + 1)) * 0x01010101u;
  }
  unsigned long long off = 0, slotOff = 0, candOff = 0;
  int cellBase = 0, tile2Base = 0, kpBase = 0, maxNode = 1;
  g->fcTilesA = g->fcTilesB = g->fcNbB = g->fcTilesA5 = 0;
  for (int l = 0; l < o->nlevels; l++) {
    LevelGeo& L = g->lv[l];
    L.w = cv_round_f((float)w * o->invScale[l]);
    L.h = cv_round_f((float)h * o->invScale[l]);
    if (L.w < 1 || L.h < 1)
      return fail(PGB_ERR_INVALID, "level %d of a %dx%d image is empty", l, w, h);
    L.pitch = round_up(L.w, 64);
    L.off = off;
    off += (unsigned long long)L.pitch * L.h;
    off = (off + 255) & ~255ull;
    L.maxBX = L.w - kMinBorder;
    L.maxBY = L.h - kMinBorder;
    const float width = (float)(L.maxBX - kMinBorder), height = (float)(L.maxBY - kMinBorder);
    L.nCols = (int)(width / 30.f);
    L.nRows = (int)(height / 30.f);
    if (L.nCols <= 0 || L.nRows <= 0) {
      // Level too small for a 30-px cell: the reference's cell loops do not execute (nRows or nCols is 0; the
      // inf -> int cell size it computes is never used) and the level contributes no keypoints.
      L.nCols = 0; L.nRows = 0; L.wCell = 1; L.hCell = 1;
    } else {
      L.wCell = (int)std::ceil(width / L.nCols);
      L.hCell = (int)std::ceil(height / L.nRows);
    }
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    L.slotCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
    L.slotBase = slotOff;
    slotOff += (unsigned long long)L.nCols * L.nRows * L.slotCap;
    L.quota = o->nPerLevel[l];
    L.nIni = L.nCols > 0 ? (int)std::round((float)(L.maxBX - kMinBorder) / (L.maxBY - kMinBorder)) : 1;
    if (L.nIni <= 0)
      return fail(PGB_ERR_INVALID, "level %d (%dx%d): width/height ratio rounds to 0 root nodes", l, L.w, L.h);
    L.hX = (float)(L.maxBX - kMinBorder) / L.nIni;
    L.candCap = L.nCols * L.nRows * L.slotCap;
    L.candBase = candOff;
    candOff += (unsigned long long)L.candCap;
    L.nodeCap = std::max(L.quota + 4, 4 * L.nIni + 1);
    maxNode = std::max(maxNode, L.nodeCap);
    L.kpBase = kpBase;
    kpBase += L.nodeCap;
    L.tiles2X = (L.w + kF2W - 1) / kF2W;
    L.tiles2Y = (L.h + kF2H - 1) / kF2H;
    L.tile2Base = tile2Base;
    tile2Base += L.tiles2X * L.tiles2Y;
    // fused FAST + cell NMS kernel: a tile is one row of fcKc whole cells
    L.fcKc = std::max(1, kFcMaxFrame / L.wCell);
    L.fcNb = (L.hCell + 7) / 8;
    L.fcClassB = (L.wCell > 32 || L.hCell > 40) ? 1 : (L.hCell > 32 ? 2 : 0);
    L.fcRecip = (65536u + (unsigned)L.wCell - 1u) / (unsigned)L.wCell;
    if (L.wCell > kFcMaxFrame || L.fcNb > 8)
      return fail(PGB_ERR_INVALID, "level %d: FAST cell %dx%d exceeds the kernel's tile", l, L.wCell, L.hCell);
    {
      const int tiles = L.nRows * ((L.nCols + L.fcKc - 1) / L.fcKc);
      if (L.fcClassB == 1) { g->fcTilesB += tiles; g->fcNbB = std::max(g->fcNbB, L.fcNb); }
      else if (L.fcClassB == 2) g->fcTilesA5 += tiles;
      else g->fcTilesA += tiles;
    }
    L.scale = o->scale[l];
    L.patchSize = (int)(31 * o->scale[l]);
    if (L.maxBX - kMinBorder > 4095 || L.maxBY - kMinBorder > 4095)
      return fail(PGB_ERR_INVALID, "images wider/taller than 4127 px are not supported");
  }
  g->totalCells = cellBase;
  g->totalTiles2 = tile2Base;
  g->kpCapInternal = kpBase;
  g->maxNodeCap = maxNode;
  g->frameStride = off;
  g->slotsPerFrame = slotOff;
  g->candPerFrame = candOff;
  return PGB_OK;
}

// cv::resize(INTER_LINEAR) coefficient tables (OpenCV imgproc resize.cpp)
void make_resize_tab(int src, int dst, bool clampCoef, std::vector<ResizeTab>& out) {
  out.resize(dst);
  const double scale = 1.0 / ((double)dst / src);
  for (int d = 0; d < dst; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (clampCoef) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= src - 1) { f = 0; s = src - 1; }
    }
    out[d].s = (short)s;
    out[d].a0 = (short)cv_round_f((1.f - f) * 2048);
    out[d].a1 = (short)cv_round_f(f * 2048);
    out[d].pad = 0;
  }
}

int upload_tabs(pgb_orb* o) {
  std::vector<ResizeTab> xs, ys, t;
  std::vector<int2> tx, ty;  // per tile column {first staged source byte (16-aligned), 16-byte vectors}, per tile row {first source row, rows}
  o->xtabOff.assign(o->nlevels, 0);
  o->ytabOff.assign(o->nlevels, 0);
  o->tileXOff.assign(o->nlevels, 0);
  o->tileYOff.assign(o->nlevels, 0);
  for (int l = 1; l < o->nlevels; l++) {
    const LevelGeo& S = o->geo.lv[l - 1];
    const LevelGeo& D = o->geo.lv[l];
    o->xtabOff[l] = (int)xs.size();
    make_resize_tab(S.w, D.w, true, t);
    xs.insert(xs.end(), t.begin(), t.end());
    o->tileXOff[l] = (int)tx.size();
    for (int x0 = 0; x0 < D.w; x0 += kPyW) {
      const int xl = std::min(x0 + kPyW - 1, D.w - 1);
      const int sxLo = t[x0].s, sxHi = std::min((int)t[xl].s + 1, S.w - 1);
      const int ax = sxLo & ~15;
      tx.push_back(make_int2(ax, (sxHi - ax + 16) >> 4));
    }
    o->ytabOff[l] = (int)ys.size();
    make_resize_tab(S.h, D.h, false, t);
    ys.insert(ys.end(), t.begin(), t.end());
    o->tileYOff[l] = (int)ty.size();
    for (int y0 = 0; y0 < D.h; y0 += kPyH) {
      const int yl = std::min(y0 + kPyH - 1, D.h - 1);
      const int syLo = std::min(std::max((int)t[y0].s, 0), S.h - 1), syHi = std::min(std::max((int)t[yl].s + 1, 0), S.h - 1);
      ty.push_back(make_int2(syLo, syHi - syLo + 1));
    }
  }
  if ((o->tileX.n < tx.size() && o->tileX.alloc(tx.size() + 1)) || (o->tileY.n < ty.size() && o->tileY.alloc(ty.size() + 1))) return PGB_ERR_CUDA;
  if (!tx.empty()) PGB_CUDA(cudaMemcpyAsync(o->tileX.p, tx.data(), tx.size() * sizeof(int2), cudaMemcpyHostToDevice, o->stream));
  if (!ty.empty()) PGB_CUDA(cudaMemcpyAsync(o->tileY.p, ty.data(), ty.size() * sizeof(int2), cudaMemcpyHostToDevice, o->stream));
  if (xs.size() > o->xtab.n || ys.size() > o->ytab.n) return fail(PGB_ERR_CAPACITY, "resize tables exceed capacity");
  if (!xs.empty()) PGB_CUDA(cudaMemcpyAsync(o->xtab.p, xs.data(), xs.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, o->stream));
  if (!ys.empty()) PGB_CUDA(cudaMemcpyAsync(o->ytab.p, ys.data(), ys.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));  // the host vectors die here
  return PGB_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// u32 views of every pyramid level (TMA load, 68x70 halo box) and score-map level (TMA store, 64x64 box).
int build_tmaps(pgb_orb* o) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    PGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = (EncodeTiledFn)fn;
  }
  const OrbGeo& g = o->geo;
  for (int l = 0; l < g.nlevels; l++) {
    const LevelGeo& L = g.lv[l];
    const cuuint64_t dims[3] = {(cuuint64_t)(L.pitch / 4), (cuuint64_t)L.h, (cuuint64_t)o->maxBatch};
    const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.frameStride};
    const cuuint32_t es[3] = {1, 1, 1};
    const cuuint32_t boxIn[3] = {(cuuint32_t)kF2InWords, (cuuint32_t)kF2InRows, 1};
    const cuuint32_t boxOut[3] = {(cuuint32_t)(kF2W / 4), 8, 1};  // one warp's band
    CUresult r = encode(&o->tmaps.in[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, o->pyr.p + L.off, dims, strides, boxIn, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(load, level %d) failed: %d", l, (int)r);
    if (o->score.p) {  // the score map exists only on the unfused / stage-debugging path (ensure_score)
      r = encode(&o->tmaps.out[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, o->score.p + L.off, dims, strides, boxOut, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(store, level %d) failed: %d", l, (int)r);
    }
    const cuuint32_t boxFc[3] = {(cuuint32_t)kFcInWords, (cuuint32_t)(8 * (L.fcClassB == 1 ? g.fcNbB : L.fcClassB == 2 ? 5 : 4) + 6), 1};
    r = encode(&o->tmapsFc.in[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, o->pyr.p + L.off, dims, strides, boxFc, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(fused load, level %d) failed: %d", l, (int)r);
    const cuuint32_t boxPy[3] = {(cuuint32_t)(kPySrcPitch / 4), (cuuint32_t)kPySrcRows, 1};
    r = encode(&o->tmapsPy.in[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, o->pyr.p + L.off, dims, strides, boxPy, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(pyramid source, level %d) failed: %d", l, (int)r);
  }
  return PGB_OK;
}

// Level 0 in place (PGB_IN_DEVICE input that is 16-byte aligned): kernels read the caller's frames where they lie.
int use_external_level0(pgb_orb* o, const uint8_t* gray, size_t pitch, size_t frame_stride, int n_frames) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    PGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = (EncodeTiledFn)fn;
  }
  const LevelGeo& L = o->geo.lv[0];
  o->tmapsCur = o->tmaps;
  const cuuint64_t dims[3] = {(cuuint64_t)(pitch / 4), (cuuint64_t)L.h, (cuuint64_t)n_frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  const cuuint32_t es[3] = {1, 1, 1};
  const cuuint32_t boxIn[3] = {(cuuint32_t)kF2InWords, (cuuint32_t)kF2InRows, 1};
  CUresult r = encode(&o->tmapsCur.in[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(gray), dims, strides, boxIn, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(external level 0) failed: %d", (int)r);
  o->tmapsFcCur = o->tmapsFc;
  const cuuint32_t boxFc[3] = {(cuuint32_t)kFcInWords, (cuuint32_t)(8 * (L.fcClassB == 1 ? o->geo.fcNbB : L.fcClassB == 2 ? 5 : 4) + 6), 1};
  r = encode(&o->tmapsFcCur.in[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(gray), dims, strides, boxFc, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(external level 0, fused) failed: %d", (int)r);
  o->tmapsPyCur = o->tmapsPy;
  const cuuint32_t boxPy[3] = {(cuuint32_t)(kPySrcPitch / 4), (cuuint32_t)kPySrcRows, 1};
  r = encode(&o->tmapsPyCur.in[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(gray), dims, strides, boxPy, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PGB_ERR_CUDA, "cuTensorMapEncodeTiled(external level 0, pyramid) failed: %d", (int)r);
  o->geo.ext0 = gray;
  o->geo.ext0Stride = frame_stride;
  o->geo.ext0Pitch = (int)pitch;
  return PGB_OK;
}

int set_geometry(pgb_orb* o, int w, int h) {
  if (w == o->curW && h == o->curH) return PGB_OK;
  OrbGeo g;
  int rc = build_geo(o, w, h, &g);
  if (rc) return rc;
  const OrbGeo& c = o->capGeo;
  if (g.frameStride > c.frameStride || g.slotsPerFrame > c.slotsPerFrame || g.candPerFrame > c.candPerFrame ||
      g.totalCells > c.totalCells || g.kpCapInternal > c.kpCapInternal || g.maxNodeCap > c.maxNodeCap)
    return fail(PGB_ERR_CAPACITY, "%dx%d frames need more scratch than the %dx%d the handle was created for", w, h,
                o->maxW, o->maxH);
  o->geo = g;
  o->curW = w;
  o->curH = h;
  {
    rc = build_tmaps(o);
    if (rc) return rc;
    std::vector<int4> tab;
    for (int l = 0; l < g.nlevels; l++)
      for (int ty = 0; ty < g.lv[l].tiles2Y; ty++)
        for (int tx = 0; tx < g.lv[l].tiles2X; tx++) tab.push_back(make_int4(l, tx * kF2W, ty * kF2H, 0));
    std::vector<int> ctab;
    for (int l = 0; l < g.nlevels; l++)
      for (int i = 0; i < g.lv[l].nRows; i++)
        for (int j = 0; j < g.lv[l].nCols; j++) ctab.push_back(l | (i << 8) | (j << 20));
    if (o->cellTab.n < ctab.size() && o->cellTab.alloc(ctab.size())) return PGB_ERR_CUDA;
    PGB_CUDA(cudaMemcpyAsync(o->cellTab.p, ctab.data(), ctab.size() * sizeof(int), cudaMemcpyHostToDevice, o->stream));
    PGB_CUDA(cudaStreamSynchronize(o->stream));  // ctab dies at the end of this block
    if (o->tileTab.n < tab.size() && o->tileTab.alloc(tab.size())) return PGB_ERR_CUDA;
    PGB_CUDA(cudaMemcpyAsync(o->tileTab.p, tab.data(), tab.size() * sizeof(int4), cudaMemcpyHostToDevice, o->stream));
    PGB_CUDA(cudaStreamSynchronize(o->stream));
    // fused kernel: one tile per (cell row, group of fcKc cell columns); first tested pixel of cell (i, j) is
    // (19 + j * wCell, 19 + i * hCell) (ORBextractor.cc:789-806: iniX + 3, iniY + 3)
    std::vector<int4> fa, fb, fa5;
    for (int l = 0; l < g.nlevels; l++) {
      const LevelGeo& L = g.lv[l];
      for (int i = 0; i < L.nRows; i++)
        for (int j0 = 0; j0 < L.nCols; j0 += L.fcKc)
          (L.fcClassB == 1 ? fb : L.fcClassB == 2 ? fa5 : fa).push_back(make_int4(l, kEdge + j0 * L.wCell, kEdge + i * L.hCell, i | (j0 << 16)));
    }
    if ((int)fa.size() != g.fcTilesA || (int)fb.size() != g.fcTilesB || (int)fa5.size() != g.fcTilesA5)
      return fail(PGB_ERR_INVALID, "internal: fused tile count");
    if (o->fcTabA5.n < fa5.size() + 1 && o->fcTabA5.alloc(fa5.size() + 1)) return PGB_ERR_CUDA;
    if (!fa5.empty()) PGB_CUDA(cudaMemcpyAsync(o->fcTabA5.p, fa5.data(), fa5.size() * sizeof(int4), cudaMemcpyHostToDevice, o->stream));
    if (o->fcTabA.n < fa.size() + 1 && o->fcTabA.alloc(fa.size() + 1)) return PGB_ERR_CUDA;
    if (o->fcTabB.n < fb.size() + 1 && o->fcTabB.alloc(fb.size() + 1)) return PGB_ERR_CUDA;
    if (!fa.empty()) PGB_CUDA(cudaMemcpyAsync(o->fcTabA.p, fa.data(), fa.size() * sizeof(int4), cudaMemcpyHostToDevice, o->stream));
    if (!fb.empty()) PGB_CUDA(cudaMemcpyAsync(o->fcTabB.p, fb.data(), fb.size() * sizeof(int4), cudaMemcpyHostToDevice, o->stream));
    PGB_CUDA(cudaStreamSynchronize(o->stream));
  }
  return upload_tabs(o);
}

int check_err_flag(pgb_orb* o) {
  int e = 0;
  PGB_CUDA(cudaMemcpyAsync(&e, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  if (e) {
    cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream);
    return fail(PGB_ERR_CAPACITY, "device-side capacity flag 0x%x (1=candidates 2=octree nodes 4=output cap 8=cell chunks)", e);
  }
  return PGB_OK;
}

// The FAST score map in HBM (same layout as the pyramid) is only needed by the unfused kernel pair (PGB_UNFUSED=1,
// pgb_orb_run_stage(2), pgb_orb_get_score_map): allocated on first use, then the TMA store maps are encoded on it.
int ensure_score(pgb_orb* o) {
  if (o->score.p) return PGB_OK;
  const size_t bytes = (size_t)o->maxBatch * o->capGeo.frameStride;
  if (o->score.alloc(bytes)) return PGB_ERR_CUDA;
  PGB_CUDA(cudaMemsetAsync(o->score.p, 0, bytes, o->stream));
  if (o->curW > 0) {
    int rc = build_tmaps(o);
    if (rc) return rc;
    if (o->geo.ext0) {
      rc = use_external_level0(o, o->geo.ext0, (size_t)o->geo.ext0Pitch, (size_t)o->geo.ext0Stride, o->curFrames);
      if (rc) return rc;
    } else {
      o->tmapsCur = o->tmaps;
      o->tmapsFcCur = o->tmapsFc;
      o->tmapsPyCur = o->tmapsPy;
    }
  }
  return PGB_OK;
}

// Stages `from`..`to` over frames [f0, f0+n) of the resident batch; kps/desc/counts are the bases of the WHOLE
// batch's output arrays (frame f0 writes at f0*cap).
int run_stages(pgb_orb* o, int from, int to, pgb_keypoint* kps, uint8_t* desc, int* counts, int cap, int f0, int n,
               cudaStream_t st = nullptr) {
  if (!st) st = o->stream;
  if (n <= 0) return PGB_OK;
  OrbGeo g = o->geo;  // kernels index frames relative to f0: advance the in-place level 0 like the other bases
  if (g.ext0) g.ext0 += (size_t)f0 * g.ext0Stride;
  uint8_t* pyr = o->pyr.p + (size_t)f0 * g.frameStride;
  uint8_t* score = o->score.p ? o->score.p + (size_t)f0 * g.frameStride : nullptr;
  uint32_t* slots = o->slots.p + (size_t)f0 * g.slotsPerFrame;
  int* cellCnt = o->cellCnt.p + (size_t)f0 * g.totalCells;
  unsigned long long* cand = o->cand.p + (size_t)f0 * g.candPerFrame;
  StagedKp* staged = o->staged.p + (size_t)f0 * g.kpCapInternal;
  int* lvlCnt = o->lvlCnt.p + (size_t)f0 * g.nlevels;
  static const char* kStageName[5] = {"pgb:orb:pyramid", "pgb:orb:fast_cells", "pgb:orb:unfused_fast+cells", "pgb:orb:octree", "pgb:orb:orient_desc"};
  for (int s = from; s <= to; s++) {
    NvtxRange range(kStageName[s]);
    switch (s) {
      case 0:
        for (int l = 1; l < g.nlevels; l++)
          launch_pyramid_level(g, o->tmapsPyCur, l, f0, n, pyr, o->xtab.p + o->xtabOff[l], o->ytab.p + o->ytabOff[l],
                               o->tileX.p + o->tileXOff[l], o->tileY.p + o->tileYOff[l], st);
        break;
      case 1:
      {
        // hot path: the fused kernel (score -> per-cell NMS -> candidates, no score map); PGB_UNFUSED=1: the score map kernel
        int rc;
        if (o->unfused) {
          rc = launch_fast_score(g, o->tmapsCur, o->tileTab.p, f0, n, st);
          if (f0 == 0 && n == o->curFrames) o->scoreValid = true;
        } else {
          rc = launch_fast_cells(g, o->tmapsFcCur, o->fcTabA.p, o->fcTabB.p, o->fcTabA5.p, n, slots, cellCnt, o->err.p, st, f0, o->fcSide,
                                 o->evFcFork, o->evFcJoin);
        }
        if (rc) return rc;
        break;
      }
      case 2:  // the round-1 pair, kept for A/B timing and stage-by-stage parity: on the fused path it recomputes the same slots
        if (!o->unfused && !(from == 2 && to == 2)) break;  // ... only when asked for by itself (pgb_orb_run_stage(2))
        if (!o->unfused) {
          int rc = ensure_score(o);
          if (rc) return rc;
          score = o->score.p + (size_t)f0 * g.frameStride;
          rc = launch_fast_score(g, o->tmapsCur, o->tileTab.p, f0, n, st);
          if (rc) return rc;
        }
        launch_cells(g, n, o->cellTab.p, score, slots, cellCnt, o->err.p, st);
        if (f0 == 0 && n == o->curFrames) o->scoreValid = true;
        break;
      case 3:
        if (o->octSide && n > 4) {  // (a handful of frames: the fork / join costs more latency than the priority buys)
          PGB_CUDA(cudaEventRecord(o->evOctFork, st));
          PGB_CUDA(cudaStreamWaitEvent(o->octSide, o->evOctFork, 0));
          launch_octree(g, n, slots, cellCnt, cand, staged, lvlCnt, o->err.p, o->octSide);
          PGB_CUDA(cudaEventRecord(o->evOctJoin, o->octSide));
          PGB_CUDA(cudaStreamWaitEvent(st, o->evOctJoin, 0));
        } else {
          launch_octree(g, n, slots, cellCnt, cand, staged, lvlCnt, o->err.p, st);
        }
        break;
      case 4:
        launch_orient_desc(g, n, pyr, staged, lvlCnt, kps + (size_t)f0 * cap, desc + (size_t)f0 * cap * 32, counts + f0,
                           cap, o->err.p, st);
        break;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PGB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return PGB_OK;
}

// Host frames -> level 0, pipelined: chunks are copied on a dedicated copy stream while the compute stream works
// on the previous chunk, so the PCIe transfer hides behind the kernels (or the other way round).
int extract_host_pipelined(pgb_orb* o, const uint8_t* gray, int n_frames, int width, int height, size_t pitch,
                           size_t frame_stride, pgb_keypoint* kps, uint8_t* desc, int* counts, int cap) {
  const OrbGeo& g = o->geo;
  const int chunk = std::max(1, std::min(o->h2dChunk, n_frames));
  // the copy stream may not overwrite level 0 before the previous call's kernels are done with it
  PGB_CUDA(cudaEventRecord(o->evDone, o->stream));
  PGB_CUDA(cudaStreamWaitEvent(o->copyStream, o->evDone, 0));
  // Chunk schedule (frames): 4, 8, 16, 16, ..., 8, 4 (+ remainder).  Small chunks first so the kernels start while the bulk
  // of the batch is still on the bus; small chunks last so that what the caller waits for after the last byte has
  // arrived is only the kernels of a small chunk (their latency floor, ~0.2 ms, is the same for 1 or 8 frames).
  // Chunks alternate over three compute streams: the latency-bound kernels of one chunk (octree, small grids) overlap
  // the throughput-bound ones of its neighbours, so chunking does not cost kernel efficiency.
  cudaStream_t cs[3] = {o->stream, o->auxStream[0], o->auxStream[1]};
  for (int a = 1; a < 3; a++) PGB_CUDA(cudaStreamWaitEvent(cs[a], o->evDone, 0));
  int used = 0;
  for (int f0 = 0, k = 0, n = 0; f0 < n_frames; f0 += n, k++) {
    n = std::min(chunk, std::max(std::min(o->h2dMinChunk, n_frames - f0), (n_frames - f0) / 2));
    if (k < 2) n = std::min(n, std::max(1, chunk >> (2 - k)));
    cudaStream_t st = cs[k % 3];
    used = std::max(used, std::min(k, 2));
    if (pitch == (size_t)width && g.lv[0].pitch == width) {
      // contiguous frames: ONE 2-D copy per chunk whose "rows" are whole frames (a DMA descriptor per frame costs ~10 us)
      PGB_CUDA(cudaMemcpy2DAsync(o->pyr.p + (size_t)f0 * g.frameStride + g.lv[0].off, g.frameStride,
                                 gray + (size_t)f0 * frame_stride, frame_stride, (size_t)width * height, n,
                                 cudaMemcpyHostToDevice, o->copyStream));
    } else {
      for (int f = f0; f < f0 + n; f++)
        PGB_CUDA(cudaMemcpy2DAsync(o->pyr.p + (size_t)f * g.frameStride + g.lv[0].off, g.lv[0].pitch,
                                   gray + (size_t)f * frame_stride, pitch, width, height, cudaMemcpyHostToDevice,
                                   o->copyStream));
    }
    cudaEvent_t ev = o->evChunk[k % kMaxChunkEvents];
    PGB_CUDA(cudaEventRecord(ev, o->copyStream));
    PGB_CUDA(cudaStreamWaitEvent(st, ev, 0));
    int rc = run_stages(o, 0, 4, kps, desc, counts, cap, f0, n, st);
    if (rc) return rc;
  }
  for (int a = 1; a <= used; a++) {  // join: everything after this call on the handle's stream sees all chunks
    PGB_CUDA(cudaEventRecord(o->evAux[a - 1], cs[a]));
    PGB_CUDA(cudaStreamWaitEvent(o->stream, o->evAux[a - 1], 0));
  }
  return PGB_OK;
}

// Device-resident batches: optionally (PGB_RES_CHUNK = frames per chunk) the same three compute streams, no copies: the
// batch is cut into chunks that alternate over the streams so that the latency-bound kernels of one chunk (octree: 8 x n
// CTAs, the small pyramid levels, every kernel's tail wave) can overlap the throughput-bound kernels of its neighbours.
// OFF by default -- measured on 128-frame steps (frames/s): single pass 47.6 k, chunks of 64: 47.0 k, 32: 45.1 k, 16: 44.2 k.
// The big kernels are issue-bound, so co-running them only splits the SMs, and the smaller launches have longer tails.
int run_stages_resident(pgb_orb* o, pgb_keypoint* kps, uint8_t* desc, int* counts, int cap, int n_frames) {
  const int chunk = o->resChunk;
  if (chunk <= 0 || n_frames < 2 * chunk) return run_stages(o, 0, 4, kps, desc, counts, cap, 0, n_frames, o->stream);
  cudaStream_t cs[3] = {o->stream, o->auxStream[0], o->auxStream[1]};
  PGB_CUDA(cudaEventRecord(o->evDone, o->stream));  // level 0 in place / copied, previous consumers of the outputs done
  for (int a = 1; a < 3; a++) PGB_CUDA(cudaStreamWaitEvent(cs[a], o->evDone, 0));
  int used = 0;
  for (int f0 = 0, k = 0; f0 < n_frames; k++) {
    int n = std::min(chunk, n_frames - f0);
    if (n_frames - f0 - n < chunk / 2) n = n_frames - f0;  // no tiny last chunk
    used = std::max(used, std::min(k, 2));
    int rc = run_stages(o, 0, 4, kps, desc, counts, cap, f0, n, cs[k % 3]);
    if (rc) return rc;
    f0 += n;
  }
  for (int a = 1; a <= used; a++) {
    PGB_CUDA(cudaEventRecord(o->evAux[a - 1], cs[a]));
    PGB_CUDA(cudaStreamWaitEvent(o->stream, o->evAux[a - 1], 0));
  }
  return PGB_OK;
}

}  // namespace

extern "C" {

const char* pgb_last_error(void) { return g_last_error.c_str(); }
int pgb_version(void) { return 100; }
uint64_t pgb_launch_count(void) { return g_launches.load(); }

pgb_orb* pgb_orb_create(int device, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                        int max_width, int max_height, int max_batch, void* stream) {
  if (nfeatures <= 0 || nlevels <= 0 || nlevels > kMaxLevels || !(scale_factor > 1.0f) || max_width <= 0 ||
      max_height <= 0 || max_batch <= 0 || ini_th_fast < min_th_fast || min_th_fast < 0 || ini_th_fast > 255) {
    fail(PGB_ERR_INVALID, "pgb_orb_create: invalid argument");
    return nullptr;
  }
  if (use_device(device)) return nullptr;
  pgb_orb* o = new pgb_orb;
  o->device = device;
  o->nfeatures = nfeatures; o->nlevels = nlevels; o->iniTh = ini_th_fast; o->minTh = min_th_fast;
  o->scaleFactor = scale_factor;
  o->maxW = max_width; o->maxH = max_height; o->maxBatch = max_batch;
  // scale tables and per-level quotas, ORBextractor.cc:415-447 (the fork sizes them nlevels+1; [0,nlevels) is read)
  o->scale.resize(nlevels + 1); o->sigma2.resize(nlevels + 1);
  o->invScale.resize(nlevels + 1); o->invSigma2.resize(nlevels + 1);
  o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
  for (int i = 1; i <= nlevels; i++) {
    o->scale[i] = o->scale[i - 1] * scale_factor;
    o->sigma2[i] = o->scale[i] * o->scale[i];
  }
  for (int i = 0; i <= nlevels; i++) {
    o->invScale[i] = 1.0f / o->scale[i];
    o->invSigma2[i] = 1.0f / o->sigma2[i];
  }
  o->nPerLevel.resize(nlevels + 1);
  const float factor = 1.0f / scale_factor;
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels; l++) {
    o->nPerLevel[l] = cv_round_f(nDesired);
    sum += o->nPerLevel[l];
    nDesired *= factor;
  }
  o->nPerLevel[nlevels] = std::max(nfeatures - sum, 0);

  auto bail = [&](const char* what) -> pgb_orb* {
    std::string keep = g_last_error;
    pgb_orb_destroy(o);
    g_last_error = keep.empty() ? what : keep;
    return nullptr;
  };
  if (build_geo(o, max_width, max_height, &o->capGeo)) return bail("geometry");
  if (configure_fast_score()) return bail("cudaFuncSetAttribute failed");
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) o->numSMs = prop.multiProcessorCount;
  }
  if (stream) o->stream = (cudaStream_t)stream;
  else {
    if (cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
    o->ownStream = true;
  }
  if (cudaStreamCreateWithFlags(&o->copyStream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&o->evDone, cudaEventDisableTiming) != cudaSuccess)
    return bail("cudaStreamCreate/cudaEventCreate failed");
  for (int k = 0; k < kMaxChunkEvents; k++)
    if (cudaEventCreateWithFlags(&o->evChunk[k], cudaEventDisableTiming) != cudaSuccess) return bail("cudaEventCreate failed");
  for (int a = 0; a < 2; a++)
    if (cudaStreamCreateWithFlags(&o->auxStream[a], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&o->evAux[a], cudaEventDisableTiming) != cudaSuccess)
      return bail("cudaStreamCreate/cudaEventCreate failed");
  if (!getenv("PGB_OCT_NO_SIDE")) {
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&o->octSide, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&o->evOctFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&o->evOctJoin, cudaEventDisableTiming) != cudaSuccess)
      return bail("cudaStreamCreateWithPriority/cudaEventCreate failed");
  }
  if (!getenv("PGB_FC_NO_SIDE") &&
      (cudaStreamCreateWithFlags(&o->fcSide, cudaStreamNonBlocking) != cudaSuccess ||
       cudaEventCreateWithFlags(&o->evFcFork, cudaEventDisableTiming) != cudaSuccess ||
       cudaEventCreateWithFlags(&o->evFcJoin, cudaEventDisableTiming) != cudaSuccess))
    return bail("cudaStreamCreate/cudaEventCreate failed");
  if (const char* e = getenv("PGB_H2D_CHUNK")) o->h2dChunk = std::max(1, atoi(e));
  if (const char* e = getenv("PGB_H2D_MIN_CHUNK")) o->h2dMinChunk = std::max(1, atoi(e));
  if (const char* e = getenv("PGB_RES_CHUNK")) o->resChunk = std::max(0, atoi(e));
  const OrbGeo& c = o->capGeo;
  const size_t B = (size_t)max_batch;
  o->outCap = 0;
  for (int l = 0; l < nlevels; l++) o->outCap += c.lv[l].nodeCap;
  int tabX = 0, tabY = 0;
  for (int l = 1; l < nlevels; l++) { tabX += c.lv[l].w + 8; tabY += c.lv[l].h + 8; }
  if (const char* e = getenv("PGB_UNFUSED")) o->unfused = atoi(e) != 0;
  if (o->pyr.alloc(B * c.frameStride) || (o->unfused && o->score.alloc(B * c.frameStride)) || o->slots.alloc(B * c.slotsPerFrame) ||
      o->cellCnt.alloc(B * c.totalCells) || o->lvlCnt.alloc(B * nlevels) || o->err.alloc(1) || o->counts.alloc(B) ||
      o->cand.alloc(B * c.candPerFrame) || o->staged.alloc(B * c.kpCapInternal) || o->xtab.alloc(tabX + 8) ||
      o->ytab.alloc(tabY + 8) || o->kps.alloc(B * o->outCap) || o->desc.alloc(B * o->outCap * 32) ||
      o->tmpLevel.alloc((size_t)max_width * max_height))
    return bail("cudaMalloc failed");
  if (cudaMemsetAsync(o->err.p, 0, sizeof(int), o->stream) != cudaSuccess ||
      cudaMemsetAsync(o->pyr.p, 0, B * c.frameStride, o->stream) != cudaSuccess ||
      (o->score.p && cudaMemsetAsync(o->score.p, 0, B * c.frameStride, o->stream) != cudaSuccess) ||
      cudaStreamSynchronize(o->stream) != cudaSuccess)
    return bail("cudaMemset failed");
  return o;
}

void pgb_orb_destroy(pgb_orb* o) {
  if (!o) return;
  cudaSetDevice(o->device);
  if (o->stream) cudaStreamSynchronize(o->stream);
  if (o->copyStream) { cudaStreamSynchronize(o->copyStream); cudaStreamDestroy(o->copyStream); }
  if (o->evDone) cudaEventDestroy(o->evDone);
  for (int a = 0; a < 2; a++) {
    if (o->auxStream[a]) { cudaStreamSynchronize(o->auxStream[a]); cudaStreamDestroy(o->auxStream[a]); }
    if (o->evAux[a]) cudaEventDestroy(o->evAux[a]);
    if (a == 0) {
      if (o->fcSide) { cudaStreamSynchronize(o->fcSide); cudaStreamDestroy(o->fcSide); }
      if (o->evFcFork) cudaEventDestroy(o->evFcFork);
      if (o->evFcJoin) cudaEventDestroy(o->evFcJoin);
      if (o->octSide) { cudaStreamSynchronize(o->octSide); cudaStreamDestroy(o->octSide); }
      if (o->evOctFork) cudaEventDestroy(o->evOctFork);
      if (o->evOctJoin) cudaEventDestroy(o->evOctJoin);
    }
  }
  for (int k = 0; k < kMaxChunkEvents; k++)
    if (o->evChunk[k]) cudaEventDestroy(o->evChunk[k]);
  if (o->ownStream && o->stream) cudaStreamDestroy(o->stream);
  delete o;
}

int pgb_orb_levels(const pgb_orb* o) { return o ? o->nlevels : PGB_ERR_INVALID; }
float pgb_orb_scale_factor(const pgb_orb* o) { return o ? o->scaleFactor : 0.f; }
int pgb_orb_scale_factors(const pgb_orb* o, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  for (int i = 0; i < o->nlevels; i++) {
    if (scale) scale[i] = o->scale[i];
    if (inv_scale) inv_scale[i] = o->invScale[i];
    if (sigma2) sigma2[i] = o->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = o->invSigma2[i];
  }
  return PGB_OK;
}
int pgb_orb_features_per_level(const pgb_orb* o, int32_t* n) {
  if (!o || !n) return fail(PGB_ERR_INVALID, "null argument");
  for (int i = 0; i < o->nlevels; i++) n[i] = o->nPerLevel[i];
  return PGB_OK;
}
int pgb_orb_max_keypoints(const pgb_orb* o) { return o ? o->outCap : PGB_ERR_INVALID; }
int pgb_orb_level_size(const pgb_orb* o, int width, int height, int level, int* w, int* h) {
  if (!o || level < 0 || level >= o->nlevels) return fail(PGB_ERR_INVALID, "bad level");
  *w = cv_round_f((float)width * o->invScale[level]);
  *h = cv_round_f((float)height * o->invScale[level]);
  return PGB_OK;
}
void* pgb_orb_stream(pgb_orb* o) { return o ? (void*)o->stream : nullptr; }

int pgb_orb_extract(pgb_orb* o, const uint8_t* gray, int is_device, int n_frames, int width, int height, size_t pitch,
                    size_t frame_stride, pgb_keypoint* kps, uint8_t* desc, int32_t* counts, int cap) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  NvtxRange range("pgb_orb_extract");
  if (n_frames < 0 || n_frames > o->maxBatch) return fail(PGB_ERR_INVALID, "n_frames %d outside [0,%d]", n_frames, o->maxBatch);
  if (!counts || (cap > 0 && (!kps || !desc)) || cap < 0) return fail(PGB_ERR_INVALID, "null output buffer");
  PGB_CUDA(cudaSetDevice(o->device));
  if (n_frames == 0) return PGB_OK;
  if (width == 0 || height == 0) {  // empty image: the reference returns without touching the outputs (:1045)
    if (is_device & PGB_OUT_DEVICE) PGB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_frames, o->stream));
    else memset(counts, 0, sizeof(int32_t) * n_frames);
    return PGB_OK;
  }
  if (!gray || width < 0 || height < 0 || pitch < (size_t)width) return fail(PGB_ERR_INVALID, "bad image arguments");
  int rc = set_geometry(o, width, height);
  if (rc) return rc;
  const OrbGeo& g = o->geo;
  o->curFrames = n_frames;
  const bool inDev = (is_device & PGB_IN_DEVICE) != 0, outDev = (is_device & PGB_OUT_DEVICE) != 0;
  pgb_keypoint* dk = outDev ? kps : o->kps.p;
  uint8_t* dd = outDev ? desc : o->desc.p;
  int* dc = outDev ? counts : o->counts.p;
  const int dcap = outDev ? cap : o->outCap;  // a caller capacity below pgb_orb_max_keypoints() raises the device flag
  o->geo.ext0 = nullptr;
  o->tmapsCur = o->tmaps;
  o->tmapsFcCur = o->tmapsFc;
  o->tmapsPyCur = o->tmapsPy;
  o->scoreValid = false;
  if (inDev && ((size_t)gray & 15) == 0 && (pitch & 15) == 0 && (frame_stride & 15) == 0 && frame_stride >= pitch * (size_t)height) {
    // level 0 is read in place: no copy; the frames must stay valid until the next extract call on this handle
    rc = use_external_level0(o, gray, pitch, frame_stride, n_frames);
    if (rc) return rc;
    rc = run_stages_resident(o, dk, dd, dc, dcap, n_frames);
  } else if (inDev) {
    if (pitch == (size_t)width && g.lv[0].pitch == width) {
      // contiguous frames: ONE 2-D copy whose "rows" are whole frames (128 separate 2 MB copies cost 0.8 ms of gaps)
      PGB_CUDA(cudaMemcpy2DAsync(o->pyr.p + g.lv[0].off, g.frameStride, gray, frame_stride, (size_t)width * height, n_frames,
                                 cudaMemcpyDeviceToDevice, o->stream));
    } else {
      for (int f = 0; f < n_frames; f++)
        PGB_CUDA(cudaMemcpy2DAsync(o->pyr.p + (size_t)f * g.frameStride + g.lv[0].off, g.lv[0].pitch,
                                   gray + (size_t)f * frame_stride, pitch, width, height, cudaMemcpyDeviceToDevice,
                                   o->stream));
    }
    rc = run_stages_resident(o, dk, dd, dc, dcap, n_frames);
  } else {
    rc = extract_host_pipelined(o, gray, n_frames, width, height, pitch, frame_stride, dk, dd, dc, dcap);
  }
  if (rc || outDev) return rc;
  std::vector<int32_t> hc(n_frames);
  PGB_CUDA(cudaMemcpyAsync(hc.data(), o->counts.p, sizeof(int32_t) * n_frames, cudaMemcpyDeviceToHost, o->stream));
  rc = check_err_flag(o);  // synchronises
  if (rc) return rc;
  for (int f = 0; f < n_frames; f++) {
    if (hc[f] > cap) return fail(PGB_ERR_CAPACITY, "frame %d produced %d keypoints, caller cap is %d", f, hc[f], cap);
    counts[f] = hc[f];
    if (hc[f] > 0) {
      PGB_CUDA(cudaMemcpyAsync(kps + (size_t)f * cap, o->kps.p + (size_t)f * o->outCap, sizeof(pgb_keypoint) * hc[f],
                               cudaMemcpyDeviceToHost, o->stream));
      PGB_CUDA(cudaMemcpyAsync(desc + (size_t)f * cap * 32, o->desc.p + (size_t)f * o->outCap * 32, (size_t)32 * hc[f],
                               cudaMemcpyDeviceToHost, o->stream));
    }
  }
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  return PGB_OK;
}

int pgb_orb_check(pgb_orb* o) {
  if (!o) return fail(PGB_ERR_INVALID, "null handle");
  PGB_CUDA(cudaSetDevice(o->device));
  return check_err_flag(o);
}

int pgb_orb_run_stage(pgb_orb* o, int which) {
  if (!o || which < 0 || which > 4) return fail(PGB_ERR_INVALID, "bad stage");
  if (o->curFrames <= 0) return fail(PGB_ERR_INVALID, "no frames resident: call pgb_orb_extract first");
  PGB_CUDA(cudaSetDevice(o->device));
  return run_stages(o, which, which, o->kps.p, o->desc.p, o->counts.p, o->outCap, 0, o->curFrames);
}

static int copy_level_out(pgb_orb* o, const uint8_t* base, int frame, int level, uint8_t* out, int* w, int* h) {
  if (!o || frame < 0 || frame >= o->curFrames || level < 0 || level >= o->nlevels)
    return fail(PGB_ERR_INVALID, "bad frame/level");
  PGB_CUDA(cudaSetDevice(o->device));
  const LevelGeo& L = o->geo.lv[level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  if (!out) return PGB_OK;
  if (level == 0 && base == o->pyr.p && o->geo.ext0)  // level 0 read in place from the caller's (still valid) frames
    PGB_CUDA(cudaMemcpy2DAsync(out, L.w, o->geo.ext0 + (size_t)frame * o->geo.ext0Stride, o->geo.ext0Pitch, L.w, L.h,
                               cudaMemcpyDeviceToHost, o->stream));
  else
    PGB_CUDA(cudaMemcpy2DAsync(out, L.w, base + (size_t)frame * o->geo.frameStride + L.off, L.pitch, L.w, L.h,
                               cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  return PGB_OK;
}

int pgb_orb_get_level(pgb_orb* o, int frame, int level, uint8_t* out, int* w, int* h) {
  return copy_level_out(o, o ? o->pyr.p : nullptr, frame, level, out, w, h);
}
int pgb_orb_get_score_map(pgb_orb* o, int frame, int level, uint8_t* out, int* w, int* h) {
  if (!o || frame < 0 || frame >= o->curFrames || level < 0 || level >= o->nlevels) return fail(PGB_ERR_INVALID, "bad frame/level");
  PGB_CUDA(cudaSetDevice(o->device));
  if (out && !o->scoreValid) {  // the hot path never materialises the score map: produce it now with the score kernel
    int rc = ensure_score(o);
    if (rc) return rc;
    rc = launch_fast_score(o->geo, o->tmapsCur, o->tileTab.p, 0, o->curFrames, o->stream);
    if (rc) return rc;
    o->scoreValid = true;
  }
  return copy_level_out(o, o->score.p, frame, level, out, w, h);
}

int pgb_orb_get_blurred_level(pgb_orb* o, int frame, int level, uint8_t* out, int* w, int* h) {
  if (!o || frame < 0 || frame >= o->curFrames || level < 0 || level >= o->nlevels)
    return fail(PGB_ERR_INVALID, "bad frame/level");
  PGB_CUDA(cudaSetDevice(o->device));
  const LevelGeo& L = o->geo.lv[level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  if (!out) return PGB_OK;
  launch_blur_level(o->geo, level, frame, o->pyr.p, o->tmpLevel.p, o->stream);
  PGB_CUDA(cudaGetLastError());
  PGB_CUDA(cudaMemcpyAsync(out, o->tmpLevel.p, (size_t)L.w * L.h, cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  return PGB_OK;
}

int pgb_orb_get_candidates(pgb_orb* o, int frame, int level, int32_t* xyr, int cap, int32_t* n) {
  if (!o || frame < 0 || frame >= o->curFrames || level < 0 || level >= o->nlevels || !n)
    return fail(PGB_ERR_INVALID, "bad frame/level");
  PGB_CUDA(cudaSetDevice(o->device));
  const OrbGeo& g = o->geo;
  const LevelGeo& L = g.lv[level];
  const int nCells = L.nCols * L.nRows;
  std::vector<int> cnt(nCells);
  PGB_CUDA(cudaMemcpyAsync(cnt.data(), o->cellCnt.p + (size_t)frame * g.totalCells + L.cellBase, sizeof(int) * nCells,
                           cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  long total = 0;
  for (int c : cnt) total += c;
  *n = (int32_t)total;
  if (!xyr) return PGB_OK;
  if (total > cap) return fail(PGB_ERR_CAPACITY, "%ld candidates, cap %d", total, cap);
  std::vector<uint32_t> sl((size_t)nCells * L.slotCap);
  PGB_CUDA(cudaMemcpyAsync(sl.data(), o->slots.p + (size_t)frame * g.slotsPerFrame + L.slotBase,
                           sizeof(uint32_t) * sl.size(), cudaMemcpyDeviceToHost, o->stream));
  PGB_CUDA(cudaStreamSynchronize(o->stream));
  size_t k = 0;
  for (int c = 0; c < nCells; c++)
    for (int q = 0; q < cnt[c]; q++) {
      const uint32_t p = sl[(size_t)c * L.slotCap + q];
      xyr[3 * k] = (int)(p & 0xfff);
      xyr[3 * k + 1] = (int)((p >> 12) & 0xfff);
      xyr[3 * k + 2] = (int)(p >> 24);
      k++;
    }
  return PGB_OK;
}

}  // extern "C"
