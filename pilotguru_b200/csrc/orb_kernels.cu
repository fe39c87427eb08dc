// sm_100a kernels of the ORB extraction path.  All arithmetic is integer or fp32 with explicit round-to-nearest
// intrinsics (no FMA contraction), so results are bit-exact against the CPU oracle.
//
// Reference semantics (file:line in waiwnf/pilotguru, thirdparty/orb-slam2/src/ORBextractor.cc):
//   k_pyramid      ComputePyramid :1106-1131 -> cv::resize(INTER_LINEAR) fixed-point bilinear
//   k_fast_score   cv::FAST(TYPE_9_16) corner score (call sites :809,:814)
//   k_cells        per-cell threshold iniThFAST/minThFAST + 3x3 NMS :789-829
//   k_octree       DistributeOctTree :539-763, DivideNode :481-537
//   k_orient_desc  IC_Angle :77-104, GaussianBlur :1084-1085, computeOrbDescriptor :108-147, rescale :1094-1100
#include <cuda_runtime.h>

#include "../../include/pgb200_orb_pattern.h"
#include <cstdlib>

#include "common.cuh"
#include "fast_common.cuh"
#include "orb_kernels.cuh"

namespace pgb {

// =========================================================================================== K1 pyramid
// One thread produces 4 horizontally adjacent destination pixels (one 32-bit store).
__global__ void __launch_bounds__(256) k_pyramid(OrbGeo g, int level, uint8_t* __restrict__ pyr,
                                                 const ResizeTab* __restrict__ xtab,
                                                 const ResizeTab* __restrict__ ytab) {
  const LevelGeo& D = g.lv[level];
  const LevelGeo& S = g.lv[level - 1];
  const int wq = (D.w + 3) >> 2;
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xq >= wq) return;
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.frameStride;
  int sPitch;
  const uint8_t* src = level_base(g, pyr, level - 1, blockIdx.z, &sPitch);
  const ResizeTab ty = ytab[y];
  const int sy0 = min(max((int)ty.s, 0), S.h - 1), sy1 = min(max((int)ty.s + 1, 0), S.h - 1);
  const uint8_t* S0 = src + (size_t)sy0 * sPitch;
  const uint8_t* S1 = src + (size_t)sy1 * sPitch;
  uint32_t packed = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = xq * 4 + i;
    if (x < D.w) {
      const ResizeTab tx = xtab[x];
      const int sx = tx.s, sx1 = min(sx + 1, S.w - 1);
      const int r0 = (int)__ldg(S0 + sx) * tx.a0 + (int)__ldg(S0 + sx1) * tx.a1;
      const int r1 = (int)__ldg(S1 + sx) * tx.a0 + (int)__ldg(S1 + sx1) * tx.a1;
      const int v = ((((int)ty.a0 * (r0 >> 4)) >> 16) + (((int)ty.a1 * (r1 >> 4)) >> 16) + 2) >> 2;
      packed |= (uint32_t)(v & 0xff) << (8 * i);
    }
  }
  *reinterpret_cast<uint32_t*>(frame + D.off + (size_t)y * D.pitch + (size_t)xq * 4) = packed;
}

// Tiled variant: a CTA produces a 256 x 32 destination tile; the source footprint of every tile column / tile row
// comes from small host-built tables (one independent load each: the round-1 profile showed 56 % of the kernel's
// stall samples before the first barrier, on the dependent xtab/ytab -> address -> load chain).  The source rows it needs are staged in shared memory
// with 16-byte loads, the horizontal pass runs once per source row (not once per destination row, a 1.3x saving at
// scale 1.2) and the vertical pass reads its two rows from shared memory.  Same integer arithmetic as k_pyramid.

__global__ void __launch_bounds__(256) k_pyramid_tiled(OrbGeo g, int level, uint8_t* __restrict__ pyr,
                                                       const ResizeTab* __restrict__ xtab,
                                                       const ResizeTab* __restrict__ ytab,
                                                       const int2* __restrict__ tileX, const int2* __restrict__ tileY) {
  __shared__ __align__(16) uint8_t s_src[kPySrcRows * kPySrcPitch];
  __shared__ uint16_t s_h[kPySrcRows * kPyW];
  __shared__ ResizeTab s_ty[kPyH];
  const LevelGeo& D = g.lv[level];
  const LevelGeo& S = g.lv[level - 1];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kPyW, y0 = blockIdx.y * kPyH;
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.frameStride;
  int sPitch;
  const uint8_t* src = level_base(g, pyr, level - 1, blockIdx.z, &sPitch);
  const int2 fx = __ldg(&tileX[blockIdx.x]), fy = __ldg(&tileY[blockIdx.y]);
  const int ax = fx.x, nvec = fx.y, syLo = fy.x, nrows = fy.y;  // host guarantees nvec*16 <= pitch, nrows <= rows
  if (tid < kPyH) s_ty[tid] = ytab[min(y0 + tid, D.h - 1)];
  {  // staging: warp = source row (stride 8), lane = 16-byte vector (nvec <= 26): no index division
    const int warp = tid >> 5, lane = tid & 31;
    if (lane < nvec) {
      const int gx = ax + lane * 16;
      const bool in = gx < sPitch;
      for (int r = warp; r < nrows; r += 8) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (in) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(syLo + r) * sPitch + gx));
        *reinterpret_cast<uint4*>(s_src + r * kPySrcPitch + lane * 16) = v;
      }
    }
  }
  __syncthreads();
  // horizontal pass: thread = destination column, 4 source rows in flight
  {
    const int x = min(x0 + tid, D.w - 1);
    const ResizeTab tx = xtab[x];
    const int o0 = tx.s - ax, o1 = min((int)tx.s + 1, S.w - 1) - ax;
    const int a0 = tx.a0, a1 = tx.a1;
    const uint8_t* c0 = s_src + o0;
    const uint8_t* c1 = s_src + o1;
    uint16_t* hcol = s_h + tid;
    int r = 0;
    for (; r + 4 <= nrows; r += 4) {
      int v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = (int)c0[(r + u) * kPySrcPitch] * a0 + (int)c1[(r + u) * kPySrcPitch] * a1;
#pragma unroll
      for (int u = 0; u < 4; u++) hcol[(r + u) * kPyW] = (uint16_t)(v[u] >> 4);
    }
    for (; r < nrows; r++) hcol[r * kPyW] = (uint16_t)(((int)c0[r * kPySrcPitch] * a0 + (int)c1[r * kPySrcPitch] * a1) >> 4);
  }
  __syncthreads();
  // vertical pass: thread = 4 adjacent columns x 8 rows; the 4 u16 of a row come in one 8-byte load, one 32-bit store
  // per row.  Columns beyond the level's width land in the pitch padding (pitch is a multiple of 64 >= w), which no
  // consumer reads, so they are not masked.
  {
    const int q = tid & 63, rr = tid >> 6;
    const int gx = x0 + q * 4;
    if (gx < D.w) {
#pragma unroll
      for (int k = 0; k < kPyH / 4; k++) {
        const int ry = rr + 4 * k, gy = y0 + ry;
        if (gy >= D.h) break;
        const ResizeTab ty = s_ty[ry];
        const int r0 = min(max((int)ty.s, 0), S.h - 1) - syLo, r1 = min(max((int)ty.s + 1, 0), S.h - 1) - syLo;
        const uint2 h0 = *reinterpret_cast<const uint2*>(s_h + r0 * kPyW + q * 4);
        const uint2 h1 = *reinterpret_cast<const uint2*>(s_h + r1 * kPyW + q * 4);
        const int b0 = ty.a0, b1 = ty.a1;
        const int v0 = (((b0 * (int)(h0.x & 0xffffu)) >> 16) + ((b1 * (int)(h1.x & 0xffffu)) >> 16) + 2) >> 2;
        const int v1 = (((b0 * (int)(h0.x >> 16)) >> 16) + ((b1 * (int)(h1.x >> 16)) >> 16) + 2) >> 2;
        const int v2 = (((b0 * (int)(h0.y & 0xffffu)) >> 16) + ((b1 * (int)(h1.y & 0xffffu)) >> 16) + 2) >> 2;
        const int v3 = (((b0 * (int)(h0.y >> 16)) >> 16) + ((b1 * (int)(h1.y >> 16)) >> 16) + 2) >> 2;
        // every v is in [0, 255] (a convex combination of bytes), so the bytes can be packed without masking
        *reinterpret_cast<uint32_t*>(frame + D.off + (size_t)gy * D.pitch + gx) =
            (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
      }
    }
  }
}

// Round-2 kernel (the default).  The round-1 profile of k_pyramid_tiled (profiles/r01e_other_kernels_sass_regions.md) showed
// 4.7 M warp-instructions per frame, 42 % of them in a horizontal pass that went through shared memory as u16 and 40 % in
// a vertical pass that read it back; a first rewrite with one byte load per source pixel (thread = one destination column)
// cut the instructions by 18 % but saturated the shared-memory pipe instead (l1tex 94 %, profiles/r02_kernels_ncu_full.md).
// This one is built around LOAD count:
//   * the source footprint of a 256 x 32 tile comes in with ONE TMA load (no staging instructions);
//   * a thread owns FOUR adjacent destination columns and 8 destination rows.  At scale <= 1.25 the 8 source bytes its four
//     columns need from a source row lie inside a 12-byte aligned window: three 32-bit shared loads per source row (instead
//     of 8 byte loads), two funnel shifts align the window to the first column's left pixel, and two PRMT with per-thread
//     selectors (constant down the tile) produce (left, right) byte pairs, one pair per 16-bit half;
//   * IDP.2A (dp2a) multiplies a pair by the column's (a0, a1) coefficient pair: one instruction per horizontal
//     interpolation; the vertical interpolation is two IMAD + shifts; the four results leave as one 32-bit global store
//     (a warp writes 128 contiguous bytes).
// Same integer arithmetic as k_pyramid, bit for bit (all values are non-negative and below 2^27).
#ifndef PGB_PY_RPT
#define PGB_PY_RPT 16
#endif
constexpr int kPyRpt = PGB_PY_RPT;  // destination rows per thread (a thread's set-up -- 12 table loads, selectors -- is paid once per kPyRpt rows)
constexpr int kPyThreads = 64 * (kPyH / kPyRpt);
__global__ void __launch_bounds__(kPyThreads) k_pyramid_walk(const __grid_constant__ OrbGeo g, const __grid_constant__ TmapIn tm,
                                                      int level, int frame0, uint8_t* __restrict__ pyr,
                                                      const ResizeTab* __restrict__ xtab, const ResizeTab* __restrict__ ytab,
                                                      const int2* __restrict__ tileX, const int2* __restrict__ tileY) {
  __shared__ __align__(128) uint8_t s_src[kPySrcRows * kPySrcPitch];
  __shared__ uint4 s_row[kPyH];
  __shared__ __align__(8) uint64_t s_bar;
  using namespace fastk;
  const LevelGeo& D = g.lv[level];
  const LevelGeo& S = g.lv[level - 1];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kPyW, y0 = blockIdx.y * kPyH;
  const int2 fx = __ldg(&tileX[blockIdx.x]), fy = __ldg(&tileY[blockIdx.y]);
  const int ax = fx.x, syLo = fy.x;  // first staged source byte (16-aligned) / row; the box is kPySrcPitch x kPySrcRows
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_bar, kPySrcRows * kPySrcPitch);
    tma_load_3d(s_src, &tm.in[level - 1], ax >> 2, syLo, blockIdx.z + frame0, &s_bar);
  }
  if (tid < kPyH) {  // per destination row: byte offsets of its two source rows inside the staged box + the two coefficients
    const ResizeTab ty = ytab[min(y0 + tid, D.h - 1)];
    const int r0 = min(max((int)ty.s, 0), S.h - 1) - syLo, r1 = min(max((int)ty.s + 1, 0), S.h - 1) - syLo;
    // bit 31 of .y: the row's upper source row is the previous destination row's lower one (the common case at scale 1.2:
    // consecutive destination rows advance by one source row), so its interpolation is already in registers -- except for
    // the first row of a thread's group of 8
    const ResizeTab tp = ytab[min(max(y0 + tid - 1, 0), D.h - 1)];
    const int p1 = min(max((int)tp.s + 1, 0), S.h - 1) - syLo;
    const bool reuse = (tid % kPyRpt) != 0 && y0 + tid < D.h && p1 == r0;
    s_row[tid] = make_uint4((uint32_t)(r0 * kPySrcPitch), (uint32_t)(r1 * kPySrcPitch) | (reuse ? 0x80000000u : 0u), (uint32_t)ty.a0,
                            (uint32_t)ty.a1);
  }
  // this thread's four columns: coefficient pairs (a0 | a1 << 16) and the byte offsets of their left pixels inside the
  // 8-byte window that starts at the first column's left pixel.  The right neighbour is always the next byte: where
  // cv::resize clamps it to the last column its coefficient a1 is 0 (make_resize_tab, clampCoef) and the byte read instead
  // is still inside the staged box.
  const int cg = tid & 63, rg = tid >> 6;
  const int gx = x0 + 4 * cg;
  // Columns beyond the level's width land in the pitch padding (a multiple of 64 >= w).  Threads whose four columns or eight
  // rows lie entirely outside the level have nothing to compute or write: at 1080p the edge tiles make them 15 % of all
  // threads (43 % on level 7, whose 536 columns take three 256-wide tiles).  They leave before the coefficient set-up;
  // the barrier below counts the threads that are still alive (warp 0 -- the TMA issue and the row table -- never leaves).
  if (gx >= D.pitch || y0 + rg * kPyRpt >= D.h) return;
  uint32_t coef[4];
  int off[4];
  int s0 = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const ResizeTab tx = xtab[min(gx + k, D.w - 1)];
    if (k == 0) s0 = tx.s;
    coef[k] = (uint32_t)(uint16_t)tx.a0 | ((uint32_t)(uint16_t)tx.a1 << 16);
    off[k] = tx.s - s0;  // 0 <= off <= 6 for scale factors the tiled path accepts (pyramid_tile_fits: <= 1.5)
  }
  const int rel = s0 - ax;                       // first needed byte inside a staged row
  const uint8_t* wbase = s_src + (rel & ~3);     // its aligned word
  const uint32_t shift = 8u * (uint32_t)(rel & 3);
  // PRMT selectors on the aligned 8-byte window (v0 = bytes 0..3, v1 = bytes 4..7): [l_k, r_k, l_k+1, r_k+1]
  const uint32_t selAB = (uint32_t)off[0] | ((uint32_t)(off[0] + 1) << 4) | ((uint32_t)off[1] << 8) | ((uint32_t)(off[1] + 1) << 12);
  const uint32_t selCD = (uint32_t)off[2] | ((uint32_t)(off[2] + 1) << 4) | ((uint32_t)off[3] << 8) | ((uint32_t)(off[3] + 1) << 12);
  __syncthreads();
  while (!mbar_try_wait(&s_bar, 0)) {
  }
  // horizontal interpolation (>> 4, cv::resize's intermediate) of this thread's four columns on the staged row at byte offset ro
  auto hrow = [&](uint32_t ro, uint32_t h[4]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(wbase + ro);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    const uint32_t v0 = __funnelshift_r(w0, w1, shift), v1 = __funnelshift_r(w1, w2, shift);
    const uint32_t ab = __byte_perm(v0, v1, selAB), cd = __byte_perm(v0, v1, selCD);
    h[0] = __dp2a_lo(coef[0], ab, 0u) >> 4;
    h[1] = __dp2a_hi(coef[1], ab, 0u) >> 4;
    h[2] = __dp2a_lo(coef[2], cd, 0u) >> 4;
    h[3] = __dp2a_hi(coef[3], cd, 0u) >> 4;
  };
  uint8_t* dst = pyr + (size_t)blockIdx.z * g.frameStride + D.off + (size_t)(y0 + rg * kPyRpt) * D.pitch + gx;
  size_t dpitch = (size_t)D.pitch;
  asm volatile("" : "+l"(dst), "+l"(dpitch));  // (keeps both in registers: the compiler otherwise rebuilds the address per row)
  uint32_t ha[4], hb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int j = 0; j < kPyRpt; j++) {
    const int ry = rg * kPyRpt + j;
    const uint4 rw = s_row[ry];  // warp-uniform: one broadcast load
    if (rw.y & 0x80000000u) {    // warp-uniform branch
#pragma unroll
      for (int k = 0; k < 4; k++) ha[k] = hb[k];
    } else {
      hrow(rw.x, ha);
    }
    hrow(rw.y & 0x7fffffffu, hb);
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      // (tried: masking the low 4 bits instead of shifting and taking ((h >> 4) * b) >> 16 as mul.hi(h & ~15, b << 12), which
      // moves 8 shifts per row from the ALU pipe to the FMA pipe: IMAD.HI issues at a quarter rate, 4.52 -> 4.90 us/frame)
      // the rounding constant rides on the second product (2 << 16 does not touch its low 16 bits): one add less per pixel
      uint32_t t0, t1;
      asm("mul.lo.u32 %0, %1, %2;" : "=r"(t0) : "r"(ha[k]), "r"(rw.z));
      asm("mad.lo.u32 %0, %1, %2, 0x20000;" : "=r"(t1) : "r"(hb[k]), "r"(rw.w));
      v[k] = ((t0 >> 16) + (t1 >> 16)) >> 2;  // in [0, 255]: a convex combination of bytes
    }
    uint32_t out;  // v0 | v1 << 8 | v2 << 16 | v3 << 24 by Horner on the FMA pipe (the ALU pipe carries the shifts)
    asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(out) : "r"(v[3]), "r"(v[2]));
    asm("mad.lo.u32 %0, %0, 256, %1;" : "+r"(out) : "r"(v[1]));
    asm("mad.lo.u32 %0, %0, 256, %1;" : "+r"(out) : "r"(v[0]));
    if (y0 + ry < D.h) *reinterpret_cast<uint32_t*>(dst) = out;
    dst += dpitch;  // a running pointer: two adds per row instead of a 64-bit multiply-add chain
  }
}

static bool g_pyrOld = getenv("PGB_PYR_OLD") != nullptr;  // A/B: the round-1 kernel

void launch_pyramid_level(const OrbGeo& g, const TmapIn& tm, int level, int frame0, int nFrames, uint8_t* pyr,
                          const ResizeTab* xtab, const ResizeTab* ytab, const int2* tileX, const int2* tileY,
                          cudaStream_t st) {
  const LevelGeo& D = g.lv[level];
  const LevelGeo& S = g.lv[level - 1];
  // does the source footprint of a 256x32 tile fit the static staging buffers?  (true for scale factors <= ~1.5)
  const bool fits = pyramid_tile_fits(S.w, S.h, D.w, D.h);
  if (fits) {
    dim3 grid((D.w + kPyW - 1) / kPyW, (D.h + kPyH - 1) / kPyH, nFrames);
    if (g_pyrOld) k_pyramid_tiled<<<grid, 256, 0, st>>>(g, level, pyr, xtab, ytab, tileX, tileY);
    else k_pyramid_walk<<<grid, kPyThreads, 0, st>>>(g, tm, level, frame0, pyr, xtab, ytab, tileX, tileY);
  } else {
    dim3 grid((((D.w + 3) >> 2) + 255) / 256, D.h, nFrames);
    k_pyramid<<<grid, 256, 0, st>>>(g, level, pyr, xtab, ytab);
  }
  PGB_LAUNCHED();
}

// K2 (FAST-9 score) lives in fast_score.cu.

// =========================================================================================== K3 cells
// One warp per FAST cell (ORBextractor.cc:789-828: cv::FAST at iniThFAST, again at minThFAST if the cell is empty).
// A pixel survives the cell's 3x3 NMS iff its score is strictly greater than its 8 neighbours' scores, neighbours
// outside the cell's tested rectangle counting 0.  Because non-corners at the cell's threshold also count 0 and a
// survivor's own score is >= the threshold, the survivors at iniTh are exactly the survivors at minTh with
// score >= iniTh: one NMS pass serves both thresholds.
//
// Lane = row of the cell's tested rectangle (<= 64 rows: two passes of 32).  A lane reads its row with 16-byte loads,
// turns the non-zero bytes inside [rx0, rx1) into a 64-bit candidate mask (every non-zero score is >= minTh by
// construction of the score map), runs the neighbour test on its own candidates (byte loads that hit L1: the rows
// were just read by the neighbouring lanes) and keeps two 64-bit survivor masks (all / score >= iniTh).  Lane order
// = row order and bit order = x order, so one warp scan of the per-lane counts places the survivors in the
// reference's row-major order.  (Round-1 profile of the previous word-scan version: 1120 warp-instructions per cell,
// a third of them in the candidate scan; this one needs about a third of that.)
#ifndef PGB_CELL_WARPS
#define PGB_CELL_WARPS 4
#endif
constexpr int kCellWarps = PGB_CELL_WARPS;
constexpr int kCellListCap = 256;  // window over a cell's candidate sequence (natural images: one window)

__global__ void __launch_bounds__(kCellWarps * 32) k_cells(OrbGeo g, const int* __restrict__ cellTab,
                                                           const uint8_t* __restrict__ score,
                                                           uint32_t* __restrict__ slots, int* __restrict__ cellCnt,
                                                           int* __restrict__ err) {
  __shared__ uint16_t s_list[kCellWarps][kCellListCap];
  __shared__ __align__(16) uint32_t s_keep[kCellWarps][64 * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x * kCellWarps + warp;
  if (cell >= g.totalCells) return;
  const int f = blockIdx.y;
  const int tv = __ldg(&cellTab[cell]);  // level | row i << 8 | column j << 20 (host table: no search, no division)
  const int level = tv & 0xff, i = (tv >> 8) & 0xfff, j = (tv >> 20) & 0xfff;
  const LevelGeo& L = g.lv[level];
  const int ci = cell - L.cellBase;
  int* cnt = cellCnt + (size_t)f * g.totalCells + cell;
  const int iniX = kMinBorder + j * L.wCell, iniY = kMinBorder + i * L.hCell;
  if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6) {
    if (lane == 0) *cnt = 0;
    return;
  }
  const int maxX = min(iniX + L.wCell + 6, L.maxBX), maxY = min(iniY + L.hCell + 6, L.maxBY);
  const int rx0 = iniX + 3, rx1 = maxX - 3, ry0 = iniY + 3, ry1 = maxY - 3;
  const int tw = rx1 - rx0, th = ry1 - ry0;
  if (tw <= 0 || th <= 0) {
    if (lane == 0) *cnt = 0;
    return;
  }
  if (tw > 64 || th > 64) {  // cells are W = 30 nominal, so < 60 in each direction (ORBextractor.cc:769-778)
    if (lane == 0) { atomicOr(err, kErrCellChunks); *cnt = 0; }
    return;
  }
  const uint8_t* S = score + (size_t)f * g.frameStride + L.off;
  const int pitch = L.pitch;
  const int a16 = rx0 & ~15;
  const int nvec = (rx1 - a16 + 15) >> 4;  // <= 5
  const int iniTh = g.iniTh;

  // ---- 1. per-lane candidate masks (bit = x - rx0), two passes of 32 rows
  unsigned long long cand[2] = {0ull, 0ull};
#pragma unroll
  for (int c = 0; c < 2; c++) {
    if (32 * c >= th) break;  // warp-uniform
    const int row = lane + 32 * c;
    if (row < th) {
      const uint8_t* rowp = S + (size_t)(ry0 + row) * pitch;
      unsigned long long cm = 0ull;
      for (int v = 0; v < nvec; v++) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(rowp + a16 + 16 * v));
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
        uint32_t m16 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t wd = w4[k];
          // non-zero bytes -> 4 adjacent bits: msb flags, then a multiply gathers them at bits 21..24
          const uint32_t nz = (((wd & 0x7f7f7f7fu) + 0x7f7f7f7fu) | wd) & 0x80808080u;
          m16 |= ((((nz >> 7) * 0x00204081u) >> 21) & 0xfu) << (4 * k);
        }
        const int off = a16 + 16 * v - rx0;  // x of the vector's byte 0 relative to the rectangle: -15 .. 63
        cm |= off >= 0 ? ((unsigned long long)m16 << off) : ((unsigned long long)(m16 >> (-off)));
      }
      if (tw < 64) cm &= (1ull << tw) - 1ull;
      cand[c] = cm;
    }
  }
  // ---- 2. compact the candidates into a list so that the neighbour test keeps all 32 lanes busy (the candidates of a
  //         cell sit in a few rows: per-lane loops ran at 6 of 32 lanes in the round-1 profile).  The list is a
  //         window of kCellListCap entries over the row-major candidate sequence (one window on natural images);
  //         survivors are reported back to per-row masks in shared memory.  Shared memory is kept small on purpose:
  //         the neighbour loads live on L1 hits, and L1 is what the shared-memory carve-out leaves.
  uint16_t* list = s_list[warp];              // row | bx << 6
  uint32_t* rowKeep = s_keep[warp];           // [64 rows][4]: keep lo, keep hi, keep20 lo, keep20 hi
#pragma unroll
  for (int c = 0; c < 2; c++) {
    uint4* rk = reinterpret_cast<uint4*>(rowKeep + 4 * (lane + 32 * c));
    *rk = make_uint4(0u, 0u, 0u, 0u);
  }
  const int cnt0 = __popcll(cand[0]), cnt1 = __popcll(cand[1]);
  int incl = cnt0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int tot0 = __shfl_sync(0xffffffffu, incl, 31);
  int pos0 = incl - cnt0, pos1 = 0, total = tot0;
  if (th > 32) {
    int in1 = cnt1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, in1, d);
      if (lane >= d) in1 += v;
    }
    pos1 = tot0 + in1 - cnt1;
    total = tot0 + __shfl_sync(0xffffffffu, in1, 31);
  }
  __syncwarp();
  for (int base = 0; base < total; base += kCellListCap) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      if (32 * c >= th) break;
      int pos = (c ? pos1 : pos0) - base;
      const uint32_t rowcode = (uint32_t)(lane + 32 * c);
      for (int hf = 0; hf < 2; hf++) {
        uint32_t m = hf ? (uint32_t)(cand[c] >> 32) : (uint32_t)cand[c];
        while (m) {
          const uint32_t low = m & (0u - m);
          m ^= low;
          if (pos >= 0 && pos < kCellListCap) list[pos] = (uint16_t)(rowcode | ((uint32_t)(31 - __clz(low) + 32 * hf) << 6));
          pos++;
        }
      }
    }
    __syncwarp();
    // ---- 3. neighbour test, one candidate per lane.  Neighbours outside the rectangle count 0; the loads themselves
    //         are always inside the level (rectangles start >= 19 px from every border), so they are issued
    //         unconditionally and masked afterwards.
    const int nwin = min(kCellListCap, total - base);
    for (int i = lane; i < nwin; i += 32) {
      const uint32_t e = list[i];
      const int row = e & 63, bx = (e >> 6) & 63;
      const uint8_t* p = S + (size_t)(ry0 + row) * pitch + rx0 + bx;
      const int sc = __ldg(p);
      int nw = __ldg(p - pitch - 1), n = __ldg(p - pitch), ne = __ldg(p - pitch + 1);
      const int w = __ldg(p - 1), ea = __ldg(p + 1);
      int sw = __ldg(p + pitch - 1), so = __ldg(p + pitch), se = __ldg(p + pitch + 1);
      if (row == 0) { nw = 0; n = 0; ne = 0; }
      if (row == th - 1) { sw = 0; so = 0; se = 0; }
      int left = max(max(nw, w), sw), right = max(max(ne, ea), se);
      if (bx == 0) left = 0;
      if (bx == tw - 1) right = 0;
      if (max(max(left, right), max(n, so)) < sc) {
        uint32_t* rk = rowKeep + 4 * row + (bx >> 5);
        atomicOr(rk, 1u << (bx & 31));
        if (sc >= iniTh) atomicOr(rk + 2, 1u << (bx & 31));
      }
    }
    __syncwarp();
  }
  // ---- 4. emit the survivors (at iniTh if there is any, else at minTh): lane order = row order, bit order = x order
  uint4 rk[2];
#pragma unroll
  for (int c = 0; c < 2; c++) rk[c] = *reinterpret_cast<const uint4*>(rowKeep + 4 * (lane + 32 * c));
  const bool any20 = __any_sync(0xffffffffu, (rk[0].z | rk[0].w | rk[1].z | rk[1].w) != 0u);
  uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase + (size_t)ci * L.slotCap;
  const int nHalf = tw > 32 ? 2 : 1;
  int basei = 0;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    if (32 * c >= th) break;
    const uint32_t s0 = any20 ? rk[c].z : rk[c].x, s1 = any20 ? rk[c].w : rk[c].y;
    const int n = __popc(s0) + __popc(s1);
    int in2 = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, in2, d);
      if (lane >= d) in2 += v;
    }
    int pos = basei + in2 - n;
    basei += __shfl_sync(0xffffffffu, in2, 31);
    const int y = ry0 + lane + 32 * c;
    for (int hf = 0; hf < nHalf; hf++) {
      uint32_t m = hf ? s1 : s0;
      while (m) {
        const uint32_t low = m & (0u - m);
        m ^= low;
        const int x = rx0 + 31 - __clz(low) + 32 * hf;
        if (pos < L.slotCap)
          slot[pos] = (uint32_t)(x - kMinBorder) | ((uint32_t)(y - kMinBorder) << 12) | ((uint32_t)__ldg(S + (size_t)y * pitch + x) << 24);
        pos++;
      }
    }
  }
  if (lane == 0) {
    if (basei > L.slotCap) { atomicOr(err, kErrCandOverflow); basei = L.slotCap; }
    *cnt = basei;
  }
}

void launch_cells(const OrbGeo& g, int nFrames, const int* cellTab, const uint8_t* score, uint32_t* slots, int* cellCnt,
                  int* err, cudaStream_t st) {
  dim3 grid((g.totalCells + kCellWarps - 1) / kCellWarps, nFrames);
  k_cells<<<grid, kCellWarps * 32, 0, st>>>(g, cellTab, score, slots, cellCnt, err);
  PGB_LAUNCHED();
}

// =========================================================================================== K4 octree
// One CTA per (level, frame).  The std::list of the reference is an array in list order (index 0 = front);
// every pass rebuilds it: children of the nodes split in the pass, most recently created first (push_front),
// then the untouched nodes in their old order.  Points carry their node id; a pass is
//   (a) choose the nodes to split (all expandable ones front-to-back, or -- in the "careful" phase -- the
//       previous pass's children sorted by (size, creation seq) descending),
//   (b) count each chosen node's points per quadrant in parallel,
//   (c) one thread replays the sequential bookkeeping (list size, stop the instant size >= N),
//   (d) points of the nodes actually split move to their child.
constexpr int kOctThreads = 256;

struct OctNode {
  short x0, y0, x1, y1;
  int cnt;
  int seq;
};

__device__ __forceinline__ int oct_mid(int a, int b) { return a + ((b - a + 1) >> 1); }  // a + ceil((b-a)/2)

__global__ void __launch_bounds__(kOctThreads) k_octree(OrbGeo g, const uint32_t* __restrict__ slots,
                                                        const int* __restrict__ cellCnt,
                                                        unsigned long long* __restrict__ candAll,
                                                        StagedKp* __restrict__ staged, int* __restrict__ lvlCnt,
                                                        int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const LevelGeo& L = g.lv[level];
  const int cap = L.nodeCap;
  // shared arrays (all sized by nodeCap)
  OctNode* nodeA = reinterpret_cast<OctNode*>(smem_raw);
  OctNode* nodeB = nodeA + cap;
  int* E = reinterpret_cast<int*>(nodeB + cap);  // processing order: node indices
  int* eidx = E + cap;                           // node -> position in E or -1
  int* cc = eidx + cap;                          // [cap][4] child counts
  int* childNew = cc + 4 * cap;                  // [cap][4] new index of child
  int* remap = childNew + 4 * cap;               // node -> new index (survivors)
  int* vlist = remap + cap;                      // children with >1 points created last pass, creation order
  unsigned long long* best = reinterpret_cast<unsigned long long*>(vlist + cap);  // [cap], 8-aligned by layout
  __shared__ int s_scan[kOctThreads];
  __shared__ int s_n, s_alive, s_nE, s_nV, s_P, s_seq, s_mode, s_finish, s_prevSize, s_created;
  __shared__ int s_wsum[kOctThreads / 32];

  const int lane = tid & 31, wid = tid >> 5;
  auto block_scan = [&](int v, int& carry) {  // exclusive prefix of v over the block (+ carry); carry += block total
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kOctThreads / 32; w++) {
      const int x = s_wsum[w];
      if (w < wid) wbase += x;
      total += x;
    }
    __syncthreads();
    const int ex = carry + wbase + incl - v;
    carry += total;
    return ex;
  };

  const int nCells = L.nCols * L.nRows;
  const int* cnt = cellCnt + (size_t)f * g.totalCells + L.cellBase;
  const uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase;
  unsigned long long* cand = candAll + (size_t)f * g.candPerFrame + L.candBase;

  // ---- gather the per-cell candidate lists into one array, cell row-major then in-cell order
  const int chunk = (nCells + kOctThreads - 1) / kOctThreads;
  const int c0 = min(tid * chunk, nCells), c1 = min(c0 + chunk, nCells);
  int mySum = 0;
  for (int c = c0; c < c1; c++) mySum += cnt[c];
  {
    int total = 0;
    s_scan[tid] = block_scan(mySum, total);
    if (tid == 0) {
      s_n = total;
      if (total > L.candCap) { atomicOr(err, kErrCandOverflow); s_n = 0; }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n > 0) {
    int o = s_scan[tid];
    for (int c = c0; c < c1; c++) {
      const int k = cnt[c];
      for (int q = 0; q < k; q++) cand[o + q] = (unsigned long long)slot[(size_t)c * L.slotCap + q];
      o += k;
    }
  }
  int* myCnt = lvlCnt + (size_t)f * g.nlevels + level;
  if (n == 0) {
    if (tid == 0) *myCnt = 0;
    return;
  }
  __syncthreads();

  // ---- roots
  const int N = L.quota;
  const float hX = L.hX;
  const int H = L.maxBY - kMinBorder;
  for (int i = tid; i < L.nIni; i += kOctThreads) {
    nodeB[i].x0 = (short)(int)(hX * (float)i);
    nodeB[i].x1 = (short)(int)(hX * (float)(i + 1));
    nodeB[i].y0 = 0; nodeB[i].y1 = (short)H;
    nodeB[i].cnt = 0; nodeB[i].seq = i;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kOctThreads) {
    const uint32_t p = (uint32_t)cand[i];
    int r = (int)(__fdiv_rn((float)(p & 0xfff), hX));
    r = min(r, L.nIni - 1);
    cand[i] = (unsigned long long)p | ((unsigned long long)r << 32);
    atomicAdd(&nodeB[r].cnt, 1);
  }
  __syncthreads();
  if (tid == 0) {  // erase empty roots
    int a = 0;
    for (int i = 0; i < L.nIni; i++)
      if (nodeB[i].cnt > 0) { nodeA[a] = nodeB[i]; remap[i] = a; a++; } else remap[i] = -1;
    s_alive = a; s_seq = L.nIni; s_mode = 0; s_finish = 0; s_nV = 0;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kOctThreads) {
    const unsigned long long v = cand[i];
    cand[i] = (v & 0xffffffffull) | ((unsigned long long)remap[(int)(v >> 32)] << 32);
  }
  OctNode* cur = nodeA;
  OctNode* nxt = nodeB;
  __syncthreads();

  // ---- main loop
  while (true) {
    const int alive = s_alive, mode = s_mode;
    // (a) processing order
    for (int i = tid; i < alive; i += kOctThreads) eidx[i] = -1;
    __syncthreads();
    if (mode == 0) {  // every node with more than one point, in list order
      int ne = 0;
      for (int base = 0; base < alive; base += kOctThreads) {
        const int i = base + tid;
        const bool split = i < alive && cur[i].cnt > 1;
        const int ex = block_scan(split ? 1 : 0, ne);
        if (split) { E[ex] = i; eidx[i] = ex; }
      }
      if (tid == 0) { s_nE = ne; s_prevSize = alive; }
    } else {
      const int nV = s_nV;
      for (int t = tid; t < nV; t += kOctThreads) {
        const int me = vlist[t];
        const long long km = ((long long)cur[me].cnt << 32) | (unsigned)cur[me].seq;
        int rank = 0;
        for (int u = 0; u < nV; u++) {
          const int o = vlist[u];
          const long long ko = ((long long)cur[o].cnt << 32) | (unsigned)cur[o].seq;
          rank += (ko > km);
        }
        E[rank] = me; eidx[me] = rank;
      }
      if (tid == 0) { s_nE = nV; s_prevSize = alive; }
    }
    __syncthreads();
    const int nE = s_nE;
    for (int i = tid; i < nE * 4; i += kOctThreads) cc[i] = 0;
    __syncthreads();
    // (b) quadrant counts
    for (int i = tid; i < n; i += kOctThreads) {
      const unsigned long long v = cand[i];
      const int node = (int)((v >> 32) & 0x3fffffff);
      const int e = eidx[node];
      if (e >= 0) {
        const int x = (int)(v & 0xfff), y = (int)((v >> 12) & 0xfff);
        const OctNode nd = cur[node];
        const int mx = oct_mid(nd.x0, nd.x1), my = oct_mid(nd.y0, nd.y1);
        const int q = (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
        atomicAdd(&cc[e * 4 + q], 1);
        cand[i] = (v & 0x3fffffffffffffffull) | ((unsigned long long)q << 62);
      }
    }
    __syncthreads();
    // (c) bookkeeping of the pass, in parallel (it used to be replayed by one thread while 255 waited: 35 % of the kernel's
    // stall samples sat at the barrier behind it).  The reference's list order is reproduced by prefix sums:
    //   * expanded node p of the processing order E creates k_p children (its non-empty quadrants), in quadrant order;
    //     the pass stops after the node with which the list reaches N entries (mode 1: ORBextractor.cc:699-737);
    //   * children are pushed to the FRONT of the list one by one, so the child with creation ordinal `ord` ends up at
    //     index created - 1 - ord; the nodes that were not expanded follow in their old order;
    //   * children with more than one point, in creation order, are the next pass's candidates (vlist).
    {
      int* kEx = remap;                          // [nE] exclusive prefix of k (free until the survivors are renumbered)
      int* gEx = reinterpret_cast<int*>(best);   // [nE] exclusive prefix of the (c > 1) counts (best is unused until the end)
      if (tid == 0) s_P = nE;
      __syncthreads();
      int carry = 0;
      for (int base = 0; base < nE; base += kOctThreads) {
        const int p = base + tid;
        int k = 0, gq = 0;
        if (p < nE) {
#pragma unroll
          for (int q = 0; q < 4; q++) { const int c = cc[p * 4 + q]; k += (c > 0); gq += (c > 1); }
        }
        const int ex = block_scan(k | (gq << 16), carry);  // both sums stay below 2^15 (at most 4 per node)
        if (p < nE) {
          kEx[p] = ex & 0xffff;
          gEx[p] = ex >> 16;
          // list size after node p: alive + sum_{j <= p} (k_j - 1); mode 1 stops at the first node that reaches N
          if (mode == 1 && alive + (ex & 0xffff) + k - (p + 1) >= N) atomicMin(&s_P, p + 1);
        }
      }
      __syncthreads();
      const int P = s_P;
      const int seq0 = s_seq;
      if (P > 0 && tid == ((P - 1) & (kOctThreads - 1))) {  // the last expanded node closes the sums
        int k = 0, gq = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) { const int c = cc[(P - 1) * 4 + q]; k += (c > 0); gq += (c > 1); }
        s_created = kEx[P - 1] + k;
        s_nV = gEx[P - 1] + gq;
      }
      if (P == 0 && tid == 0) { s_created = 0; s_nV = 0; }
      __syncthreads();
      const int created = s_created, size = alive + created - P, nv = s_nV;
      if (size > cap) {
        if (tid == 0) { atomicOr(err, kErrNodeOverflow); s_finish = 2; }
      } else {
        for (int p = tid; p < P; p += kOctThreads) {
          const OctNode nd = cur[E[p]];
          const int mx = oct_mid(nd.x0, nd.x1), my = oct_mid(nd.y0, nd.y1);
          int ord = kEx[p], g = gEx[p];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int c = cc[p * 4 + q];
            if (c <= 0) { childNew[p * 4 + q] = -1; continue; }
            const int idx = created - 1 - ord;
            OctNode ch;
            ch.x0 = (q & 1) ? (short)mx : nd.x0;
            ch.x1 = (q & 1) ? nd.x1 : (short)mx;
            ch.y0 = (q & 2) ? (short)my : nd.y0;
            ch.y1 = (q & 2) ? nd.y1 : (short)my;
            ch.cnt = c;
            ch.seq = seq0 + ord;
            nxt[idx] = ch;
            childNew[p * 4 + q] = idx;
            if (c > 1) vlist[g++] = idx;
            ord++;
          }
        }
        __syncthreads();  // kEx (= remap) has been read: the survivors may renumber now
        int scarry = 0;
        for (int base = 0; base < alive; base += kOctThreads) {
          const int i = base + tid;
          bool keep = false;
          if (i < alive) { const int e = eidx[i]; keep = !(e >= 0 && e < P); }
          const int ex = block_scan(keep ? 1 : 0, scarry);
          if (i < alive) {
            if (keep) { nxt[created + ex] = cur[i]; remap[i] = created + ex; }
            else remap[i] = -1;
          }
        }
        if (tid == 0) {
          s_seq = seq0 + created;
          s_alive = size; s_P = P;
          // loop control (ORBextractor.cc:663-737)
          if (size >= N || size == s_prevSize) s_finish = 1;
          else if (mode == 0 && (size + nv * 3) > N) s_mode = 1;
        }
      }
    }
    __syncthreads();
    if (s_finish == 2) {
      if (tid == 0) *myCnt = 0;
      return;
    }
    // (d) move points
    const int P = s_P;
    for (int i = tid; i < n; i += kOctThreads) {
      const unsigned long long v = cand[i];
      const int node = (int)((v >> 32) & 0x3fffffff);
      const int e = eidx[node];
      int nn;
      if (e >= 0 && e < P) nn = childNew[e * 4 + (int)(v >> 62)];
      else nn = remap[node];
      cand[i] = (v & 0xffffffffull) | ((unsigned long long)nn << 32);
    }
    OctNode* tmp = cur; cur = nxt; nxt = tmp;
    __syncthreads();
    if (s_finish) break;
  }

  // ---- best point per node: max response, first in candidate order on ties (:743-760)
  const int alive = s_alive;
  for (int i = tid; i < alive; i += kOctThreads) best[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += kOctThreads) {
    const unsigned long long v = cand[i];
    const int node = (int)((v >> 32) & 0x3fffffff);
    const unsigned long long key = ((unsigned long long)((uint32_t)v >> 24) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    atomicMax(&best[node], key);
  }
  __syncthreads();
  StagedKp* out = staged + (size_t)f * g.kpCapInternal + L.kpBase;
  for (int i = tid; i < alive; i += kOctThreads) {
    const uint32_t pi = 0xffffffffu - (uint32_t)(best[i] & 0xffffffffull);
    const uint32_t p = (uint32_t)cand[pi];
    StagedKp k;
    k.x = (int)(p & 0xfff) + kMinBorder;
    k.y = (int)((p >> 12) & 0xfff) + kMinBorder;
    k.score = (int)(p >> 24);
    k.level = level;
    out[i] = k;
  }
  if (tid == 0) *myCnt = alive;
}

static size_t octree_smem_bytes(int cap) {
  // 2 node arrays (12 B each) + E, eidx, remap, vlist (4 ints) + cc, childNew (8 ints) + best (8 B)
  size_t b = (size_t)cap * (2 * sizeof(OctNode) + 12 * sizeof(int));
  b = (b + 7) & ~(size_t)7;
  return b + (size_t)cap * 8;
}

void launch_octree(const OrbGeo& g, int nFrames, const uint32_t* slots, const int* cellCnt, unsigned long long* cand,
                   StagedKp* staged, int* lvlCnt, int* err, cudaStream_t st) {
  const size_t smem = octree_smem_bytes(g.maxNodeCap);
  static DynSmemLimit lim;
  lim.ensure(k_octree, smem);  // a failure surfaces as the launch error run_stages() checks
  dim3 grid(g.nlevels, nFrames);
  k_octree<<<grid, kOctThreads, smem, st>>>(g, slots, cellCnt, cand, staged, lvlCnt, err);
  PGB_LAUNCHED();
}

// =========================================================================================== K5-K7 orientation + blur + rBRIEF
// One warp per keypoint.  The warp stages the 45x45 neighbourhood of the keypoint in shared memory (with the
// level's BORDER_REFLECT_101 applied), takes the intensity-centroid angle from the unblurred pixels, blurs the
// inner 39x39 patch with the fixed-point 7x7 kernel and samples the 256 rotated test pairs from it.  The blurred
// level image of the reference is never materialised: its value at a pixel depends only on the 7x7 neighbourhood.
constexpr int kOdWarps = 4;
#ifndef PGB_OD_OCC
#define PGB_OD_OCC 8
#endif
constexpr int kPR = 22;            // patch radius: 19 (pattern reach) + 3 (blur)
constexpr int kPW = 2 * kPR + 1;   // 45
constexpr int kPP = 52;            // shared-memory row pitch of the patch: 13 aligned words cover 45 px at any phase (odd word
                                   // count: column walks are bank-conflict free)
constexpr int kBR = 19;
constexpr int kBW = 2 * kBR + 1;   // 39
constexpr int kHP = 40;            // row pitch of the horizontally blurred patch (u16): even, so two outputs go in one store

__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
__device__ const signed char d_pattern[1024] = PGB200_ORB_PATTERN_INIT;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// cv::fastAtan2 (fp32 polynomial, no FMA)
__device__ __forceinline__ float fast_atan2_dev(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
  const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// (float)cos((double)th), (float)sin((double)th) of computeOrbDescriptor (ORBextractor.cc:113; the reference's `float angle`
// promotes to the double overloads) for th in [0, 2 pi]: one quadrant reduction (k <= 4, two-piece pi/2: exact first
// product) and the fdlibm kernel polynomials on [-pi/4, pi/4], ~30 DFMA instead of the 330 instructions of libdevice's
// full-range cos() + sin().  Max error 1.1e-16 = 1 ulp of double; the FLOAT results equal glibc's on every 7th float
// angle in [0, 360) (162 M values, checked on the host with the same expression: tools/probe/sincos_check.c).
__device__ __forceinline__ void sincos_rbrief(double x, float* s, float* c) {
  const double k = rint(__dmul_rn(x, 0.63661977236758134308));
  double r = __fma_rn(-k, 1.57079632673412561417e+00, x);
  r = __fma_rn(-k, 6.07710050650619224932e-11, r);
  const double z = __dmul_rn(r, r);
  const double ps = __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08),
                                                     2.75573137070700676789e-06), -1.98412698298579493134e-04), 8.33333333332248946124e-03);
  const double sr = __fma_rn(__dmul_rn(z, r), __fma_rn(z, ps, -1.66666666666666324348e-01), r);
  const double pc = __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, __fma_rn(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09),
                                                                -2.75573143513906633035e-07), 2.48015872894767294178e-05),
                                         -1.38888888888741095749e-03), 4.16666666666666019037e-02);
  const double cr = __dsub_rn(1.0, __fma_rn(0.5, z, -__dmul_rn(__dmul_rn(z, z), pc)));
  const int q = (int)k & 3;
  const double sv = (q & 1) ? cr : sr, cv = (q & 1) ? sr : cr;
  *s = (float)((q & 2) ? -sv : sv);
  *c = (float)((q == 1 || q == 2) ? -cv : cv);
}

__global__ void __launch_bounds__(kOdWarps * 32, PGB_OD_OCC) k_orient_desc(OrbGeo g, const uint8_t* __restrict__ pyr,
                                                               const StagedKp* __restrict__ staged,
                                                               const int* __restrict__ lvlCnt,
                                                               pgb_keypoint* __restrict__ kps,
                                                               uint8_t* __restrict__ desc, int* __restrict__ counts,
                                                               int cap, int* __restrict__ err) {
  __shared__ __align__(4) uint8_t s_patch[kOdWarps][kPW * kPP];
  __shared__ __align__(4) uint16_t s_h[kOdWarps][kPW * kHP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * kOdWarps + warp;
  const int f = blockIdx.y;
  if (slot >= g.kpCapInternal) return;
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < g.nlevels; l++)
    if (slot >= g.lv[l].kpBase) level = l;
  const LevelGeo& L = g.lv[level];
  const int j = slot - L.kpBase;
  const int* lc = lvlCnt + (size_t)f * g.nlevels;
  int before = 0, total = 0;
  for (int l = 0; l < g.nlevels; l++) {
    const int c = lc[l];
    if (l < level) before += c;
    total += c;
  }
  if (slot == 0 && lane == 0) {
    counts[f] = min(total, cap);
    if (total > cap) atomicOr(err, kErrOutCap);
  }
  if (j >= lc[level]) return;
  const int oi = before + j;
  if (oi >= cap) return;
  const StagedKp kp = staged[(size_t)f * g.kpCapInternal + slot];
  int iPitch;
  const uint8_t* img = level_base(g, pyr, level, f, &iPitch);
  uint8_t* Pw = s_patch[warp];
  // stage the patch.  Interior keypoints (all 45x45 px inside the level) copy aligned 32-bit words and keep the
  // row's byte phase; keypoints within 22 px of a border take the byte path with BORDER_REFLECT_101.
  const int xs = kp.x - kPR, ys = kp.y - kPR;
  int po = 0;
  if (xs >= 0 && ys >= 0 && xs + 2 * kPR < L.w && ys + 2 * kPR < L.h) {
    // 45 rows x 4 vectors of 16 bytes = 180 loads, 6 per lane, all issued before the first shared-memory store (the
    // round-1 profile had 39 % of this kernel's stall samples on a word-at-a-time version of this loop); the 13 words
    // of a row that hold the patch are then stored at the odd 13-word pitch
    const int a = xs & ~15, a4 = xs & ~3;
    po = xs - a4;
    const int woff = (a4 - a) >> 2;  // first useful word of the 16 loaded per row
    uint4 v[6];
#pragma unroll
    for (int u = 0; u < 6; u++) {
      const int i = lane + 32 * u, r = i >> 2, k = i & 3;
      v[u] = make_uint4(0, 0, 0, 0);
      if (r < kPW && a + 16 * k < iPitch) v[u] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(ys + r) * iPitch + a) + k);
    }
#pragma unroll
    for (int u = 0; u < 6; u++) {
      const int i = lane + 32 * u, r = i >> 2, k = i & 3;
      if (r < kPW) {
        uint32_t* row = reinterpret_cast<uint32_t*>(Pw + r * kPP);
        const int w0 = 4 * k - woff;  // shared-memory word index of v[u].x
        if (w0 >= 0 && w0 < 13) row[w0] = v[u].x;
        if (w0 + 1 >= 0 && w0 + 1 < 13) row[w0 + 1] = v[u].y;
        if (w0 + 2 >= 0 && w0 + 2 < 13) row[w0 + 2] = v[u].z;
        if (w0 + 3 >= 0 && w0 + 3 < 13) row[w0 + 3] = v[u].w;
      }
    }
  } else {
    for (int i = lane; i < kPW * kPW; i += 32) {
      const int r = i / kPW, c = i - r * kPW;
      const int y = reflect101(ys + r, L.h), x = reflect101(xs + c, L.w);
      Pw[r * kPP + c] = __ldg(img + (size_t)y * iPitch + x);
    }
  }
  const uint8_t* P = Pw + po;  // P[r * kPP + c] = level pixel (ys + r, xs + c)
  __syncwarp();
  // intensity centroid over the radius-15 disc (integer sums: order-free)
  int m10 = 0, m01 = 0;
  for (int r = lane; r < 2 * kHalfPatch + 1; r += 32) {
    const int v = r - kHalfPatch;
    const int d = c_umax[v < 0 ? -v : v];
    const uint8_t* row = P + (kPR + v) * kPP + kPR;
    int rs = 0;
    for (int u = -d; u <= d; u++) { const int px = row[u]; m10 += u * px; rs += px; }
    m01 += v * rs;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
  const float angle = fast_atan2_dev((float)m01, (float)m10);
  // horizontal then vertical pass of the 7x7 sigma=2 fixed-point Gaussian {18,34,48,56,48,34,18}, with register
  // sliding windows instead of 7 shared-memory loads per output (the round-1 profile had this kernel at 91 % of the
  // shared-memory pipe).  Horizontal: lane = patch row; the row's 13 words are realigned to the patch phase (po is
  // warp-uniform), adjacent pixels are packed as two 16-bit halves, and one multiply-add chain yields two outputs
  // (a half never exceeds 256 * 255).  Vertical: lane = (column, third of the rows).
  uint16_t* Hh = s_h[warp];
  for (int r = lane; r < kPW; r += 32) {
    const uint32_t* wrow = reinterpret_cast<const uint32_t*>(Pw + r * kPP);
    uint32_t w[13], al[12];
#pragma unroll
    for (int k = 0; k < 13; k++) w[k] = wrow[k];
#pragma unroll
    for (int k = 0; k < 12; k++) al[k] = __funnelshift_r(w[k], w[k + 1], 8 * po);  // bytes 4k .. 4k+3 of the patch row
    uint32_t pr[46];  // pr[j] = pixel j | pixel j+1 << 16
#pragma unroll
    for (int j = 0; j < 46; j++) {
      const int k = j >> 2;
      switch (j & 3) {
        case 0: pr[j] = __byte_perm(al[k], 0u, 0x4140); break;
        case 1: pr[j] = __byte_perm(al[k], 0u, 0x4241); break;
        case 2: pr[j] = __byte_perm(al[k], 0u, 0x4342); break;
        default: pr[j] = __byte_perm(al[k], al[k + 1], 0x0403) & 0x00ff00ffu; break;  // j <= 43 here: k + 1 <= 11
      }
    }
    uint32_t* hrow = reinterpret_cast<uint32_t*>(Hh + r * kHP);
#pragma unroll
    for (int c = 0; c < kBW + 1; c += 2)  // columns c and c+1 (column 39 is padding)
      hrow[c >> 1] = 18u * (pr[c] + pr[c + 6]) + 34u * (pr[c + 1] + pr[c + 5]) + 48u * (pr[c + 2] + pr[c + 4]) + 56u * pr[c + 3];
  }
  __syncwarp();
  uint8_t* Bl = Pw;  // the blurred 39 x 39 patch takes the place of the staged pixels (dead after the horizontal pass): 9 CTAs per SM
#pragma unroll 1
  for (int id = lane; id < 3 * kBW; id += 32) {
    const int t = id / kBW, c = id - t * kBW;
    const uint16_t* hc = Hh + (13 * t) * kHP + c;
    int h[19];
#pragma unroll
    for (int k = 0; k < 19; k++) h[k] = hc[k * kHP];
    uint8_t* out = Bl + (13 * t) * kBW + c;
#pragma unroll
    for (int k = 0; k < 13; k++) {
      const int acc = 18 * (h[k] + h[k + 6]) + 34 * (h[k + 1] + h[k + 5]) + 48 * (h[k + 2] + h[k + 4]) + 56 * h[k + 3];
      out[k * kBW] = (uint8_t)((acc + 32768) >> 16);
    }
  }
  __syncwarp();
  // rBRIEF: lane = descriptor byte
  const float factorPI = (float)(3.14159265358979323846 / 180.0);
  const float th = __fmul_rn(angle, factorPI);
  float a, b;
  sincos_rbrief((double)th, &b, &a);
  const uint8_t* C = Bl + kBR * kBW + kBR;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const signed char* pt = d_pattern + (lane * 16 + 2 * k) * 2;
    const float x0 = (float)pt[0], y0 = (float)pt[1], x1 = (float)pt[2], y1 = (float)pt[3];
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = C[r0 * kBW + c0], t1 = C[r1 * kBW + c1];
    val |= (t0 < t1) << k;
  }
  desc[((size_t)f * cap + oi) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    pgb_keypoint o;
    const float s = L.scale;
    o.x = (level == 0) ? (float)kp.x : __fmul_rn((float)kp.x, s);
    o.y = (level == 0) ? (float)kp.y : __fmul_rn((float)kp.y, s);
    o.size = (float)L.patchSize;
    o.angle = angle;
    o.response = (float)kp.score;
    o.octave = level;
    o.class_id = -1;
    kps[(size_t)f * cap + oi] = o;
  }
}

void launch_orient_desc(const OrbGeo& g, int nFrames, const uint8_t* pyr, const StagedKp* staged, const int* lvlCnt,
                        pgb_keypoint* kps, uint8_t* desc, int* counts, int cap, int* err, cudaStream_t st) {
  dim3 grid((g.kpCapInternal + kOdWarps - 1) / kOdWarps, nFrames);
  k_orient_desc<<<grid, kOdWarps * 32, 0, st>>>(g, pyr, staged, lvlCnt, kps, desc, counts, cap, err);
  PGB_LAUNCHED();
}

// =========================================================================================== full-level blur
// GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) of a whole level (ORBextractor.cc:1084-1085), for callers
// that want the reference's blurred working image; the descriptor path above does not need it.
__global__ void __launch_bounds__(256) k_blur_level(OrbGeo g, int level, int frame, const uint8_t* __restrict__ pyr,
                                                    uint8_t* __restrict__ out) {
  const LevelGeo& L = g.lv[level];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= L.w) return;
  int iPitch;
  const uint8_t* img = level_base(g, pyr, level, frame, &iPitch);
  const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  int acc = 0;
#pragma unroll
  for (int dy = -3; dy <= 3; dy++) {
    const uint8_t* row = img + (size_t)reflect101(y + dy, L.h) * iPitch;
    int h = 0;
#pragma unroll
    for (int dx = -3; dx <= 3; dx++) h += k[dx + 3] * (int)__ldg(row + reflect101(x + dx, L.w));
    acc += k[dy + 3] * h;
  }
  out[(size_t)y * L.w + x] = (uint8_t)((acc + 32768) >> 16);
}

void launch_blur_level(const OrbGeo& g, int level, int frame, const uint8_t* pyr, uint8_t* out, cudaStream_t st) {
  const LevelGeo& L = g.lv[level];
  dim3 grid((L.w + 255) / 256, L.h);
  k_blur_level<<<grid, 256, 0, st>>>(g, level, frame, pyr, out);
  PGB_LAUNCHED();
}

}  // namespace pgb
