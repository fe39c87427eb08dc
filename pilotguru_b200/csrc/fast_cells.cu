// K2+K3 fused: FAST-9 score, per-cell 3x3 non-maximum suppression and the per-cell iniThFAST / minThFAST decision in
// ONE kernel for sm_100a -- the score map never reaches HBM.
// Reference: ORBextractor::ComputeKeyPointsOctTree, thirdparty/orb-slam2/src/ORBextractor.cc:765-829 (30-px cell grid,
// cv::FAST(cell view, iniThFAST, nms=true), again with minThFAST when the cell came back empty) over cv::FAST(TYPE_9_16).
//
// Two kernels live here:
//   * k_fast_cells2<bands> (second half of the file): the hot path.  It follows the reference's own order -- an iniThFAST
//     pass over the whole tile, a minThFAST pass only for the cells that came back empty -- with both thresholds' flags from
//     one prefilter, a tile-wide pooled candidate queue of 16-bit entries, scoring rounds dealt round-robin to the warps,
//     9 CTAs per SM.  Levels whose cells are at most 32 px wide and 8 * bands px high (every level of a 1080p pyramid).
//   * k_fast_cells<false, 1> (first half): the single-pass kernel of the first half of round 2 -- everything scored at
//     minThFAST, the thresholds sorted out at emission -- kept as the generic instantiation for levels with larger cells
//     (other resolutions / scale factors), at one CTA per SM.
// Both produce the slots / cellCnt layout of the round-1 pair (k_fast_score -> k_cells), so the octree kernel is unchanged,
// and both are held to that pair bit for bit (tests/test_gpu_orb.py::test_fused_fast_cells_equals_the_unfused_pair).
//
// Why the fused shape.  The round-1 pair wrote a 6.4 MB/frame score map (96.5 % zeros) and read it back: 12.8 MB/frame of
// HBM traffic and 4.9 us/frame in k_cells, most of it spent finding the non-zero bytes again
// (profiles/r01e_other_kernels_sass_regions.md).  What makes the fusion clean is the reference's own cell semantics:
// cv::FAST runs on a cell VIEW, so a neighbour outside the cell's tested rectangle counts as score 0 in the 3x3 NMS,
// and the tested rectangles of the cells tile the level exactly (pitch wCell x hCell from (19, 19)).
// A tile made of WHOLE cells therefore needs no score halo at all:
//   * one CTA = one row of up to 8 FAST cells: <= 249 x hCell tested pixels.  The input box (72 words x (8 * bands + 6)
//     rows, 3-px ring halo) comes in with one 3-D TMA load; the box must start on a 16-byte boundary, the tile does not, so
//     the 256-px "lane frame" (8 px per lane) starts at the 8-byte boundary at or below the tile's first pixel and
//     per-lane validity masks cut the frame down to the tested rectangle;
//   * prefilter / candidate expansion / exact score per 8-row band descend from the round-1 score kernel (fast_score.cu);
//     the score goes into a shared-memory tile with a zero guard ring;
//   * barrier; NMS over the corner lists, one corner per lane: 8 neighbour bytes from the shared tile, neighbours across a
//     cell boundary masked to 0; survivors go to one list per tile;
//   * barrier; emission: a survivor's slot in its cell is its rank in the reference's row-major order; the packed
//     candidates go to the cell's slots.
// Algorithmic bytes per 1080p frame: 6,419,321 B read + 4 B x (cells + candidates) written (SURVEY.md 8d's fused figure).
#include <cuda_runtime.h>

#include <type_traits>

#include "common.cuh"
#include "fast_common.cuh"
#include "orb_kernels.cuh"

namespace pgb {

namespace {

using namespace fastk;

constexpr int kRowB = kFcInWords * 4;  // 288 bytes per staged input row
constexpr int kSP = kFcTilePitch;      // score tile: 16 pad bytes (x = -16 .. -1) + 256 px per row
constexpr int kQCap = kFcQueueCap;

__device__ __forceinline__ int fast_bam_minmax(const uint8_t* c) { return fastk::fast_bam_minmax<kRowB>(c); }

// One corner against its 8 neighbours in the shared score tile; p points at the corner's byte.  Rows above / below the
// tile are zero guard rows; `first` / `last`: the corner sits in the first / last column of its cell.
template <int kPitch = kSP>
__device__ __forceinline__ bool nms_keep(const uint8_t* p, int s, bool first, bool last) {
  const int nw = p[-kPitch - 1], n = p[-kPitch], ne = p[-kPitch + 1];
  const int w = p[-1], e = p[1];
  const int sw = p[kPitch - 1], so = p[kPitch], se = p[kPitch + 1];
  int left = max(max(nw, w), sw), right = max(max(ne, e), se);
  if (first) left = 0;
  if (last) right = 0;
  return max(max(left, right), max(n, so)) < s;
}

}  // namespace

// kA = true: class A levels (wCell <= 32, hCell <= 32): 4 bands per tile, one per warp, 32-bit cell windows.
// kA = false: any cell size (bands looped over the warps, 64-bit cell windows); used by the small levels whose cells
// are larger (1080p: level 7 only) at a lower occupancy.
template <bool kA, int kOcc>
__global__ void __launch_bounds__(kFcThreads, kOcc) k_fast_cells(const __grid_constant__ OrbGeo g,
                                                                 const __grid_constant__ TmapIn tm,
                                                                 const int4* __restrict__ tileTab, int nbTile, int frame0,
                                                                 uint32_t* __restrict__ slots, int* __restrict__ cellCnt,
                                                                 int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nb = kA ? 4 : nbTile;  // bands the shared-memory layout holds
  const FcSmem lay = fc_smem_layout(nb);
  const uint8_t* sInB = smem;
  uint8_t* sTile = smem + lay.tile + kSP + 16;  // (row 0, x 0) of the score tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.misc);
  int* sCorner = reinterpret_cast<int*>(smem + lay.misc + 16);  // per band: corners in the list, or -1 = use the slow NMS

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int4 te = __ldg(&tileTab[blockIdx.x]);
  const int level = te.x, X0 = te.y, Y0 = te.z, ci0 = te.w & 0xffff, cj0 = te.w >> 16;
  const LevelGeo& L = g.lv[level];
  const int wCell = L.wCell, hCell = L.hCell;
  const int kc = min(L.fcKc, L.nCols - cj0);
  const int X1 = min(X0 + kc * wCell, L.w - kEdge), Y1 = min(Y0 + hCell, L.h - kEdge);
  int* cnt = cellCnt + (size_t)f * g.totalCells + L.cellBase + ci0 * L.nCols + cj0;
  if (X1 <= X0 || Y1 <= Y0) {  // cells the reference skips (ORBextractor.cc:796-805) or whose view is too small for FAST
    if (tid < kc) cnt[tid] = 0;
    return;
  }
  const int xa = (X0 - 3) & ~15;           // first byte of the TMA box (level x), 16-byte aligned, <= X0 - 3
  const int o = (X0 - xa) & ~7;            // lane frame: staged bytes [o, o + 256) of every row, 8-byte aligned
  const int xoff = X0 - xa - o;            // tile's first tested pixel inside the lane frame: 0 .. 7
  const int tw = X1 - X0, th = Y1 - Y0;    // tested pixels of the tile

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)(kRowB * (nb * 8 + 6)));
    tma_load_3d(smem, &tm.in[level], xa >> 2, Y0 - 3, f + frame0, bar);  // the maps index frames from the batch's base
  }
  // every warp zeroes the score rows of its bands (+ the guard row above the first / below the last band)
  for (int b = warp; b < nb; b += kFcWarps) {
    uint4* z = reinterpret_cast<uint4*>(smem + lay.tile + (8 * b + 1) * kSP);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int i = lane; i < 8 * kSP / 16; i += 32) z[i] = zero;
    if (b == 0 && lane < kSP / 16) reinterpret_cast<uint4*>(smem + lay.tile)[lane] = zero;
    if (b == nb - 1 && lane < kSP / 16) reinterpret_cast<uint4*>(smem + lay.tile + (8 * nb + 1) * kSP)[lane] = zero;
  }
  if (tid < kc) cnt[tid] = 0;  // cells without survivors; the others are overwritten after the NMS (ordered by the barriers)
  if (tid == 0) sCorner[8] = 0;  // survivor count of the tile
  __syncthreads();  // mbarrier initialised before anyone polls it
  while (!mbar_try_wait(bar, 0)) {
  }

  // ================================================================ score: per 8-row band (fast_score.cu's phases)
  for (int b = warp; b < nb; b += kFcWarps) {
    const int r0 = 8 * b;
    uint32_t* q = reinterpret_cast<uint32_t*>(smem + lay.queue + b * kFcQueueBytes);  // one-hot flag, later corner entries
    uint8_t* qc = reinterpret_cast<uint8_t*>(q + kQCap);                              // position code, later the band's bitmap
    int nCorner = 0;
    if (r0 < th) {
      // validity masks of this lane's 8x8 block in the flag layout: byte b of a half-register holds pixels b (word A,
      // bits 7,5,3,1 for rows 0..3 of the half) and 4+b (word B, bits 6,4,2,0)
      uint32_t vmLo, vmHi;
      {
        const int a = min(max(xoff - 8 * lane, 0), 8), e = min(max(xoff + tw - 8 * lane, 0), 8);
        const uint32_t m8 = e > a ? ((1u << e) - 1u) & ~((1u << a) - 1u) : 0u;  // bit i = pixel i of the lane is tested
        const uint32_t sa = ((m8 & 15u) * 0x00204081u) & 0x01010101u;          // bit 0 of byte b = pixel b
        const uint32_t sb = ((m8 >> 4) * 0x00204081u) & 0x01010101u;           // bit 0 of byte b = pixel 4+b
        const uint32_t xm = sa * 0xAAu + sb * 0x55u;
        uint32_t rl = 0, rh = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (r0 + j < th) rl |= 0xC0C0C0C0u >> (2 * j);
          if (r0 + 4 + j < th) rh |= 0xC0C0C0C0u >> (2 * j);
        }
        vmLo = xm & rl;
        vmHi = xm & rh;
      }
      // ---------------- phase 1: 64 prefilter flags per lane
      uint32_t lo = 0, hi = 0;
      if (__any_sync(0xffffffffu, (vmLo | vmHi) != 0)) {
        uint32_t ra[14], rb[14];
        const uint32_t* col = reinterpret_cast<const uint32_t*>(sInB + r0 * kRowB + o) + 2 * lane;
#pragma unroll
        for (int i = 0; i < 14; i++) {
          const uint2 v = *reinterpret_cast<const uint2*>(col + i * kFcInWords);
          ra[i] = v.x;
          rb[i] = v.y;
        }
        const uint32_t one = g.one;
        const uint32_t M = g.absMask;  // 0x80 - (minTh + 1) in every byte: x + M has its msb set iff x > minTh (x < 0x80)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int c = j + 3;
          const uint32_t* crow = col + c * kFcInWords;
          const uint32_t wl = crow[-1], wr = crow[2];  // lane 0 / o = 0: the word before the row -- only feeds untested pixels
          const uint32_t cA = ra[c], cB = rb[c];
          const uint32_t vA = __vabsdiffu4(ra[j], cA) | __vabsdiffu4(cA, ra[c + 3]);
          const uint32_t vB = __vabsdiffu4(rb[j], cB) | __vabsdiffu4(cB, rb[c + 3]);
          const uint32_t hA = __vabsdiffu4(cA, __byte_perm(wl, cA, 0x4321)) | __vabsdiffu4(cA, __byte_perm(cA, cB, 0x6543));
          const uint32_t hB = __vabsdiffu4(cB, __byte_perm(cA, cB, 0x4321)) | __vabsdiffu4(cB, __byte_perm(cB, wr, 0x6543));
          // msb of a byte: the absolute difference exceeds minTh.  x + M sets the msb for x in (minTh, 0x7f]; "| x"
          // covers x >= 0x80; a carry out of a byte can only turn a neighbour's flag ON (over-accepting is harmless)
          const uint32_t fA = (mad1(vA, one, M) | vA) & (mad1(hA, one, M) | hA);
          const uint32_t fB = (mad1(vB, one, M) | vB) & (mad1(hB, one, M) | hB);
          const int s = 2 * (j & 3);
          const uint32_t bits = ((fA >> s) & (0x80808080u >> s)) | ((fB >> (s + 1)) & (0x80808080u >> (s + 1)));
          if (j < 4) lo |= bits; else hi |= bits;
        }
        lo &= vmLo;
        hi &= vmHi;
      }
      // ---------------- expansion: 4x4 byte transpose inside lane quads evens out the per-lane counts
      {
        const uint32_t sel1 = (lane & 1) ? 0x3715u : 0x6240u, sel2 = (lane & 2) ? 0x3276u : 0x5410u;
        uint32_t x = __shfl_xor_sync(0xffffffffu, lo, 1), y = __shfl_xor_sync(0xffffffffu, hi, 1);
        lo = __byte_perm(lo, x, sel1);
        hi = __byte_perm(hi, y, sel1);
        x = __shfl_xor_sync(0xffffffffu, lo, 2);
        y = __shfl_xor_sync(0xffffffffu, hi, 2);
        lo = __byte_perm(lo, x, sel2);
        hi = __byte_perm(hi, y, sel2);
      }
      const int cntL = __popc(lo) + __popc(hi);
      int incl = cntL;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      // Normally the band's candidates fit the queue in one pass; otherwise (noise images) one pass per row (<= 256) and
      // the NMS of this band takes the slow path (the corner list cannot live in a queue that is refilled)
      const int nParts = total <= kQCap ? 1 : 8;
      uint8_t* band = sTile + r0 * kSP;
      for (int part = 0; part < nParts; part++) {
        uint32_t mlo = lo, mhi = hi;
        int pos = incl - cntL, T = total;
        if (nParts > 1) {
          const uint32_t rm = 0xC0C0C0C0u >> (2 * (part & 3));
          mlo = part < 4 ? (lo & rm) : 0u;
          mhi = part < 4 ? 0u : (hi & rm);
          const int c2 = __popc(mlo) + __popc(mhi);
          int in2 = c2;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, in2, d);
            if (lane >= d) in2 += t;
          }
          T = __shfl_sync(0xffffffffu, in2, 31);
          pos = in2 - c2;
        }
        const uint8_t pcode = (uint8_t)((lane & 28) * 8 + (lane & 3));  // x of (source-lane quad, column); bit 2 = half
        uint32_t* qp = q + pos;
        uint8_t* qcp = qc + pos;
        while (mlo) {
          const uint32_t low = mlo & (0u - mlo);
          mlo ^= low;
          *qp++ = low;
          *qcp++ = pcode;
        }
        while (mhi) {
          const uint32_t low = mhi & (0u - mhi);
          mhi ^= low;
          *qp++ = low;
          *qcp++ = (uint8_t)(pcode | 4);
        }
        __syncwarp();
        // ---------------- phase 2: exact score, one candidate per lane; true corners are compacted in place
        for (int base = 0; base < T; base += 32) {
          const int i = base + lane;
          bool corner = false;
          uint32_t entry = 0;
          if (i < T) {
            const uint32_t low = q[i], c = qc[i];
            const uint32_t bit = 31u - (uint32_t)__clz(low), u = bit ^ 7u;  // u & 7 = 2 * (row in half) + word
            const int row = (int)(c & 4u) + (int)((u >> 1) & 3u);
            const int x = (int)((c & 0xE3u) + (bit & 0x18u) + ((u & 1u) << 2));  // (lane quad)*32 + (source lane)*8 + word*4 + byte
            const int bam = fast_bam_minmax(sInB + (r0 + row + 3) * kRowB + o + x);
            if (bam > g.minTh) {
              band[row * kSP + x] = (uint8_t)(bam - 1);
              corner = true;
              entry = (uint32_t)x | ((uint32_t)(r0 + row) << 8) | ((uint32_t)(bam - 1) << 16);
            }
          }
          if (nParts == 1) {
            const uint32_t bal = __ballot_sync(0xffffffffu, corner);
            __syncwarp();  // this round's queue reads (all lanes) are ordered before the in-place writes below
            if (corner) q[nCorner + __popc(bal & ((1u << lane) - 1u))] = entry;
            nCorner += __popc(bal);
          }
        }
        __syncwarp();
      }
      if (nParts > 1) nCorner = -1;
    }
    // the band's survivor bitmap (8 rows x 256 bits) lives in the queue's code bytes, dead from here on
    __syncwarp();
    reinterpret_cast<uint2*>(qc)[lane] = make_uint2(0u, 0u);
    if (lane == 0) sCorner[b] = nCorner;
  }
  __syncthreads();  // every score of the tile is in shared memory

  // ================================================================ NMS: 3x3 inside the corner's own cell
  // Survivors go to ONE list per tile (in the input stage, dead since the barrier): cell | row | x | score.  A tile has
  // ~20 of them on natural images; the list holds kFcListCap.  If it overflows (noise), the NMS is repeated into per-row
  // bitmaps and the cells are emitted from those (the general, slower path below).
  const uint32_t recip = L.fcRecip;  // (x * recip) >> 16 = x / wCell for x < 1024
  uint32_t* sList = reinterpret_cast<uint32_t*>(smem);  // [kFcListCap] survivors, then [kFcListCap] packed rank counters
  uint32_t* sRank = sList + kFcListCap;                 // (four 8-bit counts: the list holds <= 128 survivors)
  int* sN = sCorner + 8;
  if (tid < kFcListCap) sRank[tid] = 0u;
  auto nms_pass = [&](const bool toBitmap) {
    for (int b = warp; b < nb; b += kFcWarps) {
      const int nC = sCorner[b];
      const uint32_t* q = reinterpret_cast<const uint32_t*>(smem + lay.queue + b * kFcQueueBytes);
      uint32_t* bm = reinterpret_cast<uint32_t*>(smem + lay.queue + b * kFcQueueBytes + kQCap * 4);  // [8 rows][8 words]
      auto test = [&](int x, int row, int sc) {
        const int rel = x - xoff;
        const int c = (int)(((uint32_t)rel * recip) >> 16), c0 = c * wCell;
        if (nms_keep(sTile + row * kSP + x, sc, rel == c0, rel == c0 + wCell - 1)) {
          if (toBitmap) atomicOr(bm + (row & 7) * 8 + (x >> 5), 1u << (x & 31));
          else {
            const int pos = atomicAdd(sN, 1);
            if (pos < kFcListCap) sList[pos] = ((uint32_t)c << 24) | ((uint32_t)row << 16) | ((uint32_t)x << 8) | (uint32_t)sc;
          }
        }
      };
      if (nC >= 0) {
        for (int i = lane; i < nC; i += 32) {
          const uint32_t e = q[i];
          test((int)(e & 0xff), (int)((e >> 8) & 0xff), (int)(e >> 16));
        }
      } else {  // the band's queue was refilled per row (> kQCap candidates): every non-zero score of the band
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
          const uint8_t* rowp = sTile + (8 * b + j) * kSP;
          const uint2 v = *reinterpret_cast<const uint2*>(rowp + 8 * lane);
          uint32_t nzLo = (((v.x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v.x) & 0x80808080u;
          uint32_t nzHi = (((v.y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v.y) & 0x80808080u;
          while (nzLo | nzHi) {
            const bool inLo = nzLo != 0;
            uint32_t& m = inLo ? nzLo : nzHi;
            const int k = (31 - __clz(m & (0u - m))) >> 3;
            m &= m - 1;
            const int x = 8 * lane + k + (inLo ? 0 : 4);
            test(x, 8 * b + j, (int)rowp[x]);
          }
        }
      }
    }
  };
  nms_pass(false);
  __syncthreads();  // every survivor is in the list (or the list overflowed)

  const int iniTh = g.iniTh;
  const int nSurv = *sN;
  if (nSurv <= kFcListCap) {
    // ================================================================ emission from the list
    // A survivor's slot in its cell = number of survivors of the same cell that precede it in the reference's row-major
    // order, counted among those with score >= iniTh if the cell has any such (then the others are dropped: the cell was
    // detected at iniThFAST), among all otherwise (the cell went to the minThFAST retry).  The n x n comparison is split
    // over the CTA: lane = survivor i (chunks of 32), warp = a quarter of the partners j (shared-memory broadcasts);
    // the four partial counts meet in shared-memory atomics.  (One warp doing all of it was a serial tail: the other
    // three had exited but the CTA's shared memory stayed allocated.)
    const int jq = (nSurv + kFcWarps - 1) / kFcWarps, j0 = warp * jq, j1 = min(j0 + jq, nSurv);
    for (int base = 0; base < nSurv; base += 32) {
      const int i = base + lane;
      const uint32_t ai = (i < nSurv ? sList[i] : 0xffffffffu) >> 8;  // cell | row | x
      int cntAll = 0, cnt20 = 0, lessAll = 0, less20 = 0;
      for (int j = j0; j < j1; j++) {
        const uint32_t e = sList[j], a = e >> 8;
        const bool same = (a ^ ai) < 0x10000u, hi = (int)(e & 0xffu) >= iniTh, less = a < ai;
        cntAll += same;
        cnt20 += same && hi;
        lessAll += same && less;
        less20 += same && hi && less;
      }
      if (i < nSurv && j1 > j0) atomicAdd(&sRank[i], (uint32_t)cntAll | ((uint32_t)cnt20 << 8) | ((uint32_t)lessAll << 16) | ((uint32_t)less20 << 24));
    }
    __syncthreads();
    for (int i = tid; i < nSurv; i += kFcThreads) {
      const uint32_t ei = sList[i], rk = sRank[i];
      const int c = (int)(ei >> 24), row = (int)((ei >> 16) & 0xffu), x = (int)((ei >> 8) & 0xffu), sc = (int)(ei & 0xffu);
      const int cntAll = rk & 0xff, cnt20 = (rk >> 8) & 0xff, lessAll = (rk >> 16) & 0xff, less20 = rk >> 24;
      const bool any20 = cnt20 > 0;
      if (!any20 || sc >= iniTh) {
        const int pos = any20 ? less20 : lessAll, total = any20 ? cnt20 : cntAll;
        uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase + (size_t)(ci0 * L.nCols + cj0 + c) * L.slotCap;
        if (pos < L.slotCap)
          slot[pos] = (uint32_t)(X0 - xoff + x - kMinBorder) | ((uint32_t)(Y0 + row - kMinBorder) << 12) | ((uint32_t)sc << 24);
        if (pos == 0) {  // the cell's first candidate also reports the cell's count (cells without survivors keep the 0 written at the start)
          if (total > L.slotCap) atomicOr(err, kErrCandOverflow);
          cnt[c] = min(total, L.slotCap);
        }
      }
    }
    return;
  }
  nms_pass(true);
  __syncthreads();  // every survivor bit is set

  // ================================================================ emission from the bitmaps: warp = cell, lane = tested row
  for (int c = warp; c < kc; c += kFcWarps) {
    const int cx0 = xoff + c * wCell, cw = min(wCell, xoff + tw - cx0);  // the cell's columns in the lane frame
    uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase + (size_t)(ci0 * L.nCols + cj0 + c) * L.slotCap;
    if (cw <= 0) {
      if (lane == 0) cnt[c] = 0;
      continue;
    }
    const int wi = cx0 >> 5, sh = cx0 & 31;
    int basei = 0;
    // pass 1 decides the threshold (any survivor with score >= iniTh in the whole cell), pass 2 emits.  Class A: cells
    // are at most 32 x 32, one 32-bit window per row; otherwise up to 59 x 59: 64-bit windows, two passes of 32 rows.
    typedef typename std::conditional<kA, uint32_t, uint64_t>::type mask_t;
    constexpr int kPasses = kA ? 1 : 2;
    mask_t keep[kPasses], keep20[kPasses];
    bool any20 = false;
    const int nPass = kA ? 1 : (th + 31) >> 5;  // warp-uniform
#pragma unroll
    for (int p = 0; p < kPasses; p++) {
      keep[p] = 0; keep20[p] = 0;
      if (p >= nPass) continue;
      const int r = lane + 32 * p;
      if (r < th) {
        // bits [cx0, cx0 + cw) of the row's 256-bit survivor bitmap; words past the row's end are clamped re-reads whose
        // bits land at or above cw (cx0 + cw <= 256) and fall to the width mask
        const uint32_t* rowBm = reinterpret_cast<const uint32_t*>(smem + lay.queue + (r >> 3) * kFcQueueBytes + kQCap * 4) + (r & 7) * 8;
        const uint32_t w0 = rowBm[wi], w1 = rowBm[min(wi + 1, 7)];
        mask_t m = __funnelshift_r(w0, w1, sh);
        if (!kA) m |= (mask_t)((uint64_t)__funnelshift_r(w1, rowBm[min(wi + 2, 7)], sh) << 32);
        m &= cw >= (int)(8 * sizeof(mask_t)) ? ~(mask_t)0 : (((mask_t)1 << cw) - (mask_t)1);
        keep[p] = m;
        mask_t m20 = 0;
        const uint8_t* rowp = sTile + r * kSP + cx0;
        for (mask_t t = m; t; t &= t - 1) {
          const int bx = kA ? __ffs((int)t) - 1 : __ffsll((long long)t) - 1;
          if (rowp[bx] >= iniTh) m20 |= (mask_t)1 << bx;
        }
        keep20[p] = m20;
      }
      any20 = __any_sync(0xffffffffu, keep20[p] != 0) || any20;
    }
    const int xbase = X0 - xoff + cx0 - kMinBorder;  // level x of the cell's first column, relative to the 16-px border
#pragma unroll
    for (int p = 0; p < kPasses; p++) {
      if (p >= nPass) continue;
      const mask_t sel = any20 ? keep20[p] : keep[p];
      const int n = kA ? __popc((uint32_t)sel) : __popcll((uint64_t)sel);
      int in2 = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, in2, d);
        if (lane >= d) in2 += v;
      }
      int pos = basei + in2 - n;
      basei += __shfl_sync(0xffffffffu, in2, 31);
      const int r = lane + 32 * p;
      const uint8_t* rowp = sTile + r * kSP + cx0;
      const uint32_t ybits = (uint32_t)(Y0 + r - kMinBorder) << 12;
      for (mask_t t = sel; t; t &= t - 1) {  // lane order = row order, bit order = x order: the reference's row-major order
        const int bx = kA ? __ffs((int)t) - 1 : __ffsll((long long)t) - 1;
        if (pos < L.slotCap) slot[pos] = (uint32_t)(xbase + bx) | ybits | ((uint32_t)rowp[bx] << 24);
        pos++;
      }
    }
    if (lane == 0) {
      if (basei > L.slotCap) { atomicOr(err, kErrCandOverflow); basei = L.slotCap; }
      cnt[c] = basei;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-tier variant (the hot path for levels whose cells are at most 32 px wide and 8 * kNb px high: every level of a
// 1080p pyramid).  One warp per 8-row band, kNb warps.  Same tile / lane-frame geometry as above; what changes is the
// ORDER of the work, which now follows the reference's own (ORBextractor.cc:807-829: cv::FAST at iniThFAST, and only for
// a cell that came back empty again at minThFAST):
//   * the prefilter produces the flags of BOTH thresholds from the same absolute differences (four more IMAD + LOP3 per
//     word and row);
//   * pass 0 scores the iniThFAST candidates only (natural images: ~45 % fewer than at minThFAST), NMS, survivors to the
//     tile's list, cells with a survivor are marked;
//   * pass 1 (only if the tile has a cell without survivors; warp-uniform): the minThFAST flags restricted to the
//     columns of those cells are scored (few: such cells are the texture-poor ones), NMS among them, survivors appended
//     to the same list.  A cell that was empty at iniThFAST has no survivor with score >= iniThFAST at minThFAST either
//     (a higher-scored corner is suppressed by the same neighbour in both runs), so the emission below is unchanged;
//   * overflow (noise images: > kFcListCap survivors): every cell is rescored at minThFAST and the NMS goes to per-row
//     bitmaps, emitted by the general path (one NMS serves both thresholds there, as in the kernel above).
// Results are identical to the single-pass kernel's (tests/test_gpu_orb.py::test_fused_fast_cells_equals_the_unfused_pair).
template <int kNb>
__global__ void __launch_bounds__(32 * kNb, kNb == 4 ? kFcOccA : 6) k_fast_cells2(const __grid_constant__ OrbGeo g,
                                                                                 const __grid_constant__ TmapIn tm,
                                                                                 const int4* __restrict__ tileTab, int frame0,
                                                                                 uint32_t* __restrict__ slots,
                                                                                 int* __restrict__ cellCnt, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kT = 32 * kNb, kQ2 = kFc2QueueCap, kQ2Bytes = kFc2QueueCap * 3;
  constexpr int kSP = kFc2TilePitch;  // (shadows the single-pass kernel's pitch)
  constexpr FcSmem lay = fc2_smem_layout(kNb);
  const uint8_t* sInB = smem;
  uint8_t* sTile = smem + lay.tile + kSP + 8;  // (row 0, x 0) of the score tile: 8 pad bytes in front of every row
  uint32_t* sList = reinterpret_cast<uint32_t*>(smem + lay.misc);  // [kFcListCap] survivors
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.misc + kFcListCap * 4);
  int* sCorner = reinterpret_cast<int*>(smem + lay.misc + kFcListCap * 4 + 16);  // per band: corners in the list, or -1 = use the slow NMS
  int* sN = sCorner + 8;                                                        // survivors of the tile
  uint32_t* sCells = reinterpret_cast<uint32_t*>(sCorner + 9);                  // bit c: cell c of the tile has a survivor
  int* sQn = sCorner + 10;                                                      // [2] candidates in the tile's pooled queue, per pass
  uint16_t* sLut = reinterpret_cast<uint16_t*>(sCorner + 12);                   // [32] flag bit -> x | row << 8 inside a transposed flag word
  // pooled queue of the tile (fast path): one 16-bit entry x | row << 8 per candidate, later per corner
  constexpr int kTileQ = kFc2QueueCap * kNb * 3 / 2;
  uint16_t* tq = reinterpret_cast<uint16_t*>(smem + lay.queue);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y;
  const int4 te = __ldg(&tileTab[blockIdx.x]);
  const int level = te.x, X0 = te.y, Y0 = te.z, ci0 = te.w & 0xffff, cj0 = te.w >> 16;
  const LevelGeo& L = g.lv[level];
  const int wCell = L.wCell, hCell = L.hCell;
  const int kc = min(L.fcKc, L.nCols - cj0);
  const int X1 = min(X0 + kc * wCell, L.w - kEdge), Y1 = min(Y0 + hCell, L.h - kEdge);
  int* cnt = cellCnt + (size_t)f * g.totalCells + L.cellBase + ci0 * L.nCols + cj0;
  if (X1 <= X0 || Y1 <= Y0) {  // cells the reference skips (ORBextractor.cc:796-805) or whose view is too small for FAST
    if (tid < kc) cnt[tid] = 0;
    return;
  }
  const int xa = (X0 - 3) & ~15;           // first byte of the TMA box (level x), 16-byte aligned, <= X0 - 3
  const int o = (X0 - xa) & ~7;            // lane frame: staged bytes [o, o + 256) of every row, 8-byte aligned
  const int xoff = X0 - xa - o;            // tile's first tested pixel inside the lane frame: 0 .. 7
  const int tw = X1 - X0, th = Y1 - Y0;    // tested pixels of the tile

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)(kRowB * (kNb * 8 + 6)));
    tma_load_3d(smem, &tm.in[level], xa >> 2, Y0 - 3, f + frame0, bar);  // the maps index frames from the batch's base
  }
  {  // zero the score tile (guard rows included)
    uint4* z = reinterpret_cast<uint4*>(smem + lay.tile);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (8 * kNb + 2) * kSP / 16; i += kT) z[i] = zero;
  }
  if (tid < kc) cnt[tid] = 0;  // cells without survivors; the others are overwritten by the emission
  if (tid == 0) { *sN = 0; *sCells = 0u; sQn[0] = 0; sQn[1] = 0; }
  if (tid < 32) {  // bit b of a transposed flag word: byte (b >> 3) = source lane inside the quad, bit 7 - b % 8 = 2 * row + word
    const uint32_t u = (uint32_t)tid ^ 7u;
    sLut[tid] = (uint16_t)((tid & 0x18) + ((u & 1u) << 2) + (((u >> 1) & 3u) << 8));
  }
  __syncthreads();  // mbarrier initialised before anyone polls it
  while (!mbar_try_wait(bar, 0)) {
  }

  const int b = warp, r0 = 8 * warp;
  const bool bandOn = r0 < th;  // warp-uniform
  uint16_t* q = reinterpret_cast<uint16_t*>(smem + lay.queue + b * kQ2Bytes);  // general path: this band's queue (candidates, later
                                                                               // corners), followed by its survivor bitmap
  // column masks in the flag layout: byte k of a half-register holds pixels k (word A, bits 7,5,3,1 for rows 0..3 of
  // the half) and 4+k (word B, bits 6,4,2,0)
  auto expand_cols = [](uint32_t m8) {  // bit i of m8 = pixel i of the lane
    const uint32_t sa = ((m8 & 15u) * 0x00204081u) & 0x01010101u;  // bit 0 of byte k = pixel k
    const uint32_t sb = ((m8 >> 4) * 0x00204081u) & 0x01010101u;   // bit 0 of byte k = pixel 4+k
    return sa * 0xAAu + sb * 0x55u;
  };
  // ================================================================ prefilter: 64 flags per lane and threshold
  uint32_t f7lo = 0, f7hi = 0, f20lo = 0, f20hi = 0;
  if (bandOn) {
    uint32_t vmLo, vmHi;
    {
      const int a = min(max(xoff - 8 * lane, 0), 8), e = min(max(xoff + tw - 8 * lane, 0), 8);
      const uint32_t xm = expand_cols(e > a ? ((1u << e) - 1u) & ~((1u << a) - 1u) : 0u);
      uint32_t rl = 0, rh = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (r0 + j < th) rl |= 0xC0C0C0C0u >> (2 * j);
        if (r0 + 4 + j < th) rh |= 0xC0C0C0C0u >> (2 * j);
      }
      vmLo = xm & rl;
      vmHi = xm & rh;
    }
    if (__any_sync(0xffffffffu, (vmLo | vmHi) != 0)) {
      uint32_t ra[14], rb[14];
      const uint32_t* col = reinterpret_cast<const uint32_t*>(sInB + r0 * kRowB + o) + 2 * lane;
#pragma unroll
      for (int i = 0; i < 14; i++) {
        const uint2 v = *reinterpret_cast<const uint2*>(col + i * kFcInWords);
        ra[i] = v.x;
        rb[i] = v.y;
      }
      uint32_t dvA[11], dvB[11];
#pragma unroll
      for (int i = 0; i < 11; i++) {
        dvA[i] = __vabsdiffu4(ra[i], ra[i + 3]);
        dvB[i] = __vabsdiffu4(rb[i], rb[i + 3]);
      }
      const uint32_t one = g.one;
      const uint32_t M7 = g.absMask, M20 = g.absMaskIni;  // 0x80 - (th + 1) in every byte: x + M has its msb set iff x > th (x < 0x80)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int c = j + 3;
        const uint32_t* crow = col + c * kFcInWords;
        const uint32_t wl = crow[-1], wr = crow[2];  // lane 0 / o = 0: the word before the row -- only feeds untested pixels
        const uint32_t cA = ra[c], cB = rb[c];
        // |row y - row y-3| serves row y (its upper ring pixel) and row y-3 (its lower one): 22 differences per band, not 32
        const uint32_t vA = dvA[j] | dvA[j + 3];
        const uint32_t vB = dvB[j] | dvB[j + 3];
        const uint32_t hA = __vabsdiffu4(cA, __byte_perm(wl, cA, 0x4321)) | __vabsdiffu4(cA, __byte_perm(cA, cB, 0x6543));
        const uint32_t hB = __vabsdiffu4(cB, __byte_perm(cA, cB, 0x4321)) | __vabsdiffu4(cB, __byte_perm(cB, wr, 0x6543));
        // msb of a byte: the (OR of the two) absolute differences exceeds the threshold -- a necessary condition for a
        // 9-arc, which contains one pixel of every opposite pair.  x + M sets the msb for x in (th, 0x7f]; "| x" covers
        // x >= 0x80; a carry out of a byte can only turn a neighbour's flag ON (over-accepting is harmless)
        const uint32_t fA = (mad1(vA, one, M7) | vA) & (mad1(hA, one, M7) | hA);
        const uint32_t fB = (mad1(vB, one, M7) | vB) & (mad1(hB, one, M7) | hB);
        const uint32_t gA = (mad1(vA, one, M20) | vA) & (mad1(hA, one, M20) | hA);
        const uint32_t gB = (mad1(vB, one, M20) | vB) & (mad1(hB, one, M20) | hB);
        const int s = 2 * (j & 3);
        const uint32_t b7 = ((fA >> s) & (0x80808080u >> s)) | ((fB >> (s + 1)) & (0x80808080u >> (s + 1)));
        const uint32_t b20 = ((gA >> s) & (0x80808080u >> s)) | ((gB >> (s + 1)) & (0x80808080u >> (s + 1)));
        if (j < 4) { f7lo |= b7; f20lo |= b20; } else { f7hi |= b7; f20hi |= b20; }
      }
      f7lo &= vmLo; f7hi &= vmHi;
      f20lo &= f7lo; f20hi &= f7hi;  // (a carry may have switched an iniTh flag on where the minTh flag is off: keep the sets nested)
    }
  }

  const int iniTh = g.iniTh;
  const uint32_t recip = L.fcRecip;  // (x * recip) >> 16 = x / wCell for x < 1024
  const uint32_t allCells = (1u << kc) - 1u;
  int pass = 0;
  uint32_t cellSel = allCells;  // cells whose corners this pass scores
  bool toBitmap = false;
  int nSurv = 0;
  uint32_t myCells = 0u;
  // One corner against its cell-local 3x3 neighbourhood; survivors go to the tile's list (or, general path after a list
  // overflow, to the per-row bitmaps in the band queues' code bytes)
  auto test = [&](int x, int row, int sc) {
    const int rel = x - xoff;
    const int c = (int)(((uint32_t)rel * recip) >> 16), c0 = c * wCell;
    if (!((cellSel >> c) & 1u)) return;  // (slow path of a later pass: scores of the cells already decided)
    if (nms_keep<kSP>(sTile + row * kSP + x, sc, rel == c0, rel == c0 + wCell - 1)) {
      if (toBitmap) {
        uint32_t* bm = reinterpret_cast<uint32_t*>(smem + lay.queue + (row >> 3) * kQ2Bytes + kQ2 * 2);  // [8 rows][8 words]
        atomicOr(bm + (row & 7) * 8 + (x >> 5), 1u << (x & 31));
      } else {
        const int pos = atomicAdd(sN, 1);
        if (pos < kFcListCap) sList[pos] = ((uint32_t)c << 24) | ((uint32_t)row << 16) | ((uint32_t)x << 8) | (uint32_t)sc;
        myCells |= 1u << c;
      }
    }
  };
  // this lane's pixels inside the selected cells, in the flag layout (a lane's 8 pixels touch at most two cells: wCell >= 30)
  auto cell_cols = [&]() {
    const int relFirst = 8 * lane - xoff;
    const int c0 = (int)(((uint32_t)max(relFirst, 0) * recip) >> 16);
    const int bnd = min(max((c0 + 1) * wCell - relFirst, 0), 8);
    const uint32_t lowm = (1u << bnd) - 1u;
    return expand_cols((((cellSel >> c0) & 1u) ? lowm : 0u) | (((cellSel >> (c0 + 1)) & 1u) ? (0xffu & ~lowm) : 0u));
  };

  // ================================================================ fast path: the tile's candidates in ONE pooled queue
  // Every warp appends its band's candidates (one atomicAdd per warp for the base), then the 32-candidate rounds are
  // dealt round-robin to the warps -- bands with many candidates no longer keep the others waiting at the barrier --
  // and each round's corners are compacted into the round's own 32 slots.  Leaves for the general per-band loop below
  // when the pooled queue or the survivor list overflows (noise).
  bool general = false;
  for (;;) {
    uint32_t lo = 0, hi = 0;
    const int thr = pass == 0 ? iniTh : g.minTh;  // (every warp scores rounds of the pooled queue, also one whose own band is empty)
    if (bandOn) {
      if (pass == 0) {
        lo = f20lo; hi = f20hi;
      } else {
        const uint32_t xm = cell_cols();
        lo = f7lo & xm; hi = f7hi & xm;
      }
    }
    if (__any_sync(0xffffffffu, (lo | hi) != 0)) {
      {  // 4x4 byte transpose inside lane quads evens out the per-lane counts
        const uint32_t sel1 = (lane & 1) ? 0x3715u : 0x6240u, sel2 = (lane & 2) ? 0x3276u : 0x5410u;
        uint32_t x = __shfl_xor_sync(0xffffffffu, lo, 1), y = __shfl_xor_sync(0xffffffffu, hi, 1);
        lo = __byte_perm(lo, x, sel1);
        hi = __byte_perm(hi, y, sel1);
        x = __shfl_xor_sync(0xffffffffu, lo, 2);
        y = __shfl_xor_sync(0xffffffffu, hi, 2);
        lo = __byte_perm(lo, x, sel2);
        hi = __byte_perm(hi, y, sel2);
      }
      const int cntL = __popc(lo) + __popc(hi);
      int incl = cntL;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      int base = 0;
      if (lane == 31) base = atomicAdd(&sQn[pass], incl);
      base = __shfl_sync(0xffffffffu, base, 31);
      if (base + __shfl_sync(0xffffffffu, incl, 31) <= kTileQ) {
        // x of (source-lane quad, column) | first row of the band << 8; the hi word holds rows 4..7
        uint32_t b16 = (uint32_t)((lane & 28) * 8 + (lane & 3) + (warp << 11));
        uint32_t qa = smem_u32(tq + base + incl - cntL);
        const uint32_t lutA = smem_u32(sLut);
        asm volatile("" : "+r"(b16), "+r"(qa));  // (keeps the loop-invariant values in registers)
#pragma unroll 1
        while (lo) {  // highest flag first (one FLO); the order inside the queue does not matter
          uint32_t fb, v;
          asm("bfind.u32 %0, %1;" : "=r"(fb) : "r"(lo));
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lutA + 2u * fb));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(v + b16) : "memory");
          qa += 2;
          lo ^= 1u << fb;
        }
        b16 += 0x400u;
#pragma unroll 1
        while (hi) {
          uint32_t fb, v;
          asm("bfind.u32 %0, %1;" : "=r"(fb) : "r"(hi));
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lutA + 2u * fb));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "r"(v + b16) : "memory");
          qa += 2;
          hi ^= 1u << fb;
        }
      }
    }
    __syncthreads();  // the pooled queue is complete
    const int nQ = sQn[pass];
    if (nQ > kTileQ) {  // does not fit: this pass (and what follows) per band
      general = true;
      break;
    }
    // scoring rounds, dealt round-robin; a warp compacts the corners it finds into the slots of its own rounds, in
    // order (its k-th corner -> slot k % 32 of its (k / 32)-th round: never ahead of the round being read)
    int nCw = 0;
    for (int rbase = 32 * warp; rbase < nQ; rbase += 32 * kNb) {
      const int i = rbase + lane;
      bool corner = false;
      uint32_t e = 0;
      if (i < nQ) {
        e = tq[i];
        const int x = (int)(e & 0xffu), row = (int)(e >> 8);
        const int bam = fast_bam_minmax(sInB + (row + 3) * kRowB + o + x);
        if (bam > thr) {
          sTile[row * kSP + x] = (uint8_t)(bam - 1);
          corner = true;
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, corner);
      __syncwarp();  // this round's queue reads (all lanes) are ordered before the in-place writes below
      if (corner) {
        const int k = nCw + __popc(bal & ((1u << lane) - 1u));
        tq[32 * (warp + kNb * (k >> 5)) + (k & 31)] = (uint16_t)e;
      }
      nCw += __popc(bal);
    }
    __syncthreads();  // every score of the pass is in shared memory
    for (int k = lane; k < nCw; k += 32) {
      const uint32_t e = tq[32 * (warp + kNb * (k >> 5)) + (k & 31)];
      const int x = (int)(e & 0xffu), row = (int)(e >> 8);
      test(x, row, (int)sTile[row * kSP + x]);
    }
    if (pass == 0) {
      myCells = __reduce_or_sync(0xffffffffu, myCells);
      if (lane == 0 && myCells) atomicOr(sCells, myCells);
    }
    __syncthreads();  // every survivor of the pass is in the list (or the list overflowed)
    nSurv = *sN;
    if (nSurv > kFcListCap) {  // noise: rescore every cell at minThFAST, NMS into the bitmaps, general emission
      toBitmap = true;
      cellSel = allCells;
      pass = 1;
      general = true;
      break;
    }
    if (pass == 0) {
      const uint32_t empty = allCells & ~*sCells;
      if (empty) {
        cellSel = empty;
        pass = 1;
        continue;
      }
    }
    break;
  }

  // ================================================================ general path: per band, queues refilled row by row
  while (general) {
    // ================================================================ score this band's candidates of the pass
    int nCorner = 0;
    if (bandOn) {
      uint32_t lo, hi;
      int thr;
      if (pass == 0) {
        lo = f20lo; hi = f20hi; thr = iniTh;
      } else {
        const uint32_t xm = cell_cols();
        lo = f7lo & xm; hi = f7hi & xm; thr = g.minTh;
      }
      if (__any_sync(0xffffffffu, (lo | hi) != 0)) {
        // ---------------- expansion: 4x4 byte transpose inside lane quads evens out the per-lane counts
        {
          const uint32_t sel1 = (lane & 1) ? 0x3715u : 0x6240u, sel2 = (lane & 2) ? 0x3276u : 0x5410u;
          uint32_t x = __shfl_xor_sync(0xffffffffu, lo, 1), y = __shfl_xor_sync(0xffffffffu, hi, 1);
          lo = __byte_perm(lo, x, sel1);
          hi = __byte_perm(hi, y, sel1);
          x = __shfl_xor_sync(0xffffffffu, lo, 2);
          y = __shfl_xor_sync(0xffffffffu, hi, 2);
          lo = __byte_perm(lo, x, sel2);
          hi = __byte_perm(hi, y, sel2);
        }
        const int cntL = __popc(lo) + __popc(hi);
        int incl = cntL;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        // Normally the band's candidates fit the queue in one pass; otherwise (noise images) one pass per row (<= 256) and
        // the NMS of this band takes the slow path (the corner list cannot live in a queue that is refilled)
        const int nParts = total <= kQ2 ? 1 : 8;
        for (int part = 0; part < nParts; part++) {
          uint32_t mlo = lo, mhi = hi;
          int pos = incl - cntL, T = total;
          if (nParts > 1) {
            const uint32_t rm = 0xC0C0C0C0u >> (2 * (part & 3));
            mlo = part < 4 ? (lo & rm) : 0u;
            mhi = part < 4 ? 0u : (hi & rm);
            const int c2 = __popc(mlo) + __popc(mhi);
            int in2 = c2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, in2, d);
              if (lane >= d) in2 += t;
            }
            T = __shfl_sync(0xffffffffu, in2, 31);
            pos = in2 - c2;
          }
          // 16-bit entries x | row << 8 through the bit -> (x, row) table, as in the pooled queue
          uint32_t b16 = (uint32_t)((lane & 28) * 8 + (lane & 3) + (warp << 11));
          uint16_t* qp = q + pos;
          while (mlo) {
            const uint32_t fb = 31u - (uint32_t)__clz(mlo);
            mlo ^= 1u << fb;
            *qp++ = (uint16_t)(b16 + sLut[fb]);
          }
          b16 += 0x400u;
          while (mhi) {
            const uint32_t fb = 31u - (uint32_t)__clz(mhi);
            mhi ^= 1u << fb;
            *qp++ = (uint16_t)(b16 + sLut[fb]);
          }
          __syncwarp();
          // ---------------- exact score, one candidate per lane; true corners are compacted in place
          for (int base = 0; base < T; base += 32) {
            const int i = base + lane;
            bool corner = false;
            uint32_t e = 0;
            if (i < T) {
              e = q[i];
              const int x = (int)(e & 0xffu), row = (int)(e >> 8);
              const int bam = fast_bam_minmax(sInB + (row + 3) * kRowB + o + x);
              if (bam > thr) {
                sTile[row * kSP + x] = (uint8_t)(bam - 1);
                corner = true;
              }
            }
            if (nParts == 1) {
              const uint32_t bal = __ballot_sync(0xffffffffu, corner);
              __syncwarp();  // this round's queue reads (all lanes) are ordered before the in-place writes below
              if (corner) q[nCorner + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)e;
              nCorner += __popc(bal);
            }
          }
          __syncwarp();
        }
        if (nParts > 1) nCorner = -1;
      }
    }
    // the band's survivor bitmap (8 rows x 256 bits) follows its queue
    __syncwarp();
    reinterpret_cast<uint2*>(q + kQ2)[lane] = make_uint2(0u, 0u);
    if (lane == 0) sCorner[b] = nCorner;
    __syncthreads();  // every score of the pass is in shared memory

    // ================================================================ NMS: 3x3 inside the corner's own cell
    {
      myCells = 0u;
      if (nCorner >= 0) {
        for (int i = lane; i < nCorner; i += 32) {
          const uint32_t e = q[i];
          const int x = (int)(e & 0xffu), row = (int)(e >> 8);
          test(x, row, (int)sTile[row * kSP + x]);
        }
      } else {  // the band's queue was refilled per row (> kQ2 candidates): every non-zero score of the band
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
          const uint8_t* rowp = sTile + (r0 + j) * kSP;
          const uint2 v = *reinterpret_cast<const uint2*>(rowp + 8 * lane);
          uint32_t nzLo = (((v.x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v.x) & 0x80808080u;
          uint32_t nzHi = (((v.y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v.y) & 0x80808080u;
          while (nzLo | nzHi) {
            const bool inLo = nzLo != 0;
            uint32_t& m = inLo ? nzLo : nzHi;
            const int k = (31 - __clz(m & (0u - m))) >> 3;
            m &= m - 1;
            const int x = 8 * lane + k + (inLo ? 0 : 4);
            if (x >= xoff && x < xoff + tw) test(x, r0 + j, (int)rowp[x]);
          }
        }
      }
      if (pass == 0) {
        myCells = __reduce_or_sync(0xffffffffu, myCells);
        if (lane == 0 && myCells) atomicOr(sCells, myCells);
      }
    }
    __syncthreads();  // every survivor of the pass is in the list / the bitmaps (or the list overflowed)
    if (toBitmap) break;
    nSurv = *sN;
    if (nSurv > kFcListCap) {  // noise: rescore every cell at minThFAST, NMS into the bitmaps, general emission
      toBitmap = true;
      cellSel = allCells;
      pass = 1;
      continue;
    }
    if (pass == 0) {
      const uint32_t empty = allCells & ~*sCells;
      if (empty) {
        cellSel = empty;
        pass = 1;
        continue;
      }
    }
    break;
  }

  if (!toBitmap) {
    // ================================================================ emission from the list
    // A survivor's slot in its cell = number of survivors of the same cell that precede it in the reference's row-major
    // order.  (No threshold logic here: a cell's survivors are either all from the iniThFAST pass or all from the
    // minThFAST pass of a cell that was empty at iniThFAST.)  One thread per survivor (the list holds <= kFcListCap <= the
    // CTA's threads), the partners are shared-memory broadcasts.
    static_assert(kFcListCap <= kT, "one thread per list entry");
    if (tid < nSurv) {
      const uint32_t ei = sList[tid], ai = ei >> 8;  // cell | row | x
      int total = 0, pos = 0;
      for (int j = 0; j < nSurv; j++) {
        const uint32_t a = sList[j] >> 8;
        const bool same = (a ^ ai) < 0x10000u;
        total += same;
        pos += same && a < ai;
      }
      const int c = (int)(ei >> 24), row = (int)((ei >> 16) & 0xffu), x = (int)((ei >> 8) & 0xffu), sc = (int)(ei & 0xffu);
      uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase + (size_t)(ci0 * L.nCols + cj0 + c) * L.slotCap;
      if (pos < L.slotCap)
        slot[pos] = (uint32_t)(X0 - xoff + x - kMinBorder) | ((uint32_t)(Y0 + row - kMinBorder) << 12) | ((uint32_t)sc << 24);
      if (pos == 0) {  // the cell's first candidate also reports the cell's count (cells without survivors keep the 0 written at the start)
        if (total > L.slotCap) atomicOr(err, kErrCandOverflow);
        cnt[c] = min(total, L.slotCap);
      }
    }
    return;
  }

  // ================================================================ emission from the bitmaps: warp = cell, lane = tested row
  for (int c = warp; c < kc; c += kNb) {
    const int cx0 = xoff + c * wCell, cw = min(wCell, xoff + tw - cx0);  // the cell's columns in the lane frame
    uint32_t* slot = slots + (size_t)f * g.slotsPerFrame + L.slotBase + (size_t)(ci0 * L.nCols + cj0 + c) * L.slotCap;
    if (cw <= 0) {
      if (lane == 0) cnt[c] = 0;
      continue;
    }
    const int wi = cx0 >> 5, sh = cx0 & 31;
    int basei = 0;
    // pass 1 decides the threshold (any survivor with score >= iniTh in the whole cell), pass 2 emits; cells are at most
    // 32 px wide (one 32-bit window per row) and 8 * kNb rows high
    constexpr int kPasses = (8 * kNb + 31) / 32;
    uint32_t keep[kPasses], keep20[kPasses];
    bool any20 = false;
    const int nPass = (th + 31) >> 5;  // warp-uniform
#pragma unroll
    for (int p = 0; p < kPasses; p++) {
      keep[p] = 0; keep20[p] = 0;
      if (p >= nPass) continue;
      const int r = lane + 32 * p;
      if (r < th) {
        const uint32_t* rowBm = reinterpret_cast<const uint32_t*>(smem + lay.queue + (r >> 3) * kQ2Bytes + kQ2 * 2) + (r & 7) * 8;
        const uint32_t w0 = rowBm[wi], w1 = rowBm[min(wi + 1, 7)];
        uint32_t m = __funnelshift_r(w0, w1, sh);
        m &= cw >= 32 ? ~0u : ((1u << cw) - 1u);
        keep[p] = m;
        uint32_t m20 = 0;
        const uint8_t* rowp = sTile + r * kSP + cx0;
        for (uint32_t t = m; t; t &= t - 1) {
          const int bx = __ffs((int)t) - 1;
          if (rowp[bx] >= iniTh) m20 |= 1u << bx;
        }
        keep20[p] = m20;
      }
      any20 = __any_sync(0xffffffffu, keep20[p] != 0) || any20;
    }
    const int xbase = X0 - xoff + cx0 - kMinBorder;  // level x of the cell's first column, relative to the 16-px border
#pragma unroll
    for (int p = 0; p < kPasses; p++) {
      if (p >= nPass) continue;
      const uint32_t sel = any20 ? keep20[p] : keep[p];
      const int n = __popc(sel);
      int in2 = n;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, in2, d);
        if (lane >= d) in2 += v;
      }
      int pos = basei + in2 - n;
      basei += __shfl_sync(0xffffffffu, in2, 31);
      const int r = lane + 32 * p;
      const uint8_t* rowp = sTile + r * kSP + cx0;
      const uint32_t ybits = (uint32_t)(Y0 + r - kMinBorder) << 12;
      for (uint32_t t = sel; t; t &= t - 1) {  // lane order = row order, bit order = x order: the reference's row-major order
        const int bx = __ffs((int)t) - 1;
        if (pos < L.slotCap) slot[pos] = (uint32_t)(xbase + bx) | ybits | ((uint32_t)rowp[bx] << 24);
        pos++;
      }
    }
    if (lane == 0) {
      if (basei > L.slotCap) { atomicOr(err, kErrCandOverflow); basei = L.slotCap; }
      cnt[c] = basei;
    }
  }
}


int configure_fast_cells(int nbGeneric) {
  static DynSmemLimit limA, limA5, limB;
  if (int rc = limA.ensure(k_fast_cells2<4>, fc2_smem_layout(4).total)) return rc;
  if (int rc = limA5.ensure(k_fast_cells2<5>, fc2_smem_layout(5).total)) return rc;
  if (nbGeneric > 0)
    if (int rc = limB.ensure(k_fast_cells<false, 1>, fc_smem_layout(nbGeneric).total)) return rc;
  return PGB_OK;
}

int launch_fast_cells(const OrbGeo& g, const TmapIn& tm, const int4* tileTabA, const int4* tileTabB, const int4* tileTabA5,
                      int nFrames, uint32_t* slots, int* cellCnt, int* err, cudaStream_t st, int frame0, cudaStream_t side,
                      cudaEvent_t evFork, cudaEvent_t evJoin) {
  if (nFrames <= 0) return PGB_OK;
  if (int rc = configure_fast_cells(g.fcTilesB > 0 ? g.fcNbB : 0)) return rc;
  // The 4-band launch holds ~97 % of a 1080p pyramid's tiles; the small ones (level 7's 34-px cell rows, levels with
  // cells wider than 32 px) go to a side stream so that their partial waves fill in next to it instead of running alone.
  const bool fork = side && nFrames > 4 && g.fcTilesA > 0 && (g.fcTilesA5 > 0 || g.fcTilesB > 0);  // (not worth two event hops for a few frames)
  cudaStream_t s2 = fork ? side : st;
  if (fork) {
    PGB_CUDA(cudaEventRecord(evFork, st));
    PGB_CUDA(cudaStreamWaitEvent(side, evFork, 0));
  }
  if (g.fcTilesA5 > 0) {
    dim3 grid(g.fcTilesA5, nFrames);
    k_fast_cells2<5><<<grid, 160, fc2_smem_layout(5).total, s2>>>(g, tm, tileTabA5, frame0, slots, cellCnt, err);
    PGB_LAUNCHED();
  }
  if (g.fcTilesB > 0) {
    dim3 grid(g.fcTilesB, nFrames);
    k_fast_cells<false, 1><<<grid, kFcThreads, fc_smem_layout(g.fcNbB).total, s2>>>(g, tm, tileTabB, g.fcNbB, frame0, slots, cellCnt, err);
    PGB_LAUNCHED();
  }
  if (fork) PGB_CUDA(cudaEventRecord(evJoin, side));
  if (g.fcTilesA > 0) {
    dim3 grid(g.fcTilesA, nFrames);
    k_fast_cells2<4><<<grid, 128, fc2_smem_layout(4).total, st>>>(g, tm, tileTabA, frame0, slots, cellCnt, err);
    PGB_LAUNCHED();
  }
  if (fork) PGB_CUDA(cudaStreamWaitEvent(st, evJoin, 0));
  return PGB_OK;
}

}  // namespace pgb
