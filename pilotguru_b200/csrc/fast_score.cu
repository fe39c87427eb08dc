// K2: FAST-9 score kernel for sm_100a.  Score map = (max arc threshold) where
// the pixel is a FAST-9 corner at minThFAST inside [19, w-19) x [19, h-19), else 0.
// Reference: cv::FAST(TYPE_9_16) as called at ORBextractor.cc:809,:814.
//
// The round-1 ncu capture of the previous design (profiles/r01_fast_score_ncu_full.md: per-row ballots, quantised
// signed prefilter, entry lists, direct global stores) showed 82 % of the issue slots busy at 16 % of DRAM bandwidth:
// the kernel is bound by instruction issue, and on this part LOP3/PRMT/SHF/VIMNMX/VABSDIFF4 share one half-rate pipe
// while IMAD runs on the other (profiles/r01_pipe_probe.txt).  This design therefore minimises instructions on the
// ALU pipe:
//   * one CTA per 256x32 tile (4 warps); one elected thread issues a 3-D TMA load of the 72-word x 38-row halo box; every warp
//     owns an 8-row band of a shared-memory score tile, zeroes it, and at the end stores it with ONE TMA store
//     (clipped by the tensor map) -- no per-thread global stores, no bounds arithmetic on the output side.
//   * phase 1 (prefilter), 8 px per lane per row, no quantisation: VABSDIFF4 gives |c - p| for 4 pixels per
//     instruction; "(p0 or p8) and (p4 or p12) differ from the centre by more than t'" with t' = 2^k - 1 <= minTh is
//     a bit test on the OR of two absolute differences.  It ignores polarity, which costs 3.6 % more candidates than
//     the 6-bit signed test it replaced (measured on the synthetic frames) and saves the whole quantisation pass.
//     The vertical differences |row r - row r+3| are shared between the two centre rows that use them.
//     The 64 flags of a lane's 8x8 block stay in two registers; there are no per-row ballots or list stores.
//   * expansion: a 4x4 byte transpose inside lane quads evens out the blobs, one warp scan of the per-lane counts,
//     then every lane appends the isolated flag bits of its candidates to the warp's queue (m & -m: no FLO/BREV in
//     the divergent loop; the bit index is recovered once per scoring round) -- no per-candidate ballots, no CTA
//     barrier, no cross-warp rebalancing.
//   * phase 2 (exact score), one candidate per lane: ring pixel p becomes (p, -p) in the two s16 halves of a
//     register with one IMAD (FMA pipe); the circular 9-wide sliding MAX is two rounds of VIMNMX3.S16x2, the MIN
//     over the 16 arcs a 3-input tree; dark = v - min_arcs(max_arc p), bright = max_arcs(min_arc p) - v.
#include <cuda_runtime.h>

#include "common.cuh"
#include "fast_common.cuh"
#include "orb_kernels.cuh"

namespace pgb {

namespace {

using namespace fastk;
constexpr int kRowB = kF2InWords * 4;  // 288 bytes per staged input row
// (Tried: 16-row bands per warp to amortise the per-band overhead -- 14 % fewer instructions but half the resident warps;
// 5.0 -> 5.8 us/frame.)
// (Tile height: 64 rows / 8 warps per CTA 4.97 us/frame, 32 rows / 4 warps 4.59, 16 rows / 2 warps 4.71 -- a CTA lives as
// long as its slowest band, so smaller CTAs keep more warps resident; below 32 rows the halo and per-CTA set-up win.)
// (Tried: a 192-entry queue so that 5 CTAs fit an SM, with a two-pass split for bands above 192 candidates (64-row tiles): 4.97 -> 5.13
// us/frame -- the extra resident warps do not pay for the second passes and the smaller L1.)
__device__ __forceinline__ int fast_bam_minmax(const uint8_t* c) { return fastk::fast_bam_minmax<kRowB>(c); }

}  // namespace

// dynamic shared memory layout (bytes):
//   [0, kInStage)               input tile, kF2InRows x kF2InWords u32
//   [kInStage, +kF2W * kF2H)    score tile 32 x 256 u8 (one 2 KB band per warp, each the source of one TMA store)
//   [.., +warps * 384 * 4)      per-warp candidate queues: one-hot flag bit of the candidate in its register
//   [.., +warps * 384)          ... and the producer's part of the candidate position
//   [.., +16)                   the mbarrier
constexpr int kFsInStage = (kF2InBytes + 127) / 128 * 128;
// Queue entry formats (A/B build macro PGB_FS_Q64): 0 = u32 one-hot flag + u8 code in two arrays (5 B/entry, cap 384);
// 1 = one 8-byte {flag, code} entry written with a single STS.64 and read with a single LDS.64 (cap 224, same bytes).
// Measured: 1 shortens the divergent append loops from 11 to 8 instructions per flag (they are 15 % of the kernel's
// warp-instructions at 7 of 32 active lanes, profiles/r01e_fast_score_sass_regions.md) but runs 4.77 vs 4.60 us/frame:
// bands above 224 candidates need a second pass and the 8-byte stores of divergent lanes conflict more.  Default 0.
#ifndef PGB_FS_Q64
#define PGB_FS_Q64 0
#endif
constexpr int kFsQueueCap = PGB_FS_Q64 ? 224 : 384;
constexpr int kFsQueueEntryBytes = PGB_FS_Q64 ? 8 : 5;
constexpr int kFsWarps = kF2Threads / 32;
#ifndef PGB_FS_OCC
#define PGB_FS_OCC 8
#endif
constexpr int kFsOcc = PGB_FS_OCC;  // resident CTAs per SM the kernel is compiled for
constexpr int kFsSmem = kFsInStage + kF2W * kF2H + kFsWarps * kFsQueueCap * kFsQueueEntryBytes + 16;

template <int kOcc>
__global__ void __launch_bounds__(kF2Threads, kOcc) k_fast_score(const __grid_constant__ OrbGeo g,
                                                                 const __grid_constant__ TmapPack tm,
                                                                 const int4* __restrict__ tileTab, int frame0) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t* sIn = reinterpret_cast<const uint32_t*>(smem);
  uint8_t* sScore = smem + kFsInStage;
  uint32_t* sQueue = reinterpret_cast<uint32_t*>(sScore + kF2W * kF2H);
#if PGB_FS_Q64
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQueue + 2 * kFsWarps * kFsQueueCap);
#else
  uint8_t* sQCode = reinterpret_cast<uint8_t*>(sQueue + kFsWarps * kFsQueueCap);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQCode + kFsWarps * kFsQueueCap);
#endif

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.y + frame0;
  const int4 te = __ldg(&tileTab[blockIdx.x]);
  const int level = te.x, x0 = te.y, y0 = te.z;
  const LevelGeo& L = g.lv[level];

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, kF2InBytes);
    tma_load_3d(smem, &tm.in[level], x0 / 4 - 4, y0 - 3, f, bar);
  }
  const int r0 = warp * 8;
  const bool bandLive = y0 + r0 < L.h;  // else the whole band lies below the level (warp-uniform): nothing to store
  uint8_t* band = sScore + r0 * kF2W;
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 4; i++) *reinterpret_cast<uint4*>(band + i * 512 + lane * 16) = z;
  }

  // validity masks of this lane's 8x8 block in the flag layout: byte b of a half-register holds pixels b (word A,
  // bits 7,5,3,1 for rows 0..3 of the half) and 4+b (word B, bits 6,4,2,0).  Interior bands (warp-uniform) skip it.
  const int gx = x0 + lane * 8;
  uint32_t vmLo = 0xffffffffu, vmHi = 0xffffffffu;
  if (!(x0 >= kEdge && x0 + kF2W <= L.w - kEdge && y0 + r0 >= kEdge && y0 + r0 + 8 <= L.h - kEdge)) {
    const int a = min(max(kEdge - gx, 0), 8), b = min(max(L.w - kEdge - gx, 0), 8);
    const uint32_t m8 = b > a ? ((1u << b) - 1u) & ~((1u << a) - 1u) : 0u;  // bit i = pixel i of the lane is testable
    const uint32_t sa = ((m8 & 15u) * 0x00204081u) & 0x01010101u;          // bit 0 of byte b = pixel b
    const uint32_t sb = ((m8 >> 4) * 0x00204081u) & 0x01010101u;           // bit 0 of byte b = pixel 4+b
    const uint32_t xm = sa * 0xAAu + sb * 0x55u;
    const int gy = y0 + r0;
    uint32_t rl = 0, rh = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (gy + j >= kEdge && gy + j < L.h - kEdge) rl |= 0xC0C0C0C0u >> (2 * j);
      if (gy + 4 + j >= kEdge && gy + 4 + j < L.h - kEdge) rh |= 0xC0C0C0C0u >> (2 * j);
    }
    vmLo = xm & rl;
    vmHi = xm & rh;
  }

  __syncthreads();  // mbarrier initialised before anyone polls it
  if (!bandLive) return;
  while (!mbar_try_wait(bar, 0)) {
  }

  // ---------------- phase 1: 64 prefilter flags per lane
  uint32_t lo = 0, hi = 0;
  if (__any_sync(0xffffffffu, (vmLo | vmHi) != 0)) {
    uint32_t ra[14], rb[14];
    const uint32_t* col = sIn + r0 * kF2InWords + 4 + 2 * lane;
#pragma unroll
    for (int i = 0; i < 14; i++) {
      const uint2 v = *reinterpret_cast<const uint2*>(col + i * kF2InWords);
      ra[i] = v.x;
      rb[i] = v.y;
    }
    const uint32_t one = g.one;
    const uint32_t M = g.absMask;  // K = 0x80 - 2^k in every byte, 2^k - 1 = largest such value <= minTh
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = j + 3;
      const uint32_t* crow = col + c * kF2InWords;
      const uint32_t wl = crow[-1], wr = crow[2];
      const uint32_t cA = ra[c], cB = rb[c];
      const uint32_t vA = __vabsdiffu4(ra[j], cA) | __vabsdiffu4(cA, ra[c + 3]);
      const uint32_t vB = __vabsdiffu4(rb[j], cB) | __vabsdiffu4(cB, rb[c + 3]);
      const uint32_t hA = __vabsdiffu4(cA, __byte_perm(wl, cA, 0x4321)) | __vabsdiffu4(cA, __byte_perm(cA, cB, 0x6543));
      const uint32_t hB = __vabsdiffu4(cB, __byte_perm(cA, cB, 0x4321)) | __vabsdiffu4(cB, __byte_perm(cB, wr, 0x6543));
      // msb of a byte: the absolute difference has a bit at or above k, i.e. exceeds 2^k - 1
      // (the additions are issued as IMAD x*1+M: the FMA pipe idles while LOP3/PRMT/VABSDIFF4 saturate the ALU pipe)
      // x + K sets the msb for x in [2^k, 0x7f]; "| x" covers x >= 0x80; a carry out of a byte >= 0x88 can only turn a
      // neighbour's flag ON (over-accepting is harmless here), never off.
      const uint32_t fA = (mad1(vA, one, M) | vA) & (mad1(hA, one, M) | hA);
      const uint32_t fB = (mad1(vB, one, M) | vB) & (mad1(hB, one, M) | hB);
      const int s = 2 * (j & 3);
      const uint32_t bits = ((fA >> s) & (0x80808080u >> s)) | ((fB >> (s + 1)) & (0x80808080u >> (s + 1)));
      if (j < 4) lo |= bits; else hi |= bits;
    }
    lo &= vmLo;
    hi &= vmHi;
  }

  // ---------------- expansion + phase 2
  // Candidates come in blobs, so a few lanes hold most of a band's flags.  A 4x4 byte transpose inside every lane
  // quad (2 SHFL + 2 PRMT per register) hands lane i of the quad column i of all four lanes, which evens the
  // per-lane counts before the (divergent) append loops.  Byte j of a transposed register = source lane 4q + j.
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
  const uint8_t* sInB = reinterpret_cast<const uint8_t*>(sIn);
#if PGB_FS_Q64
  uint2* q = reinterpret_cast<uint2*>(sQueue) + warp * kFsQueueCap;   // {one-hot flag, code}
#else
  uint32_t* q = sQueue + warp * kFsQueueCap;                 // one-hot flag of the candidate inside its register
  uint8_t* qc = sQCode + warp * kFsQueueCap;                 // (lane & 28) * 8 + (lane & 3) + 4 * half
#endif
  const int cnt = __popc(lo) + __popc(hi);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
#if PGB_FS_Q64
  // Normally the band's candidates fit the queue in one pass; otherwise one pass per half band (4 rows) if both halves
  // fit, else (noise images) one pass per (row, word): at most 4 flags per lane = 128 per pass.
  int nParts = 1;
  if (total > kFsQueueCap) {
    const int nLo = __reduce_add_sync(0xffffffffu, __popc(lo));
    nParts = (nLo <= kFsQueueCap && total - nLo <= kFsQueueCap) ? 2 : 16;
  }
  for (int part = 0; part < nParts; part++) {
    uint32_t mlo = lo, mhi = hi;
    int pos = incl - cnt, T = total;
    if (nParts > 1) {
      const uint32_t rm = nParts == 2 ? 0xffffffffu : (0x80808080u >> (part & 7));
      const bool first = nParts == 2 ? part == 0 : part < 8;
      mlo = first ? (lo & rm) : 0u;
      mhi = first ? 0u : (hi & rm);
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
    const uint32_t pcode = (uint32_t)((lane & 28) * 8 + (lane & 3));  // x of (source-lane quad, column); bit 2 = half
    uint2* qp = q + pos;
    while (mlo) {
      const uint32_t low = mlo & (0u - mlo);
      mlo ^= low;
      *qp++ = make_uint2(low, pcode);
    }
    while (mhi) {
      const uint32_t low = mhi & (0u - mhi);
      mhi ^= low;
      *qp++ = make_uint2(low, pcode | 4u);
    }
#else
  // Normally the band's candidates fit the queue in one pass; otherwise (noise images) one pass per row (<= 256).
  const int nParts = total <= kFsQueueCap ? 1 : 8;
  for (int part = 0; part < nParts; part++) {
    uint32_t mlo = lo, mhi = hi;
    int pos = incl - cnt, T = total;
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
#endif
    __syncwarp();
    for (int i = lane; i < T; i += 32) {
#if PGB_FS_Q64
      const uint2 e = q[i];
      const uint32_t low = e.x, c = e.y;
#else
      const uint32_t low = q[i], c = qc[i];
#endif
      const uint32_t bit = 31u - (uint32_t)__clz(low), u = bit ^ 7u;  // u & 7 = 2 * (row in half) + word
      const int row = (int)(c & 4u) + (int)((u >> 1) & 3u);
      const int x = (int)((c & 0xE3u) + (bit & 0x18u) + ((u & 1u) << 2));  // (lane quad)*32 + (source lane)*8 + word*4 + byte
      const int bam = fast_bam_minmax(sInB + (r0 + row + 3) * kRowB + 16 + x);
      if (bam > g.minTh) band[row * kF2W + x] = (uint8_t)(bam - 1);
    }
    __syncwarp();
  }

  // ---------------- store the band
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(&tm.out[level], x0 / 4, y0 + r0, f, band);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// Function attributes are per device: called once per extractor handle (after its device was made current).
int configure_fast_score() {
  PGB_CUDA(cudaFuncSetAttribute(k_fast_score<kFsOcc>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFsSmem));
  return PGB_OK;
}

int launch_fast_score(const OrbGeo& g, const TmapPack& tm, const int4* tileTab, int frame0, int nFrames,
                      cudaStream_t st) {
  if (g.totalTiles2 <= 0 || nFrames <= 0) return PGB_OK;
  dim3 grid(g.totalTiles2, nFrames);
  k_fast_score<kFsOcc><<<grid, kF2Threads, kFsSmem, st>>>(g, tm, tileTab, frame0);
  PGB_LAUNCHED();
  return PGB_OK;
}

}  // namespace pgb
