// K2 v2: FAST-9 score kernel for sm_100a -- TMA-staged tiles, byte-SIMD prefilter on rolling registers,
// warp-private candidate lists, packed-polarity exact scoring.  Produces exactly the score map of
// k_fast_score (orb_kernels.cu): score = (max arc threshold) where the pixel is a FAST-9 corner at minThFAST inside
// [19, w-19) x [19, h-19), else 0.  Reference: cv::FAST(TYPE_9_16) as called at ORBextractor.cc:809,:814.
//
// Per 256x64 tile (one CTA; 6 CTAs are resident per SM, so one CTA's load latency is covered by the others):
//   * one elected thread issues a 3-D TMA load (cp.async.bulk.tensor) of the 72-word x 70-row halo box of the level
//     image into shared memory, signalled through an mbarrier; out-of-image words arrive as zeros.  Meanwhile
//     every warp writes the zeros of its 8-row band of the score tile straight to global memory (16-byte stores).
//   * phase 1, per warp (8 rows x 256 px, 8 px per lane): the 14 rows the band touches are loaded once as 64-bit
//     words and quantised to 6 bits; for every row the compass test "(p0|p8)&(p4|p12) all darker / all brighter than
//     the centre by more than t" is 8 subtractions per 4 px whose per-byte MSBs are the comparison results.  Lanes
//     with a surviving pixel append one 32-bit entry (8 flags + row + lane) to the warp's own list via one ballot.
//   * phase 2 (after the only CTA barrier): the tile's entries are split evenly over the warps, expanded into
//     per-pixel candidates through a small circular queue (ballot compaction, one level per extra pixel of an
//     entry) and scored exactly 32 at a time; scores are byte stores to global memory.  v-p_k and p_k-v ride in the
//     two s16 halves of one register, produced by a single IMAD per ring pixel ((v-p)*(1-2^16)); the circular
//     9-wide sliding minimum is two rounds of 3-input VIMNMX3.S16x2, the maximum a 3-input tree.
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.cuh"
#include "orb_kernels.cuh"

namespace pgb {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t quant6(uint32_t w) { return (w >> 2) & 0x3f3f3f3fu; }

// prefilter of one centre word: returns per-byte MSB flags
__device__ __forceinline__ uint32_t compass(uint32_t qc, uint32_t q0, uint32_t q8, uint32_t q4, uint32_t q12, uint32_t K) {
  const uint32_t V = qc + K;   // per byte 128 + qc - qth            (no carry: <= 191)
  const uint32_t C = K - qc;   // per byte 128 - qth - qc            (no borrow: >= 1)
  const uint32_t d0 = V - q0, d8 = V - q8, d4 = V - q4, d12 = V - q12;  // msb <=> qc - q >= qth (darker ring)
  const uint32_t b0 = q0 + C, b8 = q8 + C, b4 = q4 + C, b12 = q12 + C;  // msb <=> q - qc >= qth (brighter ring)
  return ((d0 | d8) & (d4 | d12)) | ((b0 | b8) & (b4 | b12));
}

// Exact bam of the pixel at byte pointer c inside the staged tile (row stride kF2InWords*4 bytes).
__device__ __forceinline__ int fast_bam_packed(const uint8_t* c) {
  constexpr int S = kF2InWords * 4;
  const uint32_t v = c[0];
  const uint32_t vK = v - (v << 16);  // v * (1 - 2^16)
  uint32_t w[16];
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
#pragma unroll
  for (int k = 0; k < 16; k++) w[k] = (uint32_t)c[dy[k] * S + dx[k]] * 65535u + vK;  // (v-p)*(1-2^16): lo = v-p, hi = p-v-[p>v]
  uint32_t t3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) t3[k] = __vimin3_s16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
  uint32_t m9[16];
#pragma unroll
  for (int k = 0; k < 16; k++) m9[k] = __vimin3_s16x2(t3[k], t3[(k + 3) & 15], t3[(k + 6) & 15]);
  uint32_t a = __vimax3_s16x2(m9[0], m9[1], m9[2]);
  uint32_t b = __vimax3_s16x2(m9[3], m9[4], m9[5]);
  uint32_t cc = __vimax3_s16x2(m9[6], m9[7], m9[8]);
  uint32_t d2 = __vimax3_s16x2(m9[9], m9[10], m9[11]);
  uint32_t e = __vimax3_s16x2(m9[12], m9[13], m9[14]);
  a = __vimax3_s16x2(a, b, cc);
  d2 = __vimax3_s16x2(d2, e, m9[15]);
  a = __vmaxs2(a, d2);
  const int lo = (int)(short)(a & 0xffff);
  const int hi = ((int)a >> 16) + 1;  // undo the -1 carried by every positive p-v (monotone, so min/max commute)
  return max(lo, hi);
}

}  // namespace

// dynamic shared memory layout (bytes):
//   [0, kInStage)              input tile, kF2InRows x kF2InWords u32
//   [.., +8 warps * 256 * 4)   warp-private entry lists
//   [.., +8 warps * 512 * 2)   per-warp circular candidate queues (u16 codes: row<<8 | x)
//   [.., +8)                   the mbarrier, then the 8 per-warp entry counts
constexpr int kInStage = (kF2InBytes + 127) / 128 * 128;
constexpr int kListPerWarp = 8 * 32;
constexpr int kQueueCap = 512;  // at most 31 + 256 candidates are ever queued
constexpr int kF2Smem = kInStage + 8 * kListPerWarp * 4 + 8 * kQueueCap * 2 + 16 + 32;

template <int kOcc>
__global__ void __launch_bounds__(kF2Threads, kOcc) k_fast_score_v2(const __grid_constant__ OrbGeo g,
                                                                 const __grid_constant__ TmapPack tm,
                                                                 const int4* __restrict__ tileTab,
                                                                 uint8_t* __restrict__ score, int frame0) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t* sIn = reinterpret_cast<const uint32_t*>(smem);
  uint32_t* sList = reinterpret_cast<uint32_t*>(smem + kInStage);
  uint16_t* sQueue = reinterpret_cast<uint16_t*>(sList + 8 * kListPerWarp);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQueue + 8 * kQueueCap);
  int* sCnt = reinterpret_cast<int*>(bar + 2);

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
  uint8_t* out = score + (size_t)f * g.frameStride + L.off;
  // zeros of this warp's band of the score tile (8 rows x 256 B, 64 B per lane), clipped to the level's pitch/height
  {
    const int gx = x0 + lane * 8;
    const uint2 z = make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int gy = y0 + r0 + j;
      if (gy < L.h && gx < L.pitch) *reinterpret_cast<uint2*>(out + (size_t)gy * L.pitch + gx) = z;
    }
  }
  __syncthreads();  // mbarrier initialised before anyone polls it
  while (!mbar_try_wait(bar, 0)) {
  }

  // ---------------- phase 1
  const uint32_t K = 0x80808080u - (uint32_t)g.qTh * 0x01010101u;
  uint32_t* myList = sList + warp * kListPerWarp;
  int cnt = 0;
  {
    const int gx = x0 + lane * 8;
    uint32_t xmA = 0, xmB = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      if (gx + b >= kEdge && gx + b < L.w - kEdge) xmA |= 0x80u << (8 * b);
      if (gx + 4 + b >= kEdge && gx + 4 + b < L.w - kEdge) xmB |= 0x80u << (8 * b);
    }
    const bool bandLive = (y0 + r0 + 7 >= kEdge) && (y0 + r0 < L.h - kEdge);
    if (bandLive && __any_sync(0xffffffffu, (xmA | xmB) != 0)) {
      // quantised own words of the 14 smem rows r0 .. r0+13 (global rows y0+r0-3 .. y0+r0+10)
      uint32_t qa[14], qb[14];
      const uint32_t* col = sIn + r0 * kF2InWords + 4 + 2 * lane;
#pragma unroll
      for (int i = 0; i < 14; i++) {
        const uint2 v = *reinterpret_cast<const uint2*>(col + i * kF2InWords);
        qa[i] = quant6(v.x);
        qb[i] = quant6(v.y);
      }
      const uint32_t lt = (1u << lane) - 1u;
      const uint32_t code = (uint32_t)lane << 8;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int gy = y0 + r0 + j;
        if (gy < kEdge || gy >= L.h - kEdge) continue;  // warp-uniform
        const uint32_t* crow = sIn + (r0 + j + 3) * kF2InWords + 3 + 2 * lane;
        const uint32_t qL = quant6(crow[0]), qR = quant6(crow[3]);
        const uint32_t cA = qa[j + 3], cB = qb[j + 3];
        const uint32_t a12 = __byte_perm(qL, cA, 0x4321), a4 = __byte_perm(cA, cB, 0x6543);
        const uint32_t b12 = __byte_perm(cA, cB, 0x4321), b4 = __byte_perm(cB, qR, 0x6543);
        const uint32_t mA = compass(cA, qa[j + 6], qa[j], a4, a12, K) & xmA;
        const uint32_t mB = compass(cB, qb[j + 6], qb[j], b4, b12, K) & xmB;
        const bool any = (mA | mB) != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, any);
        if (any) myList[cnt + __popc(bal & lt)] = mA | (mB >> 1) | code | (uint32_t)j;
        cnt += __popc(bal);
      }
    }
  }
  if (lane == 0) sCnt[warp] = cnt;
  __syncthreads();  // entry lists complete; every band's zeros are ordered before any score store

  // ---------------- phase 2
  {
    int pre[9];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) pre[w + 1] = pre[w] + sCnt[w];
    const int tot = pre[8];
    const int lo = (int)(((long long)tot * warp) >> 3), hi = (int)(((long long)tot * (warp + 1)) >> 3);
    uint16_t* q = sQueue + warp * kQueueCap;
    const uint8_t* sInB = reinterpret_cast<const uint8_t*>(sIn);
    const uint32_t lt = (1u << lane) - 1u;
    int head = 0, tail = 0;
    auto score_round = [&](int n) {  // the first n queued candidates, one per lane
      if (lane < n) {
        const int code = q[(head + lane) & (kQueueCap - 1)];
        const int r = code >> 8, x = code & 255;
        const int bam = fast_bam_packed(sInB + (r + 3) * (kF2InWords * 4) + 16 + x);
        if (bam > g.minTh) out[(size_t)(y0 + r) * L.pitch + x0 + x] = (uint8_t)(bam - 1);
      }
    };
    for (int e0 = lo; e0 < hi; e0 += 32) {
      const int e = e0 + lane;
      uint32_t flags = 0;
      int base = 0;
      if (e < hi) {
        int w = 0, wbase = 0;
#pragma unroll
        for (int k = 1; k < 8; k++)
          if (e >= pre[k]) { w = k; wbase = pre[k]; }
        const uint32_t entry = sList[w * kListPerWarp + (e - wbase)];
        flags = entry & 0xC0C0C0C0u;
        base = ((w * 8 + (int)(entry & 7)) << 8) | ((int)((entry >> 8) & 31) * 8);
      }
      // one compaction level per flagged pixel of the fullest entry
      uint32_t bal;
      while ((bal = __ballot_sync(0xffffffffu, flags != 0)) != 0) {
        if (flags) {
          const int bit = __ffs(flags) - 1;
          flags &= flags - 1;
          const int xo = (bit >> 3) + (((bit & 7) == 6) ? 4 : 0);
          q[(tail + __popc(bal & lt)) & (kQueueCap - 1)] = (uint16_t)(base + xo);
        }
        tail += __popc(bal);
      }
      __syncwarp();
      while (tail - head >= 32) {
        score_round(32);
        head += 32;
      }
      __syncwarp();
    }
    if (tail > head) score_round(tail - head);
  }
}

int launch_fast_score_v2(const OrbGeo& g, const TmapPack& tm, const int4* tileTab, uint8_t* score, int frame0,
                         int nFrames, cudaStream_t st) {
  static int occ = 0;
  if (!occ) {
    const char* e = getenv("PGB_FAST_OCC");  // resident CTAs per SM the kernel is compiled for (register budget)
    occ = (e && atoi(e) == 6) ? 6 : 5;
    PGB_CUDA(cudaFuncSetAttribute(k_fast_score_v2<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2Smem));
    PGB_CUDA(cudaFuncSetAttribute(k_fast_score_v2<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2Smem));
  }
  if (g.totalTiles2 <= 0 || nFrames <= 0) return PGB_OK;
  dim3 grid(g.totalTiles2, nFrames);
  if (occ == 6) k_fast_score_v2<6><<<grid, kF2Threads, kF2Smem, st>>>(g, tm, tileTab, score, frame0);
  else k_fast_score_v2<5><<<grid, kF2Threads, kF2Smem, st>>>(g, tm, tileTab, score, frame0);
  PGB_LAUNCHED();
  return PGB_OK;
}

}  // namespace pgb
