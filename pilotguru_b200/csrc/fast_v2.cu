// K2 v2: FAST-9 score kernel for sm_100a -- persistent CTAs, TMA-staged tiles, byte-SIMD prefilter on rolling
// registers, warp-private candidate lists, packed-polarity exact scoring.  Produces exactly the score map of
// k_fast_score (orb_kernels.cu): score = (max arc threshold) where the pixel is a FAST-9 corner at minThFAST inside
// [19, w-19) x [19, h-19), else 0.  Reference: cv::FAST(TYPE_9_16) as called at ORBextractor.cc:809,:814.
//
// Per 256x64 tile (one CTA iteration):
//   * one elected thread issues a 3-D TMA load (cp.async.bulk.tensor) of the 72-word x 70-row halo box of the level
//     image into a 2-stage shared-memory ring, signalled through an mbarrier; out-of-image words arrive as zeros.
//     The load of tile i+1 is in flight while tile i is processed.
//   * phase 1, per warp (8 rows x 256 px, 8 px per lane): the 14 rows the band touches are loaded once as 64-bit
//     words and quantised to 6 bits; for every row the compass test "(p0|p8)&(p4|p12) all darker / all brighter than
//     the centre by more than t" is 8 subtractions per 4 px whose per-byte MSBs are the comparison results.  Lanes
//     with a surviving pixel append one 32-bit entry (8 flags + row + lane) to the warp's own list via one ballot.
//   * phase 2: the tile's entries are split evenly over the warps, expanded into per-pixel candidates (warp scan
//     + circular queue) and scored exactly 32 at a time.  v-p_k and p_k-v ride in the
//     two s16 halves of one register, produced by a single IMAD per ring pixel ((v-p)*(1-2^16)); the circular
//     9-wide sliding minimum is two rounds of 3-input VIMNMX3.S16x2, the maximum a 3-input tree.
//   * the zero-initialised 256x64 output tile leaves through a TMA store, which also clips it to the image.
#include <cuda_runtime.h>

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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm),
               "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
//   [0, 2*kInStage)            input ring, each stage kF2InRows x kF2InWords u32 (stage size rounded up to 128)
//   [.., +kF2W*kF2H)           output tile
//   [.., +8 warps * 256 * 4)   warp-private entry lists
//   [.., +8 warps * 512 * 2)   per-warp circular candidate queues
//   [.., +16)                  two mbarriers, then the 8 per-warp entry counts
constexpr int kInStage = (kF2InBytes + 127) / 128 * 128;
constexpr int kOutBytes = kF2W * kF2H;
constexpr int kListPerWarp = 8 * 32;
constexpr int kQueueCap = 512;  // u16 candidate codes per warp; at most 31 + 256 are ever queued
constexpr int kF2Smem = 2 * kInStage + kOutBytes + 8 * kListPerWarp * 4 + 8 * kQueueCap * 2 + 16 + 32;

__global__ void __launch_bounds__(kF2Threads, 3) k_fast_score_v2(const __grid_constant__ OrbGeo g,
                                                                 const __grid_constant__ TmapPack tm, int nFrames) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* sIn0 = reinterpret_cast<uint32_t*>(smem);
  uint8_t* sOut = smem + 2 * kInStage;
  uint32_t* sList = reinterpret_cast<uint32_t*>(sOut + kOutBytes);
  uint16_t* sQueue = reinterpret_cast<uint16_t*>(sList + 8 * kListPerWarp);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQueue + 8 * kQueueCap);
  int* sCnt = reinterpret_cast<int*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = g.totalTiles2 * nFrames;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto decode = [&](int t, int& level, int& x0, int& y0, int& f) {
    f = t / g.totalTiles2;
    int tl = t - f * g.totalTiles2;
    level = 0;
#pragma unroll 1
    for (int l = 1; l < g.nlevels; l++)
      if (tl >= g.lv[l].tile2Base) level = l;
    tl -= g.lv[level].tile2Base;
    const int ty = tl / g.lv[level].tiles2X;
    x0 = (tl - ty * g.lv[level].tiles2X) * kF2W;
    y0 = ty * kF2H;
  };
  auto issue_load = [&](int t, int stage) {
    int level, x0, y0, f;
    decode(t, level, x0, y0, f);
    mbar_expect_tx(&bars[stage], kF2InBytes);
    tma_load_3d(smem + stage * kInStage, &tm.in[level], x0 / 4 - 4, y0 - 3, f, &bars[stage]);
  };

  int t = blockIdx.x;
  if (t < total && tid == 0) issue_load(t, 0);

  const uint32_t K = 0x80808080u - (uint32_t)g.qTh * 0x01010101u;
  uint32_t* myList = sList + warp * kListPerWarp;
  const int r0 = warp * 8;

  for (int it = 0; t < total; it++, t += gridDim.x) {
    const int stage = it & 1;
    const uint32_t parity = (uint32_t)(it >> 1) & 1u;
    if (tid == 0) {
      const int tn = t + gridDim.x;
      if (tn < total) issue_load(tn, stage ^ 1);  // stage^1 was released by the barrier that ended iteration it-1
      tma_store_wait_read0();                     // the previous tile's store has finished reading sOut
    }
    __syncthreads();
    int level, x0, y0, f;
    decode(t, level, x0, y0, f);
    const LevelGeo& L = g.lv[level];
    // zero this warp's band of the output tile: 8 rows x 256 B = 32 lanes x 64 B
    {
      uint4* o = reinterpret_cast<uint4*>(sOut + r0 * kF2W) + lane * 4;
      const uint4 z = make_uint4(0, 0, 0, 0);
      o[0] = z; o[1] = z; o[2] = z; o[3] = z;
    }
    while (!mbar_try_wait(&bars[stage], parity)) {
    }
    const uint32_t* sIn = sIn0 + stage * (kInStage / 4);

    // ---------------- phase 1
    int cnt = 0;
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
        if (bal) {
          if (any) myList[cnt + __popc(bal & lt)] = mA | (mB >> 1) | code | (uint32_t)j;
          cnt += __popc(bal);
        }
      }
    }
    __syncwarp();

    // ---------------- phase 2: the tile's entries (all warps' lists, concatenated) are split evenly over the
    // warps; each warp expands its share into per-pixel candidates through a small circular queue and scores them
    // 32 at a time, so lanes stay full no matter how the candidates cluster.
    if (lane == 0) sCnt[warp] = cnt;
    __syncthreads();
    {
      int pre[9];
      pre[0] = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) pre[w + 1] = pre[w] + sCnt[w];
      const int tot = pre[8];
      const int lo = (int)(((long long)tot * warp) >> 3), hi = (int)(((long long)tot * (warp + 1)) >> 3);
      uint16_t* q = sQueue + warp * kQueueCap;
      const uint8_t* sInB = reinterpret_cast<const uint8_t*>(sIn);
      int head = 0, tail = 0;
      auto score_round = [&](int n) {  // the first n queued candidates, one per lane
        if (lane < n) {
          const int code = q[(head + lane) & (kQueueCap - 1)];
          const int r = code >> 8, x = code & 255;
          const int bam = fast_bam_packed(sInB + (r + 3) * (kF2InWords * 4) + 16 + x);
          if (bam > g.minTh) sOut[r * kF2W + x] = (uint8_t)(bam - 1);
        }
      };
      for (int e0 = lo; e0 < hi; e0 += 32) {
        const int e = e0 + lane;
        uint32_t flags = 0;
        int rbase = 0, xbase = 0;
        if (e < hi) {
          int w = 0, wbase = 0;
#pragma unroll
          for (int k = 1; k < 8; k++)
            if (e >= pre[k]) { w = k; wbase = pre[k]; }
          const uint32_t entry = sList[w * kListPerWarp + (e - wbase)];
          flags = entry & 0xC0C0C0C0u;
          rbase = (w * 8 + (int)(entry & 7)) << 8;
          xbase = (int)((entry >> 8) & 31) * 8;
        }
        // exclusive scan of the per-lane candidate counts
        const int c = __popc(flags);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += v;
        }
        int pos = tail + inc - c;
        tail += __shfl_sync(0xffffffffu, inc, 31);
        while (flags) {
          const int bit = __ffs(flags) - 1;
          flags &= flags - 1;
          const int xo = (bit >> 3) + (((bit & 7) == 6) ? 4 : 0);
          q[pos & (kQueueCap - 1)] = (uint16_t)(rbase | (xbase + xo));
          pos++;
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
    fence_proxy_async();
    __syncthreads();  // every warp is done with sIn[stage] and with its band of sOut
    if (tid == 0) tma_store_3d(&tm.out[level], sOut, x0 / 4, y0, f);
  }
  if (tid == 0) tma_store_wait_all();
}

int launch_fast_score_v2(const OrbGeo& g, const TmapPack& tm, int nFrames, int numSMs, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PGB_CUDA(cudaFuncSetAttribute(k_fast_score_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, kF2Smem));
    configured = true;
  }
  const int total = g.totalTiles2 * nFrames;
  const int grid = std::min(total, numSMs * 3);
  if (grid <= 0) return PGB_OK;
  k_fast_score_v2<<<grid, kF2Threads, kF2Smem, st>>>(g, tm, nFrames);
  PGB_LAUNCHED();
  return PGB_OK;
}

}  // namespace pgb
