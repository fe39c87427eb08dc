// Device helpers shared by the two FAST-9 kernels (fast_score.cu: score map in HBM, the stage-by-stage / debug path;
// fast_cells.cu: the fused score + per-cell NMS kernel of the hot path): mbarrier / TMA wrappers and the exact
// corner score.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace pgb {
namespace fastk {

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
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int x, int y, int z, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(x),
               "r"(y), "r"(z), "r"(smem_u32(src))
               : "memory");
}

// a * one + c with `one` = 1 opaque to the compiler: the addition is issued as IMAD on the FMA pipe
__device__ __forceinline__ uint32_t mad1(uint32_t a, uint32_t one, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
  return d;
}

// Exact bam (largest threshold at which the pixel is a FAST-9 corner, + 1) of the pixel at byte pointer c inside a staged
// tile whose rows are kRowB bytes apart.  Ring pixel p becomes (p, -p) in the two s16 halves of a register with one
// IMAD; the circular 9-wide sliding MAX is two rounds of VIMNMX3.S16x2, the MIN over the 16 arcs a 3-input tree;
// dark = v - min_arcs(max_arc p), bright = max_arcs(min_arc p) - v.
// (Tried: encoding p as fp16-compatible halves so that part of the min/max tree runs as HMNMX2 on the FMA pipes;
// ptxas fuses the pairs into 3-input VHMNMX on the ALU pipe again, and splitting them costs issue slots -- no gain.)
template <int kRowB>
__device__ __forceinline__ int fast_bam_minmax(const uint8_t* c) {
  const int v = c[0];
  const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const uint32_t p = c[dy[k] * kRowB + dx[k]];
    asm("mul.lo.u32 %0, %1, 0xFFFF0001;" : "=r"(w[k]) : "r"(p));  // lo16 = p, hi16 = -p
  }
  uint32_t t3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) t3[k] = __vimax3_s16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
  uint32_t m9[16];  // lo: max of p over the arc starting at k; hi: -(min of p over the arc)
#pragma unroll
  for (int k = 0; k < 16; k++) m9[k] = __vimax3_s16x2(t3[k], t3[(k + 3) & 15], t3[(k + 6) & 15]);
  uint32_t a = __vimin3_s16x2(m9[0], m9[1], m9[2]);
  uint32_t b = __vimin3_s16x2(m9[3], m9[4], m9[5]);
  uint32_t cc = __vimin3_s16x2(m9[6], m9[7], m9[8]);
  uint32_t d = __vimin3_s16x2(m9[9], m9[10], m9[11]);
  uint32_t e = __vimin3_s16x2(m9[12], m9[13], m9[14]);
  a = __vimin3_s16x2(a, b, cc);
  d = __vimin3_s16x2(d, e, m9[15]);
  a = __vmins2(a, d);
  const int lo = (int)(a & 0xffffu);  // min over arcs of (max p)
  const int hi = (int)a >> 16;        // -(max over arcs of (min p))
  return max(v - lo, -hi - v);
}

}  // namespace fastk
}  // namespace pgb
