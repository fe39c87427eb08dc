// Issue-rate probe for the integer/video ops the FAST-9 kernel is built from (sm_100a).
// Prints warp-instructions per clock per SM for each op and for the mixes the kernel uses.
// usage: pipe_probe        (results feed DESIGN.md section 4; not part of the product)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

constexpr int kIters = 512, kAcc = 8;

template <int OP>
__device__ __forceinline__ void step(unsigned (&a)[kAcc], unsigned b, unsigned c) {
#pragma unroll
  for (int i = 0; i < kAcc; i++) {
    if (OP == 0) a[i] = __vabsdiffu4(a[i], b);
    if (OP == 1) a[i] = __vimin3_s16x2(a[i], b, c);
    if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
    if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    if (OP == 5) a[i] = __byte_perm(a[i], b, 0x4321);
    if (OP == 6) asm volatile("shr.u32 %0, %0, 2;" : "+r"(a[i]));
    if (OP == 7) { a[i] = __vabsdiffu4(a[i], b); i++; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 8) { a[i] = __vimin3_s16x2(a[i], b, c); i++; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 9) { a[i] = __vimin3_s16x2(a[i], b, c); i++; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 10) { a[i] = __vabsdiffu4(a[i], b); i++; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 11) { a[i] = __vimin3_s16x2(a[i], b, c); i++; a[i] = __vabsdiffu4(a[i], b); }
    if (OP == 12) a[i] = __vimax3_u16x2(a[i], b, c);
    if (OP == 13) a[i] = __popc(a[i]) + b;
    if (OP == 15) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    if (OP == 16) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); i++; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 17) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); i++; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 18) { asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); i++; a[i] = __vimin3_s16x2(a[i], b, c); }
    if (OP == 19) { float f = __uint_as_float(a[i]); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); a[i] = __float_as_uint(f); }
    if (OP == 20) { float f = __uint_as_float(a[i]); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); a[i] = __float_as_uint(f); i++; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 21) a[i] = a[i] + a[(i + 1) % kAcc] + b;
    if (OP == 22) a[i] = __funnelshift_r(a[i], a[(i + 1) % kAcc], 3);
    if (OP == 23) a[i] = __brev(a[i]) ^ b;
    if (OP == 24) a[i] = __ffs(a[i]) + b;
    if (OP == 25) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
    if (OP == 26) { a[i] = a[i] + a[(i + 1) % kAcc] + b; i++; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 27) { a[i] = a[i] + a[(i + 1) % kAcc] + b; i++; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
    if (OP == 28) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    if (OP == 29) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(a[i]) : "r"(b), "r"(c));
    if (OP == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); i++; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
  }
}

template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned b, unsigned c, long long* cyc) {
  unsigned a[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; i++) a[i] = threadIdx.x * 2654435761u + i;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < kIters; it++) step<OP>(a, b, c);
  __syncthreads();
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < kAcc; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, unsigned* out, long long* cyc, int sms) {
  k<OP><<<sms, 1024>>>(out, 0x01020304u, 0x7f7f7f7fu, cyc);
  CK(cudaDeviceSynchronize());
  k<OP><<<sms, 1024>>>(out, 0x01020304u, 0x7f7f7f7fu, cyc);
  CK(cudaDeviceSynchronize());
  long long h[256];
  CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += h[i];
  avg /= sms;
  const double winst = 32.0 * kIters * kAcc;  // warp instructions per SM (32 warps)
  printf("%-28s %8.0f cycles  %.3f warp-inst/clk/SM  (%.2f per SMSP)\n", name, avg, winst / avg, winst / avg / 4);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  unsigned* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
  CK(cudaMalloc(&cyc, 256 * 8));
  printf("%s, %d SMs\n", p.name, sms);
  run<0>("VABSDIFF4.U8", out, cyc, sms);
  run<1>("VIMNMX3.S16x2", out, cyc, sms);
  run<12>("VIMNMX3.U16x2", out, cyc, sms);
  run<2>("LOP3", out, cyc, sms);
  run<3>("IMAD", out, cyc, sms);
  run<4>("IADD", out, cyc, sms);
  run<5>("PRMT", out, cyc, sms);
  run<6>("SHF", out, cyc, sms);
  run<13>("POPC+IADD", out, cyc, sms);
  run<14>("LOP3 + IMAD", out, cyc, sms);
  run<7>("VABSDIFF4 + IMAD", out, cyc, sms);
  run<8>("VIMNMX3 + IMAD", out, cyc, sms);
  run<9>("VIMNMX3 + LOP3", out, cyc, sms);
  run<10>("VABSDIFF4 + LOP3", out, cyc, sms);
  run<11>("VIMNMX3 + VABSDIFF4", out, cyc, sms);
  run<15>("HMNMX2", out, cyc, sms);
  run<16>("HMNMX2 + LOP3", out, cyc, sms);
  run<17>("HMNMX2 + IMAD", out, cyc, sms);
  run<18>("HMNMX2 + VIMNMX3", out, cyc, sms);
  run<19>("FMNMX3", out, cyc, sms);
  run<20>("FMNMX3 + LOP3", out, cyc, sms);
  run<21>("IADD3 (a+a'+b)", out, cyc, sms);
  run<26>("IADD3 + IMAD", out, cyc, sms);
  run<27>("IADD3 + LOP3", out, cyc, sms);
  run<22>("SHF (funnel)", out, cyc, sms);
  run<23>("BREV+LOP3", out, cyc, sms);
  run<24>("FFS(BREV+FLO)+IADD", out, cyc, sms);
  run<25>("SHFL.BFLY", out, cyc, sms);
  run<28>("IMAD.HI", out, cyc, sms);
  run<29>("VIADDMNMX", out, cyc, sms);
  return 0;
}
