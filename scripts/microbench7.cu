// FP64 pipe on B200: DFMA rate, and whether it overlaps with IMAD.WIDE / IADD3 (exploration).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
// V: 0 = 16 DFMA; 1 = 16 DFMA + 16 IMAD.WIDE; 2 = 16 DFMA + 16 IADD3; 3 = 16 DFMA + 16 IMAD.WIDE carry-chain; 4 = 16 IMAD.WIDE; 5 = 8 DFMA + 16 IMAD.WIDE
template <int V>
__global__ void __launch_bounds__(256) k(double* sink, uint32_t iters, uint32_t a0) {
  double x = 1.0 + a0 * 1e-9, y = 0.5 + a0 * 1e-10;
  double d[16];
  uint64_t acc[16];
  uint32_t u[16];
  for (int j = 0; j < 16; ++j) { d[j] = threadIdx.x + j; acc[j] = ((uint64_t)threadIdx.x << 20) + j; u[j] = threadIdx.x + j; }
  uint32_t xi = a0 | 1u, yi = (a0 * 2654435761u) | 1u;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (V == 0 || V == 1 || V == 2 || V == 3 || (V == 5 && (j & 1))) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(x), "d"(y));
        if (V == 1 || V == 4 || V == 5) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(xi), "r"(yi));
        if (V == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[j]) : "r"(yi));
        if (V == 3) { uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
                      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(xi), "r"(yi)); acc[j] = ((uint64_t)hi << 32) | lo; }
      }
  }
  double s = 0; uint64_t t = 0;
  for (int j = 0; j < 16; ++j) { s += d[j]; t ^= acc[j] ^ u[j]; }
  if (s == 1.2345 || t == 0x123456789abcdefull) sink[0] = s + (double)t;
}
template <int V>
void run(const char* name) {
  double* sink; cudaMalloc(&sink, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 4096; double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<V><<<148 * 8, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double slots = 8.0 * 8 / 4 * iters * 64;
  printf("%-40s %8.3f ms  -> %.2f SMSP-cycles per warp-slot\n", name, best, best * 1e-3 * 1.965e9 / slots);
}
int main() {
  run<0>("DFMA"); run<4>("IMAD.WIDE"); run<1>("DFMA + IMAD.WIDE"); run<2>("DFMA + IADD3"); run<3>("DFMA + IMAD.WIDE(carry)"); run<5>("0.5 DFMA + IMAD.WIDE");
  return 0;
}
