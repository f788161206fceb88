// Do IMAD.WIDE and ALU-pipe instructions overlap?  (exploration)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
// V: 0 = 16 IMAD.WIDE; 1 = 16 IMAD.WIDE + 16 IADD3; 2 = 16 IMAD.WIDE + 16 LOP3; 3 = 16 IADD3 only;
//    4 = 16 IMAD(lo) + 16 IADD3; 5 = 16 IMAD.WIDE + 16 SHF; 6 = 16 IMAD (lo) only; 7 = 16 IMAD.WIDE + 32 IADD3
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t iters, uint32_t a0) {
  uint32_t x = a0 | 1u, y = (a0 * 2654435761u) | 1u;
  uint64_t acc[16];
  uint32_t u[16], w[16];
  for (int j = 0; j < 16; ++j) { acc[j] = ((uint64_t)threadIdx.x << 20) + j * 977u + a0; u[j] = threadIdx.x + j; w[j] = a0 ^ j; }
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (V == 0 || V == 1 || V == 2 || V == 5 || V == 7) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(x), "r"(y));
        if (V == 4 || V == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[j]) : "r"(x), "r"(y));
        if (V == 1 || V == 3 || V == 4 || V == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[j]) : "r"(y));
        if (V == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(w[j]) : "r"(x));
        if (V == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[j]) : "r"(y));
        if (V == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[j]) : "r"(y));
      }
  }
  uint64_t s = 0;
  for (int j = 0; j < 16; ++j) s ^= acc[j] ^ u[j] ^ ((uint64_t)w[j] << 32);
  if (s == 0x123456789abcdefull) sink[0] = (uint32_t)s;
}
template <int V>
void run(const char* name) {
  uint32_t* sink; cudaMalloc(&sink, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 4096; double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<V><<<148 * 8, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  // cycles per SMSP per "slot" (one j-iteration of one warp)
  double slots = 8.0 * 8 /*warps per block*/ / 4 /*SMSP*/ * iters * 64;
  printf("%-36s %8.3f ms  -> %.2f SMSP-cycles per warp-slot\n", name, best, best * 1e-3 * 1.965e9 / slots);
}
int main() {
  run<0>("IMAD.WIDE"); run<6>("IMAD lo"); run<3>("IADD3"); run<1>("IMAD.WIDE + IADD3"); run<2>("IMAD.WIDE + LOP3");
  run<5>("IMAD.WIDE + SHF"); run<4>("IMAD lo + IADD3"); run<7>("IMAD.WIDE + 2 IADD3");
  return 0;
}
