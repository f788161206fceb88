// IMAD.WIDE carry-variant throughput (exploration; not product code).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// V=0: pair with carry-out only, carry consumed by an addc on the alu pipe
// V=1: chain of 4 pairs (carry in + out), as field.cuh
// V=2: plain mad.wide (no carry)
// V=3: carry-out only, carry dropped (mad.lo.cc + madc.hi, no consumer)
template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t iters, uint32_t a0) {
  uint32_t x[8];
  for (int j = 0; j < 8; ++j) x[j] = (a0 * (2 * j + 3)) | 1u;
  uint32_t lo[8], hi[8], cc[8];
  for (int j = 0; j < 8; ++j) { lo[j] = threadIdx.x + j; hi[j] = a0 + j; cc[j] = j; }
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (V == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                       : "+r"(lo[j]), "+r"(hi[j]), "+r"(cc[j]) : "r"(x[j]), "r"(x[(j + rep) & 7]));
      } else if (V == 1) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                       "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                       "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                       "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                       : "+r"(lo[4 * h]), "+r"(hi[4 * h]), "+r"(lo[4 * h + 1]), "+r"(hi[4 * h + 1]), "+r"(lo[4 * h + 2]),
                         "+r"(hi[4 * h + 2]), "+r"(lo[4 * h + 3]), "+r"(hi[4 * h + 3]), "+r"(cc[h])
                       : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4 + rep]));
      } else if (V == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint64_t acc = ((uint64_t)hi[j] << 32) | lo[j];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[j]), "r"(x[(j + rep) & 7]));
          lo[j] = (uint32_t)acc; hi[j] = (uint32_t)(acc >> 32);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(lo[j]), "+r"(hi[j]) : "r"(x[j]), "r"(x[(j + rep) & 7]));
      }
    }
  }
  uint32_t s = 0;
  for (int j = 0; j < 8; ++j) s ^= lo[j] ^ hi[j] ^ cc[j];
  if (s == 0x12345678u) sink[0] = s;
}

template <int V>
void run(const char* name) {
  uint32_t* sink; cudaMalloc(&sink, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 4096; double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<V><<<148 * 8, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double r = 148.0 * 8 * 256 * iters * 32.0 / (ms * 1e-3);  // 32 multiplies (IMAD.WIDE) per iteration
    if (r > best) best = r;
  }
  printf("%-44s %8.3f Tmul/s  (%.2f IMAD.WIDE per clk per SM @1.965GHz)\n", name, best / 1e12, best / 148 / 1.965e9);
}

int main() {
  run<2>("plain IMAD.WIDE");
  run<3>("IMAD.WIDE carry-out, dropped");
  run<0>("IMAD.WIDE carry-out + IADD3.X consumer");
  run<1>("IMAD.WIDE.X chain of 4 (carry in+out)");
  return 0;
}
