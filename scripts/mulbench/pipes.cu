// Integer-multiply pipe rates on sm_100a, one instruction form per kernel, operands that change every
// iteration (each accumulator feeds its own multiplicand) so that ptxas can neither hoist the product out of
// the loop nor split the multiply-add: check with  cuobjdump -sass pipes | grep -c IMAD.WIDE .
// Round 1's "plain IMAD.WIDE" benchmark multiplied two loop-invariant registers; ptxas hoisted the product and
// the loop measured 64-bit ADDS (IADD3 + IADD3.X), which is where its 64 lanes/clk/SM came from.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int WHICH>
__global__ void __launch_bounds__(256) pipe(uint32_t* sink, int iters, uint32_t a0) {
  uint64_t acc[16];
  uint32_t y = (a0 * 2654435761u) | 1u;
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = ((uint64_t)(threadIdx.x + 1) << 20) + j * 977u + a0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
        if (WHICH == 0) {         // IMAD.WIDE.U32 Rd, Ra, Rb, Rc  (32x32 + 64 -> 64)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(lo), "r"(y));
        } else if (WHICH == 1) {  // IMAD.WIDE.U32 Rd, Ra, Rb, RZ  (32x32 -> 64, no addend)
          const uint32_t other = (uint32_t)(acc[(j + 1) & 15] >> 32);
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"(lo), "r"(other));
        } else if (WHICH == 2) {  // IMAD (32x32 + 32 -> low 32)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(y), "r"(hi));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else if (WHICH == 3) {  // IMAD.HI.U32
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(y), "r"(hi));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else if (WHICH == 4) {  // the pair ptxas fuses into one IMAD.WIDE.U32 with a 64-bit addend
          asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;\n\tmadc.hi.u32 %1, %1, %2, %1;" : "+r"(lo), "+r"(hi) : "r"(y));
          acc[j] = ((uint64_t)hi << 32) | lo;
        } else {                   // 64-bit add (IADD3 + IADD3.X): what round 1's benchmark really timed
          asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(y), "r"(lo));
          acc[j] = ((uint64_t)hi << 32) | lo;
        }
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= acc[j];
  if (s == 0x12345678ull) sink[0] = (uint32_t)s;
}

// carry chains: 8 x (mad.lo.cc / madc.hi.cc) = 4 IMAD.WIDE.U32.X per chain, as in the Montgomery product
__global__ void __launch_bounds__(256) pipe_chain(uint32_t* sink, int iters, uint32_t a0) {
  uint32_t c[2][9];
  uint32_t x = a0 | 1u, y = (a0 * 2654435761u) | 1u;
#pragma unroll
  for (int j = 0; j < 9; ++j) c[0][j] = c[1][j] = threadIdx.x + j + a0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t* C = c[s];
        asm volatile(
            "mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]), "+r"(C[4]), "+r"(C[5]), "+r"(C[6]), "+r"(C[7]), "+r"(C[8])
            : "r"(x), "r"(y), "r"(x + 2), "r"(y + 2), "r"(x + 4 + s));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 9; ++j) s ^= c[0][j] ^ c[1][j];
  if (s == 0x12345678u) sink[0] = s;
}

template <class K>
static double run(K kern, int blocks, int iters, double per_iter, uint32_t* sink, float* ms_out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * 256 * iters * per_iter / (ms * 1e-3);
    if (rate > best) { best = rate; *ms_out = ms; }
  }
  return best;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  uint32_t* sink; cudaMalloc(&sink, 64);
  const int sms = pr.multiProcessorCount, blocks = sms * 8, iters = 4096;
  const double ghz = clk_khz / 1e6;
  const char* names[] = {"IMAD.WIDE.U32 Rd,Ra,Rb,Rc (mad.wide.u32, 64-bit addend)", "IMAD.WIDE.U32 Rd,Ra,Rb,RZ (mul.wide.u32)",
                         "IMAD (mad.lo.u32)", "IMAD.HI.U32 (mad.hi.u32)", "mad.lo.cc + madc.hi pair -> IMAD.WIDE.U32 64-bit addend",
                         "64-bit add (IADD3 + IADD3.X)"};
  printf("%s, %d SMs, max clock %.3f GHz\n", pr.name, sms, ghz);
  float ms;
  double r[7];
  r[0] = run(pipe<0>, blocks, iters, 64, sink, &ms);
  r[1] = run(pipe<1>, blocks, iters, 64, sink, &ms);
  r[2] = run(pipe<2>, blocks, iters, 64, sink, &ms);
  r[3] = run(pipe<3>, blocks, iters, 64, sink, &ms);
  r[4] = run(pipe<4>, blocks, iters, 64, sink, &ms);
  r[5] = run(pipe<5>, blocks, iters, 64, sink, &ms);
  for (int i = 0; i < 6; ++i)
    printf("%-62s %7.2f T/s = %5.1f lanes/clk/SM = %.2f cycles per warp instruction per SMSP\n", names[i], r[i] / 1e12,
           r[i] / sms / (ghz * 1e9), 32.0 * 4 / (r[i] / sms / (ghz * 1e9)));
  r[6] = run(pipe_chain, blocks, iters / 4, 64, sink, &ms);
  printf("%-62s %7.2f T/s = %5.1f lanes/clk/SM = %.2f cycles per warp instruction per SMSP\n",
         "IMAD.WIDE.U32.X carry chains (mad.lo.cc / madc.hi.cc)", r[6] / 1e12, r[6] / sms / (ghz * 1e9),
         32.0 * 4 / (r[6] / sms / (ghz * 1e9)));
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
