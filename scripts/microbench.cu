// Exploration microbenchmarks for the integer pipe (not product code, not a test).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/microbench scripts/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../halo2-pse_b200/csrc/field.cuh"
using namespace h2b;

template <int WHICH>
__global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t iters, uint32_t a0) {
  uint32_t x[8];
  for (int j = 0; j < 8; ++j) x[j] = (a0 * (2 * j + 3)) | 1u;
  if (WHICH == 0 || WHICH == 1 || WHICH == 2 || WHICH == 5) {
    uint64_t acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = ((uint64_t)threadIdx.x << 20) + j * 977u + a0;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t lo = (uint32_t)acc[j], hi = (uint32_t)(acc[j] >> 32);
          if (WHICH == 0) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(x[0]), "r"(x[1])); acc[j] = ((uint64_t)hi << 32) | lo; }
          else if (WHICH == 1) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(x[0]), "r"(x[1])); acc[j] = ((uint64_t)hi << 32) | lo; }
          else if (WHICH == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(x[0]), "r"(x[1]));
          else asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(x[j & 7]), "r"(x[(j * 3 + rep) & 7]));
        }
    }
    uint64_t s = 0;
    for (int j = 0; j < 16; ++j) s ^= acc[j];
    if (s == 0x123456789abcdefull) sink[0] = (uint32_t)s;
  } else if (WHICH == 3 || WHICH == 6) {
    // carry chains: 4 independent chains of 8 (mad.lo.cc, madc.hi.cc ...) per rep
    uint32_t c[4][9];
    for (int q = 0; q < 4; ++q) for (int j = 0; j < 9; ++j) c[q][j] = threadIdx.x + j + a0 + q;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (WHICH == 3)
            chain_mad_top<false>(c[q][0], c[q][1], c[q][2], c[q][3], c[q][4], c[q][5], c[q][6], c[q][7], c[q][8],
                                 x[0], x[2], x[4], x[6], x[q], 0u, 0u);
          else
            chain_mad(c[q][0], c[q][1], c[q][2], c[q][3], c[q][4], c[q][5], c[q][6], c[q][7],
                                 x[0], x[2], x[4], x[6], x[q]);
        }
    }
    uint32_t s = 0;
    for (int q = 0; q < 4; ++q) for (int j = 0; j < 9; ++j) s ^= c[q][j];
    if (s == 0x12345678u) sink[0] = s;
  } else if (WHICH == 4 || WHICH == 7) {
    constexpr int NCH = WHICH == 4 ? 2 : 4;
    Fr a[NCH], b;
    for (int q = 0; q < NCH; ++q) for (int j = 0; j < 8; ++j) a[q].v[j] = threadIdx.x * (q + 1) + j + a0;
    for (int j = 0; j < 8; ++j) b.v[j] = FrParams::one(j) ^ (a0 & 0xff);
    for (int q = 0; q < NCH; ++q) a[q].v[7] &= 0x0fffffffu;
    b.v[7] &= 0x0fffffffu;
    for (uint32_t it = 0; it < iters; ++it)
#pragma unroll
      for (int q = 0; q < NCH; ++q) a[q] = mul(a[q], b);
    uint32_t s = 0;
    for (int q = 0; q < NCH; ++q) for (int j = 0; j < 8; ++j) s ^= a[q].v[j];
    if (s == 0x12345678u) sink[0] = s;
  }
}

template <int W>
void run(const char* name, double per_iter, uint32_t iters, int blocks_per_sm, int threads) {
  uint32_t* sink; cudaMalloc(&sink, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<W><<<148 * blocks_per_sm, threads>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double r = 148.0 * blocks_per_sm * threads * iters * per_iter / (ms * 1e-3);
    if (r > best) best = r;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<W>);
  printf("%-34s blocks/SM %d thr %d regs %d: %8.3f T/s  (%.2f per clk per SM @1.965GHz)\n", name, blocks_per_sm, threads, fa.numRegs, best / 1e12, best / 148 / 1.965e9);
  cudaFree(sink);
}

int main() {
  run<0>("IMAD lo (mad.lo.u32)", 64, 4096, 8, 256);
  run<1>("IMAD.HI (mad.hi.u32)", 64, 4096, 8, 256);
  run<2>("IMAD.WIDE invariant operands", 64, 4096, 8, 256);
  run<5>("IMAD.WIDE varying operands", 64, 4096, 8, 256);
  run<3>("carry chain w/ top (9 instr)", 2 * 4 * 8, 4096, 8, 256);
  run<6>("carry chain (8 instr)", 2 * 4 * 8, 4096, 8, 256);
  for (int b : {1, 2, 4, 8}) run<4>("fr mul x2 chains (mulmods)", 2, 2048, b, 256);
  for (int b : {1, 2, 4}) run<7>("fr mul x4 chains (mulmods)", 4, 2048, b, 256);
  return 0;
}
