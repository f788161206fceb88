// mul29 (carry-free 9 x 29-bit Montgomery product) vs mul (8 x 32 carry chains): throughput.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "field29.cuh"
using namespace h2b;

template <int NCH, int W>
__global__ void __launch_bounds__(256) k(uint32_t* sink, uint32_t iters, uint32_t a0) {
  if (W == 0) {
    Fq29 a[NCH], b;
    for (int q = 0; q < NCH; ++q) for (int j = 0; j < 9; ++j) a[q].l[j] = (threadIdx.x * (q + 1) + j + a0) & MASK29;
    for (int j = 0; j < 9; ++j) b.l[j] = (Fq29Params::one(j) ^ (threadIdx.x * 77 + a0)) & MASK29;
    for (uint32_t it = 0; it < iters; ++it)
#pragma unroll
      for (int q = 0; q < NCH; ++q) a[q] = mul29(a[q], b);
    uint32_t s = 0;
    for (int q = 0; q < NCH; ++q) for (int j = 0; j < 9; ++j) s ^= a[q].l[j];
    if (s == 0x12345678u) sink[0] = s;
  } else {
    Fq a[NCH], b;
    for (int q = 0; q < NCH; ++q) for (int j = 0; j < 8; ++j) a[q].v[j] = threadIdx.x * (q + 1) + j + a0;
    for (int j = 0; j < 8; ++j) b.v[j] = FqParams::one(j) ^ (threadIdx.x * 77 + a0);
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

template <int NCH, int W>
void run(const char* name, int bps) {
  uint32_t* sink; cudaMalloc(&sink, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint32_t iters = 2048; double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<NCH, W><<<148 * bps, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double r = 148.0 * bps * 256 * iters * NCH / (ms * 1e-3);
    if (r > best) best = r;
  }
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<NCH, W>);
  printf("%-28s chains %d blocks/SM %d regs %3d: %7.2f G mulmod/s (%.3f per clk per SM)\n", name, NCH, bps, fa.numRegs, best / 1e9, best / 148 / 1.965e9);
}

int main() {
  run<2, 1>("mul 8x32 (carry chains)", 4);
  run<1, 0>("mul29 9x29 (carry-free)", 4);
  run<2, 0>("mul29 9x29 (carry-free)", 2);
  run<2, 0>("mul29 9x29 (carry-free)", 4);
  run<2, 0>("mul29 9x29 (carry-free)", 8);
  run<4, 0>("mul29 9x29 (carry-free)", 4);
  return 0;
}
