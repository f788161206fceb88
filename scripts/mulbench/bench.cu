// Register-only throughput of the modular multipliers: the 8 x 32 carry-chain product of csrc/field.cuh against
// the carry-free 9 x 29 product of mul29.cuh.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mulbench bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../halo2-pse_b200/csrc/field.cuh"
#include "mul29.cuh"
using namespace h2b;

template <int CH>
__global__ void __launch_bounds__(256) k_v0(uint32_t* sink, int iters, uint32_t seed) {
  Fq a[CH], b;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 8; ++j) a[c].v[j] = FqParams::one(j) ^ (threadIdx.x * 131 + c * 7 + seed) & 0x0fffffff;
  for (int j = 0; j < 8; ++j) b.v[j] = FqParams::r2(j) ^ (blockIdx.x + seed) & 0x0fffffff;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = mul(a[c], b);
  uint32_t s = 0;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 8; ++j) s ^= a[c].v[j];
  if (s == 0x12345678u) sink[0] = s;
}

template <int CH, int VARIANT>
__global__ void __launch_bounds__(256) k_v3(uint32_t* sink, int iters, uint32_t seed) {
  L29<Fq29> a[CH], b;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 9; ++j) a[c].l[j] = (Fq29::ONE(j) ^ (threadIdx.x * 131 + c * 7 + seed)) & MASK29;
  for (int j = 0; j < 9; ++j) b.l[j] = (Fq29::UP(j) ^ (blockIdx.x + seed)) & MASK29;
  for (int c = 0; c < CH; ++c) a[c].l[8] &= 0xffffff;
  b.l[8] &= 0xffffff;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = mul29<Fq29, VARIANT>(a[c], b);
  uint32_t s = 0;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 9; ++j) s ^= a[c].l[j];
  if (s == 0x12345678u) sink[0] = s;
}

// correctness on the device: out[i] = mul29(a[i], b[i]) for the host to compare
__global__ void k_check(const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  L29<Fq29> x, y;
  for (int j = 0; j < 9; ++j) { x.l[j] = a[i * 9 + j]; y.l[j] = b[i * 9 + j]; }
  L29<Fq29> z = mul29<Fq29, 1>(x, y);
  L29<Fq29> w = mul29<Fq29, 0>(x, y);
  for (int j = 0; j < 9; ++j) out[i * 9 + j] = z.l[j] ^ (z.l[j] ^ w.l[j]);  // == w; differs only if the variants disagree
}

template <class K>
static double run(K kern, int blocks, int iters, double muls_per_iter, uint32_t* sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(sink, iters, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * 256 * iters * muls_per_iter / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  return best;
}

int main() {
  int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  uint32_t* sink; cudaMalloc(&sink, 64);
  const int sms = pr.multiProcessorCount;
  // device correctness vs the host path of the same template
  {
    const int n = 4096;
    uint32_t *ha = new uint32_t[n * 9], *hb = new uint32_t[n * 9], *ho = new uint32_t[n * 9];
    uint64_t s = 88172645463325252ull;
    auto nx = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (int i = 0; i < n * 9; ++i) { ha[i] = nx() & MASK29; hb[i] = nx() & ((1u << 30) - 1); }
    for (int i = 0; i < n; ++i) { ha[i * 9 + 8] &= 0x3ffffff; hb[i * 9 + 8] &= 0x3ffffff; }
    uint32_t *da, *db, *dout;
    cudaMalloc(&da, n * 36); cudaMalloc(&db, n * 36); cudaMalloc(&dout, n * 36);
    cudaMemcpy(da, ha, n * 36, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 36, cudaMemcpyHostToDevice);
    k_check<<<(n + 127) / 128, 128>>>(da, db, dout, n);
    cudaMemcpy(ho, dout, n * 36, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) {
      L29<Fq29> x, y;
      for (int j = 0; j < 9; ++j) { x.l[j] = ha[i * 9 + j]; y.l[j] = hb[i * 9 + j]; }
      L29<Fq29> z = mul29<Fq29, 0>(x, y);
      for (int j = 0; j < 9; ++j) bad += z.l[j] != ho[i * 9 + j];
    }
    printf("device mul29 vs host mul29 (checked against big integers in scripts/mulbench/host_test): %s\n", bad ? "MISMATCH" : "ok");
  }
  for (int bps : {4, 8}) {
    const int blocks = sms * bps;
    printf("blocks/SM requested %d\n", bps);
    printf("  v0 8x32 carry chains  CH=1: %.2f G mulmod/s\n", run(k_v0<1>, blocks, 2048, 1, sink) / 1e9);
    printf("  v0 8x32 carry chains  CH=2: %.2f G mulmod/s\n", run(k_v0<2>, blocks, 1024, 2, sink) / 1e9);
    printf("  v3 9x29 carry-free(a) CH=1: %.2f G mulmod/s\n", run(k_v3<1, 0>, blocks, 2048, 1, sink) / 1e9);
    printf("  v3 9x29 carry-free(a) CH=2: %.2f G mulmod/s\n", run(k_v3<2, 0>, blocks, 1024, 2, sink) / 1e9);
    printf("  v3 9x29 carry-free(b) CH=1: %.2f G mulmod/s\n", run(k_v3<1, 1>, blocks, 2048, 1, sink) / 1e9);
    printf("  v3 9x29 carry-free(b) CH=2: %.2f G mulmod/s\n", run(k_v3<2, 1>, blocks, 1024, 2, sink) / 1e9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
