// Register-only throughput of the FP64-assisted Montgomery product (csrc/field_dfma.cuh) against the integer one
// (csrc/field.cuh), and a device-side bit-exactness check of the two on random and edge operands.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dfma_bench dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field_dfma.cuh"
using namespace h2b;

template <class F, int CH, int V>
__global__ void __launch_bounds__(256) k_mul(uint32_t* sink, int iters, uint32_t seed) {
  F a[CH], b;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 8; ++j) a[c].v[j] = F::one().v[j] ^ (threadIdx.x * 131 + c * 7 + seed) & 0x0fffffff;
  for (int j = 0; j < 8; ++j) b.v[j] = F::r2().v[j] ^ (blockIdx.x + seed) & 0x0fffffff;
  F bw = from_mont(b), bs = shoup_companion(b);
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int c = 0; c < CH; ++c) a[c] = V == 5 ? mul_shoup(a[c], bw, bs) : V == 0 ? mul(a[c], b) : V == 1 ? mul_dfma(a[c], b) : V == 2 ? sqr_dfma(a[c]) : V == 4 ? sqr(a[c]) : mul_dfma(a[c], a[(c + 1) % CH]);
  uint32_t s = 0;
  for (int c = 0; c < CH; ++c) for (int j = 0; j < 8; ++j) s ^= a[c].v[j];
  if (s == 0x12345678u) sink[0] = s;
}

// radix-2 butterfly chain (a, c) -> (a + w c, a - w c): the transform's inner operation, Montgomery or Shoup product
template <class F, int V>
__global__ void __launch_bounds__(256) k_bfly(uint32_t* sink, int iters, uint32_t seed) {
  F a, c, b;
  for (int j = 0; j < 8; ++j) { a.v[j] = F::one().v[j] ^ (threadIdx.x * 131 + seed) & 0x0fffffff; c.v[j] = F::r2().v[j] ^ (threadIdx.x * 17 + seed) & 0x0fffffff; }
  for (int j = 0; j < 8; ++j) b.v[j] = F::r2().v[j] ^ (blockIdx.x + seed) & 0x0fffffff;
  F bw = from_mont(b), bs = shoup_companion(b);
  for (int i = 0; i < iters; ++i) {
    F t = V == 0 ? mul(c, b) : mul_shoup(c, bw, bs);
    F na = add(a, t);
    c = sub(a, t);
    a = na;
  }
  uint32_t s = 0;
  for (int j = 0; j < 8; ++j) s ^= a.v[j] ^ c.v[j];
  if (s == 0x12345678u) sink[0] = s;
}

// out[i] = number of differing words between mul and mul_dfma (and sqr) on operand pair i
template <class F>
__global__ void k_check(const uint32_t* a, const uint32_t* b, uint32_t* bad, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x, y;
  for (int j = 0; j < 8; ++j) { x.v[j] = a[i * 8 + j]; y.v[j] = b[i * 8 + j]; }
  for (int k = 0; k < 4; ++k) { reduce_once(x); reduce_once(y); }
  F r0 = mul(x, y), r1 = mul_dfma(x, y), s0 = mul(x, x), s1 = sqr_dfma(x), s2 = sqr(x);
  F yy = y; if (yy.is_zero()) yy = F::one();
  F sh = mul_shoup(x, from_mont(yy), shoup_companion(yy)), sh0 = mul(x, yy);
  uint32_t d = 0;
  for (int j = 0; j < 8; ++j) d += (r0.v[j] != r1.v[j]) + (s0.v[j] != s1.v[j]) + (s0.v[j] != s2.v[j]) + (sh.v[j] != sh0.v[j]);
  if (d) atomicAdd(bad, 1u);
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

template <class F>
static int check(const char* name) {
  const int n = 1 << 20;
  uint32_t *ha = new uint32_t[n * 8], *hb = new uint32_t[n * 8];
  uint64_t s = 88172645463325252ull;
  auto nx = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
  for (int i = 0; i < n * 8; ++i) { ha[i] = nx(); hb[i] = nx(); }
  for (int i = 0; i < n; ++i) {
    ha[i * 8 + 7] &= 0x3fffffff; hb[i * 8 + 7] &= 0x3fffffff;
    if (i < 64) for (int j = 0; j < 8; ++j) ha[i * 8 + j] = (i & 1) ? 0xffffffffu >> (j == 7 ? 2 : 0) : (i & 2 ? 0u : F::one().v[j]);
    if (i < 64 && (i & 4)) for (int j = 0; j < 8; ++j) hb[i * 8 + j] = ha[i * 8 + j];
    if (i >= 64 && i < 128) for (int j = 0; j < 8; ++j) ha[i * 8 + j] = j == (i & 7) ? 0xffffffffu >> (j == 7 ? 2 : 0) : 0u;  // single full limb
    if (i >= 128 && i < 192) for (int j = 0; j < 8; ++j) { uint32_t m = F().v[0]; (void)m; ha[i * 8 + j] = j == 0 ? (uint32_t)(i - 128) : 0u; }
  }
  uint32_t *da, *db, *dbad;
  cudaMalloc(&da, n * 32); cudaMalloc(&db, n * 32); cudaMalloc(&dbad, 4);
  cudaMemcpy(da, ha, n * 32, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 32, cudaMemcpyHostToDevice);
  cudaMemset(dbad, 0, 4);
  k_check<F><<<(n + 127) / 128, 128>>>(da, db, dbad, n);
  uint32_t bad = 1;
  cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost);
  printf("%s: mul_dfma / sqr_dfma / sqr / mul_shoup vs mul on %d operand pairs: %u mismatches\n", name, n, bad);
  return bad != 0;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  uint32_t* sink; cudaMalloc(&sink, 64);
  int rc = check<Fr>("Fr") + check<Fq>("Fq");
  const int sms = pr.multiProcessorCount;
  for (int bps : {2, 4, 8}) {
    const int blocks = sms * bps;
    printf("blocks/SM requested %d (256 threads)\n", bps);
    printf("  integer 8x32 mul          CH=1: %.2f G mulmod/s\n", run(k_mul<Fq, 1, 0>, blocks, 2048, 1, sink) / 1e9);
    printf("  integer 8x32 mul          CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 0>, blocks, 1024, 2, sink) / 1e9);
    printf("  dfma mul (b unpack hoisted) CH=1: %.2f G mulmod/s\n", run(k_mul<Fq, 1, 1>, blocks, 2048, 1, sink) / 1e9);
    printf("  dfma mul (b unpack hoisted) CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 1>, blocks, 1024, 2, sink) / 1e9);
    printf("  dfma mul (b unpack hoisted) CH=4: %.2f G mulmod/s\n", run(k_mul<Fq, 4, 1>, blocks, 512, 4, sink) / 1e9);
    printf("  dfma mul (both unpacked)    CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 3>, blocks, 1024, 2, sink) / 1e9);
    printf("  dfma mul (both unpacked)    CH=4: %.2f G mulmod/s\n", run(k_mul<Fq, 4, 3>, blocks, 512, 4, sink) / 1e9);
    printf("  dfma sqr                    CH=1: %.2f G mulmod/s\n", run(k_mul<Fq, 1, 2>, blocks, 2048, 1, sink) / 1e9);
    printf("  dfma sqr                    CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 2>, blocks, 1024, 2, sink) / 1e9);
    printf("  shoup mul (fixed multiplier) CH=1: %.2f G mulmod/s\n", run(k_mul<Fq, 1, 5>, blocks, 2048, 1, sink) / 1e9);
    printf("  shoup mul (fixed multiplier) CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 5>, blocks, 1024, 2, sink) / 1e9);
    printf("  butterfly, Montgomery product: %.2f G butterflies/s\n", run(k_bfly<Fq, 0>, blocks, 2048, 1, sink) / 1e9);
    printf("  butterfly, Shoup product:      %.2f G butterflies/s\n", run(k_bfly<Fq, 1>, blocks, 2048, 1, sink) / 1e9);
    printf("  integer sqr (36 + 72 wide)  CH=1: %.2f G mulmod/s\n", run(k_mul<Fq, 1, 4>, blocks, 2048, 1, sink) / 1e9);
    printf("  integer sqr (36 + 72 wide)  CH=2: %.2f G mulmod/s\n", run(k_mul<Fq, 2, 4>, blocks, 1024, 2, sink) / 1e9);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return rc || e != cudaSuccess;
}
