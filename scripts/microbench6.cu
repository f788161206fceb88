// Random 64-byte gathers from a large table (the MSM's base loads): achievable rate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
struct __align__(32) V8 { uint32_t v[8]; };
__device__ __forceinline__ V8 ldnc(const V8* p) {
  V8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st(V8* p, const V8& r) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p) : "memory");
}
__global__ void mkidx(uint32_t* idx, uint64_t n, uint64_t tab) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t x = i * 0x9e3779b97f4a7c15ull + 12345; x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 32;
    idx[i] = (uint32_t)(x % tab);
  }
}
template <int HALF>
__global__ void gather(const uint32_t* idx, const V8* table, V8* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const V8* src = table + 2ull * idx[i];
    V8 a = ldnc(src);
    st(out + 2 * i, a);
    if (!HALF) { V8 b = ldnc(src + 1); st(out + 2 * i + 1, b); }
  }
}
int main() {
  const uint64_t tab = 12ull << 24, n = 201326592ull / 2;  // 12 GiB table, 100M gathers
  V8 *table, *out; uint32_t* idx;
  cudaMalloc(&table, tab * 64); cudaMalloc(&out, n * 64); cudaMalloc(&idx, n * 4);
  cudaMemset(table, 1, tab * 64);
  mkidx<<<148 * 8, 256>>>(idx, n, tab);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    for (int half = 0; half < 2; ++half) {
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (half) gather<1><<<blocks, 256>>>(idx, table, out, n); else gather<0><<<blocks, 256>>>(idx, table, out, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("blocks %5d x256 %s: %7.3f ms for %.0fM gathers -> %.1f G gathers/s, %.2f TB/s read\n", blocks, half ? "32B" : "64B", best, n / 1e6, n / best / 1e6, n * (half ? 32.0 : 64.0) / best / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
