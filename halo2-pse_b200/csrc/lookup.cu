// The lookup argument's prover-side row work on the device:
//   permute_expression_pair            halo2_proofs/src/plonk/lookup/prover.rs:390-475
//   the grand-product fractions        halo2_proofs/src/plonk/lookup/prover.rs:146-199
//
// The reference sorts the compressed input column, walks it with a BTreeMap of the table's values and a
// stack of repeated rows.  The same permutation is a data-parallel pipeline:
//   1. canonical (non-Montgomery) copies of both columns; 256-bit LSD radix sort = eight stable
//      cub::DeviceRadixSort passes over one 32-bit limb each (ascending canonical order = `Ord for Fr`);
//   2. first-of-run flags on the sorted input S and the sorted table T; a first-of-run of S must occur in
//      T (binary search, else Error::ConstraintSystemFailure); a first-of-run of T that occurs in S is the
//      one copy the BTreeMap decrements, every other table element is "left over";
//   3. exclusive scans rank the repeated rows of S and the left-over elements of T (both ascending);
//   4. permuted_table[row] = S[row] on first occurrences, else leftover[m - 1 - rank(row)]: the reference
//      pops repeated rows from the back while it walks the left-over values upwards.
#include "common.cuh"

#include <string.h>

#ifndef H2B_EMU
#include <cub/cub.cuh>
#else
#include <algorithm>
#include <numeric>
#endif

namespace h2b {

__global__ void lk_canon_kernel(const Fr* in, Fr* canon, uint32_t* perm, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    st_fp(canon + i, from_mont(ld_fp(in + i)));
    perm[i] = i;
  }
}

__global__ void lk_limb_kernel(const Fr* canon, const uint32_t* perm, uint32_t* key, uint32_t limb, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    key[i] = reinterpret_cast<const uint32_t*>(canon + perm[i])[limb];
}

// sorted[i] = src[perm[i]] for two sources at once (canonical keys and the original Montgomery values)
__global__ void lk_gather_kernel(const Fr* canon, const Fr* mont, const uint32_t* perm, Fr* canon_sorted,
                                 Fr* mont_sorted, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t p = perm[i];
    st_fp(canon_sorted + i, ld_fp(canon + p));
    st_fp(mont_sorted + i, ld_fp(mont + p));
  }
}

// a < b as 256-bit little-endian integers
H2B_D bool lk_less(const Fr& a, const Fr& b) {
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (a.v[i] != b.v[i]) return a.v[i] < b.v[i];
  }
  return false;
}

H2B_D bool lk_contains(const Fr* sorted, uint32_t n, const Fr& x) {
  uint32_t lo = 0, hi = n;  // first index with sorted[idx] >= x
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (lk_less(ld_fp(sorted + mid), x)) lo = mid + 1; else hi = mid;
  }
  return lo < n && ld_fp(sorted + lo) == x;
}

// which = 0: over S (the sorted input): flag[i] = 1 on repeated rows; a first occurrence missing from T raises *err
// which = 1: over T (the sorted table): flag[j] = 1 on left-over elements (not the first of a run that occurs in S)
__global__ void lk_flags_kernel(const Fr* self, const Fr* other, uint32_t n, uint32_t* flag, int which, int* err) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Fr x = ld_fp(self + i);
    const bool first = i == 0 || ld_fp(self + i - 1) != x;
    if (which == 0) {
      flag[i] = first ? 0u : 1u;
      if (first && !lk_contains(other, n, x)) atomicOr(err, 1);
    } else {
      flag[i] = (first && lk_contains(other, n, x)) ? 0u : 1u;
    }
  }
}

__global__ void lk_compact_kernel(const Fr* t_mont_sorted, const uint32_t* flag, const uint32_t* rank, Fr* leftover,
                                  uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (flag[i]) st_fp(leftover + rank[i], ld_fp(t_mont_sorted + i));
}

__global__ void lk_assign_kernel(const Fr* s_mont_sorted, const uint32_t* rep_flag, const uint32_t* rep_rank,
                                 const Fr* leftover, uint32_t m, Fr* out_input, Fr* out_table, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Fr s = ld_fp(s_mont_sorted + i);
    st_fp(out_input + i, s);
    st_fp(out_table + i, rep_flag[i] ? ld_fp(leftover + (m - 1 - rep_rank[i])) : s);
  }
}

// phase 0: out[i] = (beta + permuted_input[i]) * (gamma + permuted_table[i])            (prover.rs:163-175)
// phase 1: out[i] *= (compressed_input[i] + beta) * (compressed_table[i] + gamma)       (prover.rs:183-191)
__global__ void lk_product_kernel(const Fr* a, const Fr* b, Fr beta, Fr gamma, Fr* out, uint64_t n, int phase) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const Fr t = mul(add(ld_fp(a + i), beta), add(ld_fp(b + i), gamma));
    st_fp(out + i, phase == 0 ? t : mul(ld_fp(out + i), t));
  }
}

}  // namespace h2b

using namespace h2b;

namespace {
Fr load_fr(const h2b_fr* x) {
  Fr r;
  memcpy(&r, x, sizeof(Fr));
  return r;
}
uint32_t grid_for(h2b_ctx* ctx, uint64_t n, uint32_t threads) {
  const uint64_t want = (n + threads - 1) / threads, cap = (uint64_t)ctx->sm_count * 16;
  return (uint32_t)(want < cap ? want : cap);
}

struct DevMem {  // scratch of one call, released on scope exit: the context's stream-ordered pool and block cache
  h2b_ctx* ctx;    // (cudaMalloc / cudaFree would synchronise the whole device 16 times per lookup)
  std::vector<void*> ptrs;
  explicit DevMem(h2b_ctx* c) : ctx(c) {}
  ~DevMem() {
    for (void* p : ptrs) h2b_device_free(ctx, p);
  }
  template <class T>
  cudaError_t get(T** out, size_t count) {
    void* p = nullptr;
    const int rc = h2b_device_alloc(ctx, (count ? count : 1) * sizeof(T), &p);
    if (rc == H2B_OK) ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return rc == H2B_OK ? cudaSuccess : cudaErrorMemoryAllocation;
  }
};

// stable sort of perm by key (ascending), both updated in place
int sort_pass(h2b_ctx* ctx, uint32_t* key, uint32_t* key_alt, uint32_t* perm, uint32_t* perm_alt, uint32_t n,
              void* tmp, size_t tmp_bytes) {
#ifndef H2B_EMU
  cub::DoubleBuffer<uint32_t> k(key, key_alt), v(perm, perm_alt);
  H2B_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, (int)n, 0, 32, ctx->stream));
  ctx->launches += 4;
  if (v.Current() != perm)
    H2B_CUDA(ctx, cudaMemcpyAsync(perm, v.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
#else
  std::vector<uint32_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
  for (uint32_t i = 0; i < n; ++i) perm_alt[i] = perm[idx[i]];
  memcpy(perm, perm_alt, (size_t)n * 4);
  (void)key_alt, (void)tmp, (void)tmp_bytes;
#endif
  return H2B_OK;
}

int exclusive_scan(h2b_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n, void* tmp, size_t tmp_bytes) {
#ifndef H2B_EMU
  H2B_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, ctx->stream));
  ctx->launches += 2;
#else
  uint32_t acc = 0;
  for (uint32_t i = 0; i < n; ++i) {
    out[i] = acc;
    acc += in[i];
  }
  (void)tmp, (void)tmp_bytes;
#endif
  return H2B_OK;
}

// canonical + Montgomery copies of in[0..n) in ascending canonical order
int sort_column(h2b_ctx* ctx, const Fr* in, uint32_t n, Fr* canon, Fr* canon_sorted, Fr* mont_sorted, uint32_t* perm,
                uint32_t* perm_alt, uint32_t* key, uint32_t* key_alt, void* tmp, size_t tmp_bytes) {
  const uint32_t g = grid_for(ctx, n, 256);
  H2B_TRY(launch(ctx, lk_canon_kernel, dim3(g), dim3(256), 0, in, canon, perm, n));
  for (uint32_t limb = 0; limb < 8; ++limb) {
    H2B_TRY(launch(ctx, lk_limb_kernel, dim3(g), dim3(256), 0, (const Fr*)canon, (const uint32_t*)perm, key, limb, n));
    H2B_TRY(sort_pass(ctx, key, key_alt, perm, perm_alt, n, tmp, tmp_bytes));
  }
  return launch(ctx, lk_gather_kernel, dim3(g), dim3(256), 0, (const Fr*)canon, in, (const uint32_t*)perm, canon_sorted,
                mont_sorted, n);
}
}  // namespace

extern "C" int h2b_lookup_permute(h2b_ctx* ctx, const h2b_fr* input_dev, const h2b_fr* table_dev, size_t usable_rows,
                                  h2b_fr* permuted_input_dev, h2b_fr* permuted_table_dev) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (usable_rows && (!input_dev || !table_dev || !permuted_input_dev || !permuted_table_dev))
    return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (usable_rows == 0) return H2B_OK;
  if (usable_rows > (1ull << 31)) return fail(ctx, H2B_ERR_ARG, "too many rows");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t n = (uint32_t)usable_rows;
  DevMem mem(ctx);
  Fr *canon, *s_canon, *s_mont, *t_canon, *t_mont, *leftover;
  uint32_t *perm, *perm_alt, *key, *key_alt, *rep_flag, *rep_rank, *left_flag, *left_rank;
  int* d_err;
  H2B_CUDA(ctx, mem.get(&canon, n));
  H2B_CUDA(ctx, mem.get(&s_canon, n));
  H2B_CUDA(ctx, mem.get(&s_mont, n));
  H2B_CUDA(ctx, mem.get(&t_canon, n));
  H2B_CUDA(ctx, mem.get(&t_mont, n));
  H2B_CUDA(ctx, mem.get(&leftover, n));
  H2B_CUDA(ctx, mem.get(&perm, n));
  H2B_CUDA(ctx, mem.get(&perm_alt, n));
  H2B_CUDA(ctx, mem.get(&key, n));
  H2B_CUDA(ctx, mem.get(&key_alt, n));
  H2B_CUDA(ctx, mem.get(&rep_flag, n + 1));
  H2B_CUDA(ctx, mem.get(&rep_rank, n + 1));
  H2B_CUDA(ctx, mem.get(&left_flag, n + 1));
  H2B_CUDA(ctx, mem.get(&left_rank, n + 1));
  H2B_CUDA(ctx, mem.get(&d_err, 1));
  size_t tmp_bytes = 1;
  void* tmp = nullptr;
#ifndef H2B_EMU
  {
    size_t a = 0, b = 0;
    cub::DoubleBuffer<uint32_t> k(key, key_alt), v(perm, perm_alt);
    H2B_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, a, k, v, (int)n, 0, 32, ctx->stream));
    H2B_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, b, rep_flag, rep_rank, (int)(n + 1), ctx->stream));
    tmp_bytes = a > b ? a : b;
  }
#endif
  unsigned char* tmp_c;
  H2B_CUDA(ctx, mem.get(&tmp_c, tmp_bytes));
  tmp = tmp_c;
  H2B_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
  const Fr* in = reinterpret_cast<const Fr*>(input_dev);
  const Fr* tab = reinterpret_cast<const Fr*>(table_dev);
  H2B_TRY(sort_column(ctx, in, n, canon, s_canon, s_mont, perm, perm_alt, key, key_alt, tmp, tmp_bytes));
  H2B_TRY(sort_column(ctx, tab, n, canon, t_canon, t_mont, perm, perm_alt, key, key_alt, tmp, tmp_bytes));
  const uint32_t g = grid_for(ctx, n, 256);
  // flags with one extra zero slot so that the scan's last output is the total count
  H2B_CUDA(ctx, cudaMemsetAsync(rep_flag + n, 0, 4, ctx->stream));
  H2B_CUDA(ctx, cudaMemsetAsync(left_flag + n, 0, 4, ctx->stream));
  H2B_TRY(launch(ctx, lk_flags_kernel, dim3(g), dim3(256), 0, (const Fr*)s_canon, (const Fr*)t_canon, n, rep_flag, 0, d_err));
  H2B_TRY(launch(ctx, lk_flags_kernel, dim3(g), dim3(256), 0, (const Fr*)t_canon, (const Fr*)s_canon, n, left_flag, 1, d_err));
  H2B_TRY(exclusive_scan(ctx, rep_flag, rep_rank, n + 1, tmp, tmp_bytes));
  H2B_TRY(exclusive_scan(ctx, left_flag, left_rank, n + 1, tmp, tmp_bytes));
  int err = 0;
  uint32_t m_rep = 0, m_left = 0;
  H2B_CUDA(ctx, cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaMemcpyAsync(&m_rep, rep_rank + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaMemcpyAsync(&m_left, left_rank + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (err || m_rep != m_left)  // an input value that the table does not contain (prover.rs:425-431)
    return fail(ctx, H2B_ERR_CONSTRAINT, "Error::ConstraintSystemFailure: lookup input not in the table");
  H2B_TRY(launch(ctx, lk_compact_kernel, dim3(g), dim3(256), 0, (const Fr*)t_mont, (const uint32_t*)left_flag,
                 (const uint32_t*)left_rank, leftover, n));
  H2B_TRY(launch(ctx, lk_assign_kernel, dim3(g), dim3(256), 0, (const Fr*)s_mont, (const uint32_t*)rep_flag,
                 (const uint32_t*)rep_rank, (const Fr*)leftover, m_rep, reinterpret_cast<Fr*>(permuted_input_dev),
                 reinterpret_cast<Fr*>(permuted_table_dev), n));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_lookup_product_fractions(h2b_ctx* ctx, const h2b_fr* permuted_input, const h2b_fr* permuted_table,
                                            const h2b_fr* compressed_input, const h2b_fr* compressed_table,
                                            const h2b_fr* beta, const h2b_fr* gamma, size_t n, h2b_fr* out_dev) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!permuted_input || !permuted_table || !compressed_input || !compressed_table || !beta || !gamma || !out_dev)
    return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr b = load_fr(beta), g = load_fr(gamma);
  Fr* out = reinterpret_cast<Fr*>(out_dev);
  const uint32_t grid = grid_for(ctx, n, 256);
  H2B_TRY(launch(ctx, lk_product_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<const Fr*>(permuted_input),
                 reinterpret_cast<const Fr*>(permuted_table), b, g, out, (uint64_t)n, 0));
  H2B_TRY(h2b_batch_invert(ctx, out_dev, H2B_DEVICE, n));  // prover.rs:179
  H2B_TRY(launch(ctx, lk_product_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<const Fr*>(compressed_input),
                 reinterpret_cast<const Fr*>(compressed_table), b, g, out, (uint64_t)n, 1));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}
