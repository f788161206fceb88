// Device pieces of create_proof between the transforms and the commitments, so that
// a proof is produced without a polynomial ever returning to the host:
//   Fr::random / from_bytes_wide (512-bit little-endian integer mod r)    halo2curves 0.3.1 bn256/fr.rs
//       used for blinding rows and the vanishing argument's random polynomial
//                                                halo2_proofs/src/plonk/vanishing/prover.rs:49-53
//   the permutation argument's per-row fractions   halo2_proofs/src/plonk/permutation/prover.rs:96-144
//   acc = acc * a + p * b (Polynomial * F + &Polynomial folds)
//                                                halo2_proofs/src/plonk/vanishing/prover.rs:131-135,
//                                                halo2_proofs/src/poly/kzg/multiopen/gwc/prover.rs:62-76
#include "common.cuh"

#include <string.h>

namespace h2b {

// x < 2^256 -> x mod r  (2^256 / r < 6: at most five subtractions)
H2B_D Fr reduce_256(Fr x) {
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = FrParams::mod(i);
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    uint32_t t[8];
    const uint32_t borrow = sub8(t, x.v, m);
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] = borrow ? x.v[i] : t[i];
  }
  return x;
}

// out[i] = Montgomery form of (lo + hi * 2^256) mod r, in[i] = sixteen u32 = eight u64, little-endian.
// from_u512: d0 * R2 + d1 * R3 in Montgomery products.
__global__ void fr_from_u512_kernel(const uint32_t* in, Fr* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fr lo, hi;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lo.v[j] = in[i * 16 + j];
      hi.v[j] = in[i * 16 + 8 + j];
    }
    lo = reduce_256(lo);
    hi = reduce_256(hi);
    const Fr r2 = Fr::r2();
    st_fp(out + i, add(mul(lo, r2), mul(mul(hi, r2), r2)));
  }
}

// The CounterRng stream (halo2-pse_b200/prover.py): word w (1-based) = splitmix64 finaliser of seed + w * golden.
H2B_D uint64_t counter_rng_word(uint64_t seed, uint64_t w) {
  uint64_t z = seed + w * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// out[i] = Fr::random drawn from the counter stream at words ctr + 8 i + 1 .. ctr + 8 i + 8
__global__ void fr_random_counter_kernel(uint64_t seed, uint64_t ctr, Fr* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fr lo, hi;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t a = counter_rng_word(seed, ctr + 8 * i + j + 1), b = counter_rng_word(seed, ctr + 8 * i + 4 + j + 1);
      lo.v[2 * j] = (uint32_t)a, lo.v[2 * j + 1] = (uint32_t)(a >> 32);
      hi.v[2 * j] = (uint32_t)b, hi.v[2 * j + 1] = (uint32_t)(b >> 32);
    }
    lo = reduce_256(lo);
    hi = reduce_256(hi);
    const Fr r2 = Fr::r2();
    st_fp(out + i, add(mul(lo, r2), mul(mul(hi, r2), r2)));
  }
}

struct PermFracArgs {
  const Fr* const* values;
  const Fr* const* sigma;
  uint32_t n_cols;
  Fr beta, gamma, deltaomega;  // deltaomega = DELTA^(index of the chunk's first column)
  Fr delta;
  const Fr *tw_lo, *tw_hi;     // omega^i two-level table
  uint32_t tw_h;
};

// phase 0: out[i] = prod_j (beta * sigma_j[i] + gamma + v_j[i])                      (prover.rs:99-116)
// phase 1: out[i] *= prod_j (delta^j * deltaomega * omega^i * beta + gamma + v_j[i])  (prover.rs:121-144)
__global__ void perm_fraction_kernel(PermFracArgs q, Fr* out, uint64_t n, int phase) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (phase == 0) {
      Fr acc = Fr::one();
      for (uint32_t c = 0; c < q.n_cols; ++c)
        acc = mul(acc, add(add(mul(q.beta, ld_fp(q.sigma[c] + i)), q.gamma), ld_fp(q.values[c] + i)));
      st_fp(out + i, acc);
    } else {
      Fr acc = ld_fp(out + i);
      Fr dw = mul(mul(q.deltaomega, q.beta),
                  mul(ld_fp_nc(q.tw_lo + (i & ((1ull << q.tw_h) - 1))), ld_fp_nc(q.tw_hi + (i >> q.tw_h))));
      for (uint32_t c = 0; c < q.n_cols; ++c) {
        acc = mul(acc, add(add(dw, q.gamma), ld_fp(q.values[c] + i)));
        dw = mul(dw, q.delta);
      }
      st_fp(out + i, acc);
    }
  }
}

// acc[i] = acc[i] * a + p[i] * b
__global__ void poly_fma_kernel(Fr* acc, const Fr* p, uint64_t n, Fr a, Fr b, int a_is_one, int b_is_one) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fr x = ld_fp(acc + i), y = ld_fp(p + i);
    if (!a_is_one) x = mul(x, a);
    if (!b_is_one) y = mul(y, b);
    st_fp(acc + i, add(x, y));
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
Fr fr_from_canonical(const uint64_t l[4]) {
  Fr a;
  for (int i = 0; i < 4; ++i) {
    a.v[2 * i] = (uint32_t)l[i];
    a.v[2 * i + 1] = (uint32_t)(l[i] >> 32);
  }
  return to_mont(a);
}
const uint64_t kDelta[4] = {0x870e56bbe533e9a2ull, 0x5b5f898e5e963f25ull, 0x64ec26aad4c86e71ull, 0x09226b6e22c6f0caull};
uint32_t grid_for(h2b_ctx* ctx, uint64_t n, uint32_t threads) {
  const uint64_t want = (n + threads - 1) / threads, cap = (uint64_t)ctx->sm_count * 16;
  return (uint32_t)(want < cap ? want : cap);
}
}  // namespace

extern "C" int h2b_fr_from_u512(h2b_ctx* ctx, const uint64_t* wide, int loc, size_t n, h2b_fr* out_dev) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!wide || !out_dev)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t* d_in = reinterpret_cast<const uint32_t*>(wide);
  if (loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 0, n * 64));
    H2B_CUDA(ctx, cudaMemcpyAsync(ctx->stage[0], wide, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    d_in = reinterpret_cast<const uint32_t*>(ctx->stage[0]);
  }
  H2B_TRY(launch(ctx, fr_from_u512_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0, d_in,
                 reinterpret_cast<Fr*>(out_dev), (uint64_t)n));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_fr_random_counter(h2b_ctx* ctx, uint64_t seed, uint64_t ctr, size_t n, h2b_fr* out_dev) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && !out_dev) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_TRY(launch(ctx, fr_random_counter_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0, seed, ctr,
                 reinterpret_cast<Fr*>(out_dev), (uint64_t)n));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_permutation_fractions(h2b_domain* dom, const h2b_fr* const* values, const h2b_fr* const* sigma,
                                         uint32_t n_cols, uint32_t first_column, const h2b_fr* beta,
                                         const h2b_fr* gamma, h2b_fr* out_dev) {
  if (!dom) return H2B_ERR_ARG;
  h2b_ctx* ctx = dom->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!values || !sigma || !beta || !gamma || !out_dev || n_cols == 0) return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t n = 1ull << dom->k;
  std::vector<const void*> host(2 * (size_t)n_cols);
  for (uint32_t i = 0; i < n_cols; ++i) {
    if (!values[i] || !sigma[i]) return fail(ctx, H2B_ERR_ARG, "null column pointer");
    host[i] = values[i];
    host[n_cols + i] = sigma[i];
  }
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, dom->omega, dom->k, &tw));
  void* d_list = nullptr;
  H2B_CUDA(ctx, cudaMalloc(&d_list, host.size() * sizeof(void*)));
  cudaError_t e = cudaMemcpyAsync(d_list, host.data(), host.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaFree(d_list);
    H2B_CUDA(ctx, e);
  }
  PermFracArgs q;
  q.values = reinterpret_cast<const Fr* const*>(d_list);
  q.sigma = q.values + n_cols;
  q.n_cols = n_cols;
  q.beta = load_fr(beta), q.gamma = load_fr(gamma);
  q.delta = fr_from_canonical(kDelta);
  q.deltaomega = pow_u64(q.delta, first_column);
  q.tw_lo = tw->d_lo, q.tw_hi = tw->d_hi, q.tw_h = tw->h;
  Fr* out = reinterpret_cast<Fr*>(out_dev);
  int rc = launch(ctx, perm_fraction_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0, q, out, n, 0);
  // invert the denominators (zeros stay zero, as ff::BatchInvert)            prover.rs:119
  if (rc == H2B_OK) rc = h2b_batch_invert(ctx, out_dev, H2B_DEVICE, n);
  if (rc == H2B_OK) rc = launch(ctx, perm_fraction_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0, q, out, n, 1);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_list);
  if (rc != H2B_OK) return rc;
  H2B_CUDA(ctx, e2);
  return H2B_OK;
}

extern "C" int h2b_poly_fma(h2b_ctx* ctx, h2b_fr* acc_dev, const h2b_fr* a, const h2b_fr* p_dev, const h2b_fr* b,
                            size_t n) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!acc_dev || !p_dev || !a || !b)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr fa = load_fr(a), fb = load_fr(b);
  H2B_TRY(launch(ctx, poly_fma_kernel, dim3(grid_for(ctx, n, 256)), dim3(256), 0, reinterpret_cast<Fr*>(acc_dev),
                 reinterpret_cast<const Fr*>(p_dev), (uint64_t)n, fa, fb, (int)(fa == Fr::one()),
                 (int)(fb == Fr::one())));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}
