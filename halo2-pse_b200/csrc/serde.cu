// Wire formats of G1 points on the device: the three SerdeFormat variants of
// halo2_proofs/src/helpers.rs:8-52 for the 2 x 2^k points of ParamsKZG::read_custom / write_custom
// (halo2_proofs/src/poly/kzg/commitment.rs:142-244).
//   RawBytes / RawBytesUnchecked : the in-memory Montgomery limbs, 64 B per point -> zero-copy upload; RawBytes
//                                  additionally checks every point is on the curve (SerdeObject::read_raw).
//   Processed                    : G1Affine::to_bytes / from_bytes, 32 B per point: little-endian canonical x
//                                  with the parity of y in one spare bit of byte 31, identity = zeros.  Which bit
//                                  is halo2curves' choice (0.3.1: bit 7); it is a parameter here.
#include "common.cuh"

namespace h2b {

// limb i of (q + 1) / 4, the square-root exponent in Fq (q = 3 mod 4: adding one does not carry out of limb 0)
H2B_D uint32_t sqrt_exp_limb(int i) {
  const uint32_t lo = FqParams::mod(i) + (i == 0 ? 1u : 0u);
  const uint32_t hi = i < 7 ? FqParams::mod(i + 1) : 0u;
  return (lo >> 2) | (hi << 30);
}

H2B_D Fq fq_three() {
  const Fq one = Fq::one();
  return add(add(one, one), one);
}

H2B_D bool g1_on_curve(const G1Affine& p) {
  if (p.is_identity()) return true;
  return sqr(p.y) == add(mul(sqr(p.x), p.x), fq_three());
}

// limbs < q: SerdeObject::read_raw rejects non-canonical coordinates before it checks the curve equation.  A
// residue such as x + q passes the Montgomery curve check, but the group law's is_zero() / == (P + P, P - P and
// identity detection in ec.cuh) assume reduced limbs.
H2B_D bool fq_is_canonical(const Fq& a) {
  uint32_t m[8], t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = FqParams::mod(j);
  return sub8(t, a.v, m) != 0;  // borrow: a < q
}

__global__ void g1_check_kernel(const G1Affine* pts, uint64_t n, int* err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    G1Affine p;
    p.x = ld_fp(&pts[i].x);
    p.y = ld_fp(&pts[i].y);
    if (!fq_is_canonical(p.x) || !fq_is_canonical(p.y) || !g1_on_curve(p)) atomicOr(err, 1);
  }
}

__global__ void g1_compress_kernel(const G1Affine* pts, uint32_t* out, uint64_t n, uint32_t sign_bit) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    G1Affine p;
    p.x = ld_fp(&pts[i].x);
    p.y = ld_fp(&pts[i].y);
    Fq x = Fq::zero();
    if (!p.is_identity()) {
      x = from_mont(p.x);
      x.v[7] |= (from_mont(p.y).v[0] & 1u) << (24 + sign_bit);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) out[i * 8 + j] = x.v[j];
  }
}

__global__ void g1_decompress_kernel(const uint32_t* in, G1Affine* out, uint64_t n, uint32_t sign_bit, int* err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fq x;
#pragma unroll
    for (int j = 0; j < 8; ++j) x.v[j] = in[i * 8 + j];
    G1Affine p;
    p.x = Fq::zero();
    p.y = Fq::zero();
    if (!x.is_zero()) {
      const uint32_t sign = (x.v[7] >> (24 + sign_bit)) & 1u;
      x.v[7] &= ~(1u << (24 + sign_bit));
      // canonical: x < q
      uint32_t m[8], t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = FqParams::mod(j);
      if (sub8(t, x.v, m) == 0) atomicOr(err, 1);  // no borrow: x >= q
      p.x = to_mont(x);
      const Fq rhs = add(mul(sqr(p.x), p.x), fq_three());
      Fq y = Fq::one();
      for (int limb = 7; limb >= 0; --limb) {
        const uint32_t e = sqrt_exp_limb(limb);
        for (int bit = 31; bit >= 0; --bit) {
          y = sqr(y);
          if ((e >> bit) & 1u) y = mul(y, rhs);
        }
      }
      if (sqr(y) != rhs) atomicOr(err, 1);  // x^3 + 3 is not a square: not a point
      if ((from_mont(y).v[0] & 1u) != sign) y = neg(y);
      p.y = y;
    }
    st_fp(&out[i].x, p.x);
    st_fp(&out[i].y, p.y);
  }
}

}  // namespace h2b

using namespace h2b;

namespace {
uint32_t grid_for(h2b_ctx* ctx, uint64_t n, uint32_t threads) {
  const uint64_t want = (n + threads - 1) / threads, cap = (uint64_t)ctx->sm_count * 16;
  return (uint32_t)(want < cap ? want : cap);
}
int read_flag(h2b_ctx* ctx, int* d_err, int* out) {
  cudaError_t e = cudaMemcpyAsync(out, d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_err);
  H2B_CUDA(ctx, e);
  return H2B_OK;
}
}  // namespace

extern "C" int h2b_g1_check_on_curve(h2b_ctx* ctx, const h2b_g1_affine* pts_dev, size_t n, int* all_valid) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!all_valid || (n && !pts_dev)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  *all_valid = 1;
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  int* d_err;
  H2B_CUDA(ctx, cudaMalloc((void**)&d_err, sizeof(int)));
  cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream);
  int rc = launch(ctx, g1_check_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0,
                  reinterpret_cast<const G1Affine*>(pts_dev), (uint64_t)n, d_err);
  int err = 0;
  H2B_TRY(read_flag(ctx, d_err, &err));
  H2B_TRY(rc);
  *all_valid = err ? 0 : 1;
  return H2B_OK;
}

extern "C" int h2b_g1_compress(h2b_ctx* ctx, const h2b_g1_affine* pts_dev, size_t n, uint32_t sign_bit,
                               uint8_t* out_host) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!pts_dev || !out_host)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (sign_bit > 7) return fail(ctx, H2B_ERR_ARG, "sign_bit is a bit of byte 31");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_TRY(ensure_stage(ctx, 1, n * 32));
  H2B_TRY(launch(ctx, g1_compress_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0,
                 reinterpret_cast<const G1Affine*>(pts_dev), reinterpret_cast<uint32_t*>(ctx->stage[1]), (uint64_t)n,
                 sign_bit));
  H2B_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->stage[1], n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_g1_decompress(h2b_ctx* ctx, const uint8_t* in_host, size_t n, uint32_t sign_bit,
                                 h2b_g1_affine* out_dev, int* all_valid) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!all_valid || (n && (!in_host || !out_dev))) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (sign_bit > 7) return fail(ctx, H2B_ERR_ARG, "sign_bit is a bit of byte 31");
  *all_valid = 1;
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_TRY(ensure_stage(ctx, 0, n * 32));
  H2B_CUDA(ctx, cudaMemcpyAsync(ctx->stage[0], in_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  int* d_err;
  H2B_CUDA(ctx, cudaMalloc((void**)&d_err, sizeof(int)));
  cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream);
  int rc = launch(ctx, g1_decompress_kernel, dim3(grid_for(ctx, n, 128)), dim3(128), 0,
                  reinterpret_cast<const uint32_t*>(ctx->stage[0]), reinterpret_cast<G1Affine*>(out_dev), (uint64_t)n,
                  sign_bit, d_err);
  int err = 0;
  H2B_TRY(read_flag(ctx, d_err, &err));
  H2B_TRY(rc);
  *all_valid = err ? 0 : 1;
  return H2B_OK;
}
