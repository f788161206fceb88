// Device-resident polynomial helpers around the two hot paths (SURVEY.md 8f,
// rank 2): the O(n) passes halo2_proofs runs on the host between the NTTs and
// the commitments, so that a prover can keep its polynomials on the GPU.
//
//   eval_polynomial(poly, point)          halo2_proofs/src/arithmetic.rs:304-329
//   compute_inner_product(a, b)           halo2_proofs/src/arithmetic.rs:331-345
//   kate_division(a, b)                   halo2_proofs/src/arithmetic.rs:348-367
//   Polynomial + / - / * scalar           halo2_proofs/src/poly.rs:229-305
//   batch_invert + the running product z[i] = z[i-1] * f[i-1] of the permutation / lookup grand
//   products        halo2_proofs/src/plonk/permutation/prover.rs:119, 152-158
//
// The reference evaluates with per-thread Horner chunks and divides with a
// serial recurrence q_i = a_(i+1) + b q_(i+1).  Both are the suffix Horner scan
//   S_i = a_i + x S_(i+1),  S_n = 0:   eval = S_0,  q_i = S_(i+1),
// computed here as: per-thread Horner over 8 coefficients, a Hillis-Steele
// suffix scan of the 256 thread totals in shared memory (step d multiplies by
// x^(8*2^d)), block totals scanned recursively with x^2048, and a second
// Horner sweep seeded with each thread's carry.  Exact field arithmetic: the
// result is identical to the reference's whatever the association order.
#include "common.cuh"

#include <string.h>

namespace h2b {

static const uint32_t kPolyE = 8;                // coefficients per thread
static const uint32_t kPolyChunk = 256 * kPolyE;  // coefficients per block

// Block b handles in[b*2048 .. +2048).  carry[b] (optional) is S at the end of the block.
//   totals != null : totals[b] = S at the start of the block
//   out    != null : out[i + out_shift] = S_i for every i of the block with 0 <= i + out_shift
__global__ void __launch_bounds__(256)
    poly_suffix_horner_kernel(const Fr* in, Fr* out, int64_t out_shift, uint64_t n, Fr x, Fr x8,
                              const Fr* carry, Fr* totals) {
  __shared__ Fr sc[256];
  const uint32_t t = threadIdx.x;
  const uint64_t base = (uint64_t)blockIdx.x * kPolyChunk + (uint64_t)t * kPolyE;
  Fr a[kPolyE];
#pragma unroll
  for (uint32_t e = 0; e < kPolyE; ++e) a[e] = base + e < n ? ld_fp(in + base + e) : Fr::zero();
  Fr v = a[kPolyE - 1];
#pragma unroll
  for (int e = (int)kPolyE - 2; e >= 0; --e) v = add(a[e], mul(x, v));
  const Fr cb = carry ? ld_fp(carry + blockIdx.x) : Fr::zero();
  if (t == 255 && carry) v = add(v, mul(x8, cb));  // the carry enters as the element after the block
  sc[t] = v;
  __syncthreads();
  Fr pw = x8;
  for (uint32_t off = 1; off < 256; off <<= 1) {
    Fr other = Fr::zero();
    const bool has = t + off < 256;
    if (has) other = sc[t + off];
    __syncthreads();
    if (has) sc[t] = add(sc[t], mul(pw, other));
    __syncthreads();
    pw = sqr(pw);
  }
  if (totals && t == 0) st_fp(totals + blockIdx.x, sc[0]);
  if (!out) return;
  v = t < 255 ? sc[t + 1] : cb;
#pragma unroll
  for (int e = (int)kPolyE - 1; e >= 0; --e) {
    v = add(a[e], mul(x, v));
    const int64_t o = (int64_t)(base + e) + out_shift;
    if (base + e < n && o >= 0) st_fp(out + o, v);
  }
}

// op 0: a += b, 1: a -= b, 2: a *= s, 3: a = a * b (element-wise product, for inner products)
__global__ void poly_elementwise_kernel(Fr* a, const Fr* b, uint64_t n, int op, Fr s) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Fr x = ld_fp(a + i);
    Fr r;
    if (op == 0)
      r = add(x, ld_fp(b + i));
    else if (op == 1)
      r = sub(x, ld_fp(b + i));
    else if (op == 2)
      r = mul(x, s);
    else
      r = mul(x, ld_fp(b + i));
    st_fp(a + i, r);
  }
}

// out[i] = a[i] * b[i] into a fresh buffer (inner product, first stage)
__global__ void poly_product_kernel(const Fr* a, const Fr* b, Fr* out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    st_fp(out + i, mul(ld_fp(a + i), ld_fp(b + i)));
}

// Exclusive running product: out[i] = carry[b] * prod_{j < i, j in block} in[j] (elements past n count as 1).
//   totals != null : totals[b] = product of the block's elements
//   out    != null : the running products of the block
__global__ void __launch_bounds__(256)
    poly_prefix_product_kernel(const Fr* in, Fr* out, uint64_t n, const Fr* carry, Fr* totals) {
  __shared__ Fr sc[256];
  const uint32_t t = threadIdx.x;
  const uint64_t base = (uint64_t)blockIdx.x * kPolyChunk + (uint64_t)t * kPolyE;
  Fr a[kPolyE];
#pragma unroll
  for (uint32_t e = 0; e < kPolyE; ++e) a[e] = base + e < n ? ld_fp(in + base + e) : Fr::one();
  Fr v = a[0];
#pragma unroll
  for (uint32_t e = 1; e < kPolyE; ++e) v = mul(v, a[e]);
  sc[t] = v;
  __syncthreads();
  for (uint32_t off = 1; off < 256; off <<= 1) {  // inclusive prefix products of the thread totals
    Fr other = Fr::one();
    const bool has = t >= off;
    if (has) other = sc[t - off];
    __syncthreads();
    if (has) sc[t] = mul(other, sc[t]);
    __syncthreads();
  }
  if (totals && t == 255) st_fp(totals + blockIdx.x, sc[255]);
  if (!out) return;
  v = ld_fp(carry + blockIdx.x);
  if (t > 0) v = mul(v, sc[t - 1]);
#pragma unroll
  for (uint32_t e = 0; e < kPolyE; ++e) {
    if (base + e < n) st_fp(out + base + e, v);
    v = mul(v, a[e]);
  }
}

// a[i] <- 1 / a[i] (zeros stay zero, as ff::BatchInvert): Montgomery's trick, one inversion per thread
// over the strided elements it owns, prefix products parked in `pre`.
__global__ void __launch_bounds__(128) poly_batch_invert_kernel(Fr* a, Fr* pre, uint64_t n) {
  const uint64_t T = (uint64_t)gridDim.x * blockDim.x, t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fr run = Fr::one();
  for (uint64_t i = t; i < n; i += T) {
    const Fr x = ld_fp(a + i);
    st_fp(pre + i, run);
    if (!x.is_zero()) run = mul(run, x);
  }
  Fr inv_run = inv(run);
  const uint64_t last = t + ((n - 1 - t) / T) * T;
  for (int64_t i = (int64_t)last; i >= (int64_t)t; i -= (int64_t)T) {
    const Fr x = ld_fp(a + i);
    if (x.is_zero()) continue;
    st_fp(a + i, mul(inv_run, ld_fp(pre + i)));
    inv_run = mul(inv_run, x);
  }
}

static Fr fr_pow(Fr a, uint64_t e) { return pow_u64(a, e); }

// out[i] = init * prod_{j<i} in[j], i < n; d_tmp sized by poly_tmp_elems(n)
static int prefix_product_device(h2b_ctx* ctx, const Fr* in, Fr* out, uint64_t n, const Fr& init, Fr* d_tmp) {
  const uint64_t nb = (n + kPolyChunk - 1) / kPolyChunk;
  Fr* T = d_tmp;              // nb + 1
  Fr* Cx = d_tmp + (nb + 1);  // nb + 1
  Fr* rest = Cx + (nb + 1);
  if (nb == 1) {
    H2B_CUDA(ctx, cudaMemcpyAsync(Cx, &init, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `init` may live on the caller's stack
    return launch(ctx, poly_prefix_product_kernel, dim3(1), dim3(256), 0, in, out, n, (const Fr*)Cx, (Fr*)nullptr);
  }
  H2B_TRY(launch(ctx, poly_prefix_product_kernel, dim3((uint32_t)nb), dim3(256), 0, in, (Fr*)nullptr, n,
                 (const Fr*)nullptr, T));
  H2B_TRY(prefix_product_device(ctx, T, Cx, nb, init, rest));
  return launch(ctx, poly_prefix_product_kernel, dim3((uint32_t)nb), dim3(256), 0, in, out, n, (const Fr*)Cx,
                (Fr*)nullptr);
}

// S_0 of `in` (n coefficients) at x, left in d_tmp[0]; d_tmp holds ceil(n / 2048) + 2048 elements
static int eval_device(h2b_ctx* ctx, const Fr* in, uint64_t n, Fr x, Fr* d_tmp, Fr** result) {
  Fr* bufs[2] = {d_tmp, d_tmp + ((n + kPolyChunk - 1) / kPolyChunk + 1)};
  int cur = 0;
  const Fr* src = in;
  uint64_t len = n;
  while (true) {
    const uint64_t nb = (len + kPolyChunk - 1) / kPolyChunk;
    H2B_TRY(launch(ctx, poly_suffix_horner_kernel, dim3((uint32_t)nb), dim3(256), 0, src, (Fr*)nullptr,
                   (int64_t)0, len, x, fr_pow(x, kPolyE), (const Fr*)nullptr, bufs[cur]));
    if (nb == 1) break;
    src = bufs[cur];
    len = nb;
    x = fr_pow(x, kPolyChunk);
    cur ^= 1;
  }
  *result = bufs[cur];
  return H2B_OK;
}

// Inclusive suffix scan written with out_shift; scratch d_tmp as above (sized for n).
static int scan_device(h2b_ctx* ctx, const Fr* in, Fr* out, int64_t out_shift, uint64_t n, Fr x,
                       Fr* d_tmp) {
  const uint64_t nb = (n + kPolyChunk - 1) / kPolyChunk;
  const Fr x8 = fr_pow(x, kPolyE);
  if (nb == 1)
    return launch(ctx, poly_suffix_horner_kernel, dim3(1), dim3(256), 0, in, out, out_shift, n, x, x8,
                  (const Fr*)nullptr, (Fr*)nullptr);
  // totals T_b, their inclusive scan C_b with x^2048, carry of block b = C_(b+1)
  Fr* T = d_tmp;              // nb + 1 elements (the extra one is the zero carry of the last block)
  Fr* Cs = d_tmp + (nb + 1);  // nb + 1
  Fr* rest = Cs + (nb + 1);
  H2B_TRY(launch(ctx, poly_suffix_horner_kernel, dim3((uint32_t)nb), dim3(256), 0, in, (Fr*)nullptr, (int64_t)0,
                 n, x, x8, (const Fr*)nullptr, T));
  H2B_CUDA(ctx, cudaMemsetAsync(Cs + nb, 0, sizeof(Fr), ctx->stream));
  H2B_TRY(scan_device(ctx, T, Cs, 0, nb, fr_pow(x, kPolyChunk), rest));
  return launch(ctx, poly_suffix_horner_kernel, dim3((uint32_t)nb), dim3(256), 0, in, out, out_shift, n, x, x8,
                (const Fr*)(Cs + 1), (Fr*)nullptr);
}

static size_t poly_tmp_elems(uint64_t n) {
  // T and Cs of every recursion level (geometric) + slack
  size_t total = 0;
  uint64_t len = n;
  while (len > 1) {
    const uint64_t nb = (len + kPolyChunk - 1) / kPolyChunk;
    total += 2 * (nb + 1);
    len = nb;
  }
  return total + 2 * kPolyChunk + 16;
}

}  // namespace h2b

using namespace h2b;

namespace {
const Fr* as_fr(const h2b_fr* p) { return reinterpret_cast<const Fr*>(p); }
Fr* as_fr(h2b_fr* p) { return reinterpret_cast<Fr*>(p); }

// device view of a caller buffer (the pointer itself, or a staged copy of a host slice)
int stage_in(h2b_ctx* ctx, int which, const h2b_fr* p, int loc, size_t count, const Fr** dev) {
  if (loc == H2B_DEVICE) {
    *dev = as_fr(p);
    return H2B_OK;
  }
  H2B_TRY(ensure_stage(ctx, which, (count ? count : 1) * sizeof(Fr)));
  H2B_TRY(copy_h2d_any(ctx, ctx->stage[which], p, count * sizeof(Fr), ctx->stream));
  *dev = reinterpret_cast<const Fr*>(ctx->stage[which]);
  return H2B_OK;
}
}  // namespace

extern "C" int h2b_eval_polynomial(h2b_ctx* ctx, const h2b_fr* poly, int loc, size_t n, const h2b_fr* point,
                                   h2b_fr* out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if ((!poly && n) || !point || !out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) {  // the fold over an empty slice is zero
    memset(out, 0, sizeof(h2b_fr));
    return H2B_OK;
  }
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_in;
  H2B_TRY(stage_in(ctx, 0, poly, loc, n, &d_in));
  H2B_TRY(ensure_scratch(ctx, poly_tmp_elems(n) * sizeof(Fr)));
  Fr* res;
  H2B_TRY(eval_device(ctx, d_in, n, *as_fr(point), reinterpret_cast<Fr*>(ctx->scratch), &res));
  H2B_CUDA(ctx, cudaMemcpyAsync(out, res, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_kate_division(h2b_ctx* ctx, const h2b_fr* a, int loc, size_t n, const h2b_fr* b,
                                 h2b_fr* q_out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a || !b || (!q_out && n > 1)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return fail(ctx, H2B_ERR_LENGTH, "kate_division of an empty polynomial");  // a.len() - 1 underflows
  if (n == 1) return H2B_OK;
  if (loc == H2B_DEVICE && as_fr(q_out) == as_fr(a)) return fail(ctx, H2B_ERR_ARG, "aliased output");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_in;
  H2B_TRY(stage_in(ctx, 0, a, loc, n, &d_in));
  Fr* d_out = as_fr(q_out);
  if (loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 1, (n - 1) * sizeof(Fr)));
    d_out = reinterpret_cast<Fr*>(ctx->stage[1]);
  }
  H2B_TRY(ensure_scratch(ctx, poly_tmp_elems(n) * sizeof(Fr)));
  // q_i = S_(i+1): the scan of a at b, shifted down by one
  H2B_TRY(scan_device(ctx, d_in, d_out, -1, n, *as_fr(b), reinterpret_cast<Fr*>(ctx->scratch)));
  if (loc != H2B_DEVICE)
    H2B_CUDA(ctx, cudaMemcpyAsync(q_out, d_out, (n - 1) * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_inner_product(h2b_ctx* ctx, const h2b_fr* a, const h2b_fr* b, int loc, size_t n,
                                 h2b_fr* out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if ((n && (!a || !b)) || !out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) {
    memset(out, 0, sizeof(h2b_fr));
    return H2B_OK;
  }
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr *da, *db;
  H2B_TRY(stage_in(ctx, 0, a, loc, n, &da));
  H2B_TRY(stage_in(ctx, 1, b, loc, n, &db));
  H2B_TRY(ensure_scratch(ctx, (n + poly_tmp_elems(n)) * sizeof(Fr)));
  Fr* prod = reinterpret_cast<Fr*>(ctx->scratch);
  const uint64_t want = (n + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
  H2B_TRY(launch(ctx, poly_product_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0, da, db, prod,
                 (uint64_t)n));
  Fr* res;
  H2B_TRY(eval_device(ctx, prod, n, Fr::one(), prod + n, &res));  // Horner at 1 = the plain sum
  H2B_CUDA(ctx, cudaMemcpyAsync(out, res, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

static int poly_elementwise(h2b_ctx* ctx, h2b_fr* a, const h2b_fr* b, int loc, size_t n, int op, const h2b_fr* s) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!a || (op != 2 && !b) || (op == 2 && !s))) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* da_c;
  H2B_TRY(stage_in(ctx, 0, a, loc, n, &da_c));
  Fr* da = const_cast<Fr*>(da_c);
  const Fr* db = nullptr;
  if (op != 2) H2B_TRY(stage_in(ctx, 1, b, loc, n, &db));
  Fr sc = Fr::zero();
  if (s) memcpy(&sc, s, sizeof(Fr));
  const uint64_t want = (n + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
  H2B_TRY(launch(ctx, poly_elementwise_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0, da, db,
                 (uint64_t)n, op, sc));
  if (loc != H2B_DEVICE)
    H2B_CUDA(ctx, cudaMemcpyAsync(a, da, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_poly_add(h2b_ctx* ctx, h2b_fr* lhs, const h2b_fr* rhs, int loc, size_t n) {
  return poly_elementwise(ctx, lhs, rhs, loc, n, 0, nullptr);
}
extern "C" int h2b_poly_sub(h2b_ctx* ctx, h2b_fr* lhs, const h2b_fr* rhs, int loc, size_t n) {
  return poly_elementwise(ctx, lhs, rhs, loc, n, 1, nullptr);
}
extern "C" int h2b_poly_scale(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n, const h2b_fr* scalar) {
  return poly_elementwise(ctx, a, nullptr, loc, n, 2, scalar);
}

// Fr::to_repr / Fr::from_repr over a whole polynomial (SerdeFormat::Processed of Polynomial::write / read,
// poly.rs + helpers.rs:54-94): Montgomery limbs <-> canonical little-endian integers.  from_repr rejects
// values >= r (`ok` = 0), like the reference's CtOption; with check_only the limbs are only range-checked
// (SerdeObject::read_raw of RawBytes).
namespace h2b {
__global__ void fr_repr_kernel(Fr* a, uint64_t n, int mode, int* bad) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Fr x = ld_fp(a + i);
    if (mode == 0) {
      st_fp(a + i, from_mont(x));
      continue;
    }
    uint32_t m[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = FrParams::mod(j);
    if (sub8(t, x.v, m) == 0) atomicOr(bad, 1);  // no borrow: x >= r
    if (mode == 1) st_fp(a + i, to_mont(x));
  }
}
}  // namespace h2b

extern "C" int h2b_fr_repr(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n, int mode, int* ok) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if ((n && !a) || mode < 0 || mode > 2 || (mode && !ok)) return fail(ctx, H2B_ERR_ARG, "bad argument");
  if (ok) *ok = 1;
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_c;
  H2B_TRY(stage_in(ctx, 0, a, loc, n, &d_c));
  Fr* d = const_cast<Fr*>(d_c);
  H2B_TRY(ensure_scratch(ctx, 64));
  int* d_bad = reinterpret_cast<int*>(ctx->scratch);
  H2B_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  const uint64_t want = (n + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
  H2B_TRY(launch(ctx, fr_repr_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0, d, (uint64_t)n, mode, d_bad));
  int bad = 0;
  H2B_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (loc != H2B_DEVICE && mode != 2)
    H2B_CUDA(ctx, cudaMemcpyAsync(a, d, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ok) *ok = bad ? 0 : 1;
  return H2B_OK;
}

extern "C" int h2b_batch_invert(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && !a) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_c;
  H2B_TRY(stage_in(ctx, 0, a, loc, n, &d_c));
  Fr* d = const_cast<Fr*>(d_c);
  H2B_TRY(ensure_scratch(ctx, n * sizeof(Fr)));
  // >= 256 elements per thread keep the Fermat inversion (381 multiplications) at ~1.5 per element
  uint64_t threads = (n + 255) / 256;
  const uint64_t cap = (uint64_t)ctx->sm_count * 4 * 128;
  if (threads > cap) threads = cap;
  const uint32_t blocks = (uint32_t)((threads + 127) / 128);
  H2B_TRY(launch(ctx, poly_batch_invert_kernel, dim3(blocks), dim3(128), 0, d, reinterpret_cast<Fr*>(ctx->scratch),
                 (uint64_t)n));
  if (loc != H2B_DEVICE)
    H2B_CUDA(ctx, cudaMemcpyAsync(a, d, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_running_product(h2b_ctx* ctx, const h2b_fr* in, int loc, size_t n, const h2b_fr* init,
                                   h2b_fr* out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!in || !out || !init)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  if (loc == H2B_DEVICE && as_fr(out) == as_fr(in)) return fail(ctx, H2B_ERR_ARG, "aliased output");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_in;
  H2B_TRY(stage_in(ctx, 0, in, loc, n, &d_in));
  Fr* d_out = as_fr(out);
  if (loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 1, n * sizeof(Fr)));
    d_out = reinterpret_cast<Fr*>(ctx->stage[1]);
  }
  H2B_TRY(ensure_scratch(ctx, poly_tmp_elems(n) * sizeof(Fr)));
  Fr i0;
  memcpy(&i0, init, sizeof(Fr));
  H2B_TRY(prefix_product_device(ctx, d_in, d_out, n, i0, reinterpret_cast<Fr*>(ctx->scratch)));
  if (loc != H2B_DEVICE)
    H2B_CUDA(ctx, cudaMemcpyAsync(out, d_out, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}
