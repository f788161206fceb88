/* libhalo2b200 -- C ABI of the B200-native backend for halo2_proofs' two
 * data-parallel hot paths over bn256 (MSM and NTT).
 *
 * The reference (eldenpark/halo2-pse, 100 % Rust) has no FFI: its boundary is
 * the Rust function signatures below.  Each export names the signature it
 * replaces (paths relative to /root/reference/halo2_proofs/src).  The Rust
 * shim a maintainer would add on the reference side is in INTEGRATION.md.
 *
 * Data layout (identical bytes to halo2curves 0.3.1):
 *   h2b_fr / h2b_fq : four little-endian u64 limbs holding the Montgomery
 *                     residue a * 2^256 mod p, fully reduced.
 *   h2b_g1_affine   : {x, y}; the identity is (0, 0).
 *   h2b_g1          : Jacobian {x, y, z} (x/z^2, y/z^3); identity has z = 0.
 *
 * Conventions: every function returns H2B_OK (0) or a negative error code and
 * never throws; the reference panics where this ABI returns H2B_ERR_LENGTH /
 * H2B_ERR_ARG (arithmetic.rs:133,184; domain.rs:227,244,282,311;
 * kzg/commitment.rs:290,332).  `loc` says where a data pointer lives:
 * H2B_HOST (the drop-in case: the library stages through pinned memory) or
 * H2B_DEVICE (device-resident polynomials/scalars on the context's device).
 * All calls on one context are serialised internally (thread-safe); use one
 * context per caller thread for concurrency.  There is no CPU fallback: every
 * compute entry point fails with H2B_ERR_CUDA when no device is usable.
 */
#ifndef HALO2_B200_H
#define HALO2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } h2b_fr;
typedef struct { uint64_t l[4]; } h2b_fq;
typedef struct { h2b_fq x, y; } h2b_g1_affine;
typedef struct { h2b_fq x, y, z; } h2b_g1;

typedef struct h2b_ctx h2b_ctx;       /* one device + stream + scratch            */
typedef struct h2b_bases h2b_bases;   /* device-resident, immutable affine bases  */
typedef struct h2b_domain h2b_domain; /* EvaluationDomain: constants + tables     */

enum {
  H2B_OK = 0,
  H2B_ERR_ARG = -1,       /* null pointer, bad enum, k out of range              */
  H2B_ERR_LENGTH = -2,    /* the reference's assert_eq!(len) panics              */
  H2B_ERR_CUDA = -3,      /* CUDA runtime / launch failure, or no device         */
  H2B_ERR_OOM = -4,       /* device or pinned allocation failed                  */
  H2B_ERR_BAD_OMEGA = -5, /* omega is not a primitive 2^log_n-th root of unity   */
  H2B_ERR_CONSTRAINT = -6 /* Error::ConstraintSystemFailure (lookup input not in the table) */
};
enum { H2B_HOST = 0, H2B_DEVICE = 1 };

/* ---- context ----------------------------------------------------------- */
int h2b_ctx_create(int device, h2b_ctx** out);
void h2b_ctx_destroy(h2b_ctx* ctx);
const char* h2b_last_error(const h2b_ctx* ctx);
int h2b_ctx_sync(h2b_ctx* ctx);
/* the cudaStream_t every kernel of this context is launched on */
void* h2b_ctx_stream(h2b_ctx* ctx);
/* number of kernels this context has launched so far */
uint64_t h2b_ctx_launches(const h2b_ctx* ctx);

/* ---- MSM --------------------------------------------------------------- */
/* Upload `g` / `g_lagrange` once (ParamsKZG fields, poly/kzg/commitment.rs:23-31). */
int h2b_bases_upload(h2b_ctx* ctx, const h2b_g1_affine* bases, size_t n, int loc,
                     h2b_bases** out);
void h2b_bases_free(h2b_bases* bases);
/* One-time window table of an immutable base set: T_w[i] = 2^(c*w) * bases[i] for every
 * signed-digit window w, so that all windows of later h2b_msm calls share ONE set of
 * 2^(c-1) buckets (fewer windows, a W-times smaller bucket reduction).  Costs
 * ceil(255/c) * n * 64 B of device memory.  window_bits = 0 picks c from n.  Results
 * are unchanged; h2b_msm uses the table when present. */
int h2b_bases_precompute(h2b_ctx* ctx, h2b_bases* bases, uint32_t window_bits);
uint32_t h2b_bases_table_window_bits(const h2b_bases* bases);
size_t h2b_bases_len(const h2b_bases* bases);
void* h2b_bases_device_ptr(const h2b_bases* bases);

/* best_multiexp(coeffs, &bases[offset..offset+n])          arithmetic.rs:132
 * = ParamsKZG::commit_lagrange / commit with the matching base set
 *                                  poly/kzg/commitment.rs:281-292, 327-334
 * `out` is written on the host. */
int h2b_msm(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset, const h2b_fr* scalars,
            int loc, size_t n, h2b_g1* out);
/* Same result, normalised: out_affine = to_affine(sum) ((0,0) for identity). */
int h2b_msm_affine(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                   const h2b_fr* scalars, int loc, size_t n, h2b_g1_affine* out_affine);
/* `ncols` commitments on the same base slice in ONE pass: out_affine[j] = best_multiexp(scalars_dev[j][..n],
 * &bases[offset..offset+n]).  scalars_dev is a host array of DEVICE pointers.  With a window table the columns
 * share the digit / sort / accumulate / reduce pipeline (one bucket set per column), which amortises its
 * latency-bound stages: the commitments no challenge separates -- the advice columns of a phase
 * (plonk/prover.rs:375-392), the pieces of h (plonk/vanishing/prover.rs:100-113), the witness polynomials of a
 * multi-opening (poly/kzg/multiopen/gwc/prover.rs:80-88).  Same points as ncols calls of h2b_msm_affine. */
int h2b_msm_multi_affine(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                         const h2b_fr* const* scalars_dev, uint32_t ncols, size_t n,
                         h2b_g1_affine* out_affine);
/* One-shot drop-in with the exact shape of best_multiexp (host slices).  arithmetic.rs:132 */
int h2b_best_multiexp(h2b_ctx* ctx, const h2b_fr* coeffs, const h2b_g1_affine* bases, size_t n,
                      h2b_g1* out);

/* signed-digit window width c the Pippenger kernels use for an n-point MSM
 * (ceil(255 / c) windows of 2^(c-1) buckets) */
uint32_t h2b_msm_window_bits(size_t n);
/* out[i] = [scalars[i]] G for the generator G = (1, 2): the scalar multiplications of
 * ParamsKZG::setup (poly/kzg/commitment.rs:67-116), affine outputs (identity = (0, 0)). */
int h2b_g1_mul_generator(h2b_ctx* ctx, const h2b_fr* scalars, int loc, size_t n, h2b_g1_affine* out,
                         int out_loc);
/* out = sum of n affine points, on the host: the fold of per-chunk partial sums
 * (arithmetic.rs:153) when an MSM is sharded by point range across GPUs. */
int h2b_g1_sum(const h2b_g1_affine* pts, size_t n, h2b_g1_affine* out);

/* small_multiexp(coeffs, bases): double-and-add with shared doublings over a handful of points, on the
 * host like the reference's (not a hot path).                 arithmetic.rs:105-125 */
int h2b_small_multiexp(const h2b_fr* coeffs, const h2b_g1_affine* bases, size_t n, h2b_g1* out);
/* g_to_lagrange(g.to_curve(), k): inverse FFT over the 2^k curve points g (omega^-1 butterflies with
 * group_add / group_sub / group_scale), every point scaled by 1/n, batch-normalised.  arithmetic.rs:277-301
 * The KZG caller is ParamsKZG::downsize(k): g.truncate(1 << k); g_lagrange = g_to_lagrange(g, k)
 *                                                                  poly/kzg/commitment.rs:267-275 */
int h2b_g_to_lagrange(h2b_ctx* ctx, const h2b_g1_affine* g, int loc, uint32_t k, h2b_g1_affine* out,
                      int out_loc);

/* ---- NTT --------------------------------------------------------------- */
/* best_fft(a, omega, log_n): in place, natural order in and out.  arithmetic.rs:171
 * omega must be a primitive 2^log_n-th root of unity (every non-bench caller
 * passes one: domain.rs:230,248,285-290; arithmetic.rs:285). */
int h2b_best_fft(h2b_ctx* ctx, h2b_fr* a, int loc, const h2b_fr* omega, uint32_t log_n);

/* EvaluationDomain::new(j, k)                                     poly/domain.rs:39 */
int h2b_domain_new(h2b_ctx* ctx, uint32_t j, uint32_t k, h2b_domain** out);
void h2b_domain_free(h2b_domain* dom);
uint32_t h2b_domain_k(const h2b_domain* dom);
uint32_t h2b_domain_extended_k(const h2b_domain* dom);
/* n * quotient_poly_degree: the length extended_to_coeff returns (domain.rs:299-300) */
size_t h2b_domain_quotient_len(const h2b_domain* dom);
/* which = 0 omega, 1 omega_inv, 2 extended_omega, 3 extended_omega_inv,
 *         4 g_coset, 5 g_coset_inv, 6 ifft_divisor, 7 extended_ifft_divisor,
 *         8 + i  t_evaluations[i]                           (domain.rs:19-34) */
int h2b_domain_constant(const h2b_domain* dom, uint32_t which, h2b_fr* out);

/* lagrange_to_coeff: a (2^k) in place                       poly/domain.rs:226 */
int h2b_lagrange_to_coeff(h2b_domain* dom, h2b_fr* a, int loc);
/* coeff_to_extended: in (2^k) -> out (2^extended_k)         poly/domain.rs:240 */
int h2b_coeff_to_extended(h2b_domain* dom, const h2b_fr* in, h2b_fr* out, int loc);
/* extended_to_coeff: in (2^extended_k) -> out (quotient_len); when
 * divide_by_vanishing != 0 the preceding divide_by_vanishing_poly is fused in
 * (the only call order in the reference, plonk/vanishing/prover.rs:84-87).
 * `in` is not modified.                                     poly/domain.rs:281 */
int h2b_extended_to_coeff(h2b_domain* dom, const h2b_fr* in, h2b_fr* out, int loc,
                          int divide_by_vanishing);
/* divide_by_vanishing_poly: a (2^extended_k) in place       poly/domain.rs:307 */
int h2b_divide_by_vanishing_poly(h2b_domain* dom, h2b_fr* a, int loc);

/* Batched (column-parallel) forms: `ncols` polynomials, column c at
 * base + c * stride elements (stride >= the column's length).  One launch per
 * pass covers all columns. */
int h2b_best_fft_batch(h2b_ctx* ctx, h2b_fr* a, int loc, const h2b_fr* omega, uint32_t log_n,
                       uint32_t ncols, size_t stride);
int h2b_lagrange_to_coeff_batch(h2b_domain* dom, h2b_fr* a, int loc, uint32_t ncols,
                                size_t stride);
int h2b_coeff_to_extended_batch(h2b_domain* dom, const h2b_fr* in, size_t in_stride, h2b_fr* out,
                                size_t out_stride, int loc, uint32_t ncols);
int h2b_extended_to_coeff_batch(h2b_domain* dom, const h2b_fr* in, size_t in_stride, h2b_fr* out,
                                size_t out_stride, int loc, uint32_t ncols,
                                int divide_by_vanishing);

/* ---- polynomial helpers around the hot paths (the O(n) host passes
 * between the transforms and the commitments, so polynomials can stay on the device) ---- */
/* eval_polynomial(poly, point): sum poly[i] * point^i                 arithmetic.rs:304 */
int h2b_eval_polynomial(h2b_ctx* ctx, const h2b_fr* poly, int loc, size_t n, const h2b_fr* point,
                        h2b_fr* out);
/* kate_division(a, b): quotient of a(X) by (X - b), n - 1 coefficients, the remainder a(b) is
 * dropped as in the reference; q_out has the same loc as a and must not alias it.  arithmetic.rs:348 */
int h2b_kate_division(h2b_ctx* ctx, const h2b_fr* a, int loc, size_t n, const h2b_fr* b, h2b_fr* q_out);
/* compute_inner_product(a, b)                                          arithmetic.rs:331 */
int h2b_inner_product(h2b_ctx* ctx, const h2b_fr* a, const h2b_fr* b, int loc, size_t n, h2b_fr* out);
/* Polynomial += / -= / *= scalar, in place on lhs                      poly.rs:229, 243, 278 */
int h2b_poly_add(h2b_ctx* ctx, h2b_fr* lhs, const h2b_fr* rhs, int loc, size_t n);
int h2b_poly_sub(h2b_ctx* ctx, h2b_fr* lhs, const h2b_fr* rhs, int loc, size_t n);
int h2b_poly_scale(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n, const h2b_fr* scalar);
/* Fr::to_repr / from_repr over n elements in place (Polynomial::write / read with SerdeFormat::Processed,
 * helpers.rs:54-94): mode 0 = Montgomery limbs -> canonical little-endian integers; mode 1 = canonical ->
 * Montgomery, *ok = 0 if any value is >= r (the reference's from_repr returns None); mode 2 = range check
 * only (read_raw of SerdeFormat::RawBytes), data untouched. */
int h2b_fr_repr(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n, int mode, int* ok);

/* The two data-parallel pieces of the grand-product constructions:
 * a[i] <- 1/a[i] in place, zeros stay zero (ff::BatchInvert, plonk/permutation/prover.rs:119), and
 * out[0] = init, out[i] = out[i-1] * in[i-1] for i < n (plonk/permutation/prover.rs:152-158). */
int h2b_batch_invert(h2b_ctx* ctx, h2b_fr* a, int loc, size_t n);
int h2b_running_product(h2b_ctx* ctx, const h2b_fr* in, int loc, size_t n, const h2b_fr* init, h2b_fr* out);

/* ---- quotient evaluation: Evaluator::evaluate_h, plonk/evaluation.rs:280-522 ----
 * A GraphEvaluator (evaluation.rs:193-202) is handed over as a flat word stream, one record per
 * CalculationInfo in order:  op, target, operands...  with
 *   op      0 Add(a,b) 1 Sub(a,b) 2 Mul(a,b) 3 Square(a) 4 Double(a) 5 Negate(a)
 *           6 Horner(start, factor, nparts, part_0 .. part_{nparts-1})   7 Store(a)      (evaluation.rs:110-127)
 *   operand three words (kind, i, j):  0 Constant(i) 1 Intermediate(i) 2 Fixed(column i, rotation index j)
 *           3 Advice(i, j) 4 Instance(i, j) 5 Challenge(i) 6 Beta 7 Gamma 8 Theta 9 Y 10 PreviousValue
 *                                                                                         (evaluation.rs:38-61)
 * `rotations` are GraphEvaluator::rotations (Rotation.0 values); h2b_graph_new compiles the stream once, as
 * Evaluator::new does at keygen (evaluation.rs:224-277). */
typedef struct h2b_graph h2b_graph;
int h2b_graph_new(h2b_ctx* ctx, const uint32_t* calculations, size_t n_words, const h2b_fr* constants,
                  uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations,
                  uint32_t num_intermediates, h2b_graph** out);
void h2b_graph_free(h2b_graph* graph);
uint32_t h2b_graph_num_slots(const h2b_graph* graph);        /* live intermediates after renaming */
uint32_t h2b_graph_num_instructions(const h2b_graph* graph); /* three-address instructions */

/* The polynomials evaluate_h reads, as DEVICE pointers to cosets of 2^extended_k elements each (the outputs
 * of h2b_coeff_to_extended), plus the challenges (host values). */
typedef struct {
  const h2b_fr* const* fixed;    uint32_t n_fixed;     /* pk.fixed_cosets              */
  const h2b_fr* const* advice;   uint32_t n_advice;    /* coeff_to_extended(advice)    */
  const h2b_fr* const* instance; uint32_t n_instance;  /* coeff_to_extended(instance)  */
  const h2b_fr* challenges;      uint32_t n_challenges;
  h2b_fr beta, gamma, theta, y;
} h2b_eval_columns;

/* custom gates: values[i] = custom_gates.evaluate(previous_value = values[i]) for every row of the extended
 * domain; `values` is a device array of 2^extended_k elements, in place.        evaluation.rs:336-362 */
int h2b_evaluate_h_gates(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols, h2b_fr* values);
/* The same interpreter over the 2^k rows of the Lagrange basis with rot_scale = 1 (columns are 2^k-element
 * Lagrange vectors): `evaluate(expression, n, 1, ...)` folded with theta by a Horner graph, as the lookup
 * argument compresses its expressions.   plonk/evaluation.rs:749-787, plonk/lookup/prover.rs:82-104 */
int h2b_graph_evaluate_lagrange(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols, h2b_fr* values);
/* permutation constraints folded into `values` with y.  column_type: 0 Advice, 1 Fixed, 2 Instance (plonk
 * circuit.rs `Any`); set s covers columns [s*chunk_len, (s+1)*chunk_len); n_sets = 0 is a no-op.
 *                                                                               evaluation.rs:364-444 */
int h2b_evaluate_h_permutation(h2b_domain* dom, const h2b_eval_columns* cols, const uint32_t* column_type,
                               const uint32_t* column_index, uint32_t n_columns,
                               const h2b_fr* const* sigma_cosets, const h2b_fr* const* product_cosets,
                               uint32_t n_sets, uint32_t chunk_len, uint32_t blinding_factors,
                               const h2b_fr* l0, const h2b_fr* l_last, const h2b_fr* l_active_row,
                               h2b_fr* values);
/* one lookup argument: table_value = lookup_graph.evaluate(previous_value = 0), then the five lookup
 * constraints folded into `values` with y.                                      evaluation.rs:446-519 */
int h2b_evaluate_h_lookup(h2b_domain* dom, h2b_graph* graph, const h2b_eval_columns* cols,
                          const h2b_fr* product_coset, const h2b_fr* permuted_input_coset,
                          const h2b_fr* permuted_table_coset, const h2b_fr* l0, const h2b_fr* l_last,
                          const h2b_fr* l_active_row, h2b_fr* values);

/* ---- device pieces of create_proof between the transforms and the commitments -------------------- */
/* out[i] = (wide[i] as a 512-bit little-endian integer) mod r, Montgomery form: Fr::from_u512, the body of
 * Fr::random (eight rng.next_u64() draws, low limb first) and Fr::from_bytes_wide (halo2curves 0.3.1).
 * `wide` holds 8 u64 per element (host or device per `loc`); `out_dev` is a device array. */
int h2b_fr_from_u512(h2b_ctx* ctx, const uint64_t* wide, int loc, size_t n, h2b_fr* out_dev);
/* n draws of Fr::random from the counter-mode RngCore of the host mirror (CounterRng: word w >= 1 of the stream
 * is the splitmix64 finaliser of seed + w * 0x9E3779B97F4A7C15), starting after `ctr` words already drawn:
 * the 2^k draws of the vanishing argument's random polynomial without a host round trip
 * (plonk/vanishing/prover.rs:49-53).  Same values as h2b_fr_from_u512 over the host-generated words. */
int h2b_fr_random_counter(h2b_ctx* ctx, uint64_t seed, uint64_t ctr, size_t n, h2b_fr* out_dev);
/* The permutation argument's per-row fractions for one chunk of columns
 *   out[i] = prod_j (v_j[i] + delta^(first_column+j) omega^i beta + gamma) / (v_j[i] + beta sigma_j[i] + gamma)
 * values / sigma: n_cols device arrays of 2^k Lagrange values; out_dev: 2^k elements.
 *                                                                  plonk/permutation/prover.rs:96-144 */
int h2b_permutation_fractions(h2b_domain* dom, const h2b_fr* const* values, const h2b_fr* const* sigma,
                              uint32_t n_cols, uint32_t first_column, const h2b_fr* beta, const h2b_fr* gamma,
                              h2b_fr* out_dev);
/* permute_expression_pair on the first `usable_rows` elements of two device columns: the input sorted
 * ascending (Ord for Fr = canonical integer order), each first occurrence's value beside it in the table
 * column, the unused table values spread over the repeated rows (largest row first, values ascending).
 * H2B_ERR_CONSTRAINT if an input value is missing from the table.  The caller appends the blinding rows.
 *                                                                          plonk/lookup/prover.rs:390-475 */
int h2b_lookup_permute(h2b_ctx* ctx, const h2b_fr* input_dev, const h2b_fr* table_dev, size_t usable_rows,
                       h2b_fr* permuted_input_dev, h2b_fr* permuted_table_dev);
/* out[i] = (compressed_input[i] + beta)(compressed_table[i] + gamma) / ((permuted_input[i] + beta)
 * (permuted_table[i] + gamma)), device arrays of n elements.                plonk/lookup/prover.rs:160-191 */
int h2b_lookup_product_fractions(h2b_ctx* ctx, const h2b_fr* permuted_input, const h2b_fr* permuted_table,
                                 const h2b_fr* compressed_input, const h2b_fr* compressed_table,
                                 const h2b_fr* beta, const h2b_fr* gamma, size_t n, h2b_fr* out_dev);
/* acc = acc * a + p * b on device arrays (poly * F + &poly: vanishing/prover.rs:131-135, gwc/prover.rs:62-76) */
int h2b_poly_fma(h2b_ctx* ctx, h2b_fr* acc_dev, const h2b_fr* a, const h2b_fr* p_dev, const h2b_fr* b, size_t n);

/* Four-step pieces for ONE transform sharded over several GPUs (device pointers
 * only; no counterpart in the reference, which is single-process).  The host
 * side (halo2-pse_b200/dist.py) composes them with an all-to-all over NCCL:
 *   transpose:  out[b][c][r] = in[b*in_batch_stride + r*in_row_stride + c]
 *   permute3:   out[b][a][c] = in[a][b][c]
 *   twiddle:    a[r][c] *= omega^((row0 + r) * c), omega of order 2^log_n */
int h2b_fr_transpose_batch(h2b_ctx* ctx, const h2b_fr* in, h2b_fr* out, uint32_t rows, uint32_t cols,
                           size_t in_row_stride, uint32_t nbatch, size_t in_batch_stride,
                           size_t out_batch_stride);
/* Transpose fused with the exchange step: `in` = this rank's rows of a global
 * (world*rows_local) x cols matrix; element (r, c) is stored directly into
 * peer_out[c / (cols/world)] (NVLink-mapped device pointers of every rank's destination
 * buffer, own rank included) at its place in the row-sharded transposed matrix. */
int h2b_fr_transpose_scatter(h2b_ctx* ctx, const h2b_fr* in, void* const* peer_out, uint32_t world,
                             uint32_t rank, uint32_t rows_local, uint32_t cols);
/* Row transforms of the four-step NTT fused with the distributed transpose that follows them: nrows best_fft's of
 * 2^log_n points (rows_dev, device, untouched); output Ko of row r, multiplied by big_omega^((row0 + r) * Ko) when
 * big_omega is not NULL, is stored into peer_out[Ko / cl][(Ko % cl) * total_rows + row0 + r], cl = 2^log_n / world.
 * peer_out: `world` device pointers (this rank's own buffer included), NVLink-mapped for the peers. */
int h2b_best_fft_rows_scatter(h2b_ctx* ctx, const h2b_fr* rows_dev, const h2b_fr* omega, uint32_t log_n,
                              uint32_t nrows, void* const* peer_out, uint32_t world, uint64_t row0,
                              uint64_t total_rows, const h2b_fr* big_omega, uint32_t big_log_n);
int h2b_fr_permute3(h2b_ctx* ctx, const h2b_fr* in, h2b_fr* out, uint32_t A, uint32_t B, uint32_t C);
int h2b_fr_twiddle_rows(h2b_ctx* ctx, h2b_fr* a, const h2b_fr* omega, uint32_t log_n, uint64_t row0,
                        uint32_t nrows, uint32_t ncols);

/* ---- wire formats of G1 points: SerdeFormat of helpers.rs:8-52, for ParamsKZG::read_custom / write_custom
 * (poly/kzg/commitment.rs:142-244).  RawBytes[Unchecked] = the Montgomery limbs as they are (h2b_bases_upload
 * takes them unchanged; RawBytes also checks the curve equation); Processed = G1Affine::to_bytes / from_bytes:
 * 32-byte little-endian canonical x, parity of y in bit `sign_bit` of byte 31 (halo2curves 0.3.1: 7),
 * identity = zeros.  *all_valid = 0 if some point is off the curve / not canonical. ---- */
int h2b_g1_check_on_curve(h2b_ctx* ctx, const h2b_g1_affine* pts_dev, size_t n, int* all_valid);
int h2b_g1_compress(h2b_ctx* ctx, const h2b_g1_affine* pts_dev, size_t n, uint32_t sign_bit, uint8_t* out_host);
int h2b_g1_decompress(h2b_ctx* ctx, const uint8_t* in_host, size_t n, uint32_t sign_bit, h2b_g1_affine* out_dev,
                      int* all_valid);

/* ---- device helpers for callers that keep data resident ------------------ */
int h2b_device_alloc(h2b_ctx* ctx, size_t bytes, void** out);
void h2b_device_free(h2b_ctx* ctx, void* p);
/* pinned (page-locked) host staging memory */
int h2b_host_alloc(size_t bytes, void** out);
void h2b_host_free(void* p);
int h2b_copy_h2d(h2b_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int h2b_copy_d2h(h2b_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int h2b_copy_d2d(h2b_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
int h2b_device_memset(h2b_ctx* ctx, void* dst_dev, int value, size_t bytes);

/* Synthetic benchmark inputs, generated on device:
 * uniform Fr in Montgomery form from a counter-based generator; and n valid
 * distinct G1 points P_i = [h_i] G with h_i = h2b_synth_base_scalar(seed, i). */
int h2b_synth_scalars(h2b_ctx* ctx, h2b_fr* dst_dev, size_t n, uint64_t seed, uint32_t kind);
int h2b_synth_bases(h2b_ctx* ctx, h2b_g1_affine* dst_dev, size_t n, uint64_t seed);
/* discrete log h_i of synthetic base i (P_i = [h_i] G): closed-form MSM checks */
uint64_t h2b_synth_base_scalar(uint64_t seed, uint64_t i);

/* ---- measurement / test hooks ------------------------------------------- */
/* Register-only integer-pipe microbenchmarks (the roofline denominators):
 * which = 0 IMAD (mad.lo.u32), 1 IMAD.HI, 2 IMAD.WIDE (mad.wide.u32),
 * 3 the mad.lo.cc/madc.hi.cc carry-chain pattern, 4 Fr Montgomery multiplications
 * (counted as 136 32x32 multiplies each).  Returns 32x32 multiplies/s and
 * instructions (or mulmods, for 4) per second, best of 6 launches. */
int h2b_pipe_peak(h2b_ctx* ctx, int which, double* mults_per_s, double* instr_per_s);
/* When on, MSM / NTT calls record CUDA events on the context's stream around
 * their dominant kernel; h2b_ctx_last_kernel_ms returns that duration. */
void h2b_ctx_set_profile(h2b_ctx* ctx, int on);
float h2b_ctx_last_kernel_ms(const h2b_ctx* ctx);
/* durations (ms) of the passes of the last NTT call made in profile mode; returns their number */
int h2b_ctx_last_ntt_passes(h2b_ctx* ctx, float* ms, int cap);
/* Element-wise device ops (op: 0 mul, 1 add, 2 sub, 3 sqr, 4 to_mont, 5 from_mont,
 * 6 neg, 7 inv, 8 mul_sub(a, b, a + b, a - b), 9 mul_shoup with b as the fixed multiplier); field: 0 Fr, 1 Fq.
 * Host pointers. */
int h2b_test_field_op(h2b_ctx* ctx, int field, int op, const h2b_fr* a, const h2b_fr* b,
                      h2b_fr* out, size_t n);
/* Same ops on the host code path of the same header (no device needed). */
int h2b_host_field_op(int field, int op, const h2b_fr* a, const h2b_fr* b, h2b_fr* out, size_t n);
/* Device group law: op 0: out[i] = a[i] + b[i] (affine + affine via XYZZ),
 * 1: out[i] = 2*a[i].  Host pointers; results normalised to affine. */
int h2b_test_g1_op(h2b_ctx* ctx, int op, const h2b_g1_affine* a, const h2b_g1_affine* b,
                   h2b_g1_affine* out, size_t n);
int h2b_host_g1_op(int op, const h2b_g1_affine* a, const h2b_g1_affine* b, h2b_g1_affine* out,
                   size_t n);

#ifdef __cplusplus
}
#endif
#endif /* HALO2_B200_H */
