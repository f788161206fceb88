// Shared host-side plumbing of libhalo2b200: context, scratch, error handling,
// the kernel-launch helper and the vectorised global-memory accessors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/halo2_b200.h"
#include "ec.cuh"
#include "field.cuh"

namespace h2b {

// Twiddle tables of one (omega, log_n) pair, all on device (ntt.cu).
struct TwTable {
  Fr omega;
  uint32_t k;      // omega has exact order 2^k
  uint32_t h;      // split of the two-level table: e = (e_hi << h) | e_lo
  Fr* d_lo;        // omega^i,           i < 2^h
  Fr* d_hi;        // omega^(i << h),    i < 2^(k-h)
  Fr* d_rt;        // omega_R^i, R = 2^min(k,9), i < R   (intra-pass roots)
  Fr* d_rts = nullptr;  // the same roots for mul_shoup: [2i] = plain value, [2i + 1] = floor(value * 2^256 / r)
  // single-multiplication inter-pass twiddles (optional; nullptr -> two-level lo/hi product)
  Fr* d_out = nullptr;   // first pass: w^(jr * K) at output index K * m2 + jr, n entries (optional)
  uint32_t out_s1 = 0;   // first digit width the table was built for
  Fr* d_mid = nullptr;   // (omega^(2^(k-mid_log)))^i, i < 2^mid_log               (later passes)
  uint32_t mid_log = 0;
};

struct MsmWorkspace;

}  // namespace h2b

struct h2b_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // second stream + events for host->device copies overlapped with compute (drop-in calls on host slices)
  cudaStream_t copy_stream = nullptr;
  // device->host leg of streamed host batches (PCIe is full duplex: it runs beside copy_stream's uploads)
  cudaStream_t d2h_stream = nullptr;
  // `stream` is created at the highest priority; the one long throughput-bound kernel of an MSM (bucket
  // accumulation, level 0) is launched on this LOW-priority stream instead, fenced by two events: when several
  // contexts commit at once, the short latency-bound kernels of one MSM (sort passes, scans, bucket reduction)
  // get the SM slots that free up first and run under the accumulation of another MSM instead of behind it
  cudaStream_t bulk_stream = nullptr;
  cudaEvent_t bulk_ev[2] = {nullptr, nullptr};
  cudaEvent_t copy_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // host-scalar MSMs in batches: timing events around every batch's copy (copy stream) and compute (main stream),
  // created on first use; their ratio (time to copy a point / time to accumulate it, running average) sizes the
  // batches of the next call -- 0.25 alone on PCIe 5, 0.6 when eight ranks share the host's memory path
  cudaEvent_t plan_ev[28] = {nullptr};
  float msm_copy_ratio = 0.f;
  std::recursive_mutex mu;
  std::string last_error;
  // grow-only device scratch (NTT ping buffer)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // grow-only device staging for host-pointer calls
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes[2] = {0, 0};
  std::vector<h2b::TwTable> tw;
  h2b::MsmWorkspace* msm_ws = nullptr;
  uint64_t launches = 0;  // kernels launched by this library (bench.py reports it)
  // timing of the last MSM / NTT call's dominant kernel (CUDA events on `stream`)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_kernel_ms = 0.f;
  int profile = 0;
  // per-pass timing of the last NTT call (profile mode): events around every pass launch
  cudaEvent_t pass_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float last_pass_ms[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  int last_npass = 0;
  bool ntt_attr_done = false;
  // pinned ring for copies from / to PAGEABLE host slices (a Rust Vec, a numpy array): kCopySlots
  // slots of kCopyChunk bytes, filled by a few host threads while the DMA engine drains them
  void* copy_slot[16] = {nullptr};
  cudaEvent_t copy_slot_ev[16] = {nullptr};
  bool copy_pool_ready = false;
  struct CopyWorkers* copy_workers = nullptr;  // persistent host threads of the staged copies (ctx.cu)
  // h2b_device_alloc / h2b_device_free: sizes of the live blocks and a per-size cache of released ones.  A prover
  // asks for the same few sizes (n and 2^extended_k elements) dozens of times per proof; a released block is
  // handed to the next request of its size without a trip through the driver's pool (cudaMallocAsync answered
  // in 1 - 8 ms now and then).  Stream-ordered like cudaMallocAsync: valid for work enqueued on `stream`.
  std::unordered_map<void*, size_t> block_size;
  std::unordered_map<size_t, std::vector<void*>> block_cache;
  size_t block_cache_bytes = 0;
};

struct h2b_bases {
  h2b_ctx* ctx;
  h2b::G1Affine* d_pts;
  size_t n;
  // optional window table (h2b_bases_precompute): d_table[w * n + i] = 2^(pre_c * w) * d_pts[i]
  h2b::G1Affine* d_table = nullptr;
  uint32_t pre_c = 0, pre_W = 0;
};

struct h2b_domain {
  h2b_ctx* ctx;
  uint32_t j, k, extended_k, quotient_poly_degree;
  h2b::Fr omega, omega_inv, extended_omega, extended_omega_inv;
  h2b::Fr g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor;
  std::vector<h2b::Fr> t_evaluations;  // already inverted (domain.rs:84-124)
  // device-side small tables
  h2b::Fr* d_zeta_in;    // [1, zeta, zeta^2]                 (coeff_to_extended pre-scale)
  h2b::Fr* d_ext_post;   // 1/2^ek * [1, zeta^2, zeta]        (extended_to_coeff post-scale)
  h2b::Fr* d_ifft_post;  // [1/n]
  h2b::Fr* d_t_inv;      // t_evaluations
};

namespace h2b {

inline int fail(h2b_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

#define H2B_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      char _b[512];                                                                      \
      snprintf(_b, sizeof _b, "%s:%d %s: %s", __FILE__, __LINE__, #expr,                 \
               cudaGetErrorString(_e));                                                  \
      cudaGetLastError(); /* clear the non-sticky error so that later launches are not blamed */ \
      return h2b::fail((ctx), _e == cudaErrorMemoryAllocation ? H2B_ERR_OOM : H2B_ERR_CUDA, _b); \
    }                                                                                    \
  } while (0)

#define H2B_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != H2B_OK) return _r; \
  } while (0)

// Launch `kern` on the context's stream and count it.  Under H2B_EMU (the CPU
// test build, tests/emu/cuda_runtime.h) the same call runs the kernel on the
// fiber emulator; the product build never defines H2B_EMU.
template <class... KArgs, class... Args>
inline int launch_on(h2b_ctx* ctx, cudaStream_t stream, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                     Args&&... args) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return H2B_OK;
#ifdef H2B_EMU
  (void)stream;
  emu::launch(grid, block, smem, [&]() { kern(args...); });
#else
  kern<<<grid, block, smem, stream>>>(std::forward<Args>(args)...);
  H2B_CUDA(ctx, cudaGetLastError());
#endif
  ctx->launches++;
  return H2B_OK;
}
template <class... KArgs, class... Args>
inline int launch(h2b_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                  Args&&... args) {
  return launch_on(ctx, ctx->stream, kern, grid, block, smem, std::forward<Args>(args)...);
}

#ifdef H2B_EMU
#define H2B_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#else
#define H2B_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// 256-bit global accessors (LDG.E.256 / STG.E.256 on sm_100a).  Pointers must
// be 32-byte aligned: every device buffer of this library is.
template <class P>
H2B_D Fp<P> ld_fp(const Fp<P>* p) {
  Fp<P> r;
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                 "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
#else
  r = *p;
#endif
  return r;
}
// read-only path (tables)
template <class P>
H2B_D Fp<P> ld_fp_nc(const Fp<P>* p) {
  Fp<P> r;
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                 "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
#else
  r = *p;
#endif
  return r;
}
// the same with a 64-byte L2 fill: isolated 64-byte gathers (a table point) should not pull the other half of the line
template <class P>
H2B_D Fp<P> ld_fp_nc64(const Fp<P>* p) {
  Fp<P> r;
#ifdef __CUDA_ARCH__
  asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                 "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
#else
  r = *p;
#endif
  return r;
}
template <class P>
H2B_D void st_fp(Fp<P>* p, const Fp<P>& r) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]),
               "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]),
               "l"(p)
               : "memory");
#else
  *p = r;
#endif
}

// cudaMalloc for the library's own long-lived buffers (workspaces, scratch, tables): on an out-of-memory answer the
// context's cache of released caller blocks and the stream-ordered pool are given back to the driver and the
// allocation is tried once more -- gigabytes can sit idle there.  Rule for the block cache: a block may
// only be released (h2b_device_free) after every stream that touched it has been fenced into ctx->stream; every entry
// point of this library ends in such a fence.
cudaError_t dev_malloc(h2b_ctx* ctx, void** out, size_t bytes);
int ensure_scratch(h2b_ctx* ctx, size_t bytes);
int ensure_stage(h2b_ctx* ctx, int which, size_t bytes);
// Host <-> device copies enqueued on `stream` that do not depend on the host slice being pinned:
// pinned (or small) slices take one cudaMemcpyAsync; pageable ones are staged through the context's
// pinned ring by several host threads (the driver's own pageable path is one memcpy thread, ~10 GB/s).
// h2d returns when every chunk is staged (the DMA may still be in flight on `stream`, as with
// cudaMemcpyAsync); d2h returns when the bytes are in `dst`.
int copy_h2d_any(h2b_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream);
int copy_d2h_any(h2b_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t stream);

// ntt.cu
int ntt_get_table(h2b_ctx* ctx, const Fr& omega, uint32_t log_n, const TwTable** out);
// sc != nullptr: the last pass of every row transform stores output Ko of row (row0 + batch member), times
// tw's omega^(row * Ko) if tw != nullptr, into peers[Ko / cl][(Ko % cl) * R + row] instead of d_out (four-step NTT:
// the row transforms fused with the distributed transpose that follows them).
struct NttScatter {
  Fr* peers[16];
  uint32_t cl;
  uint64_t R, row0;
  const TwTable* tw;
};
int ntt_run(h2b_ctx* ctx, const Fr* d_in, Fr* d_out, uint32_t log_n, const TwTable* tw,
            uint64_t n_in, const Fr* d_pre, uint32_t pre_mod, const Fr* d_post,
            uint32_t post_mod, uint64_t n_out, uint32_t batch, uint64_t in_stride,
            uint64_t out_stride, const NttScatter* sc = nullptr);
void ntt_free_tables(h2b_ctx* ctx);

// msm.cu
// table_stride == 0: classic per-window buckets over `d_points` (n points).
// table_stride  > 0: `d_points` is a window table (w * table_stride + i), all windows share one bucket set.
// h_scalars != nullptr: the scalars still live on the host; they are copied into d_scalars in chunks
// on the copy stream while the digit kernel consumes the chunks already there.
// ncols > 1 (window table, device scalars only): col_scalars[j] (HOST array of device pointers) are ncols
// scalar vectors over the same bases; one digit/sort/accumulate/reduce pipeline with one bucket set per
// column, out_host[0..ncols) receives the ncols sums.
int msm_run(h2b_ctx* ctx, const G1Affine* d_points, const Fr* d_scalars, size_t n, G1Xyzz* out_host,
            size_t table_stride = 0, uint32_t table_c = 0, const Fr* h_scalars = nullptr, uint32_t ncols = 1,
            const Fr* const* col_scalars = nullptr);
void msm_ws_free(h2b_ctx* ctx);

}  // namespace h2b
