// Context, device memory helpers, synthetic benchmark inputs and the
// measurement / test hooks of libhalo2b200 (include/halo2_b200.h).
#include "common.cuh"

#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

// Persistent host threads for the staged copies: spawning (and CUDA-initialising) fresh threads per
// call costs more than a 32 MiB copy takes.  Worker t >= 1 runs job(t) of the current generation;
// the caller runs job(0) itself and waits for the rest.
struct CopyWorkers {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  std::function<void(int)> job;
  uint64_t generation = 0;
  int active = 0, pending = 0;
  bool stop = false;

  explicit CopyWorkers(int n, int device) {
    for (int t = 1; t < n; ++t)
      threads.emplace_back([this, t, device] {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
          std::function<void(int)> f;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_job.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            if (t >= active) continue;
            f = job;
          }
          f(t);
          {
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) cv_done.notify_one();
          }
        }
      });
  }
  void run(int T, const std::function<void(int)>& f) {
    {
      std::lock_guard<std::mutex> lk(m);
      job = f;
      active = T;
      pending = T - 1;
      ++generation;
    }
    cv_job.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~CopyWorkers() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_job.notify_all();
    for (auto& t : threads) t.join();
  }
};

namespace h2b {

static int grow(h2b_ctx* ctx, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return H2B_OK;
  if (*p) {
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    H2B_CUDA(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
  }
  H2B_CUDA(ctx, dev_malloc(ctx, p, bytes));
  *cap = bytes;
  return H2B_OK;
}

int ensure_scratch(h2b_ctx* ctx, size_t bytes) {
  return grow(ctx, &ctx->scratch, &ctx->scratch_bytes, bytes);
}
int ensure_stage(h2b_ctx* ctx, int which, size_t bytes) {
  return grow(ctx, &ctx->stage[which], &ctx->stage_bytes[which], bytes);
}

// ---------------------------------------------------------------------------
// Copies from / to pageable host memory through a pinned ring
// ---------------------------------------------------------------------------
static constexpr size_t kCopySlotBytes = 2u << 20;  // bytes per slot
static size_t copy_chunk() {  // bytes staged per DMA: small enough that the first DMA starts early
  static size_t v = [] {
    size_t kb = 1024;
    if (const char* e = getenv("H2B_COPY_CHUNK_KB")) kb = (size_t)atoi(e);
    if (kb < 64) kb = 64;
    if (kb << 10 > kCopySlotBytes) kb = kCopySlotBytes >> 10;
    return kb << 10;
  }();
  return v;
}
static constexpr int kCopySlots = 16;           // 2 per thread, up to 8 threads
static constexpr size_t kCopyStagedMin = 4u << 20;

static int copy_threads() {
  static int n = [] {
    int v = 0;
    if (const char* e = getenv("H2B_COPY_THREADS")) v = atoi(e);
    if (v <= 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      v = hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
    }
    return v > kCopySlots / 2 ? kCopySlots / 2 : v;
  }();
  return n;
}

static bool host_is_pageable(const void* p) {
#ifdef H2B_EMU
  (void)p;
  return getenv("H2B_EMU_STAGED_COPY") != nullptr;  // the emulator has no pinned memory: opt in (tests)
#else
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
#endif
}

static int copy_pool_init(h2b_ctx* ctx) {
  if (ctx->copy_pool_ready) return H2B_OK;
  for (int i = 0; i < kCopySlots; ++i) {
    H2B_CUDA(ctx, cudaMallocHost(&ctx->copy_slot[i], kCopySlotBytes));
    H2B_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_slot_ev[i], cudaEventDisableTiming));
  }
  ctx->copy_workers = new CopyWorkers(copy_threads(), ctx->device);
  ctx->copy_pool_ready = true;
  return H2B_OK;
}

template <bool TO_DEVICE>
static int copy_staged(h2b_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  H2B_TRY(copy_pool_init(ctx));
  const size_t kCopyChunk = copy_chunk();
  const size_t nchunks = (bytes + kCopyChunk - 1) / kCopyChunk;
  // few threads for small copies (wake-ups cost more than they carry), all of them from ~64 MiB on
  const int T = (int)std::min<size_t>((size_t)copy_threads(),
                                      nchunks >= 64 ? nchunks : std::min<size_t>(4, std::max<size_t>(2, nchunks / 4)));
  cudaError_t errs[kCopySlots / 2];
  auto work = [&](int t) {
    cudaError_t e = cudaSuccess;
    size_t it = 0;
    // d2h: the host-side memcpy of a chunk runs one iteration behind its DMA
    size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};
    for (size_t c = (size_t)t; c < nchunks && e == cudaSuccess; c += (size_t)T, ++it) {
      const int slot = 2 * t + (int)(it & 1);
      const size_t off = c * kCopyChunk, len = std::min(kCopyChunk, bytes - off);
      if (it >= 2 || !TO_DEVICE) e = cudaEventSynchronize(ctx->copy_slot_ev[slot]);  // slot drained / filled
      if (e != cudaSuccess) break;
      if (TO_DEVICE) {
        memcpy(ctx->copy_slot[slot], (const char*)src + off, len);
        e = cudaMemcpyAsync((char*)dst + off, ctx->copy_slot[slot], len, cudaMemcpyHostToDevice, stream);
      } else {
        if (pend_len[it & 1]) memcpy((char*)dst + pend_off[it & 1], ctx->copy_slot[slot], pend_len[it & 1]);
        e = cudaMemcpyAsync(ctx->copy_slot[slot], (const char*)src + off, len, cudaMemcpyDeviceToHost, stream);
        pend_off[it & 1] = off;
        pend_len[it & 1] = len;
      }
      if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_slot_ev[slot], stream);
    }
    if (!TO_DEVICE)
      for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        const size_t j = it + (size_t)b;  // the two chunks still in flight, oldest first
        if (!pend_len[j & 1]) continue;
        e = cudaEventSynchronize(ctx->copy_slot_ev[2 * t + (int)(j & 1)]);
        if (e == cudaSuccess) memcpy((char*)dst + pend_off[j & 1], ctx->copy_slot[2 * t + (int)(j & 1)], pend_len[j & 1]);
        pend_len[j & 1] = 0;
      }
    errs[t] = e;
  };
  ctx->copy_workers->run(T, work);
  for (int t = 0; t < T; ++t)
    if (errs[t] != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, H2B_ERR_CUDA, std::string("staged copy: ") + cudaGetErrorString(errs[t]));
    }
  if (TO_DEVICE) {
    // the ring is reused by the next call: its slots must be drained before they are overwritten, which
    // the per-slot events guarantee (first two uses of a slot per call skip the wait only if the previous
    // call's events completed -- make that true here)
    for (int t = 0; t < T; ++t)
      for (int b = 0; b < 2; ++b) H2B_CUDA(ctx, cudaEventSynchronize(ctx->copy_slot_ev[2 * t + b]));
  }
  return H2B_OK;
}

int copy_h2d_any(h2b_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return H2B_OK;
  if (bytes < kCopyStagedMin || copy_threads() < 2 || !host_is_pageable(src_host)) {
    H2B_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
    return H2B_OK;
  }
  return copy_staged<true>(ctx, dst_dev, src_host, bytes, stream);
}

int copy_d2h_any(h2b_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return H2B_OK;
  if (bytes < kCopyStagedMin || copy_threads() < 2 || !host_is_pageable(dst_host)) {
    H2B_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
    return H2B_OK;
  }
  return copy_staged<false>(ctx, dst_host, src_dev, bytes, stream);
}

// ---------------------------------------------------------------------------
// Synthetic inputs
// ---------------------------------------------------------------------------
H2B_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// canonical value < r from a counter-based generator: 254 random bits,
// rejection-sampled (acceptance ~ 0.76), attempt number mixed into the counter
H2B_HD Fr uniform_fr_canonical(uint64_t seed, uint64_t i) {
  Fr x;
  for (uint32_t attempt = 0;; ++attempt) {
    for (int w = 0; w < 4; ++w) {
      const uint64_t r = splitmix64(seed ^ splitmix64(i * 4 + w + ((uint64_t)attempt << 58)));
      x.v[2 * w] = (uint32_t)r;
      x.v[2 * w + 1] = (uint32_t)(r >> 32);
    }
    x.v[7] &= 0x3fffffffu;
    uint32_t m[8], t[8];
    for (int k = 0; k < 8; ++k) m[k] = FrParams::mod(k);
    if (sub8(t, x.v, m)) return x;  // borrow  =>  x < r
  }
}

// kind: 0 uniform, 1 all equal (one uniform value), 2 0/1 with density 1/2,
//       3 uniform in [0, 2^16), 4 uniform with 90 % zeros
__global__ void synth_scalars_kernel(Fr* dst, uint64_t n, uint64_t seed, uint32_t kind) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Fr x = Fr::zero();
    const uint64_t r = splitmix64(seed + 0x5851f42d4c957f2dull * (i + 1));
    switch (kind) {
      case 0: x = uniform_fr_canonical(seed, i); break;
      case 1: x = uniform_fr_canonical(seed, 0); break;
      case 2: x.v[0] = (uint32_t)(r & 1); break;
      case 3: x.v[0] = (uint32_t)(r & 0xffff); break;
      default:
        if (r % 10 == 0) x = uniform_fr_canonical(seed, i);
    }
    st_fp(dst + i, to_mont(x));
  }
}

H2B_HD uint64_t synth_base_scalar(uint64_t seed, uint64_t i) { return splitmix64(seed + i) | 1ull; }

// P_i = [h_i] G with h_i = splitmix64(seed + i) | 1 : valid, (practically)
// distinct points whose discrete logs are known, so that
// sum c_i P_i = [sum c_i h_i] G gives a closed-form check at any size.
__global__ void synth_bases_kernel(G1Affine* dst, uint64_t n, uint64_t seed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = synth_base_scalar(seed, i);
  G1Affine g;
  g.x = Fq::one();
  g.y = add(Fq::one(), Fq::one());
  G1Xyzz acc = G1Xyzz::identity();
  for (int bit = 63; bit >= 0; --bit) {
    acc = xyzz_double(acc);
    if ((h >> bit) & 1) xyzz_add_affine(acc, g);
  }
  const G1Affine a = xyzz_to_affine(acc);
  st_fp(&dst[i].x, a.x);
  st_fp(&dst[i].y, a.y);
}

// ---------------------------------------------------------------------------
// Integer-pipe peaks: register-only microbenchmarks (the roofline denominators
// of the MSM / NTT kernels).  16 independent chains per thread, loop-invariant
// multiplicands, nothing but the measured instruction in the loop body.
//   which = 0  IMAD      (mad.lo.u32,   32x32 -> low 32  + 32)
//           1  IMAD.HI   (mad.hi.u32,   32x32 -> high 32 + 32)
//           2  IMAD.WIDE (mad.wide.u32, 32x32 -> 64      + 64)
//           3  the carry-chain pattern of field.cuh (mad.lo.cc / madc.hi.cc, 8 per chain)
//           4  Fr Montgomery multiplications (field.cuh mul), counted as 136 multiplies each
// ---------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) pipe_peak_kernel(uint32_t* sink, uint32_t iters, uint32_t a0) {
  uint32_t x = a0 | 1u, y = (a0 * 2654435761u) | 1u;
  if (WHICH <= 2) {
    uint64_t acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = ((uint64_t)threadIdx.x << 20) + j * 977u + a0;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#ifdef __CUDA_ARCH__
          uint32_t lo = (uint32_t)acc[j];
          if (WHICH == 0) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(x), "r"(y));
            acc[j] = lo;
          } else if (WHICH == 1) {
            asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(x), "r"(y));
            acc[j] = lo;
          } else {
            // 32 x 32 -> 64 with operands that change every iteration (the low word of this accumulator times the
            // high word of its neighbour): IMAD.WIDE.U32 Rd, Ra, Rb, RZ in SASS.  Round 1 multiplied two
            // loop-invariant registers here; ptxas hoisted the product and the loop timed 64-bit ADDS
            // (IADD3 + IADD3.X, 64 lanes/clk/SM), which was then quoted as the IMAD.WIDE peak.
            const uint32_t other = (uint32_t)(acc[(j + 1) & 15] >> 32);
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"(lo), "r"(other));
          }
#else
          acc[j] += (uint64_t)x * y;
#endif
        }
      }
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s ^= acc[j];
    if (s == 0x12345678ull) sink[0] = (uint32_t)s;
  } else if (WHICH == 3) {
    uint32_t c[2][9];
#pragma unroll
    for (int j = 0; j < 9; ++j) c[0][j] = c[1][j] = threadIdx.x + j + a0;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 4; ++rep) {
        chain_mad_top<false>(c[0][0], c[0][1], c[0][2], c[0][3], c[0][4], c[0][5], c[0][6], c[0][7],
                             c[0][8], x, y, x + 2, y + 2, x + 4, 0u, 0u);
        chain_mad_top<false>(c[1][0], c[1][1], c[1][2], c[1][3], c[1][4], c[1][5], c[1][6], c[1][7],
                             c[1][8], y, x, y + 2, x + 2, y + 4, 0u, 0u);
      }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) s ^= c[0][j] ^ c[1][j];
    if (s == 0x12345678u) sink[0] = s;
  } else {
    Fr a[2], b;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[0].v[j] = threadIdx.x + j + a0;
      a[1].v[j] = threadIdx.x * 3 + j + a0;
      b.v[j] = FrParams::one(j) ^ (a0 & 0xff);
    }
    a[0].v[7] &= 0x0fffffffu;
    a[1].v[7] &= 0x0fffffffu;
    b.v[7] &= 0x0fffffffu;
    for (uint32_t it = 0; it < iters; ++it) {
      a[0] = mul(a[0], b);
      a[1] = mul(a[1], b);
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= a[0].v[j] ^ a[1].v[j];
    if (s == 0x12345678u) sink[0] = s;
  }
}

// ---------------------------------------------------------------------------
// Element-wise test kernels
// ---------------------------------------------------------------------------
template <class F>
H2B_HD F field_op(int op, const F& a, const F& b) {
  switch (op) {
    case 0: return mul(a, b);
    case 1: return add(a, b);
    case 2: return sub(a, b);
    case 3: return sqr(a);
    case 4: return to_mont(a);
    case 5: return from_mont(a);
    case 6: return neg(a);
    case 8: return mul_sub(a, b, add(a, b), sub(a, b));  // a b - (a + b)(a - b), two products on one reduction
    case 9: return mul_shoup(a, from_mont(b), shoup_companion(b));  // b as a fixed multiplier: = mul(a, b)
    default: return inv(a);
  }
}

__global__ void field_op_kernel(int field, int op, const Fr* a, const Fr* b, Fr* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (field == 0) {
    st_fp(out + i, field_op<Fr>(op, ld_fp(a + i), ld_fp(b + i)));
  } else {
    const Fq* qa = reinterpret_cast<const Fq*>(a);
    const Fq* qb = reinterpret_cast<const Fq*>(b);
    st_fp(reinterpret_cast<Fq*>(out) + i, field_op<Fq>(op, ld_fp(qa + i), ld_fp(qb + i)));
  }
}

H2B_HD G1Affine g1_op(int op, const G1Affine& a, const G1Affine& b) {
  G1Xyzz acc = G1Xyzz::from_affine(a);
  if (op == 0) {
    xyzz_add_affine(acc, b);
  } else if (op == 1) {
    acc = xyzz_double(acc);
  } else {
    // full (non-mixed) addition of 2a and b, minus a:  2a + b - a  = a + b
    G1Xyzz t = xyzz_double(acc);
    G1Xyzz bb = G1Xyzz::from_affine(b);
    xyzz_add(t, bb);
    G1Xyzz na = G1Xyzz::from_affine(g1_neg(a));
    xyzz_add(t, na);
    acc = t;
  }
  return xyzz_to_affine(acc);
}

__global__ void g1_op_kernel(int op, const G1Affine* a, const G1Affine* b, G1Affine* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = g1_op(op, a[i], b[i]);
}

}  // namespace h2b

using namespace h2b;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" int h2b_ctx_create(int device, h2b_ctx** out) {
  if (!out) return H2B_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return H2B_ERR_CUDA;
  if (device < 0 || device >= count) return H2B_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return H2B_ERR_CUDA;
  h2b_ctx* ctx = new h2b_ctx();
  ctx->device = device;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
    ctx->sm_count = sms;
  int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
#ifndef H2B_EMU
  if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) prio_lo = prio_hi = 0;
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->bulk_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) {
    delete ctx;
    return H2B_ERR_CUDA;
  }
#else
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return H2B_ERR_CUDA;
  }
  ctx->bulk_stream = ctx->stream;
#endif
  for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&ctx->bulk_ev[i], cudaEventDisableTiming);
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return H2B_ERR_CUDA;
  }
#ifndef H2B_EMU
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = 16ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
#endif
  for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming);
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 6; ++i) cudaEventCreate(&ctx->pass_ev[i]);
  *out = ctx;
  return H2B_OK;
}

static void block_cache_flush(h2b_ctx* ctx);

extern "C" void h2b_ctx_destroy(h2b_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  block_cache_flush(ctx);
  cudaStreamSynchronize(ctx->stream);
  msm_ws_free(ctx);
  ntt_free_tables(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  for (int i = 0; i < 2; ++i)
    if (ctx->stage[i]) cudaFree(ctx->stage[i]);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 6; ++i)
    if (ctx->pass_ev[i]) cudaEventDestroy(ctx->pass_ev[i]);
  for (int i = 0; i < 28; ++i)
    if (ctx->plan_ev[i]) cudaEventDestroy(ctx->plan_ev[i]);
  for (int i = 0; i < 8; ++i)
    if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
  delete ctx->copy_workers;
  for (int i = 0; i < 16; ++i) {
    if (ctx->copy_slot[i]) cudaFreeHost(ctx->copy_slot[i]);
    if (ctx->copy_slot_ev[i]) cudaEventDestroy(ctx->copy_slot_ev[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  for (int i = 0; i < 2; ++i)
    if (ctx->bulk_ev[i]) cudaEventDestroy(ctx->bulk_ev[i]);
  if (ctx->bulk_stream && ctx->bulk_stream != ctx->stream) cudaStreamDestroy(ctx->bulk_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* h2b_last_error(const h2b_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : "null context";
}

extern "C" int h2b_ctx_sync(h2b_ctx* ctx) {
  if (!ctx) return H2B_ERR_ARG;
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" void* h2b_ctx_stream(h2b_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t h2b_ctx_launches(const h2b_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void h2b_ctx_set_profile(h2b_ctx* ctx, int on) {
  if (ctx) ctx->profile = on;
}
extern "C" float h2b_ctx_last_kernel_ms(const h2b_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }
extern "C" int h2b_ctx_last_ntt_passes(h2b_ctx* ctx, float* ms, int cap) {
  if (!ctx || !ms) return 0;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
  int n = 0;
  for (; n < ctx->last_npass && n < cap; ++n) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->pass_ev[n], ctx->pass_ev[n + 1]) != cudaSuccess) break;
    ms[n] = t;
  }
  return n;
}

// Caller-visible device buffers come from the device's stream-ordered memory pool on the context's stream:
// a prover allocates and frees dozens of polynomial-sized buffers per proof, and cudaMalloc / cudaFree would
// each synchronise the whole device.  Freed blocks stay cached in the pool (release threshold 16 GiB).
static constexpr size_t kBlockCacheMax = 24ull << 30;  // bytes kept for reuse per context

static void block_cache_flush(h2b_ctx* ctx) {
  for (auto& kv : ctx->block_cache)
    for (void* p : kv.second) {
#ifndef H2B_EMU
      cudaFreeAsync(p, ctx->stream);
#else
      cudaFree(p);
#endif
    }
  ctx->block_cache.clear();
  ctx->block_cache_bytes = 0;
}

cudaError_t h2b::dev_malloc(h2b_ctx* ctx, void** out, size_t bytes) {
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaErrorMemoryAllocation) return e;
  cudaGetLastError();
  cudaStreamSynchronize(ctx->stream);
  block_cache_flush(ctx);
#ifndef H2B_EMU
  cudaStreamSynchronize(ctx->stream);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
#endif
  return cudaMalloc(out, bytes);
}

extern "C" int h2b_device_alloc(h2b_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  if (bytes == 0) bytes = 1;
  auto hit = ctx->block_cache.find(bytes);
  if (hit != ctx->block_cache.end() && !hit->second.empty()) {
    *out = hit->second.back();
    hit->second.pop_back();
    ctx->block_cache_bytes -= bytes;
    ctx->block_size[*out] = bytes;
    return H2B_OK;
  }
#ifndef H2B_EMU
  cudaError_t e = cudaMallocAsync(out, bytes, ctx->stream);
  if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
    cudaGetLastError();
    block_cache_flush(ctx);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      cudaStreamSynchronize(ctx->stream);
      cudaMemPoolTrimTo(pool, 0);
    }
    e = cudaMallocAsync(out, bytes, ctx->stream);
  }
  H2B_CUDA(ctx, e);
#else
  H2B_CUDA(ctx, cudaMalloc(out, bytes));
#endif
  ctx->block_size[*out] = bytes;
  return H2B_OK;
}

extern "C" void h2b_device_free(h2b_ctx* ctx, void* p) {
  if (!ctx || !p) return;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  auto it = ctx->block_size.find(p);
  if (it != ctx->block_size.end()) {
    const size_t bytes = it->second;
    ctx->block_size.erase(it);
    if (bytes >= (1u << 16) && ctx->block_cache_bytes + bytes <= kBlockCacheMax) {
      ctx->block_cache[bytes].push_back(p);  // ordered after every kernel of this context that may still read it
      ctx->block_cache_bytes += bytes;
      return;
    }
  }
#ifndef H2B_EMU
  cudaFreeAsync(p, ctx->stream);  // ordered after every kernel of this context that may still read it
#else
  cudaStreamSynchronize(ctx->stream);
  cudaFree(p);
#endif
}

extern "C" int h2b_host_alloc(size_t bytes, void** out) {
  if (!out) return H2B_ERR_ARG;
  return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? H2B_OK : H2B_ERR_OOM;
}
extern "C" void h2b_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int h2b_copy_h2d(h2b_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_TRY(copy_h2d_any(ctx, dst_dev, src_host, bytes, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_copy_d2h(h2b_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_TRY(copy_d2h_any(ctx, dst_host, src_dev, bytes, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_copy_d2d(h2b_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_device_memset(h2b_ctx* ctx, void* dst_dev, int value, size_t bytes) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_CUDA(ctx, cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_synth_scalars(h2b_ctx* ctx, h2b_fr* dst_dev, size_t n, uint64_t seed,
                                 uint32_t kind) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!dst_dev) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (kind > 4) return fail(ctx, H2B_ERR_ARG, "unknown scalar distribution");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t want = (n + 255) / 256;
  const uint32_t blocks = (uint32_t)(want < (uint64_t)ctx->sm_count * 32 ? want : (uint64_t)ctx->sm_count * 32);
  return launch(ctx, synth_scalars_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<Fr*>(dst_dev),
                (uint64_t)n, seed, kind);
}

extern "C" int h2b_synth_bases(h2b_ctx* ctx, h2b_g1_affine* dst_dev, size_t n, uint64_t seed) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!dst_dev) return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  return launch(ctx, synth_bases_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0,
                reinterpret_cast<G1Affine*>(dst_dev), (uint64_t)n, seed);
}

extern "C" uint64_t h2b_synth_base_scalar(uint64_t seed, uint64_t i) { return synth_base_scalar(seed, i); }

extern "C" int h2b_pipe_peak(h2b_ctx* ctx, int which, double* mults_per_s, double* instr_per_s) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (which < 0 || which > 4) return fail(ctx, H2B_ERR_ARG, "unknown pipe benchmark");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t* sink;
  H2B_CUDA(ctx, cudaMalloc((void**)&sink, 64));
  const uint32_t blocks = ctx->sm_count * 8, threads = 256;
#ifdef H2B_EMU
  const uint32_t iters = 2;
#else
  const uint32_t iters = which == 4 ? 2048 : 8192;
#endif
  // instructions per thread per iteration, and 32x32 multiplies per instruction
  const double per_iter = which <= 2 ? 64.0 : which == 3 ? 32.0 : 2.0;  // 3: fused lo/hi pairs = IMAD.WIDE
  const double mults_per = which == 4 ? 136.0 : 1.0;
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    H2B_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = H2B_OK;
    const uint32_t a0 = 12345u + rep;
    switch (which) {
      case 0: rc = launch(ctx, pipe_peak_kernel<0>, dim3(blocks), dim3(threads), 0, sink, iters, a0); break;
      case 1: rc = launch(ctx, pipe_peak_kernel<1>, dim3(blocks), dim3(threads), 0, sink, iters, a0); break;
      case 2: rc = launch(ctx, pipe_peak_kernel<2>, dim3(blocks), dim3(threads), 0, sink, iters, a0); break;
      case 3: rc = launch(ctx, pipe_peak_kernel<3>, dim3(blocks), dim3(threads), 0, sink, iters, a0); break;
      default: rc = launch(ctx, pipe_peak_kernel<4>, dim3(blocks), dim3(threads), 0, sink, iters, a0);
    }
    if (rc != H2B_OK) {
      cudaFree(sink);
      return rc;
    }
    H2B_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    H2B_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    H2B_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    const double instr = (double)blocks * threads * iters * per_iter;
    const double rate = ms > 0 ? instr / (ms * 1e-3) : 0;
    if (rate > best) best = rate;
  }
  cudaFree(sink);
  if (instr_per_s) *instr_per_s = best;
  if (mults_per_s) *mults_per_s = best * mults_per;
  return H2B_OK;
}

static int run_elementwise(h2b_ctx* ctx, size_t bytes_each, const void* a, const void* b, void* out,
                           const std::function<int(void*, void*, void*)>& go) {
  void *da = nullptr, *db = nullptr, *dout = nullptr;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  H2B_CUDA(ctx, cudaMalloc(&da, bytes_each + 32));
  H2B_CUDA(ctx, cudaMalloc(&db, bytes_each + 32));
  H2B_CUDA(ctx, cudaMalloc(&dout, bytes_each + 32));
  H2B_CUDA(ctx, cudaMemcpyAsync(da, a, bytes_each, cudaMemcpyHostToDevice, ctx->stream));
  H2B_CUDA(ctx, cudaMemcpyAsync(db, b ? b : a, bytes_each, cudaMemcpyHostToDevice, ctx->stream));
  int rc = go(da, db, dout);
  if (rc == H2B_OK) {
    cudaError_t e = cudaMemcpyAsync(out, dout, bytes_each, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, H2B_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return rc;
}

extern "C" int h2b_test_field_op(h2b_ctx* ctx, int field, int op, const h2b_fr* a, const h2b_fr* b,
                                 h2b_fr* out, size_t n) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a || !out || field < 0 || field > 1 || op < 0 || op > 9) return fail(ctx, H2B_ERR_ARG, "bad argument");
  if (n == 0) return H2B_OK;
  return run_elementwise(ctx, n * 32, a, b, out, [&](void* da, void* db, void* dout) {
    return launch(ctx, field_op_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0, field, op,
                  (const Fr*)da, (const Fr*)db, (Fr*)dout, (uint64_t)n);
  });
}

extern "C" int h2b_host_field_op(int field, int op, const h2b_fr* a, const h2b_fr* b, h2b_fr* out,
                                 size_t n) {
  if (!a || !out || field < 0 || field > 1 || op < 0 || op > 9) return H2B_ERR_ARG;
  for (size_t i = 0; i < n; ++i) {
    if (field == 0) {
      Fr x, y;
      memcpy(&x, a + i, 32);
      memcpy(&y, (b ? b : a) + i, 32);
      Fr r = field_op<Fr>(op, x, y);
      memcpy(out + i, &r, 32);
    } else {
      Fq x, y;
      memcpy(&x, a + i, 32);
      memcpy(&y, (b ? b : a) + i, 32);
      Fq r = field_op<Fq>(op, x, y);
      memcpy(out + i, &r, 32);
    }
  }
  return H2B_OK;
}

extern "C" int h2b_test_g1_op(h2b_ctx* ctx, int op, const h2b_g1_affine* a, const h2b_g1_affine* b,
                              h2b_g1_affine* out, size_t n) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a || !out || op < 0 || op > 2) return fail(ctx, H2B_ERR_ARG, "bad argument");
  if (n == 0) return H2B_OK;
  return run_elementwise(ctx, n * 64, a, b, out, [&](void* da, void* db, void* dout) {
    return launch(ctx, g1_op_kernel, dim3((uint32_t)((n + 63) / 64)), dim3(64), 0, op,
                  (const G1Affine*)da, (const G1Affine*)db, (G1Affine*)dout, (uint64_t)n);
  });
}

extern "C" int h2b_host_g1_op(int op, const h2b_g1_affine* a, const h2b_g1_affine* b,
                              h2b_g1_affine* out, size_t n) {
  if (!a || !out || op < 0 || op > 2) return H2B_ERR_ARG;
  for (size_t i = 0; i < n; ++i) {
    G1Affine x, y;
    memcpy(&x, a + i, 64);
    memcpy(&y, (b ? b : a) + i, 64);
    G1Affine r = g1_op(op, x, y);
    memcpy(out + i, &r, 64);
  }
  return H2B_OK;
}
