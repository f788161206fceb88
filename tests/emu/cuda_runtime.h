// TEST INFRASTRUCTURE -- a minimal CUDA execution-model emulator for the CPU.
//
// The `-m "not gpu"` tests compile the product's .cu files as plain C++ against
// this header (g++ -x c++ -DH2B_EMU -I tests/emu) into tests/emu/_build/
// libhalo2b200_emu.so so that the kernels' index arithmetic, carry chains'
// host twins, barriers and warp collectives can be checked against the oracle
// without a GPU.  It is NOT a CPU fallback: the product library
// (libhalo2b200.so) is built by nvcc only, never sees H2B_EMU, and no product
// code path loads the emulator build.
//
// Model: one kernel launch runs its blocks on a small pool of OS threads; the
// threads of one block are ucontext fibers on one OS thread, scheduled
// round-robin; __syncthreads / __syncwarp / shuffles / votes park a fiber
// until all live participants arrived.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 {
  unsigned x, y;
};
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
struct alignas(16) ulonglong2 {
  unsigned long long x, y;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) {
  uint4 r = {x, y, z, w};
  return r;
}
static inline uint2 make_uint2(unsigned x, unsigned y) {
  uint2 r = {x, y};
  return r;
}

namespace emu {

enum { WAIT_NONE = 0, WAIT_BLOCK = 1, WAIT_WARP = 2 };

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  int wait = WAIT_NONE;
  unsigned wait_mask = 0;
};

struct BlockState {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = 0;
  unsigned nthreads = 0;
  dim3 bdim, gdim;
  uint3 bidx;
  unsigned char* smem = nullptr;
  size_t smem_cap = 0;
  const std::function<void()>* body = nullptr;
  unsigned long long xchg[32 * 32];  // per-warp shuffle slots (up to 32 warps)
  unsigned vote[32];
};

inline BlockState*& tl_block() {
  static thread_local BlockState* b = nullptr;
  return b;
}
struct Idx {
  uint3 threadIdx, blockIdx;
  dim3 blockDim, gridDim;
};
inline Idx& tl_idx() {
  static thread_local Idx i;
  return i;
}

static const size_t kStack = 512 * 1024;

inline void fiber_entry() {
  BlockState* b = tl_block();
  (*b->body)();
  b = tl_block();
  b->fibers[b->cur].done = true;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

inline void park(int kind, unsigned mask) {
  BlockState* b = tl_block();
  Fiber& f = b->fibers[b->cur];
  f.wait = kind;
  f.wait_mask = mask;
  swapcontext(&f.ctx, &b->sched);
}

inline void set_idx(BlockState* b, unsigned t) {
  Idx& I = tl_idx();
  I.blockDim = b->bdim;
  I.gridDim = b->gdim;
  I.blockIdx = b->bidx;
  I.threadIdx.x = t % b->bdim.x;
  I.threadIdx.y = (t / b->bdim.x) % b->bdim.y;
  I.threadIdx.z = t / (b->bdim.x * b->bdim.y);
}

inline void run_block(BlockState* b) {
  unsigned nt = b->nthreads;
  for (unsigned t = 0; t < nt; ++t) {
    Fiber& f = b->fibers[t];
    f.done = false;
    f.wait = WAIT_NONE;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  unsigned live = nt;
  while (live > 0) {
    bool progressed = false;
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = b->fibers[t];
      if (f.done || f.wait != WAIT_NONE) continue;
      b->cur = (int)t;
      set_idx(b, t);
      swapcontext(&b->sched, &f.ctx);
      progressed = true;
      if (f.done) --live;
    }
    bool released = false;
    // block barrier
    unsigned nb = 0;
    for (unsigned t = 0; t < nt; ++t)
      if (!b->fibers[t].done && b->fibers[t].wait == WAIT_BLOCK) ++nb;
    if (live > 0 && nb == live) {
      for (unsigned t = 0; t < nt; ++t) b->fibers[t].wait = WAIT_NONE;
      released = true;
    }
    // warp barriers
    for (unsigned w = 0; w * 32 < nt; ++w) {
      unsigned lo = w * 32, hi = std::min(nt, lo + 32);
      unsigned waiting = 0, want = 0, alive = 0;
      for (unsigned t = lo; t < hi; ++t) {
        Fiber& f = b->fibers[t];
        if (f.done) continue;
        alive |= 1u << (t - lo);
        if (f.wait == WAIT_WARP) {
          waiting |= 1u << (t - lo);
          want |= f.wait_mask;
        }
      }
      if (waiting && ((want & alive) & ~waiting) == 0) {
        for (unsigned t = lo; t < hi; ++t)
          if (b->fibers[t].wait == WAIT_WARP) b->fibers[t].wait = WAIT_NONE;
        released = true;
      }
    }
    if (!progressed && !released) {
      fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): barrier divergence\n", b->bidx.x,
              b->bidx.y, b->bidx.z);
      abort();
    }
  }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  unsigned nthreads = block.x * block.y * block.z;
  if (nblocks == 0 || nthreads == 0) return;
  unsigned nworkers = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 16));
  if (nblocks < nworkers) nworkers = (unsigned)nblocks;
  if (getenv("H2B_EMU_THREADS")) nworkers = std::max(1, atoi(getenv("H2B_EMU_THREADS")));
  std::atomic<size_t> next(0);
  auto worker = [&]() {
    BlockState* b = new BlockState();
    b->fibers.resize(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
      b->fibers[t].stack = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE,
                                       MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (b->fibers[t].stack == MAP_FAILED) {
        perror("mmap");
        abort();
      }
    }
    b->nthreads = nthreads;
    b->bdim = block;
    b->gdim = grid;
    b->smem_cap = smem + 64;
    b->smem = (unsigned char*)aligned_alloc(128, (b->smem_cap + 127) / 128 * 128);
    b->body = &body;
    tl_block() = b;
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      b->bidx.x = (unsigned)(i % grid.x);
      b->bidx.y = (unsigned)((i / grid.x) % grid.y);
      b->bidx.z = (unsigned)(i / ((size_t)grid.x * grid.y));
      run_block(b);
    }
    for (unsigned t = 0; t < nthreads; ++t) munmap(b->fibers[t].stack, kStack);
    free(b->smem);
    tl_block() = nullptr;
    delete b;
  };
  if (nworkers == 1) {
    // still on a fresh OS thread: thread_local __shared__ storage stays per launch
    std::thread th(worker);
    th.join();
  } else {
    std::vector<std::thread> ths;
    for (unsigned w = 0; w < nworkers; ++w) ths.emplace_back(worker);
    for (auto& th : ths) th.join();
  }
}

inline unsigned char* dyn_smem() { return tl_block()->smem; }
inline unsigned lane_id() { return (unsigned)tl_block()->cur & 31; }
inline unsigned warp_id() { return (unsigned)tl_block()->cur >> 5; }

template <class T>
inline T shfl_idx(unsigned mask, T v, unsigned src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  BlockState* b = tl_block();
  unsigned w = warp_id(), l = lane_id();
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->xchg[w * 32 + l] = raw;
  park(WAIT_WARP, mask);
  b = tl_block();
  raw = b->xchg[w * 32 + (src & 31)];
  park(WAIT_WARP, mask);
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::tl_idx().threadIdx)
#define blockIdx (emu::tl_idx().blockIdx)
#define blockDim (emu::tl_idx().blockDim)
#define gridDim (emu::tl_idx().gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::park(emu::WAIT_BLOCK, 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::park(emu::WAIT_WARP, mask); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() {}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned l = emu::lane_id();
  unsigned base = l & ~(unsigned)(width - 1);
  return emu::shfl_idx(mask, v, base + ((unsigned)src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  (void)width;
  return emu::shfl_idx(mask, v, emu::lane_id() ^ (unsigned)lanemask);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned l = emu::lane_id();
  unsigned src = l + delta;
  if ((src & ~(unsigned)(width - 1)) != (l & ~(unsigned)(width - 1))) src = l;
  return emu::shfl_idx(mask, v, src);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned l = emu::lane_id();
  unsigned src = (l & (unsigned)(width - 1)) >= delta ? l - delta : l;
  return emu::shfl_idx(mask, v, src);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  emu::BlockState* b = emu::tl_block();
  unsigned w = emu::warp_id(), l = emu::lane_id();
  b->xchg[w * 32 + l] = pred ? 1ull : 0ull;
  emu::park(emu::WAIT_WARP, mask);
  b = emu::tl_block();
  unsigned out = 0;
  unsigned nt = b->nthreads;
  for (unsigned i = 0; i < 32; ++i) {
    unsigned t = w * 32 + i;
    if (t < nt && !b->fibers[t].done && (mask >> i & 1) && b->xchg[w * 32 + i]) out |= 1u << i;
  }
  emu::park(emu::WAIT_WARP, mask);
  return out;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
  // lanes that exited do not vote
  return __ballot_sync(mask, !pred) == 0;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) {
  return (unsigned)(((unsigned long long)a * b) >> 32);
}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}

template <class T>
static inline T atomicAdd(T* p, T v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
template <class T>
static inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
template <class T>
static inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
template <class T>
static inline T atomicOr(T* p, T v) {
  return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}
template <class T>
static inline T atomicExch(T* p, T v) {
  return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
template <class T>
static inline T atomicCAS(T* p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}

// ---------------------------------------------------------------------------
// Runtime API subset (single "device", synchronous "streams")
// ---------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event {
  double t;
}* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind {
  cudaMemcpyHostToHost = 0,
  cudaMemcpyHostToDevice = 1,
  cudaMemcpyDeviceToHost = 2,
  cudaMemcpyDeviceToDevice = 3,
  cudaMemcpyDefault = 4
};
enum { cudaStreamNonBlocking = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrClockRate = 13 };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount;
  int clockRate;
  int major, minor;
  size_t totalGlobalMem;
};
static inline const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emu error");
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = attr == cudaDevAttrMultiProcessorCount ? 4 : 1000000;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  strcpy(p->name, "cuda_emu");
  p->multiProcessorCount = 4;
  p->clockRate = 1000000;
  p->major = 10;
  return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t n) {
  return cudaMalloc((void**)p, n);
}
static inline cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind,
                                          cudaStream_t = nullptr) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) {
  *s = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = nullptr;
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline double emu_now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new emu_event();
  return cudaSuccess;
}
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = new emu_event();
  return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  e->t = emu_now_ms();
  return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)(b->t - a->t);
  return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) {
  return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) {
  *f = *t = (size_t)8 << 30;
  return cudaSuccess;
}
