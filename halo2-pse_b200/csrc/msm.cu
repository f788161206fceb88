// Pippenger multi-scalar multiplication over bn256 G1 for sm_100a: replaces
// `best_multiexp` / `multiexp_serial`
// (/root/reference/halo2_proofs/src/arithmetic.rs:13-159) behind
// ParamsKZG::commit / commit_lagrange (poly/kzg/commitment.rs:281-292,327-334).
//
// Same result (sum_i coeffs[i] * bases[i]), different algorithm.  The reference
// walks 256/c+1 unsigned windows serially per rayon chunk with 2^c-1 buckets;
// here all windows are processed at once:
//
//  1. msm_digits_kernel: Montgomery -> canonical, signed c-bit digits
//     (|d| <= 2^(c-1)), one (key, value) pair per (scalar, window):
//     key = window * 2^(c-1) + |d| - 1, value = point index | sign << 31;
//     zero digits get an all-ones key (the reference skips them too, :86).
//  2. one radix sort of all pairs by key (cub::DeviceRadixSort) -- every bucket
//     of every window becomes one contiguous run, zero digits sort to the end.
//  3. bucket accumulation, load-balanced independent of the scalar
//     distribution: the sorted array is cut into fixed chunks of L entries, one
//     thread per chunk, mixed XYZZ additions.  A run (= bucket) that lies
//     wholly inside a chunk is written straight to the dense bucket array; a
//     run cut by a chunk boundary is emitted as a partial sum into a (still
//     sorted) next-level list, which is reduced the same way until no run is
//     cut.  All level sizes stay on the device: no host round trip.
//  4. bucket reduction: segmented running sums (sum_b b * B_b per window) ->
//     per-segment weights by double-and-add -> shared-memory tree sums.
//  5. the W window sums are combined on the host (c doublings per window).
#include "common.cuh"

#ifndef H2B_EMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#endif
#include <string.h>
#include <algorithm>
#include <vector>

namespace h2b {

static const int kMaxLevels = 16;

struct MsmWorkspace {
  // capacities
  size_t cap_pairs = 0, cap_chunks = 0, cap_list = 0, cap_buckets = 0, cap_seg = 0;
  uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr;
  void* cub_temp = nullptr;
  size_t cub_temp_bytes = 0;
  uint32_t *cnt = nullptr, *incl = nullptr;
  uint32_t* n_level = nullptr;  // [kMaxLevels + 2]
  uint32_t* lkeys[2] = {nullptr, nullptr};
  G1Xyzz* lpts[2] = {nullptr, nullptr};
  G1Xyzz* buckets = nullptr;
  G1Xyzz* buckets2 = nullptr;  // bucket array of the later batches of an MSM whose host scalars arrive in batches
  size_t cap_buckets2 = 0;
  G1Xyzz* seg[2] = {nullptr, nullptr};  // tree-sum ping-pong
  G1Xyzz* h_out = nullptr;              // pinned, W window sums
  G1Xyzz* rc = nullptr;                 // two-dimensional bucket reduction: partial sums, V, shifted bit sums
  size_t cap_rc = 0;
  // batched-affine accumulation (window-table MSMs)
  size_t cap_aff = 0, cap_slots = 0;
  uint32_t* akeys[2] = {nullptr, nullptr};
  G1Affine* apts[2] = {nullptr, nullptr};
  uint32_t* acnt = nullptr;   // inclusive count of merged pair slots
  Fq* apre = nullptr;         // prefix products of the batched inversion
  G1Affine* astage = nullptr;  // round-0 staging of one tile of gathered points
  uint32_t* h_scalar = nullptr;  // pinned, device->host readbacks of list sizes
};

static void ws_release(MsmWorkspace* ws) {
  cudaFree(ws->keys_in);
  cudaFree(ws->keys_out);
  cudaFree(ws->vals_in);
  cudaFree(ws->vals_out);
  cudaFree(ws->cub_temp);
  cudaFree(ws->cnt);
  cudaFree(ws->incl);
  cudaFree(ws->n_level);
  for (int i = 0; i < 2; ++i) {
    cudaFree(ws->lkeys[i]);
    cudaFree(ws->lpts[i]);
    cudaFree(ws->seg[i]);
  }
  cudaFree(ws->buckets);
  if (ws->rc) cudaFree(ws->rc);
  if (ws->buckets2) cudaFree(ws->buckets2);
  ws->buckets2 = nullptr;
  ws->cap_buckets2 = 0;
  for (int i = 0; i < 2; ++i) {
    cudaFree(ws->akeys[i]);
    cudaFree(ws->apts[i]);
  }
  cudaFree(ws->acnt);
  cudaFree(ws->apre);
  cudaFree(ws->astage);
  if (ws->h_scalar) cudaFreeHost(ws->h_scalar);
  if (ws->h_out) cudaFreeHost(ws->h_out);
  *ws = MsmWorkspace();
}

void msm_ws_free(h2b_ctx* ctx) {
  if (!ctx->msm_ws) return;
  ws_release(ctx->msm_ws);
  delete ctx->msm_ws;
  ctx->msm_ws = nullptr;
}

// ---------------------------------------------------------------------------
// 1. digits
// ---------------------------------------------------------------------------
// table_stride > 0 (window table): key = |d| - 1 for every window, val = w * table_stride + i
// Processes scalars [i0, i1); the pair of (scalar i, window w) goes to slot w * row_stride + (i - row_base)
// (row_stride = n, row_base = 0 for a whole MSM; a batch of scalars gets its own compact pair array).
__global__ void msm_digits_kernel(const Fr* scalars, uint64_t row_stride, uint32_t c, uint32_t W,
                                  uint32_t* keys, uint32_t* vals, uint64_t table_stride, uint64_t i0,
                                  uint64_t i1, uint64_t row_base, uint32_t key_base) {
  for (uint64_t i = i0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const Fr s = from_mont(ld_fp(scalars + i));  // to_repr(), arithmetic.rs:14
    uint32_t carry = 0, w = 0;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1u;
    auto emit = [&](uint32_t d) {
      d += carry;
      uint32_t negf = 0;
      if (d > half) {
        d = (1u << c) - d;
        negf = 1;
        carry = 1;
      } else {
        carry = 0;
      }
      const uint64_t slot = (uint64_t)w * row_stride + (i - row_base);
      if (table_stride) {
        keys[slot] = d ? key_base + d - 1 : 0xffffffffu;  // key_base: the bucket set of this column (multi-column MSM)
        vals[slot] = (uint32_t)(w * table_stride + i) | (negf << 31);
      } else {
        keys[slot] = d ? w * half + d - 1 : 0xffffffffu;
        vals[slot] = (uint32_t)i | (negf << 31);
      }
      ++w;
    };
    // the limbs stream through a 64-bit window (static limb indices: a run-time index into s.v would put the scalar
    // in local memory); fewer than c bits wait in `buf` when a limb joins, so 32 + c - 1 <= 64 bits are live
    uint64_t buf = 0;
    uint32_t have = 0;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      buf |= (uint64_t)s.v[l] << have;
      have += 32;
      while (have >= c && w < W) {
        emit((uint32_t)buf & mask);
        buf >>= c;
        have -= c;
      }
    }
    while (w < W) {  // the bits above 256 are zero
      emit((uint32_t)buf & mask);
      buf >>= c;
    }
  }
}

// first index whose key has a bit at or above `kb` set (= number of non-zero digits)
__global__ void msm_find_valid_kernel(const uint32_t* keys, uint64_t total, uint32_t kb,
                                      uint32_t* n_level) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint64_t lo = 0, hi = total;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if ((keys[mid] >> kb) != 0)
      hi = mid;
    else
      lo = mid + 1;
  }
  n_level[0] = (uint32_t)lo;
}

// ---------------------------------------------------------------------------
// 3. chunked, level-wise bucket accumulation
// ---------------------------------------------------------------------------
// emissions of chunk t into the next level: its head run if the previous chunk
// ends with the same key, its tail run if the next chunk starts with it
__global__ void msm_count_kernel(const uint32_t* keys, const uint32_t* n_cur, uint32_t L,
                                 uint32_t nchunks, uint32_t* cnt) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const uint32_t N = *n_cur;
  const uint64_t start = (uint64_t)t * L;
  if (start >= N) {
    cnt[t] = 0;
    return;
  }
  const uint64_t end = start + L < N ? start + L : N;
  const uint32_t kf = keys[start], kl = keys[end - 1];
  const uint32_t head = start > 0 && keys[start - 1] == kf;
  const uint32_t tail = end < N && keys[end] == kl;
  cnt[t] = (kf == kl) ? (head | tail) : head + tail;
}

// the same counts and their inclusive scan in ONE launch, for chunk lists short enough for one block (the levels after
// the first: three launches -- count, scan-state init, scan -- become one)
static const uint32_t kFusedScanMax = 32768;
__global__ void __launch_bounds__(1024)
    msm_count_scan_kernel(const uint32_t* keys, const uint32_t* n_cur, uint32_t L, uint32_t nchunks, uint32_t* cnt,
                          uint32_t* incl) {
  __shared__ uint32_t part[1024];
  const uint32_t tid = threadIdx.x, per = (nchunks + 1023) / 1024;
  const uint32_t N = *n_cur;
  const uint32_t t0 = tid * per, t1 = t0 + per < nchunks ? t0 + per : nchunks;
  uint32_t sum = 0;
  for (uint32_t t = t0; t < t1; ++t) {
    const uint64_t start = (uint64_t)t * L;
    uint32_t c = 0;
    if (start < N) {
      const uint64_t end = start + L < N ? start + L : N;
      const uint32_t kf = keys[start], kl = keys[end - 1];
      const uint32_t head = start > 0 && keys[start - 1] == kf;
      const uint32_t tail = end < N && keys[end] == kl;
      c = (kf == kl) ? (head | tail) : head + tail;
    }
    cnt[t] = c;
    sum += c;
  }
  part[tid] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {  // Hillis-Steele over the 1024 thread totals
    const uint32_t v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint32_t run = part[tid] - sum;
  for (uint32_t t = t0; t < t1; ++t) {
    run += cnt[t];
    incl[t] = run;
  }
}

#ifdef H2B_EMU
static void emu_inclusive_scan(const uint32_t* in, uint32_t* out, size_t n) {
  uint32_t s = 0;
  for (size_t i = 0; i < n; ++i) {
    s += in[i];
    out[i] = s;
  }
}
#endif

struct Level0Src {
  const uint32_t* vals;
  const G1Affine* bases;
  struct Item {
    G1Affine p;
    uint32_t v;
  };
  H2B_D Item fetch(uint64_t i) const {
    Item it;
    it.v = vals[i];
    const G1Affine* src = bases + (it.v & 0x7fffffffu);
    it.p.x = ld_fp_nc64(&src->x);
    it.p.y = ld_fp_nc64(&src->y);
    return it;
  }
  H2B_D void add(G1Xyzz& acc, Item& it) const {
    if (it.v >> 31) it.p.y = neg(it.p.y);
    xyzz_add_affine(acc, it.p);
  }
};

struct LevelNSrc {
  const G1Xyzz* pts;
  struct Item {
    G1Xyzz p;
  };
  H2B_D Item fetch(uint64_t i) const {
    Item it;
    it.p.x = ld_fp(&pts[i].x);
    it.p.y = ld_fp(&pts[i].y);
    it.p.zz = ld_fp(&pts[i].zz);
    it.p.zzz = ld_fp(&pts[i].zzz);
    return it;
  }
  H2B_D void add(G1Xyzz& acc, Item& it) const { xyzz_add(acc, it.p); }
};

H2B_D void st_xyzz(G1Xyzz* dst, const G1Xyzz& p) {
  st_fp(&dst->x, p.x);
  st_fp(&dst->y, p.y);
  st_fp(&dst->zz, p.zz);
  st_fp(&dst->zzz, p.zzz);
}

// A bucket is written exactly once per pass over a sorted pair list (at the level where its run is no longer
// cut), with a plain store: lanes reach the ends of their runs at different iterations, so anything heavier
// there (adding to an earlier batch's bucket, say) would be executed by the warp on almost every iteration.
// chunk t of the N sorted entries; its partial sums (runs cut by a chunk boundary) go to slots o, o + 1 of the next list
template <class Src>
H2B_D void msm_accum_chunk(const Src& src, const uint32_t* keys, uint32_t N, uint32_t L, uint32_t t, uint32_t o,
                           uint32_t* okeys, G1Xyzz* opts, G1Xyzz* buckets) {
  const uint64_t start = (uint64_t)t * L;
  if (start >= N) return;
  const uint64_t end = start + L < N ? start + L : N;
  uint32_t cur = keys[start];
  const bool head = start > 0 && keys[start - 1] == cur;
  const bool tail = end < N && keys[end] == keys[end - 1];
  bool first_run = true;
  G1Xyzz acc = G1Xyzz::identity();
  // the operand of entry i + 1 (a random table gather at level 0) is in flight while entry i is added
  typename Src::Item nxt = src.fetch(start);
  uint32_t knxt = cur;
  for (uint64_t i = start; i < end; ++i) {
    typename Src::Item it = nxt;
    const uint32_t k = knxt;
    if (i + 1 < end) {
      knxt = keys[i + 1];
      nxt = src.fetch(i + 1);
    }
    if (k != cur) {
      if (first_run && head) {
        okeys[o] = cur;
        st_xyzz(opts + o, acc);
        ++o;
      } else {
        st_xyzz(buckets + cur, acc);
      }
      first_run = false;
      acc = G1Xyzz::identity();
      cur = k;
    }
    src.add(acc, it);
  }
  if ((first_run && head) || tail) {
    okeys[o] = cur;
    st_xyzz(opts + o, acc);
  } else {
    st_xyzz(buckets + cur, acc);
  }
}

template <class Src>
H2B_D void msm_accum_body(const Src& src, const uint32_t* keys, const uint32_t* n_cur, uint32_t L,
                          const uint32_t* cnt, const uint32_t* incl, uint32_t nchunks,
                          uint32_t* n_next, uint32_t* okeys, G1Xyzz* opts, G1Xyzz* buckets) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  if (t == nchunks - 1) *n_next = incl[t];
  msm_accum_chunk(src, keys, *n_cur, L, t, incl[t] - cnt[t], okeys, opts, buckets);
}

#ifndef H2B_ACC0_MINB
#define H2B_ACC0_MINB 4  // 126 registers with the operand prefetch: 4 blocks per SM (3 blocks at 142 registers: 37.9 vs 36.9 ms at k = 24)
#endif
__global__ void __launch_bounds__(128, H2B_ACC0_MINB)
    msm_accum0_kernel(Level0Src src, const uint32_t* keys, const uint32_t* n_cur, uint32_t L,
                      const uint32_t* cnt, const uint32_t* incl, uint32_t nchunks, uint32_t* n_next,
                      uint32_t* okeys, G1Xyzz* opts, G1Xyzz* buckets) {
  msm_accum_body(src, keys, n_cur, L, cnt, incl, nchunks, n_next, okeys, opts, buckets);
}

__global__ void __launch_bounds__(128)
    msm_accumN_kernel(LevelNSrc src, const uint32_t* keys, const uint32_t* n_cur, uint32_t L,
                      const uint32_t* cnt, const uint32_t* incl, uint32_t nchunks, uint32_t* n_next,
                      uint32_t* okeys, G1Xyzz* opts, G1Xyzz* buckets) {
  msm_accum_body(src, keys, n_cur, L, cnt, incl, nchunks, n_next, okeys, opts, buckets);
}

// Every remaining level in ONE launch, once the list is short enough for one block (at most 128 chunks of L entries:
// in practice a handful of entries, the host only knows the bound): count, scan and accumulate per level with the
// two lists ping-ponging, until a level cuts no run.  `from` = index (0 / 1) of the list this level reads.
static const uint32_t kFinishChunks = 128;
__global__ void __launch_bounds__(128)
    msm_finish_levels_kernel(uint32_t* lkeys0, uint32_t* lkeys1, G1Xyzz* lpts0, G1Xyzz* lpts1, uint32_t from,
                             const uint32_t* n_cur, uint32_t L, G1Xyzz* buckets) {
  __shared__ uint32_t sc[kFinishChunks];
  __shared__ uint32_t sN;
  const uint32_t tid = threadIdx.x;
  uint32_t N = *n_cur;
  for (uint32_t guard = 0; guard < 32 && N > 0; ++guard) {
    const uint32_t* keys = from ? lkeys1 : lkeys0;
    const LevelNSrc src{from ? lpts1 : lpts0};
    uint32_t* okeys = from ? lkeys0 : lkeys1;
    G1Xyzz* opts = from ? lpts0 : lpts1;
    const uint64_t start = (uint64_t)tid * L;
    uint32_t c = 0;
    if (start < N) {
      const uint64_t end = start + L < N ? start + L : N;
      const uint32_t kf = keys[start], kl = keys[end - 1];
      const uint32_t head = start > 0 && keys[start - 1] == kf;
      const uint32_t tail = end < N && keys[end] == kl;
      c = (kf == kl) ? (head | tail) : head + tail;
    }
    sc[tid] = c;
    __syncthreads();
    for (uint32_t d = 1; d < kFinishChunks; d <<= 1) {
      const uint32_t v = tid >= d ? sc[tid - d] : 0;
      __syncthreads();
      sc[tid] += v;
      __syncthreads();
    }
    msm_accum_chunk(src, keys, N, L, tid, sc[tid] - c, okeys, opts, buckets);
    if (tid == kFinishChunks - 1) sN = sc[tid];
    __syncthreads();  // the next level reads what this one wrote (block-wide visibility of global stores)
    N = sN;
    from ^= 1;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// 3'. batched-affine bucket accumulation (window-table MSMs)
//
// The sorted list is reduced by rounds of pairwise additions: in a round of
// parity q the elements (2s+q, 2s+q+1) of pair slot s are added when they carry
// the same key (= lie in the same bucket), everything else is copied through,
// and the list is compacted (positions from a prefix count of the merged
// slots).  Parities alternate so that every run of equal keys keeps halving
// wherever it starts; the loop ends when neither parity merges anything, i.e.
// every bucket is down to one point.  All additions of a round are independent
// affine additions x3 = l^2 - x1 - x2, y3 = l (x1 - x3) - y1 with
// l = (y2 - y1) / (x2 - x1): 2M + 1S plus 3M for the thread's batched inversion
// (Montgomery's trick over the ~10^3 slots a thread owns, prefix products in
// HBM) instead of the 8M + 2S of a mixed XYZZ addition.  P + P, P + (-P) and
// identity operands are handled in-band (denominator 2y, or 1).
// ---------------------------------------------------------------------------
struct PairFlag {
  const uint32_t* keys;
  uint32_t n, parity;
  H2B_HD uint32_t operator()(uint32_t s) const {
    const uint64_t i = 2ull * s + parity;
    return (i + 1 < n && keys[i] == keys[i + 1]) ? 1u : 0u;
  }
};

struct PairSrc0 {  // round 0: points come from the window table through the sorted indices
  const uint32_t* vals;
  const G1Affine* table;
  H2B_D Fq load_x(uint64_t i) const { return ld_fp_nc(&table[vals[i] & 0x7fffffffu].x); }
  H2B_D G1Affine load(uint64_t i) const {
    const uint32_t v = vals[i];
    const G1Affine* src = table + (v & 0x7fffffffu);
    G1Affine p;
    p.x = ld_fp_nc(&src->x);
    p.y = ld_fp_nc(&src->y);
    if (v >> 31) p.y = neg(p.y);
    return p;
  }
};

struct PairSrcN {
  const G1Affine* pts;
  H2B_D Fq load_x(uint64_t i) const { return ld_fp(&pts[i].x); }
  H2B_D G1Affine load(uint64_t i) const {
    G1Affine p;
    p.x = ld_fp(&pts[i].x);
    p.y = ld_fp(&pts[i].y);
    return p;
  }
};

// kind: 0 chord (den = x2 - x1), 1 tangent (den = 2 y1), 2 result = p2, 3 result = p1, 4 result = identity
H2B_D int pair_classify(const G1Affine& p1, const G1Affine& p2, Fq& den) {
  den = Fq::one();
  if (p1.is_identity()) return 2;
  if (p2.is_identity()) return 3;
  const Fq d = sub(p2.x, p1.x);
  if (!d.is_zero()) {
    den = d;
    return 0;
  }
  if (p1.y == p2.y) {  // y != 0: the group has odd order
    den = dbl(p1.y);
    return 1;
  }
  return 4;
}

template <class Src>
H2B_D Fq pair_den(const Src& src, uint64_t i) {
  const Fq x1 = src.load_x(i), x2 = src.load_x(i + 1);
  const Fq d = sub(x2, x1);
  if (!d.is_zero() && !x1.is_zero() && !x2.is_zero()) return d;  // the common case needs no y
  Fq den;
  pair_classify(src.load(i), src.load(i + 1), den);
  return den;
}

H2B_D void st_affine(G1Affine* dst, const G1Affine& p) {
  st_fp(&dst->x, p.x);
  st_fp(&dst->y, p.y);
}

// Slots [s0, s1) of one round.  With `stage` (round 0) pass 1 gathers every point
// of the tile once from the window table and parks it in a sequential staging
// buffer, so that the random HBM accesses are not repeated by pass 2.
template <class Src>
H2B_D void msm_pair_round_body(const Src& src, const uint32_t* keys, uint32_t n, uint32_t parity,
                               uint32_t s0, uint32_t s1, const uint32_t* cnt_incl, Fq* pre,
                               G1Affine* stage, uint32_t* okeys, G1Affine* opts) {
  const uint32_t T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  const PairFlag flag{keys, n, parity};
  if (t == 0 && s0 == 0 && parity == 1 && n > 0) {  // element 0 has no slot in an odd round
    st_affine(opts, src.load(0));
    okeys[0] = keys[0];
  }
  const uint32_t nloc = s1 - s0;
  if (t >= nloc) return;
  // pass 1: running product of the denominators of this thread's merged slots
  Fq run = Fq::one();
  bool any = false;
  for (uint32_t j = t; j < nloc; j += T) {
    const uint32_t s = s0 + j;
    const uint64_t i = 2ull * s + parity;
    const uint32_t f = flag(s);
    Fq den;
    if (stage) {
      const G1Affine p1 = src.load(i);
      st_affine(stage + 2ull * j, p1);
      if (i + 1 < n) {
        const G1Affine p2 = src.load(i + 1);
        st_affine(stage + 2ull * j + 1, p2);
        if (f) pair_classify(p1, p2, den);
      }
    } else if (f) {
      den = pair_den(src, i);
    }
    if (!f) continue;
    st_fp(pre + j, run);
    run = mul(run, den);
    any = true;
  }
  Fq inv_run = any ? inv(run) : run;
  // pass 2, backwards: peel one inverse per slot, add, write to the compacted position
  const PairSrcN staged{stage};
  const uint32_t last = t + ((nloc - 1 - t) / T) * T;
  for (int64_t j = last; j >= (int64_t)t; j -= T) {
    const uint32_t s = s0 + (uint32_t)j;
    const uint64_t i = 2ull * s + parity;
    const uint32_t f = flag(s);
    const uint64_t o = i - (cnt_incl[s] - f);
    const G1Affine p1 = stage ? staged.load(2ull * j) : src.load(i);
    if (!f) {
      st_affine(opts + o, p1);
      okeys[o] = keys[i];
      if (i + 1 < n) {
        st_affine(opts + o + 1, stage ? staged.load(2ull * j + 1) : src.load(i + 1));
        okeys[o + 1] = keys[i + 1];
      }
      continue;
    }
    const G1Affine p2 = stage ? staged.load(2ull * j + 1) : src.load(i + 1);
    Fq den;
    const int kind = pair_classify(p1, p2, den);
    const Fq inv_s = mul(inv_run, ld_fp(pre + j));
    inv_run = mul(inv_run, den);
    G1Affine r;
    if (kind <= 1) {
      Fq num;
      if (kind == 0) {
        num = sub(p2.y, p1.y);
      } else {
        const Fq xx = sqr(p1.x);
        num = add(dbl(xx), xx);
      }
      const Fq lam = mul(num, inv_s);
      r.x = sub(sub(sqr(lam), p1.x), p2.x);
      r.y = sub(mul(lam, sub(p1.x, r.x)), p1.y);
    } else if (kind == 2) {
      r = p2;
    } else if (kind == 3) {
      r = p1;
    } else {
      r.x = Fq::zero();
      r.y = Fq::zero();
    }
    st_affine(opts + o, r);
    okeys[o] = keys[i];
  }
}

__global__ void __launch_bounds__(128)
    msm_pair_round0_kernel(PairSrc0 src, const uint32_t* keys, uint32_t n, uint32_t parity, uint32_t s0,
                           uint32_t s1, const uint32_t* cnt_incl, Fq* pre, G1Affine* stage,
                           uint32_t* okeys, G1Affine* opts) {
  msm_pair_round_body(src, keys, n, parity, s0, s1, cnt_incl, pre, stage, okeys, opts);
}

__global__ void __launch_bounds__(128)
    msm_pair_roundN_kernel(PairSrcN src, const uint32_t* keys, uint32_t n, uint32_t parity, uint32_t s0,
                           uint32_t s1, const uint32_t* cnt_incl, Fq* pre, G1Affine* stage,
                           uint32_t* okeys, G1Affine* opts) {
  msm_pair_round_body(src, keys, n, parity, s0, s1, cnt_incl, pre, stage, okeys, opts);
}

// one point per bucket left: write the dense bucket array
template <class Src>
H2B_D void msm_scatter_body(const Src& src, const uint32_t* keys, uint32_t n, G1Xyzz* buckets) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    st_xyzz(buckets + keys[i], G1Xyzz::from_affine(src.load(i)));
}
__global__ void msm_scatter0_kernel(PairSrc0 src, const uint32_t* keys, uint32_t n, G1Xyzz* buckets) {
  msm_scatter_body(src, keys, n, buckets);
}
__global__ void msm_scatterN_kernel(PairSrcN src, const uint32_t* keys, uint32_t n, G1Xyzz* buckets) {
  msm_scatter_body(src, keys, n, buckets);
}

// ---------------------------------------------------------------------------
// 4. bucket reduction
// ---------------------------------------------------------------------------
H2B_D G1Xyzz ld_xyzz(const G1Xyzz* p) {
  G1Xyzz b;
  b.x = ld_fp(&p->x);
  b.y = ld_fp(&p->y);
  b.zz = ld_fp(&p->zz);
  b.zzz = ld_fp(&p->zzz);
  return b;
}

// buckets[i] += more[i]: folds the bucket array of a later batch of host scalars into the MSM's buckets
__global__ void __launch_bounds__(128) msm_bucket_merge_kernel(G1Xyzz* buckets, const G1Xyzz* more, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G1Xyzz m = ld_xyzz(more + i);
  if (m.is_identity()) return;
  G1Xyzz b = ld_xyzz(buckets + i);
  xyzz_add(b, m);
  st_xyzz(buckets + i, b);
}

// Segment s of window w covers buckets [s*M, (s+1)*M) (bucket index b holds
// digit value b+1).  out[w*nseg + s] = sum_r (r + 1 + s*M) * B[s*M + r]:
// running sums give A = sum (r+1) B and S = sum B; then A + [s*M] S by
// double-and-add.
__global__ void __launch_bounds__(128)
    msm_bucket_seg_kernel(const G1Xyzz* buckets, uint32_t nb_per_window, uint32_t lM, uint32_t nseg,
                          uint32_t total, G1Xyzz* out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t w = t / nseg, s = t % nseg;
  const uint32_t M = 1u << lM;
  const G1Xyzz* seg = buckets + (uint64_t)w * nb_per_window + ((uint64_t)s << lM);
  G1Xyzz running = G1Xyzz::identity(), acc = G1Xyzz::identity();
  for (int r = (int)M - 1; r >= 0; --r) {  // arithmetic.rs:95-99
    const G1Xyzz b = ld_xyzz(seg + r);
    xyzz_add(running, b);
    xyzz_add(acc, running);
  }
  // acc += [s * M] running
  if (s != 0 && !running.is_identity()) {
    G1Xyzz m = G1Xyzz::identity();
    for (int bit = 31 - __clz((int)s); bit >= 0; --bit) {
      m = xyzz_double(m);
      if ((s >> bit) & 1) xyzz_add(m, running);
    }
    for (uint32_t i = 0; i < lM; ++i) m = xyzz_double(m);
    xyzz_add(acc, m);
  }
  st_xyzz(out + t, acc);
}

// out[w * nout + b] = sum of in[w * nin + b*256 .. +256)
__global__ void __launch_bounds__(256)
    msm_tree_sum_kernel(const G1Xyzz* in, uint32_t nin, uint32_t nout, G1Xyzz* out) {
  H2B_DYN_SMEM(smem_raw);
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smem_raw);
  const uint32_t w = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  const uint32_t idx = b * 256 + tid;
  G1Xyzz v = G1Xyzz::identity();
  if (idx < nin) v = ld_xyzz(in + (uint64_t)w * nin + idx);
  sm[tid] = v;
  __syncthreads();
  for (uint32_t stride = 128; stride > 0; stride >>= 1) {
    if (tid < stride) {
      G1Xyzz a = sm[tid];
      xyzz_add(a, sm[tid + stride]);
      sm[tid] = a;
    }
    __syncthreads();
  }
  if (tid == 0) st_xyzz(out + (uint64_t)w * nout + b, sm[0]);
}

// ---------------------------------------------------------------------------
// 4b. two-dimensional bucket reduction (bucket sets of 2^m >= 4096 buckets)
//
// sum_b (b + 1) B[b] with b = hi * T + lo (T = 2^h columns, R = 2^(m - h) rows) is
//   sum_lo (lo + 1) C[lo]  +  T * sum_hi hi * Rw[hi],   C = column sums, Rw = row sums:
// two PLAIN sums per bucket, every one of them independent of the others (the running sums of arithmetic.rs:95-99
// are a serial chain per segment, and a segment pays a double-and-add for its offset: as many group operations
// again as the chain itself at the segment lengths that keep the GPU busy).  The two weighted sums that are left
// have T entries each: per weight bit a subset sum (tree), shifted by doublings, then one sum of the shifted points.
// ---------------------------------------------------------------------------

// pass A: thread t < nb/J sums J rows of one column, thread nb/J <= t < 2 nb/J sums J columns of one row
// (J = 2^lj buckets per thread: 16 for large bucket sets, 4 for small ones, where the serial chain is what costs)
__global__ void __launch_bounds__(128)
    msm_bucket_rc_kernel(const G1Xyzz* buckets, uint32_t m, uint32_t h, uint32_t lj, uint32_t nsets, G1Xyzz* part) {
  const uint32_t per_set = 2u << (m - lj), J = 1u << lj;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsets * per_set) return;
  const uint32_t w = t / per_set, tl = t % per_set, half = per_set / 2, T = 1u << h;
  const G1Xyzz* B = buckets + ((uint64_t)w << m);
  const G1Xyzz* src;
  uint32_t stride;
  if (tl < half) {  // rows [rp * J, +J) of column lo
    const uint32_t lo = tl & (T - 1), rp = tl >> h;
    src = B + (((uint64_t)rp << lj) << h) + lo;
    stride = T;
  } else {  // columns [q * J, +J) of row hi
    src = B + ((uint64_t)(tl - half) << lj);
    stride = 1;
  }
  G1Xyzz acc = ld_xyzz(src);
  G1Xyzz nxt = ld_xyzz(src + stride);
  for (uint32_t j = 2; j <= J; ++j) {  // the load of element j is in flight while element j - 1 is added
    const G1Xyzz b = nxt;
    if (j < J) nxt = ld_xyzz(src + (uint64_t)j * stride);
    xyzz_add(acc, b);
  }
  st_xyzz(part + (uint64_t)w * per_set + tl, acc);
}

// pass B: a block sums 256 consecutive partials = 256 / len whole groups (a group = the len = R/J partials of a column
// or the T/J partials of a row; len is a power of two <= 256) by halving in shared memory: every level keeps whole
// warps busy until fewer than 32 additions are left (a warp per group would issue 5 levels with idle lanes per group).
// V[w][lo] = C[lo];  V[w][T + hi - 1] = Rw[hi] for hi >= 1 (weight hi = index + 1 as in the column half);
// the rest of the row half stays at the identity (cleared by the caller).
// Column partials are stored [row group][lo]: the group of column lo is strided by T, hence the gather below.
__global__ void __launch_bounds__(128)
    msm_rc_group_kernel(const G1Xyzz* part, uint32_t m, uint32_t h, uint32_t lj, uint32_t nsets, G1Xyzz* V) {
  __shared__ G1Xyzz sm[256];
  const uint32_t T = 1u << h, R = 1u << (m - h), half = 1u << (m - lj), per_set = 2 * half;
  const uint32_t blocks_per_half = half / 256, tid = threadIdx.x;
  const uint32_t w = blockIdx.x / (2 * blocks_per_half), bl = blockIdx.x % (2 * blocks_per_half);
  const bool rows = bl >= blocks_per_half;
  const uint32_t e0 = (rows ? bl - blocks_per_half : bl) * 256;  // first entry of this block, in group-major order
  const uint32_t len = rows ? T >> lj : R >> lj;
  const G1Xyzz* P = part + (uint64_t)w * per_set + (rows ? half : 0);
  for (uint32_t i = tid; i < 256; i += 128) {
    const uint32_t e = e0 + i;  // entry e = group e / len, member e % len
    sm[i] = rows ? ld_xyzz(P + e) : ld_xyzz(P + (((uint64_t)(e % len)) << h) + e / len);
  }
  __syncthreads();
  const uint32_t groups = 256 / len;
  for (uint32_t cur = len; cur > 1; cur >>= 1) {
    const uint32_t hl = cur / 2;  // group g keeps its live members at sm[g * len + 0 .. cur)
    for (uint32_t i = tid; i < groups * hl; i += 128) {
      const uint32_t g = i / hl, j = i % hl;
      G1Xyzz a = sm[g * len + j];
      xyzz_add(a, sm[g * len + j + hl]);
      sm[g * len + j] = a;
    }
    __syncthreads();
  }
  if (tid < groups) {
    const uint32_t g = e0 / len + tid;  // column lo or row hi
    G1Xyzz* out = V + (uint64_t)w * 2 * T;
    if (!rows)
      st_xyzz(out + g, sm[tid * len]);
    else if (g >= 1)
      st_xyzz(out + T + g - 1, sm[tid * len]);
  }
}

// pass C: block (bit, set * 2 + half, split) sums the entries of its slice whose weight (index + 1) has `bit` set,
// then shifts the sum by `bit` doublings (+ h for the row half: the factor T).
__global__ void __launch_bounds__(128)
    msm_rc_bits_kernel(const G1Xyzz* V, uint32_t h, uint32_t nsplit, G1Xyzz* out) {
  __shared__ G1Xyzz sm[128];
  const uint32_t T = 1u << h, bit = blockIdx.x, sa = blockIdx.y, sp = blockIdx.z, tid = threadIdx.x;
  const uint32_t slice = T / nsplit;
  const G1Xyzz* A = V + (uint64_t)sa * T;
  G1Xyzz acc = G1Xyzz::identity();
  for (uint32_t i = sp * slice + tid; i < (sp + 1) * slice; i += 128)
    if (((i + 1) >> bit) & 1) {
      const G1Xyzz b = ld_xyzz(A + i);
      xyzz_add(acc, b);
    }
  sm[tid] = acc;
  __syncthreads();
  for (uint32_t s = 64; s > 0; s >>= 1) {
    if (tid < s) {
      G1Xyzz a = sm[tid];
      xyzz_add(a, sm[tid + s]);
      sm[tid] = a;
    }
    __syncthreads();
  }
  if (tid == 0) {
    G1Xyzz r = sm[0];
    const uint32_t nd = bit + ((sa & 1) ? h : 0);
    for (uint32_t i = 0; i < nd; ++i) r = xyzz_double(r);
    st_xyzz(out + ((uint64_t)sa * gridDim.x + bit) * nsplit + sp, r);
  }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static uint32_t ceil_log2(uint64_t x) {
  uint32_t l = 0;
  while ((1ull << l) < x) ++l;
  return l;
}

struct MsmPlan {
  uint32_t c, W, Wb, kb, L0, LN, lM;  // Wb: number of bucket sets (W, or 1 with a window table)
  uint64_t pairs;
  uint32_t nb_per_window, nseg;
};

static MsmPlan msm_plan(size_t n, uint32_t table_c = 0, uint32_t ncols = 1) {
  MsmPlan p;
  const uint32_t k = ceil_log2(n < 2 ? 2 : n);
  int c = (int)k - 4;
  if (c < 4) c = 4;
  if (c > 20) c = 20;
  if (const char* e = getenv("H2B_MSM_C")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 24) c = v;
  }
  if (table_c) c = (int)table_c;
  p.c = (uint32_t)c;
  p.W = (255 + p.c - 1) / p.c;
  p.Wb = table_c ? ncols : p.W;  // one bucket set per column of a multi-column MSM on a window table
  p.nb_per_window = 1u << (p.c - 1);
  p.kb = ceil_log2((uint64_t)p.Wb * p.nb_per_window);
  p.pairs = (uint64_t)n * p.W * (table_c ? ncols : 1);
  // chunk length of level 0: long enough to amortise the two boundary partials, short enough to
  // keep >= ~64k threads in flight (tuned on B200: k = 16 / 18 / 20 / 24)
  p.L0 = p.pairs >= (1ull << 25) ? 128 : p.pairs >= (1ull << 23) ? 48 : p.pairs >= (1ull << 21) ? 24 : 16;
  if (const char* e = getenv("H2B_MSM_L0")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 4096) p.L0 = (uint32_t)v;
  }
  p.LN = p.pairs >= (1ull << 26) ? 16 : 8;  // short lists: the serial chain of a chunk is what costs (re-swept late in round 2)
  if (const char* e = getenv("H2B_MSM_LN")) {
    const int v = atoi(e);
    if (v >= 2 && v <= 4096) p.LN = (uint32_t)v;
  }
  {  // buckets per reduction segment: ~2^15 segments or more, 8..64 buckets each (tuned likewise)
    const int lb = (int)ceil_log2((uint64_t)p.Wb * p.nb_per_window) - 15;
    p.lM = (uint32_t)(lb < 3 ? 3 : (lb > 6 ? 6 : lb));
    if (p.lM > p.c - 1) p.lM = p.c - 1;
  }
  if (const char* e = getenv("H2B_MSM_LM")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 10 && (uint32_t)v <= p.c - 1) p.lM = (uint32_t)v;
  }
  p.nseg = p.nb_per_window >> p.lM;
  return p;
}

static int ws_ensure(h2b_ctx* ctx, const MsmPlan& p) {
  if (!ctx->msm_ws) ctx->msm_ws = new MsmWorkspace();
  MsmWorkspace* ws = ctx->msm_ws;
  size_t chunks0 = (size_t)((p.pairs + p.L0 - 1) / p.L0);
  size_t list = 2 * chunks0 + 16;
  size_t nbuckets = (size_t)p.Wb * p.nb_per_window;
  size_t nsegs = (size_t)p.Wb * p.nseg;
  size_t pairs = (size_t)p.pairs;
  if (pairs <= ws->cap_pairs && chunks0 <= ws->cap_chunks && list <= ws->cap_list &&
      nbuckets <= ws->cap_buckets && nsegs <= ws->cap_seg)
    return H2B_OK;
  // grow-only in EVERY dimension: plans of different shapes (one column, several columns, other chunk lengths)
  // alternate inside one proof and must not evict each other
  pairs = std::max(pairs, ws->cap_pairs);
  chunks0 = std::max(chunks0, ws->cap_chunks);
  list = std::max(list, ws->cap_list);
  nbuckets = std::max(nbuckets, ws->cap_buckets);
  nsegs = std::max(nsegs, ws->cap_seg);
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ws_release(ws);
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->keys_in, pairs * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->keys_out, pairs * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->vals_in, pairs * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->vals_out, pairs * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->cnt, chunks0 * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->incl, chunks0 * 4 + 64));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->n_level, (kMaxLevels + 2) * 4));
  for (int i = 0; i < 2; ++i) {
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->lkeys[i], list * 4));
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->lpts[i], list * sizeof(G1Xyzz)));
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->seg[i], (nsegs + 256) * sizeof(G1Xyzz)));
  }
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->buckets, nbuckets * sizeof(G1Xyzz)));
  H2B_CUDA(ctx, cudaMallocHost((void**)&ws->h_out, 64 * sizeof(G1Xyzz)));
#ifndef H2B_EMU
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, ws->keys_in, ws->keys_out, ws->vals_in,
                                  ws->vals_out, pairs, 0, 32, ctx->stream);  // worst case: any key width
  cub::DeviceScan::InclusiveSum(nullptr, t2, ws->cnt, ws->incl,
                                (int)std::max<size_t>(chunks0, (size_t)(pairs / 2) + 2), ctx->stream);
  ws->cub_temp_bytes = std::max(t1, t2) + 256;
  H2B_CUDA(ctx, dev_malloc(ctx, &ws->cub_temp, ws->cub_temp_bytes));
#endif
  ws->cap_pairs = pairs;
  ws->cap_chunks = chunks0;
  ws->cap_list = list;
  ws->cap_buckets = nbuckets;
  ws->cap_seg = nsegs;
  return H2B_OK;
}

// pair slots per round-0 launch (2^25 slots = 4 GiB of staging); H2B_MSM_TILE overrides (tests)
static size_t stage_tile() {
  if (const char* e = getenv("H2B_MSM_TILE")) {
    const long v = atol(e);
    if (v >= 1) return (size_t)v;
  }
  return (size_t)1 << 25;
}

// Batched-affine accumulation of the sorted (key, index) list into ws->buckets.
static int accumulate_affine(h2b_ctx* ctx, MsmWorkspace* ws, const MsmPlan& p, const G1Affine* d_table) {
  cudaStream_t st = ctx->stream;
  (void)st;
  // capacities: after the first round a list holds at most N0/2 + #buckets elements
  const size_t nbuckets = (size_t)p.Wb * p.nb_per_window;
  const size_t cap = (size_t)(p.pairs / 2) + nbuckets + 16;
  const size_t slots = (size_t)(p.pairs / 2) + 2;
  const size_t kStageTile = stage_tile();
  if (cap > ws->cap_aff || slots > ws->cap_slots) {
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; ++i) {
      cudaFree(ws->akeys[i]);
      cudaFree(ws->apts[i]);
      ws->akeys[i] = nullptr;
      ws->apts[i] = nullptr;
    }
    cudaFree(ws->acnt);
    cudaFree(ws->apre);
    cudaFree(ws->astage);
    ws->acnt = nullptr;
    ws->apre = nullptr;
    ws->astage = nullptr;
    ws->cap_aff = ws->cap_slots = 0;
    for (int i = 0; i < 2; ++i) {
      H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->akeys[i], cap * 4));
      H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->apts[i], cap * sizeof(G1Affine)));
    }
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->acnt, slots * 4));
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->apre, slots * sizeof(Fq)));
    H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->astage,
                             2 * std::min<size_t>(slots, kStageTile) * sizeof(G1Affine)));
    if (!ws->h_scalar) H2B_CUDA(ctx, cudaMallocHost((void**)&ws->h_scalar, 64));
    ws->cap_aff = cap;
    ws->cap_slots = slots;
  }
  auto read_u32 = [&](const uint32_t* d, uint32_t* out) -> int {
    H2B_CUDA(ctx, cudaMemcpyAsync(ws->h_scalar, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ws->h_scalar[0];
    return H2B_OK;
  };
  uint32_t N = 0;
  H2B_TRY(read_u32(ws->n_level, &N));  // number of non-zero digits
  const uint32_t* keys = ws->keys_out;
  int cur = -1;  // -1: the list still lives in (keys_out, vals_out); else index into akeys / apts
  uint32_t parity = 0;
  int idle = 0, round = 0;
  const uint32_t max_blocks = (uint32_t)ctx->sm_count * 4;
  while (N > 1) {
    const uint32_t nslots = N > parity ? (N - parity + 1) / 2 : 0;
    uint32_t merged = 0;
    if (nslots) {
      const PairFlag flag{keys, N, parity};
#ifdef H2B_EMU
      uint32_t acc = 0;
      for (uint32_t s = 0; s < nslots; ++s) {
        acc += flag(s);
        ws->acnt[s] = acc;
      }
#else
      cub::TransformInputIterator<uint32_t, PairFlag, cub::CountingInputIterator<uint32_t>> it(
          cub::CountingInputIterator<uint32_t>(0), flag);
      H2B_CUDA(ctx, cub::DeviceScan::InclusiveSum(ws->cub_temp, ws->cub_temp_bytes, it, ws->acnt,
                                                  (int)nslots, st));
      ctx->launches += 2;
#endif
      H2B_TRY(read_u32(ws->acnt + (nslots - 1), &merged));
    }
    if (merged == 0) {
      if (++idle == 2) break;
      parity ^= 1;
      continue;
    }
    idle = 0;
    const int o = cur < 0 ? 0 : cur ^ 1;
    auto nblocks = [&](uint32_t cnt) {
      uint32_t b = (cnt + 256 * 128 - 1) / (256 * 128);
      return b > max_blocks ? max_blocks : (b < 1 ? 1u : b);
    };
    if (cur < 0) {
      const PairSrc0 src{ws->vals_out, d_table};
      if (ctx->profile) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
      for (uint64_t s0 = 0; s0 < nslots; s0 += kStageTile) {
        const uint32_t s1 = (uint32_t)std::min<uint64_t>(nslots, s0 + kStageTile);
        H2B_TRY(launch(ctx, msm_pair_round0_kernel, dim3(nblocks(s1 - (uint32_t)s0)), dim3(128), 0, src, keys,
                       N, parity, (uint32_t)s0, s1, (const uint32_t*)ws->acnt, ws->apre, ws->astage,
                       ws->akeys[o], ws->apts[o]));
      }
      if (ctx->profile) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    } else {
      const PairSrcN src{ws->apts[cur]};
      H2B_TRY(launch(ctx, msm_pair_roundN_kernel, dim3(nblocks(nslots)), dim3(128), 0, src, keys, N, parity,
                     0u, nslots, (const uint32_t*)ws->acnt, ws->apre, (G1Affine*)nullptr, ws->akeys[o],
                     ws->apts[o]));
    }
    N -= merged;
    cur = o;
    keys = ws->akeys[o];
    parity ^= 1;
    if (++round > 200) return fail(ctx, H2B_ERR_ARG, "batched-affine accumulation did not converge");
  }
  if (N) {
    uint32_t blocks = (N + 255) / 256;
    if (blocks > max_blocks * 4) blocks = max_blocks * 4;
    if (cur < 0) {
      const PairSrc0 src{ws->vals_out, d_table};
      H2B_TRY(launch(ctx, msm_scatter0_kernel, dim3(blocks), dim3(256), 0, src, keys, N, ws->buckets));
    } else {
      const PairSrcN src{ws->apts[cur]};
      H2B_TRY(launch(ctx, msm_scatterN_kernel, dim3(blocks), dim3(256), 0, src, keys, N, ws->buckets));
    }
  }
  return H2B_OK;
}

int msm_run(h2b_ctx* ctx, const G1Affine* d_bases, const Fr* d_scalars, size_t n,
            G1Xyzz* out_host, size_t table_stride, uint32_t table_c, const Fr* h_scalars,
            uint32_t ncols, const Fr* const* col_scalars) {
  for (uint32_t j = 0; j < (ncols ? ncols : 1); ++j) out_host[j] = G1Xyzz::identity();
  if (n == 0) return H2B_OK;
  if (n >= (1ull << 31)) return fail(ctx, H2B_ERR_ARG, "MSM larger than 2^31 points");
  if (ncols > 1 && (!table_stride || h_scalars || !col_scalars))
    return fail(ctx, H2B_ERR_ARG, "multi-column MSM needs a window table and device scalars");
  if (ncols < 1) ncols = 1;
  const MsmPlan p = msm_plan(n, table_stride ? table_c : 0, ncols);
  if (p.pairs >= (1ull << 32) || p.Wb > 64) return fail(ctx, H2B_ERR_ARG, "multi-column MSM too large");
  if (table_stride && (uint64_t)p.W * table_stride >= (1ull << 31))
    return fail(ctx, H2B_ERR_ARG, "window table larger than 2^31 points");
  H2B_TRY(ws_ensure(ctx, p));
  MsmWorkspace* ws = ctx->msm_ws;
  cudaStream_t st = ctx->stream;
  (void)st;

  // Host scalars on a window table: the MSM is cut into NB batches of points that share the bucket array.
  // Batch b + 1 travels over PCIe (copy stream) while batch b is sorted and accumulated; batches after the
  // first fill a second bucket array that one element-wise kernel folds into the first.  One bucket
  // reduction at the end.  Everything else is one batch.
  // (B200, PCIe 5: k = 24 49.2 ms in one batch, 45.9 with 4 equal batches, 42.6 with 3 growing ones; a loss below k = 23)
  uint64_t batch_min = 1ull << 22;  // points per batch; H2B_MSM_BATCH_MIN overrides (tests), 0 disables
  if (const char* e = getenv("H2B_MSM_BATCH_MIN")) batch_min = strtoull(e, nullptr, 10);
  // Batch sizes grow geometrically: only the copy of batch 0 is exposed, and the copy of batch b + 1 hides under the
  // compute of batch b as long as it is not more than compute/copy times larger.  That ratio is measured (events
  // around every batch's copy and compute of the previous calls on this context): 4 alone on PCIe 5 (weights 1,4,16:
  // k = 24 45.5 ms with four equal batches -> 42.6), below 2 when eight ranks share the host's memory path (five
  // batches growing by 1.6).  H2B_MSM_BATCH_PLAN = "w0,w1,..." fixes the relative weights (at most 7 batches).
  uint64_t bstart[9] = {0};
  int NB = 1;
  if (h_scalars && table_stride && batch_min && ncols == 1 && n >= 2 * batch_min) {
    double wts[8] = {1, 4, 16, 0, 0, 0, 0, 0};
    int nw = n >= 4 * batch_min ? 3 : 2;
    if (const char* e = getenv("H2B_MSM_BATCH_PLAN")) {
      nw = 0;
      for (const char* q = e; *q && nw < 7;) {
        char* end = nullptr;
        const unsigned long v = strtoul(q, &end, 10);
        if (end == q) break;
        if (v) wts[nw++] = (double)v;
        q = *end ? end + 1 : end;
      }
      if (nw == 0) { wts[0] = 1; nw = 1; }
    } else if (ctx->msm_copy_ratio > 0.f && n >= 4 * batch_min) {
      double g = 0.95 / ctx->msm_copy_ratio;
      g = g > 4.0 ? 4.0 : g < 1.25 ? 1.25 : g;
      nw = g >= 3.0 ? 3 : g >= 1.8 ? 4 : 5;
      while (nw > 2 && n < (uint64_t)nw * batch_min) --nw;
      for (int i = 0; i < nw; ++i) wts[i] = i ? wts[i - 1] * g : 1.0;
    }
    double tot = 0, run = 0;
    uint64_t prev = 0;
    for (int i = 0; i < nw; ++i) tot += wts[i];
    NB = 0;
    for (int i = 0; i < nw; ++i) {
      run += wts[i];
      const uint64_t e1 = i + 1 == nw ? (uint64_t)n : ((uint64_t)((double)n * (run / tot)) & ~(uint64_t)(n >= 65536 ? 255 : 0));
      if (e1 > prev && e1 <= n) bstart[++NB] = prev = e1;
    }
    bstart[NB] = n;
  } else {
    bstart[1] = n;
  }
  const bool plan_timing = NB > 1;
  if (plan_timing && !ctx->plan_ev[0])
    for (int i = 0; i < 28; ++i) H2B_CUDA(ctx, cudaEventCreate(&ctx->plan_ev[i]));
  H2B_CUDA(ctx, cudaMemsetAsync(ws->buckets, 0, (size_t)p.Wb * p.nb_per_window * sizeof(G1Xyzz), st));
  // copy of batch b on the copy stream, event copy_ev[b]; queued right after the compute of batch b - 1, so
  // that a pageable source (staged by host threads, which blocks this thread) overlaps that compute as well
  auto queue_copy = [&](int b) -> int {
    if (b >= NB) return H2B_OK;
    const uint64_t i0 = bstart[b], i1 = bstart[b + 1];
    if (i0 >= i1) return H2B_OK;
    if (plan_timing) H2B_CUDA(ctx, cudaEventRecord(ctx->plan_ev[4 * b], ctx->copy_stream));
    H2B_TRY(copy_h2d_any(ctx, const_cast<Fr*>(d_scalars) + i0, h_scalars + i0, (i1 - i0) * sizeof(Fr), ctx->copy_stream));
    if (plan_timing) H2B_CUDA(ctx, cudaEventRecord(ctx->plan_ev[4 * b + 1], ctx->copy_stream));
    H2B_CUDA(ctx, cudaEventRecord(ctx->copy_ev[b], ctx->copy_stream));
    return H2B_OK;
  };
  if (h_scalars && NB > 1) {
    // the copy stream must not overwrite the staging buffer while an earlier call still reads it
    H2B_CUDA(ctx, cudaEventRecord(ctx->copy_ev[7], st));
    H2B_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[7], 0));
    H2B_TRY(queue_copy(0));
  }
  for (int batch = 0; batch < NB; ++batch) {
  const uint64_t b0 = bstart[batch], b1 = bstart[batch + 1];
  if (b0 >= b1) continue;
  const uint64_t nbatch = b1 - b0;
  const uint64_t pairs = NB > 1 ? nbatch * p.W : p.pairs;
  // later batches fill a second bucket array, folded into the first by one element-wise kernel
  const uint32_t nbuckets_all = p.Wb * p.nb_per_window;
  G1Xyzz* target = ws->buckets;
  if (batch > 0) {
    if (ws->cap_buckets2 < nbuckets_all) {
      H2B_CUDA(ctx, cudaStreamSynchronize(st));
      if (ws->buckets2) cudaFree(ws->buckets2);
      ws->buckets2 = nullptr;
      ws->cap_buckets2 = 0;
      H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->buckets2, (size_t)nbuckets_all * sizeof(G1Xyzz)));
      ws->cap_buckets2 = nbuckets_all;
    }
    target = ws->buckets2;
    H2B_CUDA(ctx, cudaMemsetAsync(target, 0, (size_t)nbuckets_all * sizeof(G1Xyzz), st));
  }
  // 1. digits (host scalars, single batch: chunked H2D on the copy stream, overlapped with the digit kernel)
  {
    const uint64_t cap = (uint64_t)ctx->sm_count * 16;
    const int nchunk = (NB == 1 && h_scalars && n >= (1u << 16)) ? 8 : 1;
    const uint64_t per = (nbatch + nchunk - 1) / nchunk;
    if (h_scalars && nchunk > 1) {
      H2B_CUDA(ctx, cudaEventRecord(ctx->copy_ev[0], st));
      H2B_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
    }
    if (NB > 1) H2B_CUDA(ctx, cudaStreamWaitEvent(st, ctx->copy_ev[batch], 0));
    if (plan_timing) H2B_CUDA(ctx, cudaEventRecord(ctx->plan_ev[4 * batch + 2], st));  // stamped once the copy has landed
    for (int ci = 0; ci < nchunk; ++ci) {
      const uint64_t i0 = b0 + (uint64_t)ci * per, i1 = std::min<uint64_t>(b1, i0 + per);
      if (i0 >= i1) break;
      if (h_scalars && NB == 1) {
        Fr* dst = const_cast<Fr*>(d_scalars) + i0;
        if (nchunk == 1) {
          H2B_TRY(copy_h2d_any(ctx, dst, h_scalars + i0, (i1 - i0) * sizeof(Fr), st));
        } else {
          H2B_TRY(copy_h2d_any(ctx, dst, h_scalars + i0, (i1 - i0) * sizeof(Fr), ctx->copy_stream));
          H2B_CUDA(ctx, cudaEventRecord(ctx->copy_ev[ci], ctx->copy_stream));
          H2B_CUDA(ctx, cudaStreamWaitEvent(st, ctx->copy_ev[ci], 0));
        }
      }
      const uint64_t want = (i1 - i0 + 255) / 256;
      for (uint32_t col = 0; col < ncols; ++col) {  // column `col` of a multi-column MSM: its own pair rows and bucket set
        const uint64_t off = (uint64_t)col * p.W * nbatch;
        H2B_TRY(launch(ctx, msm_digits_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0,
                       ncols > 1 ? col_scalars[col] : d_scalars, (uint64_t)nbatch, p.c, p.W, ws->keys_in + off,
                       ws->vals_in + off, (uint64_t)table_stride, i0, i1, b0, col * p.nb_per_window));
      }
    }
  }
  // 2. sort
#ifdef H2B_EMU
  {
    std::vector<std::pair<uint32_t, uint32_t>> v(pairs);
    for (uint64_t i = 0; i < pairs; ++i) v[i] = {ws->keys_in[i], ws->vals_in[i]};
    std::stable_sort(v.begin(), v.end(),
                     [](const std::pair<uint32_t, uint32_t>& a,
                        const std::pair<uint32_t, uint32_t>& b) { return a.first < b.first; });
    for (uint64_t i = 0; i < pairs; ++i) {
      ws->keys_out[i] = v[i].first;
      ws->vals_out[i] = v[i].second;
    }
  }
#else
  H2B_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ws->cub_temp, ws->cub_temp_bytes, ws->keys_in,
                                                ws->keys_out, ws->vals_in, ws->vals_out, pairs,
                                                0, (int)p.kb + 1, st));
  ctx->launches += 1 + (p.kb + 1 + 7) / 8;  // onesweep: histogram + one pass per 8 key bits
#endif
  H2B_TRY(launch(ctx, msm_find_valid_kernel, dim3(1), dim3(32), 0, (const uint32_t*)ws->keys_out,
                 (uint64_t)pairs, p.kb, ws->n_level));

  // 3. bucket accumulation: batched affine on a window table, else level-wise XYZZ chunks
  // (the batched-affine path is correct but not yet faster than the XYZZ chunks: opt-in)
  bool affine = false;
  if (const char* e = getenv("H2B_MSM_ACC")) affine = table_stride != 0 && NB == 1 && strcmp(e, "affine") == 0;
  if (affine) {
    H2B_TRY(accumulate_affine(ctx, ws, p, d_bases));
  } else {
    uint64_t nmax = pairs;
    const uint32_t* keys = ws->keys_out;
    for (int level = 0; level < kMaxLevels; ++level) {
      const uint32_t L = level == 0 ? p.L0 : p.LN;
      const uint32_t nchunks = (uint32_t)((nmax + L - 1) / L);
      const int o = level & 1;
      if (level >= 1 && nchunks <= kFinishChunks) {  // short enough for one block: all remaining levels in one launch
        H2B_TRY(launch(ctx, msm_finish_levels_kernel, dim3(1), dim3(kFinishChunks), 0, ws->lkeys[0], ws->lkeys[1],
                       ws->lpts[0], ws->lpts[1], (uint32_t)(o ^ 1), (const uint32_t*)(ws->n_level + level), L, target));
        break;
      }
      if (nchunks <= kFusedScanMax) {
        H2B_TRY(launch(ctx, msm_count_scan_kernel, dim3(1), dim3(1024), 0, keys,
                       (const uint32_t*)(ws->n_level + level), L, nchunks, ws->cnt, ws->incl));
      } else {
        H2B_TRY(launch(ctx, msm_count_kernel, dim3((nchunks + 255) / 256), dim3(256), 0, keys,
                       (const uint32_t*)(ws->n_level + level), L, nchunks, ws->cnt));
#ifdef H2B_EMU
        emu_inclusive_scan(ws->cnt, ws->incl, nchunks);
#else
        H2B_CUDA(ctx, cub::DeviceScan::InclusiveSum(ws->cub_temp, ws->cub_temp_bytes, ws->cnt,
                                                    ws->incl, (int)nchunks, st));
        ctx->launches += 2;  // scan-state init + scan
#endif
      }
      if (level == 0) {
        Level0Src src{ws->vals_out, d_bases};
        // the throughput-bound kernel goes to the low-priority stream (common.cuh: bulk_stream)
        cudaStream_t bulk = ctx->bulk_stream;
        H2B_CUDA(ctx, cudaEventRecord(ctx->bulk_ev[0], st));
        H2B_CUDA(ctx, cudaStreamWaitEvent(bulk, ctx->bulk_ev[0], 0));
        if (ctx->profile && batch == 0) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[0], bulk));
        H2B_TRY(launch_on(ctx, bulk, msm_accum0_kernel, dim3((nchunks + 127) / 128), dim3(128), 0, src, keys,
                          (const uint32_t*)(ws->n_level + level), L, (const uint32_t*)ws->cnt,
                          (const uint32_t*)ws->incl, nchunks, ws->n_level + level + 1, ws->lkeys[o],
                          ws->lpts[o], target));
        if (ctx->profile && batch == 0) H2B_CUDA(ctx, cudaEventRecord(ctx->ev[1], bulk));
        H2B_CUDA(ctx, cudaEventRecord(ctx->bulk_ev[1], bulk));
        H2B_CUDA(ctx, cudaStreamWaitEvent(st, ctx->bulk_ev[1], 0));
      } else {
        LevelNSrc src{ws->lpts[o ^ 1]};
        H2B_TRY(launch(ctx, msm_accumN_kernel, dim3((nchunks + 127) / 128), dim3(128), 0, src, keys,
                       (const uint32_t*)(ws->n_level + level), L, (const uint32_t*)ws->cnt,
                       (const uint32_t*)ws->incl, nchunks, ws->n_level + level + 1, ws->lkeys[o],
                       ws->lpts[o], target));
      }
      if (nmax <= L) break;  // a single chunk cuts no run
      nmax = 2ull * nchunks;
      keys = ws->lkeys[o];
      if (level + 1 == kMaxLevels) return fail(ctx, H2B_ERR_ARG, "MSM level overflow");
    }
  }
  if (batch > 0)
    H2B_TRY(launch(ctx, msm_bucket_merge_kernel, dim3((nbuckets_all + 127) / 128), dim3(128), 0, ws->buckets,
                   (const G1Xyzz*)ws->buckets2, nbuckets_all));
  if (plan_timing) H2B_CUDA(ctx, cudaEventRecord(ctx->plan_ev[4 * batch + 3], st));
  if (NB > 1) H2B_TRY(queue_copy(batch + 1));
  }  // batches

  // 4. bucket reduction: two-dimensional for bucket sets of >= 2^12 buckets (H2B_MSM_RC=0: the segmented running
  //    sums everywhere), else segmented running sums
  const bool rc_enabled = !(getenv("H2B_MSM_RC") && atoi(getenv("H2B_MSM_RC")) == 0);
  if (rc_enabled && p.c - 1 >= 12) {
    const uint32_t m = p.c - 1, h = (m + 1) / 2, T = 1u << h, R = 1u << (m - h);
    const uint32_t nbits = h + 1, nsplit = T >= 1024 ? 8 : T / 128 ? T / 128 : 1;  // 2 * nbits * nsplit <= 256
    const uint32_t lj = m >= 18 ? 4 : m == 17 ? 3 : 2;  // buckets per thread of the first pass: 16 / 8 / 4
    const size_t n_part = (size_t)p.Wb * ((size_t)2 << (m - lj)), n_v = (size_t)p.Wb * 2 * T,
                 n_bits = (size_t)p.Wb * 2 * nbits * nsplit;
    if (ws->cap_rc < n_part + n_v + n_bits) {
      H2B_CUDA(ctx, cudaStreamSynchronize(st));
      if (ws->rc) cudaFree(ws->rc);
      ws->rc = nullptr;
      ws->cap_rc = 0;
      H2B_CUDA(ctx, dev_malloc(ctx, (void**)&ws->rc, (n_part + n_v + n_bits) * sizeof(G1Xyzz)));
      ws->cap_rc = n_part + n_v + n_bits;
    }
    G1Xyzz* part = ws->rc;
    G1Xyzz* V = part + n_part;
    G1Xyzz* bits = V + n_v;
    H2B_CUDA(ctx, cudaMemsetAsync(V, 0, n_v * sizeof(G1Xyzz), st));
    const uint32_t ta = p.Wb * (2u << (m - lj));
    H2B_TRY(launch(ctx, msm_bucket_rc_kernel, dim3((ta + 127) / 128), dim3(128), 0, (const G1Xyzz*)ws->buckets, m, h,
                   lj, p.Wb, part));
    H2B_TRY(launch(ctx, msm_rc_group_kernel, dim3(ta / 256), dim3(128), 0, (const G1Xyzz*)part, m, h, lj, p.Wb, V));
    H2B_TRY(launch(ctx, msm_rc_bits_kernel, dim3(nbits, 2 * p.Wb, nsplit), dim3(128), 0, (const G1Xyzz*)V, h, nsplit,
                   bits));
    const uint32_t np = 2 * nbits * nsplit;  // <= 2 * 13 * 8 shifted points per set
    H2B_TRY(launch(ctx, msm_tree_sum_kernel, dim3(1, p.Wb), dim3(256), 256 * sizeof(G1Xyzz), (const G1Xyzz*)bits, np,
                   1u, ws->seg[0]));
    H2B_CUDA(ctx, cudaMemcpyAsync(ws->h_out, ws->seg[0], p.Wb * sizeof(G1Xyzz), cudaMemcpyDeviceToHost, st));
  } else {
    const uint32_t total = p.Wb * p.nseg;
    H2B_TRY(launch(ctx, msm_bucket_seg_kernel, dim3((total + 127) / 128), dim3(128), 0,
                   (const G1Xyzz*)ws->buckets, p.nb_per_window, p.lM, p.nseg, total, ws->seg[0]));
    uint32_t nin = p.nseg;
    int cur = 0;
    while (nin > 1) {
      const uint32_t nout = (nin + 255) / 256;
      H2B_TRY(launch(ctx, msm_tree_sum_kernel, dim3(nout, p.Wb), dim3(256), 256 * sizeof(G1Xyzz),
                     (const G1Xyzz*)ws->seg[cur], nin, nout, ws->seg[cur ^ 1]));
      cur ^= 1;
      nin = nout;
    }
    H2B_CUDA(ctx, cudaMemcpyAsync(ws->h_out, ws->seg[cur], p.Wb * sizeof(G1Xyzz),
                                  cudaMemcpyDeviceToHost, st));
  }
  H2B_CUDA(ctx, cudaStreamSynchronize(st));
  if (ctx->profile) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last_kernel_ms = ms;
  }
  if (plan_timing) {  // copy time per point over compute time per point, for the next call's batch sizes
    float copy_ms = 0, comp_ms = 0;
    bool ok = true;
    for (int b = 0; b < NB; ++b) {
      float a = 0, c2 = 0;
      ok = ok && cudaEventElapsedTime(&a, ctx->plan_ev[4 * b], ctx->plan_ev[4 * b + 1]) == cudaSuccess &&
           cudaEventElapsedTime(&c2, ctx->plan_ev[4 * b + 2], ctx->plan_ev[4 * b + 3]) == cudaSuccess;
      copy_ms += a;
      comp_ms += c2;
    }
    if (ok && comp_ms > 0.f && copy_ms > 0.f) {
      const float r = copy_ms / comp_ms;
      ctx->msm_copy_ratio = ctx->msm_copy_ratio > 0.f ? 0.5f * ctx->msm_copy_ratio + 0.5f * r : r;
    }
  }

  // 5. combine windows, top first: acc = 2^c * acc + R_w   (arithmetic.rs:46-49)
  G1Xyzz acc = G1Xyzz::identity();
  if (table_stride) {
    acc = ws->h_out[0];  // the table already carries the 2^(c*w) factors
    for (uint32_t j = 1; j < ncols; ++j) out_host[j] = ws->h_out[j];
  } else {
    for (int w = (int)p.W - 1; w >= 0; --w) {
      for (uint32_t i = 0; i < p.c; ++i) acc = xyzz_double(acc);
      xyzz_add(acc, ws->h_out[w]);
    }
  }
  *out_host = acc;
  return H2B_OK;
}

// ---------------------------------------------------------------------------
// window table: T_w[i] = 2^(c*w) * P_i, affine            (h2b_bases_precompute)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    msm_table_double_kernel(const G1Affine* src, G1Xyzz* dst, uint64_t n, uint32_t c) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p;
  p.x = ld_fp(&src[i].x);
  p.y = ld_fp(&src[i].y);
  G1Xyzz acc = G1Xyzz::identity();
  if (!p.is_identity()) {
    acc = xyzz_double_affine(p);
    for (uint32_t j = 1; j < c; ++j) acc = xyzz_double(acc);
  }
  st_xyzz(dst + i, acc);
}

// XYZZ -> affine with one inversion per thread (Montgomery's trick over kBatch points)
static const int kNormBatch = 16;
__global__ void __launch_bounds__(128)
    msm_table_normalize_kernel(const G1Xyzz* src, G1Affine* dst, uint64_t n) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = t * kNormBatch;
  if (i0 >= n) return;
  const int cnt = (int)(n - i0 < (uint64_t)kNormBatch ? n - i0 : (uint64_t)kNormBatch);
  Fq pre[kNormBatch];
  Fq run = Fq::one();
  for (int j = 0; j < cnt; ++j) {
    pre[j] = run;
    const Fq z = ld_fp(&src[i0 + j].zzz);
    if (!z.is_zero()) run = mul(run, z);
  }
  Fq invr = inv(run);
  for (int j = cnt - 1; j >= 0; --j) {
    const G1Xyzz p = ld_xyzz(src + i0 + j);
    G1Affine a;
    a.x = Fq::zero();
    a.y = Fq::zero();
    if (!p.zzz.is_zero()) {
      const Fq izzz = mul(invr, pre[j]);
      invr = mul(invr, p.zzz);
      const Fq izz = sqr(mul(p.zz, izzz));  // zz^2 / zzz^2 = 1 / zz
      a.x = mul(p.x, izz);
      a.y = mul(p.y, izzz);
    }
    st_fp(&dst[i0 + j].x, a.x);
    st_fp(&dst[i0 + j].y, a.y);
  }
}

}  // namespace h2b

// ===========================================================================
// C ABI
// ===========================================================================
using namespace h2b;

// ---------------------------------------------------------------------------
// [k_i] G for many scalars: the 2 * 2^k scalar multiplications of ParamsKZG::setup
// (poly/kzg/commitment.rs:67-116), double-and-add per thread, one batched
// normalisation.  Not a hot path of the prover; it makes test / benchmark SRS
// generation a GPU job instead of minutes of CPU time.
// ---------------------------------------------------------------------------
namespace h2b {
__global__ void __launch_bounds__(128)
    g1_mul_generator_kernel(const Fr* scalars, G1Xyzz* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr k = from_mont(ld_fp(scalars + i));
  G1Affine g;
  g.x = Fq::one();
  g.y = add(Fq::one(), Fq::one());
  G1Xyzz acc = G1Xyzz::identity();
  bool started = false;
  for (int limb = 7; limb >= 0; --limb) {
    for (int bit = 31; bit >= 0; --bit) {
      if (started) acc = xyzz_double(acc);
      if ((k.v[limb] >> bit) & 1) {
        xyzz_add_affine(acc, g);
        started = true;
      }
    }
  }
  st_xyzz(out + i, acc);
}
}  // namespace h2b

extern "C" int h2b_g1_mul_generator(h2b_ctx* ctx, const h2b_fr* scalars, int loc, size_t n,
                                    h2b_g1_affine* out, int out_loc) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (n && (!scalars || !out)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_sc = reinterpret_cast<const Fr*>(scalars);
  if (loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 0, n * sizeof(Fr)));
    H2B_TRY(copy_h2d_any(ctx, ctx->stage[0], scalars, n * sizeof(Fr), ctx->stream));
    d_sc = reinterpret_cast<const Fr*>(ctx->stage[0]);
  }
  G1Affine* d_out = reinterpret_cast<G1Affine*>(out);
  if (out_loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 1, n * sizeof(G1Affine)));
    d_out = reinterpret_cast<G1Affine*>(ctx->stage[1]);
  }
  H2B_TRY(ensure_scratch(ctx, n * sizeof(G1Xyzz)));
  G1Xyzz* tmp = reinterpret_cast<G1Xyzz*>(ctx->scratch);
  H2B_TRY(launch(ctx, g1_mul_generator_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0, d_sc, tmp,
                 (uint64_t)n));
  const uint32_t nblk = (uint32_t)(((n + kNormBatch - 1) / kNormBatch + 127) / 128);
  H2B_TRY(launch(ctx, msm_table_normalize_kernel, dim3(nblk), dim3(128), 0, (const G1Xyzz*)tmp, d_out, (uint64_t)n));
  if (out_loc != H2B_DEVICE)
    H2B_CUDA(ctx, cudaMemcpyAsync(out, d_out, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_bases_upload(h2b_ctx* ctx, const h2b_g1_affine* bases, size_t n, int loc,
                                h2b_bases** out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!out || (!bases && n)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  h2b_bases* b = new h2b_bases();
  b->ctx = ctx;
  b->n = n;
  b->d_pts = nullptr;
  cudaError_t e = dev_malloc(ctx, (void**)&b->d_pts, (n ? n : 1) * sizeof(G1Affine));
  if (e == cudaSuccess && n) {
    if (loc == H2B_DEVICE)
      e = cudaMemcpyAsync(b->d_pts, bases, n * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx->stream);
    else if (copy_h2d_any(ctx, b->d_pts, bases, n * sizeof(G1Affine), ctx->stream) != H2B_OK)
      e = cudaErrorInvalidValue;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    if (b->d_pts) cudaFree(b->d_pts);
    delete b;
    return fail(ctx, e == cudaErrorMemoryAllocation ? H2B_ERR_OOM : H2B_ERR_CUDA,
                cudaGetErrorString(e));
  }
  *out = b;
  return H2B_OK;
}

extern "C" void h2b_bases_free(h2b_bases* b) {
  if (!b) return;
  std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  cudaFree(b->d_pts);
  if (b->d_table) cudaFree(b->d_table);
  delete b;
}

// Builds the window table of a base set (one-time, at upload: the bases of a
// ParamsKZG are immutable).  window_bits = 0 picks the width from n.
extern "C" int h2b_bases_precompute(h2b_ctx* ctx, h2b_bases* b, uint32_t window_bits) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!b || b->ctx != ctx) return fail(ctx, H2B_ERR_ARG, "bases belong to another context");
  if (b->n == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t c = window_bits;
  if (c == 0) {
    // minimise  n * ceil(255 / c)  bucket additions  +  ~2.8 * 2^(c-1)  addition-equivalents of bucket reduction
    double best = 0;
    for (uint32_t cc = 8; cc <= 24; ++cc) {
      const double cost = (double)b->n * ((255 + cc - 1) / cc) + 2.8 * (double)(1u << (cc - 1));
      if (c == 0 || cost < best) {
        best = cost;
        c = cc;
      }
    }
  }
  if (c < 2 || c > 24) return fail(ctx, H2B_ERR_ARG, "window_bits out of range");
  const uint32_t W = (255 + c - 1) / c;
  if ((uint64_t)W * b->n >= (1ull << 31)) return fail(ctx, H2B_ERR_ARG, "window table larger than 2^31 points");
  if (b->d_table) {
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(b->d_table);
    b->d_table = nullptr;
    b->pre_c = b->pre_W = 0;
  }
  G1Affine* table = nullptr;
  G1Xyzz* tmp = nullptr;
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&table, (size_t)W * b->n * sizeof(G1Affine)));
  cudaError_t e = dev_malloc(ctx, (void**)&tmp, b->n * sizeof(G1Xyzz));
  if (e != cudaSuccess) {
    cudaFree(table);
    return fail(ctx, H2B_ERR_OOM, cudaGetErrorString(e));
  }
  int rc = H2B_OK;
  e = cudaMemcpyAsync(table, b->d_pts, b->n * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) rc = fail(ctx, H2B_ERR_CUDA, cudaGetErrorString(e));
  const uint32_t nblk = (uint32_t)((b->n + 127) / 128);
  const uint32_t nblk_norm = (uint32_t)(((b->n + kNormBatch - 1) / kNormBatch + 127) / 128);
  for (uint32_t w = 1; w < W && rc == H2B_OK; ++w) {
    rc = launch(ctx, msm_table_double_kernel, dim3(nblk), dim3(128), 0,
                (const G1Affine*)(table + (size_t)(w - 1) * b->n), tmp, (uint64_t)b->n, c);
    if (rc == H2B_OK)
      rc = launch(ctx, msm_table_normalize_kernel, dim3(nblk_norm), dim3(128), 0, (const G1Xyzz*)tmp,
                  table + (size_t)w * b->n, (uint64_t)b->n);
  }
  if (rc == H2B_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    rc = fail(ctx, H2B_ERR_CUDA, "window table build failed");
  cudaFree(tmp);
  if (rc != H2B_OK) {
    cudaFree(table);
    return rc;
  }
  b->d_table = table;
  b->pre_c = c;
  b->pre_W = W;
  return H2B_OK;
}

extern "C" uint32_t h2b_msm_window_bits(size_t n) { return msm_plan(n).c; }
extern "C" uint32_t h2b_bases_table_window_bits(const h2b_bases* b) { return b ? b->pre_c : 0; }
extern "C" size_t h2b_bases_len(const h2b_bases* b) { return b ? b->n : 0; }
extern "C" void* h2b_bases_device_ptr(const h2b_bases* b) { return b ? (void*)b->d_pts : nullptr; }

static int msm_common(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                      const h2b_fr* scalars, int loc, size_t n, G1Xyzz* acc) {
  // bases are immutable once uploaded (and precomputed): any context of the same device may read them,
  // so that independent commitments can run concurrently on several contexts (one stream each)
  if (!bases || bases->ctx->device != ctx->device) return fail(ctx, H2B_ERR_ARG, "bases live on another device");
  if (!scalars && n) return fail(ctx, H2B_ERR_ARG, "null pointer");
  // assert_eq!(coeffs.len(), bases.len()) / assert!(bases.len() >= size):
  // arithmetic.rs:133, kzg/commitment.rs:290,332
  if (base_offset > bases->n || n > bases->n - base_offset)
    return fail(ctx, H2B_ERR_LENGTH, "more scalars than bases");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const Fr* d_scalars = reinterpret_cast<const Fr*>(scalars);
  const Fr* h_scalars = nullptr;
  if (loc != H2B_DEVICE && n) {
    H2B_TRY(ensure_stage(ctx, 0, n * sizeof(Fr)));
    h_scalars = d_scalars;  // copied inside msm_run, chunk by chunk, under the digit kernel
    d_scalars = reinterpret_cast<const Fr*>(ctx->stage[0]);
  }
  if (bases->d_table && n >= 1024)  // commit on resident bases: all windows share one bucket set
    return msm_run(ctx, bases->d_table + base_offset, d_scalars, n, acc, bases->n, bases->pre_c, h_scalars);
  return msm_run(ctx, bases->d_pts + base_offset, d_scalars, n, acc, 0, 0, h_scalars);
}

static void xyzz_to_jacobian(const G1Xyzz& p, h2b_g1* out) {
  // (X/ZZ, Y/ZZZ) with Z := ZZ: X_j = X*ZZ, Y_j = Y*ZZZ (ZZ^3 = ZZZ^2)
  Fq xj = Fq::zero(), yj = Fq::one(), zj = Fq::zero();
  if (!p.is_identity()) {
    xj = mul(p.x, p.zz);
    yj = mul(p.y, p.zzz);
    zj = p.zz;
  }
  memcpy(&out->x, &xj, 32);
  memcpy(&out->y, &yj, 32);
  memcpy(&out->z, &zj, 32);
}

extern "C" int h2b_msm(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                       const h2b_fr* scalars, int loc, size_t n, h2b_g1* out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  G1Xyzz acc;
  H2B_TRY(msm_common(ctx, bases, base_offset, scalars, loc, n, &acc));
  xyzz_to_jacobian(acc, out);
  return H2B_OK;
}

extern "C" int h2b_msm_multi_affine(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                                    const h2b_fr* const* scalars_dev, uint32_t ncols, size_t n,
                                    h2b_g1_affine* out_affine) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (ncols == 0) return H2B_OK;
  if (!out_affine || !scalars_dev) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (!bases || bases->ctx->device != ctx->device) return fail(ctx, H2B_ERR_ARG, "bases live on another device");
  for (uint32_t j = 0; j < ncols; ++j)
    if (!scalars_dev[j] && n) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (base_offset > bases->n || n > bases->n - base_offset)
    return fail(ctx, H2B_ERR_LENGTH, "more scalars than bases");  // kzg/commitment.rs:290,332
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t W = bases->pre_c ? (255 + bases->pre_c - 1) / bases->pre_c : 0;
  const bool fused = bases->d_table && n >= 1024 && ncols > 1 && ncols <= 64 &&
                     (uint64_t)n * W * ncols < (1ull << 32) && !getenv("H2B_MSM_NO_MULTI");
  G1Xyzz acc[64];
  if (fused) {
    H2B_TRY(msm_run(ctx, bases->d_table + base_offset, nullptr, n, acc, bases->n, bases->pre_c, nullptr, ncols,
                    reinterpret_cast<const Fr* const*>(scalars_dev)));
    for (uint32_t j = 0; j < ncols; ++j) {
      const G1Affine a = xyzz_to_affine(acc[j]);
      memcpy(out_affine + j, &a, sizeof a);
    }
    return H2B_OK;
  }
  for (uint32_t j = 0; j < ncols; ++j) {  // no window table (or too small / too large): one MSM per column
    H2B_TRY(msm_common(ctx, bases, base_offset, scalars_dev[j], H2B_DEVICE, n, &acc[0]));
    const G1Affine a = xyzz_to_affine(acc[0]);
    memcpy(out_affine + j, &a, sizeof a);
  }
  return H2B_OK;
}

extern "C" int h2b_msm_affine(h2b_ctx* ctx, const h2b_bases* bases, size_t base_offset,
                              const h2b_fr* scalars, int loc, size_t n,
                              h2b_g1_affine* out_affine) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!out_affine) return fail(ctx, H2B_ERR_ARG, "null pointer");
  G1Xyzz acc;
  H2B_TRY(msm_common(ctx, bases, base_offset, scalars, loc, n, &acc));
  const G1Affine a = xyzz_to_affine(acc);
  memcpy(out_affine, &a, 64);
  return H2B_OK;
}

// Sum of n affine points on the host: the fold of per-shard partial results
// (arithmetic.rs:153 `results.iter().fold(identity, |a, b| a + b)`), used when
// an MSM is sharded by point range across GPUs.
extern "C" int h2b_g1_sum(const h2b_g1_affine* pts, size_t n, h2b_g1_affine* out) {
  if ((!pts && n) || !out) return H2B_ERR_ARG;
  G1Xyzz acc = G1Xyzz::identity();
  for (size_t i = 0; i < n; ++i) {
    G1Affine p;
    memcpy(&p, pts + i, 64);
    xyzz_add_affine(acc, p);
  }
  const G1Affine a = xyzz_to_affine(acc);
  memcpy(out, &a, 64);
  return H2B_OK;
}

// ---------------------------------------------------------------------------
// small_multiexp (arithmetic.rs:105-125): double-and-add with the doublings shared across
// the points.  A handful of points, host-side in the reference and here.
// ---------------------------------------------------------------------------
extern "C" int h2b_small_multiexp(const h2b_fr* coeffs, const h2b_g1_affine* bases, size_t n, h2b_g1* out) {
  if ((n && (!coeffs || !bases)) || !out) return H2B_ERR_ARG;
  std::vector<Fr> repr(n);
  std::vector<G1Affine> pts(n);
  for (size_t i = 0; i < n; ++i) {
    Fr c;
    memcpy(&c, coeffs + i, 32);
    repr[i] = from_mont(c);  // to_repr(), :106
    memcpy(&pts[i], bases + i, 64);
  }
  G1Xyzz acc = G1Xyzz::identity();
  for (int byte_idx = 31; byte_idx >= 0; --byte_idx)
    for (int bit_idx = 7; bit_idx >= 0; --bit_idx) {
      acc = xyzz_double(acc);
      const int bit = byte_idx * 8 + bit_idx;
      for (size_t i = 0; i < n; ++i)
        if ((repr[i].v[bit >> 5] >> (bit & 31)) & 1) xyzz_add_affine(acc, pts[i]);
    }
  xyzz_to_jacobian(acc, out);
  return H2B_OK;
}

// ---------------------------------------------------------------------------
// g_to_lagrange (arithmetic.rs:277-301): best_fft over curve points with omega^-1, every
// point scaled by 1/n, batch-normalised.  The one caller on the KZG side is
// ParamsKZG::downsize (poly/kzg/commitment.rs:267-275).  Radix-2 decimation in time on XYZZ
// points: bit-reversal gather, then log n butterfly sweeps; the twiddle of a butterfly
// is a 254-bit scalar multiplication of its odd input (group_scale, arithmetic.rs:214-225),
// skipped where the twiddle is 1; the 1/n scaling is one more scalar multiplication folded
// into the last sweep's outputs.  O(n log n) scalar multiplications, all independent.
// ---------------------------------------------------------------------------
namespace h2b {
H2B_D G1Xyzz xyzz_scale(const G1Xyzz& p, const Fr& k_canonical) {
  G1Xyzz acc = G1Xyzz::identity();
  if (p.is_identity()) return acc;
  bool started = false;
  for (int limb = 7; limb >= 0; --limb) {
    const uint32_t w = k_canonical.v[limb];
    if (!started && w == 0) continue;
    for (int bit = 31; bit >= 0; --bit) {
      if (started) acc = xyzz_double(acc);
      if ((w >> bit) & 1) {
        xyzz_add(acc, p);
        started = true;
      }
    }
  }
  return acc;
}

__global__ void __launch_bounds__(128)
    g1_fft_bitrev_kernel(const G1Affine* in, G1Xyzz* out, uint32_t k) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1ull << k)) return;
  uint32_t j = 0;
  for (uint32_t b = 0; b < k; ++b) j |= (((uint32_t)i >> b) & 1u) << (k - 1 - b);
  G1Affine p;
  p.x = ld_fp(&in[j].x);
  p.y = ld_fp(&in[j].y);
  st_xyzz(out + i, G1Xyzz::from_affine(p));
}

// sweep `u` (1-based): butterflies (r0, r0 + 2^(u-1)) with twiddle w^(i * 2^(k-u)); final_scale != nullptr on the
// last sweep multiplies both outputs by that scalar (canonical form)
__global__ void __launch_bounds__(128)
    g1_fft_sweep_kernel(G1Xyzz* a, uint32_t k, uint32_t u, const Fr* tw_lo, const Fr* tw_hi, uint32_t h,
                        const Fr* final_scale) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (1ull << k) / 2) return;
  const uint32_t half = 1u << (u - 1);
  const uint32_t i = (uint32_t)b & (half - 1);
  const uint64_t r0 = ((b >> (u - 1)) << u) + i, r1 = r0 + half;
  G1Xyzz x = ld_xyzz(a + r0), t = ld_xyzz(a + r1);
  if (i) {
    const uint64_t e = (uint64_t)i << (k - u);
    Fr w = mul(ld_fp_nc(tw_lo + (uint32_t)(e & ((1ull << h) - 1))), ld_fp_nc(tw_hi + (uint32_t)(e >> h)));
    t = xyzz_scale(t, from_mont(w));
  }
  G1Xyzz s = x, d = x;
  xyzz_add(s, t);
  t.y = neg(t.y);
  xyzz_add(d, t);
  if (final_scale) {
    const Fr f = ld_fp_nc(final_scale);
    s = xyzz_scale(s, f);
    d = xyzz_scale(d, f);
  }
  st_xyzz(a + r0, s);
  st_xyzz(a + r1, d);
}

__global__ void g1_scale_all_kernel(G1Xyzz* a, uint64_t n, const Fr* scale) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_xyzz(a + i, xyzz_scale(ld_xyzz(a + i), ld_fp_nc(scale)));
}
}  // namespace h2b

extern "C" int h2b_g_to_lagrange(h2b_ctx* ctx, const h2b_g1_affine* g, int loc, uint32_t k, h2b_g1_affine* out,
                                 int out_loc) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!g || !out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (k > 28) return fail(ctx, H2B_ERR_ARG, "k > 28 (Fr two-adicity)");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)1 << k;
  // omega_inv = ROOT_OF_UNITY_INV^(2^(S-k)), n_inv = TWO_INV^k        (arithmetic.rs:278-282)
  static const uint64_t kRootInv[4] = {0x0ed3e50a414e6dbaull, 0xb22625f59115aba7ull, 0x1bbe587180f34361ull,
                                       0x048127174daabc26ull};
  Fr w;
  for (int i = 0; i < 4; ++i) {
    w.v[2 * i] = (uint32_t)kRootInv[i];
    w.v[2 * i + 1] = (uint32_t)(kRootInv[i] >> 32);
  }
  w = to_mont(w);
  for (uint32_t i = k; i < 28; ++i) w = sqr(w);
  Fr n_inv = inv(add(Fr::one(), Fr::one()));
  {
    Fr acc = Fr::one();
    for (uint32_t i = 0; i < k; ++i) acc = mul(acc, n_inv);
    n_inv = from_mont(acc);  // canonical: the kernels walk its bits
  }
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, w, k, &tw));
  const G1Affine* d_in = reinterpret_cast<const G1Affine*>(g);
  G1Affine* d_out = reinterpret_cast<G1Affine*>(out);
  if (loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 0, n * sizeof(G1Affine)));
    H2B_TRY(copy_h2d_any(ctx, ctx->stage[0], g, n * sizeof(G1Affine), ctx->stream));
    d_in = reinterpret_cast<const G1Affine*>(ctx->stage[0]);
  }
  if (out_loc != H2B_DEVICE) {
    H2B_TRY(ensure_stage(ctx, 1, n * sizeof(G1Affine)));
    d_out = reinterpret_cast<G1Affine*>(ctx->stage[1]);
  }
  H2B_TRY(ensure_scratch(ctx, n * sizeof(G1Xyzz) + 64));
  G1Xyzz* pts = reinterpret_cast<G1Xyzz*>(ctx->scratch);
  Fr* d_scale = reinterpret_cast<Fr*>(pts + n);
  H2B_CUDA(ctx, cudaMemcpyAsync(d_scale, &n_inv, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  H2B_TRY(launch(ctx, g1_fft_bitrev_kernel, dim3((uint32_t)((n + 127) / 128)), dim3(128), 0, d_in, pts, k));
  const uint32_t nb = (uint32_t)((n / 2 + 127) / 128);
  for (uint32_t u = 1; u <= k; ++u)
    H2B_TRY(launch(ctx, g1_fft_sweep_kernel, dim3(nb), dim3(128), 0, pts, k, u, (const Fr*)tw->d_lo,
                   (const Fr*)tw->d_hi, tw->h, u == k ? (const Fr*)d_scale : (const Fr*)nullptr));
  if (k == 0) H2B_TRY(launch(ctx, g1_scale_all_kernel, dim3(1), dim3(32), 0, pts, (uint64_t)1, (const Fr*)d_scale));
  const uint32_t nblk = (uint32_t)(((n + kNormBatch - 1) / kNormBatch + 127) / 128);
  H2B_TRY(launch(ctx, msm_table_normalize_kernel, dim3(nblk), dim3(128), 0, (const G1Xyzz*)pts, d_out, (uint64_t)n));
  if (out_loc != H2B_DEVICE) H2B_TRY(copy_d2h_any(ctx, out, d_out, n * sizeof(G1Affine), ctx->stream));
  H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return H2B_OK;
}

extern "C" int h2b_best_multiexp(h2b_ctx* ctx, const h2b_fr* coeffs, const h2b_g1_affine* bases,
                                 size_t n, h2b_g1* out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  h2b_bases* b = nullptr;
  H2B_TRY(h2b_bases_upload(ctx, bases, n, H2B_HOST, &b));
  const int rc = h2b_msm(ctx, b, 0, coeffs, H2B_HOST, n, out);
  h2b_bases_free(b);
  return rc;
}
