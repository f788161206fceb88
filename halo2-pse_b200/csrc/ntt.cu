// Radix-2 NTT over bn256 Fr for sm_100a: replaces `best_fft`
// (/root/reference/halo2_proofs/src/arithmetic.rs:171-274) and the
// EvaluationDomain transforms built on it (poly/domain.rs:226-361).
//
// Same function (natural order in, natural order out, X[K] = sum_j a[j] w^(jK)),
// different algorithm: instead of bit-reversal + log n radix-2 sweeps, the
// transform is split into P = ceil(k/8) passes over HBM.  With digits
// s_1..s_P (R_p = 2^s_p, m_1 = n, m_(p+1) = m_p / R_p) pass p computes, for
// every sub-problem q of m_p contiguous elements and every jr < m_(p+1),
//
//   Y[q*m_p + K*m_(p+1) + jr] = w_(m_p)^(jr*K) * sum_(j1<R_p) X[q*m_p + j1*m_(p+1) + jr] * w_(R_p)^(j1*K)
//
// i.e. an R_p-point DFT down a strided column followed by the inter-pass
// twiddle.  After the last pass the element at position K_1*m_2 + K_2*m_3 + ...
// is output K_1 + R_1*K_2 + R_1*R_2*K_3 + ..., so the last pass scatters its
// results to natural order (the "bit reversal" of the reference folded into
// one pass).  A tile is R_p rows x C columns (C*32 B contiguous in HBM).
//
// The R_p-point column DFT itself is done by R_p/8 threads holding 8 elements
// each: register radix-8 -> shared memory -> register radix-8 -> shared memory
// -> radix-4/2, with a conflict-free XOR-swizzled limb-plane layout.  Scaling
// steps of the domain transforms (zeta coset, zero padding, 1/n, division by
// the vanishing polynomial, truncation) are fused into the first pass's load
// and the last pass's store.
#include "common.cuh"

#include <algorithm>

namespace h2b {

struct PeerPtrs {
  Fr* p[16];
};

struct PassParams {
  const Fr* in;
  Fr* out;
  uint64_t in_bstride, out_bstride;  // elements between batch members
  uint32_t k;                        // log2 n
  uint32_t lm;                       // log2 m_p
  uint32_t s;                        // log2 R_p
  uint32_t lc;                       // log2 C
  uint32_t first, last, single;
  // sb: a SINGLE-pass transform (2^s points, s = 6..9) run by the register kernel: the C columns of a tile are C
  // consecutive batch members instead of C consecutive strided columns of one transform
  uint32_t sb;
  // Batch-interleaved scratch between the last two passes of row transforms that end in peer stores: the pass before
  // the last one writes element o of batch member b at ((b / il) * n + o) * il + b % il (il_out = il), the last pass
  // reads it back with `il` consecutive batch members as its tile columns (il_in = il = its tile width), so that the
  // stores of a warp into the transposed matrix on a peer are runs of il * 32 contiguous bytes instead of isolated
  // 32-byte words (the run dimension of the destination is the row = batch member).
  uint32_t il_out, il_in;
  // lazy_out: this pass and the next one both run on the register kernel, which computes on lazy residues ([0, 2p),
  // field.cuh): the scratch between them keeps the unreduced values, the last pass reduces once
  uint32_t lazy_out;
  uint32_t s1;                       // log2 R_1
  uint32_t nmid;                     // number of middle digits (s_2 .. s_(P-1))
  uint32_t mid_s[4];                 // their widths, s_2 first
  uint64_t n_in, n_out;
  const Fr* pre;
  uint32_t pre_mod;
  const Fr* post;
  uint32_t post_mod;
  const Fr* tw_lo;
  const Fr* tw_hi;
  uint32_t h;
  const Fr* rt;
  uint32_t rt_log;
  // the intra-pass roots as (plain value, Shoup companion) pairs: the register kernel multiplies by them with
  // mul_shoup (92 wide multiplies instead of 128; 29 of the 37 products of a pass)
  const Fr* rts;
  // single-level inter-pass twiddle table of the later passes: w = tw1[(jr*K) << tw1_shift]
  const Fr* tw1;
  uint32_t tw1_shift;
  // first pass only: inter-pass twiddle of every OUTPUT element, tw_out[K * m_2 + jr] = w^(jr * K), read with
  // the same coalesced pattern as the store (one multiplication, no gather)
  const Fr* tw_out;
  // Last pass of a batch of ROW transforms fused with the distributed transpose that follows it in the four-step
  // NTT (sc_cl != 0): output Ko of row (sc_row0 + batch member) is multiplied by w_big^(row * Ko) (sc_lo != nullptr)
  // and stored straight into the buffer of the rank that owns it in the transposed matrix, at its final place
  // sc_peers[Ko / sc_cl][(Ko % sc_cl) * sc_R + row] -- NVLink-mapped peer pointers: the exchange IS these stores.
  PeerPtrs sc_peers;
  uint32_t sc_cl, sc_h;
  uint64_t sc_R, sc_row0;
  const Fr* sc_lo;
  const Fr* sc_hi;
};

// What a pass does around its R-point column DFTs; a template parameter of the register kernel, so that each
// instantiation carries only the multiplications it needs (the all-in-one kernel was 16.8k instructions = 268 KB
// of straight-line code, and ncu showed it stalling on instruction fetch).
enum PassKind {
  KIND_TWOLEVEL = 0,  // not last: twiddle = tw_lo[e & mask] * tw_hi[e >> h]   (first pass without a table)
  KIND_OUT_TABLE = 1, // first pass: twiddle read from tw_out at the output index
  KIND_MID_TABLE = 2, // later passes: twiddle = tw1[(jr * K) << shift]
  KIND_LAST = 3,      // last pass: natural-order scatter, optional post-scale, truncation
  KIND_LAST_PEER = 4  // last pass of row transforms, fused with the four-step exchange (peer stores, optional twiddle)
};

struct Tile {
  uint64_t in_base, in_rs, in_cs;
  uint64_t out_base, out_rs;
  uint64_t jr0;
  uint32_t out_cm;  // multiplier of the column index in the output position (0 in sb mode: the column is a batch member)
};

H2B_HD Tile tile_geom(const PassParams& p, uint64_t t) {
  Tile g;
  g.out_cm = (p.sb || p.il_in) ? 0u : 1u;
  if (p.single) {
    g.in_base = 0;
    g.in_rs = 1;
    g.in_cs = 0;
    g.out_base = 0;
    g.out_rs = 1;
    g.jr0 = 0;
  } else if (!p.last) {
    const uint32_t lmn = p.lm - p.s;  // log2 m_(p+1)
    const uint64_t jb = t & ((1ull << (lmn - p.lc)) - 1);
    const uint64_t q = t >> (lmn - p.lc);
    g.in_base = (q << p.lm) + (jb << p.lc);
    g.in_rs = 1ull << lmn;
    g.in_cs = 1;
    g.out_base = g.in_base;
    g.out_rs = g.in_rs;
    g.jr0 = jb << p.lc;
  } else {
    // columns = consecutive values of the lowest output digit K_1
    const uint64_t b1 = t & ((1ull << (p.s1 - p.lc)) - 1);
    uint64_t rest = t >> (p.s1 - p.lc);
    g.in_base = ((b1 << p.lc) << (p.k - p.s1)) + (rest << p.s);
    g.in_cs = 1ull << (p.k - p.s1);
    g.in_rs = 1;
    uint32_t shift = 0;
    for (uint32_t i = 0; i < p.nmid; ++i) shift += p.mid_s[i];
    uint64_t acc = 0;
    for (int i = (int)p.nmid - 1; i >= 0; --i) {
      const uint64_t d = rest & ((1ull << p.mid_s[i]) - 1);
      rest >>= p.mid_s[i];
      shift -= p.mid_s[i];
      acc |= d << shift;
    }
    g.out_base = (b1 << p.lc) + (acc << p.s1);
    g.out_rs = 1ull << (p.k - p.s);
    g.jr0 = 0;
  }
  return g;
}

// Out-of-line product for the twiddle multiplications of the register kernel (H2B_NTT_CALL_MUL): trades a call per
// product for a 3x smaller kernel body.
#if defined(H2B_NTT_CALL_MUL) && defined(__CUDA_ARCH__)
__device__ __noinline__ Fr mul_tw(Fr a, Fr b) { return mul(a, b); }
#else
H2B_D Fr mul_tw(const Fr& a, const Fr& b) { return mul(a, b); }
#endif
// the same without the final subtraction: a below 2p, b canonical -> below 1.38p
H2B_D Fr mul_tw_lazy(const Fr& a, const Fr& b) { return mul<FrParams, false>(a, b); }

template <bool PRE>
H2B_D Fr load_in_t(const PassParams& p, const Fr* in, uint64_t gi) {
  Fr x = (gi < p.n_in) ? ld_fp(in + gi) : Fr::zero();
  if (PRE) x = mul(x, ld_fp_nc(p.pre + (uint32_t)gi % p.pre_mod));
  return x;
}
H2B_D Fr load_in(const PassParams& p, const Fr* in, uint64_t gi) {
  return (p.first && p.pre) ? load_in_t<true>(p, in, gi) : load_in_t<false>(p, in, gi);
}

template <int KIND>
H2B_D void store_out_t(const PassParams& p, Fr* out, const Tile& g, uint32_t K, uint32_t c, Fr x, uint64_t bidx) {
  if (KIND == KIND_LAST_PEER) {
    const uint64_t Ko = g.out_base + (uint64_t)K * g.out_rs + c * g.out_cm;
    const uint64_t row = p.sc_row0 + bidx;
    if (p.sc_lo && row * Ko) {
      const uint64_t e = row * Ko;
      x = mul_tw_lazy(x, ld_fp_nc(p.sc_lo + (uint32_t)(e & ((1ull << p.sc_h) - 1))));
      x = mul_tw(x, ld_fp_nc(p.sc_hi + (uint32_t)(e >> p.sc_h)));
    } else {
      x = canon(x);  // the register kernel hands over lazy residues
    }
    const uint32_t h = (uint32_t)(Ko / p.sc_cl);
    st_fp(p.sc_peers.p[h] + (Ko - (uint64_t)h * p.sc_cl) * p.sc_R + row, x);
    return;
  }
  if (KIND == KIND_LAST) {
    const uint64_t Ko = g.out_base + (uint64_t)K * g.out_rs + c * g.out_cm;
    if (p.post)
      x = mul_tw(x, ld_fp_nc(p.post + (uint32_t)Ko % p.post_mod));
    else
      x = canon(x);  // the register kernel hands over lazy residues
    if (Ko < p.n_out) st_fp(out + Ko, x);
    return;
  }
  const uint64_t o = g.out_base + (uint64_t)K * g.out_rs + c;
  Fr* dst = out + o;
  if (p.il_out) dst = p.out + ((bidx / p.il_out) * p.out_bstride + o) * p.il_out + bidx % p.il_out;
  if (KIND == KIND_OUT_TABLE) {
    x = mul_tw_lazy(x, ld_fp_nc(p.tw_out + o));
  } else if (KIND == KIND_MID_TABLE) {
    const uint64_t idx = ((g.jr0 + c) * (uint64_t)K) << p.tw1_shift;
    x = mul_tw_lazy(x, ld_fp_nc(p.tw1 + idx));
  } else {
    const uint64_t e = ((g.jr0 + c) * (uint64_t)K) << (p.k - p.lm);
    const uint32_t elo = (uint32_t)e & ((1u << p.h) - 1u);
    const uint32_t ehi = (uint32_t)(e >> p.h);
    x = mul_tw_lazy(x, ld_fp_nc(p.tw_lo + elo));
    x = mul_tw_lazy(x, ld_fp_nc(p.tw_hi + ehi));
  }
  if (!p.lazy_out) x = canon(x);  // the next pass is the generic kernel (canonical residues only)
  st_fp(dst, x);
}

// the generic kernel keeps the run-time dispatch (tiny transforms only)
H2B_D void store_out(const PassParams& p, Fr* out, const Tile& g, uint32_t K, uint32_t c, Fr x) {
  if (p.last && p.sc_cl)
    store_out_t<KIND_LAST_PEER>(p, out, g, K, c, x, blockIdx.y);
  else if (p.last)
    store_out_t<KIND_LAST>(p, out, g, K, c, x, blockIdx.y);
  else if (p.tw_out)
    store_out_t<KIND_OUT_TABLE>(p, out, g, K, c, x, blockIdx.y);
  else if (p.tw1)
    store_out_t<KIND_MID_TABLE>(p, out, g, K, c, x, blockIdx.y);
  else
    store_out_t<KIND_TWOLEVEL>(p, out, g, K, c, x, blockIdx.y);
}

H2B_HD uint32_t bitrev32(uint32_t x, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < bits; ++i) {
    r = (r << 1) | (x & 1u);
    x >>= 1;
  }
  return r;
}

// ---------------------------------------------------------------------------
// Generic pass: any s <= 8, any C; radix-2 sweeps in shared memory.  Used for
// the digit widths the register kernel does not cover (tiny transforms and the
// odd leftover digit) -- never for the k = 16..28 benchmark shapes.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ntt_pass_generic(PassParams p) {
  H2B_DYN_SMEM(smem_raw);
  Fr* tile = reinterpret_cast<Fr*>(smem_raw);
  const uint32_t R = 1u << p.s, C = 1u << p.lc;
  const Tile g = tile_geom(p, blockIdx.x);
  const Fr* in = p.in + (uint64_t)blockIdx.y * p.in_bstride;
  Fr* out = p.out + (uint64_t)blockIdx.y * p.out_bstride;
  for (uint32_t idx = threadIdx.x; idx < R * C; idx += blockDim.x) {
    const uint32_t row = idx >> p.lc, col = idx & (C - 1);
    const uint64_t gi = g.in_base + row * g.in_rs + col * g.in_cs;
    tile[(bitrev32(row, p.s) << p.lc) + col] = load_in(p, in, gi);
  }
  __syncthreads();
  for (uint32_t u = 1; u <= p.s; ++u) {
    const uint32_t half = 1u << (u - 1);
    for (uint32_t idx = threadIdx.x; idx < (R / 2) * C; idx += blockDim.x) {
      const uint32_t col = idx & (C - 1), bi = idx >> p.lc;
      const uint32_t i = bi & (half - 1), blk = bi >> (u - 1);
      const uint32_t r0 = (blk << u) + i, r1 = r0 + half;
      Fr a = tile[(r0 << p.lc) + col], b = tile[(r1 << p.lc) + col];
      if (i) b = mul(b, ld_fp_nc(p.rt + (i << (p.rt_log - u))));
      tile[(r0 << p.lc) + col] = add(a, b);
      tile[(r1 << p.lc) + col] = sub(a, b);
    }
    __syncthreads();
  }
  for (uint32_t idx = threadIdx.x; idx < R * C; idx += blockDim.x) {
    const uint32_t K = idx >> p.lc, col = idx & (C - 1);
    store_out(p, out, g, K, col, tile[idx]);
  }
}

// ---------------------------------------------------------------------------
// Register/shared-memory pass for s in {6,7,8}
// ---------------------------------------------------------------------------
// 8-point DFT, natural order in and out: x[K] <- sum_a x[a] * w8^(a*K).
// w8 = rt[N/8], w4 = rt[N/4], w8^3 = rt[3N/8] with N = 2^rt_log.
// x * rts-root number i (a table constant: Shoup product), lazy residues in and out
H2B_D Fr mul_root(const Fr& x, const Fr* rts, uint32_t i) {
  return mul_shoup<FrParams, false>(x, ld_fp_nc(rts + 2 * i), ld_fp_nc(rts + 2 * i + 1));
}

// a + b / a - b whose only consumer is a product (Shoup with any multiplier, Montgomery with a canonical one: both take
// an operand below 4p and return a lazy residue): W leaves them uncorrected
template <bool W>
H2B_D Fr addw(const Fr& a, const Fr& b) {
  return W ? add_wide(a, b) : add_lazy(a, b);
}
template <bool W>
H2B_D Fr subw(const Fr& a, const Fr& b) {
  return W ? sub_wide(a, b) : sub_lazy(a, b);
}

// 8-point DFT on lazy residues ([0, 2p) in; out: x[0] in [0, 2p), x[1..7] below 4p if WIDE -- every one of them is
// multiplied by a twiddle next -- else in [0, 2p))
template <bool WIDE>
H2B_D void dft8(Fr* x, const Fr* rts, uint32_t rt_log) {
  // differences that feed a Shoup product directly stay uncorrected (sub_wide: below 4p)
  Fr s0 = add_lazy(x[0], x[4]), d0 = sub_lazy(x[0], x[4]);
  Fr s1 = add_lazy(x[1], x[5]), d1 = sub_wide(x[1], x[5]);
  Fr s2 = add_lazy(x[2], x[6]), d2 = sub_wide(x[2], x[6]);
  Fr s3 = add_lazy(x[3], x[7]), d3 = sub_wide(x[3], x[7]);
  const uint32_t i4 = 1u << (rt_log - 2);
  const Fr w4 = ld_fp_nc(rts + 2 * i4), w4s = ld_fp_nc(rts + 2 * i4 + 1);
  d1 = mul_root(d1, rts, 1u << (rt_log - 3));
  d2 = mul_shoup<FrParams, false>(d2, w4, w4s);
  d3 = mul_root(d3, rts, 3u << (rt_log - 3));
  // even outputs from s, odd outputs from d
  Fr e0 = add_lazy(s0, s2), f0 = sub_lazy(s0, s2);
  Fr e1 = add_lazy(s1, s3), f1 = mul_shoup<FrParams, false>(sub_wide(s1, s3), w4, w4s);
  x[0] = add_lazy(e0, e1);
  x[4] = subw<WIDE>(e0, e1);
  x[2] = addw<WIDE>(f0, f1);
  x[6] = subw<WIDE>(f0, f1);
  e0 = add_lazy(d0, d2);
  f0 = sub_lazy(d0, d2);
  e1 = add_lazy(d1, d3);
  f1 = mul_shoup<FrParams, false>(sub_wide(d1, d3), w4, w4s);
  x[1] = addw<WIDE>(e0, e1);
  x[5] = subw<WIDE>(e0, e1);
  x[3] = addw<WIDE>(f0, f1);
  x[7] = subw<WIDE>(f0, f1);
}

template <int S, int KIND, bool PRE>
__global__ void __launch_bounds__(256, 2) ntt_pass_fast(PassParams p) {
  constexpr int R = 1 << S, T = R / 8, LC = 11 - S, C = 1 << LC;
  constexpr int LM2 = S - 6, M2 = 1 << LM2;  // 8, 4, 2, 1
  constexpr int PLANE = 2048;
  H2B_DYN_SMEM(smem_raw);
  uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
  const uint32_t tid = threadIdx.x;
  const uint32_t c = tid & (C - 1), u = tid >> LC;
  const Tile g = tile_geom(p, blockIdx.x);
  // batch member of this thread's column: the grid row, or (sb) one of the C members of this tile
  const uint64_t bidx = p.sb ? (uint64_t)blockIdx.x * C + c : p.il_in ? (uint64_t)blockIdx.y * C + c : (uint64_t)blockIdx.y;
  const Fr* in = p.in + bidx * p.in_bstride;
  Fr* out = p.out + bidx * p.out_bstride;
  const uint32_t rsh = p.rt_log - S;  // w_R^e = rt[e << rsh]

  auto slot = [&](uint32_t pos) -> uint32_t {
    return ((pos ^ ((pos >> LM2) & (M2 - 1))) << LC) + c;
  };
#ifdef H2B_NTT_SMEM_PLANES
  auto put = [&](uint32_t pos, const Fr& v) {
    const uint32_t w = slot(pos);
#pragma unroll
    for (int l = 0; l < 8; ++l) sm[l * PLANE + w] = v.v[l];
  };
  auto get = [&](uint32_t pos) -> Fr {
    const uint32_t w = slot(pos);
    Fr v;
#pragma unroll
    for (int l = 0; l < 8; ++l) v.v[l] = sm[l * PLANE + w];
    return v;
  };
#else
  // an element = two 16-byte words (4 shared-memory accesses per exchange instead of 16); a 128-bit access is served
  // per quarter-warp = 8 consecutive columns, 32 bytes apart: the halves swap on bit 2 of the element index, so that
  // the eight 16-byte words land on all 32 banks
  uint4* sm4 = reinterpret_cast<uint4*>(sm);
  auto put = [&](uint32_t pos, const Fr& v) {
    const uint32_t w = slot(pos), hs = (w >> 2) & 1u;
    sm4[2 * w + hs] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
    sm4[2 * w + (hs ^ 1u)] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
  };
  auto get = [&](uint32_t pos) -> Fr {
    const uint32_t w = slot(pos), hs = (w >> 2) & 1u;
    const uint4 a = sm4[2 * w + hs], b = sm4[2 * w + (hs ^ 1u)];
    Fr v;
    v.v[0] = a.x; v.v[1] = a.y; v.v[2] = a.z; v.v[3] = a.w;
    v.v[4] = b.x; v.v[5] = b.y; v.v[6] = b.z; v.v[7] = b.w;
    return v;
  };
#endif

  if (KIND == KIND_OUT_TABLE) {
    // the twiddles of this tile's outputs are a stream read once, right before the stores: ask L2 for them now
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t K = (tid >> LC) + (uint32_t)j * T;
      const Fr* tp = p.tw_out + (g.out_base + (uint64_t)K * g.out_rs + c);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(tp));
    }
#endif
  }
  Fr x[8];
  // Two forms of the same rounds.  Measured on B200 (k = 24 / 26): the looped form is faster for the passes that are
  // not the last one at S <= 8 (1.26 -> 1.22, 1.23 -> 1.22 ms), slower for last passes (0.94 -> 0.99) and for S = 9
  // (three looped rounds spill 200 bytes): each pass kind takes its better form.
  constexpr bool LASTK = KIND == KIND_LAST || KIND == KIND_LAST_PEER;
  constexpr bool LOOPED = S <= 8 && !LASTK;
  // outputs that a twiddle product consumes next may stay below 4p (dft8<true>, addw / subw<true>): every output of a
  // round that is followed by intra-pass twiddles, and every output of a pass that is not the last one
  constexpr bool WOUT = !LASTK;
  if constexpr (!LOOPED) {
  // round 1: b = u, elements a*T + u
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const uint32_t row = a * T + u;
    if (KIND == KIND_LAST_PEER && p.il_in)  // interleaved scratch: C batch members side by side
      x[a] = ld_fp(p.in + ((uint64_t)blockIdx.y * p.in_bstride + g.in_base + row * g.in_rs) * C + c);
    else
      x[a] = load_in_t<PRE>(p, in, g.in_base + row * g.in_rs + c * g.in_cs);
  }
  dft8<true>(x, p.rts, p.rt_log);
#pragma unroll
  for (int Ka = 1; Ka < 8; ++Ka) x[Ka] = mul_root(x[Ka], p.rts, (u * Ka) << rsh);
#pragma unroll
  for (int Ka = 0; Ka < 8; ++Ka) put(Ka * T + u, x[Ka]);
  __syncthreads();
  // round 2: (Ka, b2) = (u / M2, u % M2), elements Ka*T + a2*M2 + b2
  const uint32_t Ka = u >> LM2, b2 = u & (M2 - 1);
#pragma unroll
  for (int a2 = 0; a2 < 8; ++a2) x[a2] = get(Ka * T + a2 * M2 + b2);
  dft8<(M2 != 1) || WOUT>(x, p.rts, p.rt_log);
  if (M2 == 1) {
#pragma unroll
    for (int Ka2 = 0; Ka2 < 8; ++Ka2) store_out_t<KIND>(p, out, g, Ka + 8 * Ka2, c, x[Ka2], bidx);
    return;
  }
#pragma unroll
  for (int Ka2 = 1; Ka2 < 8; ++Ka2)
    x[Ka2] = mul_root(x[Ka2], p.rts, (8 * b2 * Ka2) << rsh);
#pragma unroll
  for (int Ka2 = 0; Ka2 < 8; ++Ka2) put(Ka * T + Ka2 * M2 + b2, x[Ka2]);
  __syncthreads();
  // round 3: groups gq = Ka*8 + Ka2 (64 per column), M2-point DFT over b2
  if (M2 == 8) {
    const uint32_t gq = u;  // 64 threads per column, one group each
#pragma unroll
    for (int b = 0; b < 8; ++b) x[b] = get(gq * 8 + b);
    dft8<WOUT>(x, p.rts, p.rt_log);
    const uint32_t K0 = (gq >> 3) + 8 * (gq & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) store_out_t<KIND>(p, out, g, K0 + 64 * i, c, x[i], bidx);
  } else if (M2 == 4) {
    const uint32_t i4 = 1u << (p.rt_log - 2);
    const Fr w4 = ld_fp_nc(p.rts + 2 * i4), w4s = ld_fp_nc(p.rts + 2 * i4 + 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t gq = u + 32 * i;
      Fr v0 = get(gq * 4 + 0), v1 = get(gq * 4 + 1), v2 = get(gq * 4 + 2), v3 = get(gq * 4 + 3);
      Fr t0 = add_lazy(v0, v2), t1 = sub_lazy(v0, v2), t2 = add_lazy(v1, v3), t3 = mul_shoup<FrParams, false>(sub_wide(v1, v3), w4, w4s);
      const uint32_t K0 = (gq >> 3) + 8 * (gq & 7);
      store_out_t<KIND>(p, out, g, K0, c, addw<WOUT>(t0, t2), bidx);
      store_out_t<KIND>(p, out, g, K0 + 64, c, addw<WOUT>(t1, t3), bidx);
      store_out_t<KIND>(p, out, g, K0 + 128, c, subw<WOUT>(t0, t2), bidx);
      store_out_t<KIND>(p, out, g, K0 + 192, c, subw<WOUT>(t1, t3), bidx);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t gq = u + 16 * i;
      Fr v0 = get(gq * 2 + 0), v1 = get(gq * 2 + 1);
      const uint32_t K0 = (gq >> 3) + 8 * (gq & 7);
      store_out_t<KIND>(p, out, g, K0, c, addw<WOUT>(v0, v1), bidx);
      store_out_t<KIND>(p, out, g, K0 + 64, c, subw<WOUT>(v0, v1), bidx);
    }
  }
  } else {
  // The radix-8 rounds share ONE copy of their code (a loop that is not unrolled): the kernel is straight-line
  // multi-precision arithmetic, ~10 k instructions when every round is inlined, and ncu showed 0.7 - 1.0 stall cycles
  // per issue on instruction fetch.  Round r works in place on the shared-memory positions base + j * stride:
  //   round 0: elements j*T + u (from global memory), twiddles w_R^(u j);
  //   round 1: (Ka, b2) = (u / M2, u % M2), elements Ka*T + j*M2 + b2, twiddles w_R^(8 b2 j);
  //   round 2 (S = 9 only): elements 8u + j, no twiddles: straight to the output.
  const uint32_t Ka = u >> LM2, b2 = u & (M2 - 1);
  constexpr int NR8 = M2 == 8 ? 3 : 2;  // radix-8 rounds
  uint32_t base = u, stride = T, e = u;
#pragma unroll 1
  for (int round = 0; round < NR8; ++round) {
    if (round == 0) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const uint32_t row = a * T + u;
        if (KIND == KIND_LAST_PEER && p.il_in)  // interleaved scratch: C batch members side by side
          x[a] = ld_fp(p.in + ((uint64_t)blockIdx.y * p.in_bstride + g.in_base + row * g.in_rs) * C + c);
        else
          x[a] = load_in_t<PRE>(p, in, g.in_base + row * g.in_rs + c * g.in_cs);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 8; ++a) x[a] = get(base + a * stride);
    }
    dft8<true>(x, p.rts, p.rt_log);
    if (round == NR8 - 1 && (M2 == 1 || M2 == 8)) break;  // the last radix-8 round of S = 6 and S = 9 feeds the output
#pragma unroll
    for (int j = 1; j < 8; ++j) x[j] = mul_root(x[j], p.rts, (e * j) << rsh);
#pragma unroll
    for (int j = 0; j < 8; ++j) put(base + j * stride, x[j]);
    __syncthreads();
    if (round == 0) {
      base = Ka * T + b2;
      stride = M2;
      e = 8 * b2;
    } else {
      base = u * 8;
      stride = 1;
    }
  }
  if (M2 == 1) {
#pragma unroll
    for (int Ka2 = 0; Ka2 < 8; ++Ka2) store_out_t<KIND>(p, out, g, Ka + 8 * Ka2, c, x[Ka2], bidx);
  } else if (M2 == 8) {
    const uint32_t K0 = (u >> 3) + 8 * (u & 7);
#pragma unroll
    for (int i = 0; i < 8; ++i) store_out_t<KIND>(p, out, g, K0 + 64 * i, c, x[i], bidx);
  } else if (M2 == 4) {
    const uint32_t i4 = 1u << (p.rt_log - 2);
    const Fr w4 = ld_fp_nc(p.rts + 2 * i4), w4s = ld_fp_nc(p.rts + 2 * i4 + 1);
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const uint32_t gq = u + 32 * i;
      Fr v0 = get(gq * 4 + 0), v1 = get(gq * 4 + 1), v2 = get(gq * 4 + 2), v3 = get(gq * 4 + 3);
      Fr t0 = add_lazy(v0, v2), t1 = sub_lazy(v0, v2), t2 = add_lazy(v1, v3),
         t3 = mul_shoup<FrParams, false>(sub_wide(v1, v3), w4, w4s);
      const uint32_t K0 = (gq >> 3) + 8 * (gq & 7);
      store_out_t<KIND>(p, out, g, K0, c, addw<WOUT>(t0, t2), bidx);
      store_out_t<KIND>(p, out, g, K0 + 64, c, addw<WOUT>(t1, t3), bidx);
      store_out_t<KIND>(p, out, g, K0 + 128, c, subw<WOUT>(t0, t2), bidx);
      store_out_t<KIND>(p, out, g, K0 + 192, c, subw<WOUT>(t1, t3), bidx);
    }
  } else {
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const uint32_t gq = u + 16 * i;
      Fr v0 = get(gq * 2 + 0), v1 = get(gq * 2 + 1);
      const uint32_t K0 = (gq >> 3) + 8 * (gq & 7);
      store_out_t<KIND>(p, out, g, K0, c, addw<WOUT>(v0, v1), bidx);
      store_out_t<KIND>(p, out, g, K0 + 64, c, subw<WOUT>(v0, v1), bidx);
    }
  }
  }  // LOOPED
}

// ---------------------------------------------------------------------------
// Element-wise kernels
// ---------------------------------------------------------------------------
// a[i] *= tab[i % mod]          (divide_by_vanishing_poly, domain.rs:307-326)
__global__ void scale_mod_kernel(Fr* a, uint64_t n, const Fr* tab, uint32_t mod) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    st_fp(a + i, mul(ld_fp(a + i), ld_fp_nc(tab + (uint32_t)i % mod)));
}

// out[2i] = plain value of the Montgomery-form tab[i], out[2i + 1] = its Shoup companion (mul_shoup's operands)
__global__ void shoup_table_kernel(const Fr* tab, Fr* out, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const Fr wm = ld_fp(tab + i);
  st_fp(out + 2 * i, from_mont(wm));
  st_fp(out + 2 * i + 1, shoup_companion(wm));
}

// tab[i] = base^(i << shift), i < count
__global__ void pow_table_kernel(Fr* tab, Fr base, uint32_t shift, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  uint64_t e = (uint64_t)i << shift;
  Fr r = Fr::one(), b = base;
  while (e) {
    if (e & 1) r = mul(r, b);
    b = sqr(b);
    e >>= 1;
  }
  tab[i] = r;
}

// tab[K * m2 + jr] = w^(jr * K), K < 2^s1, jr < m2 = 2^(k - s1): the first pass's inter-pass twiddles in the order
// the pass stores its outputs (two-level product, once per (omega, k))
__global__ void tw_out_table_kernel(Fr* tab, uint32_t k, uint32_t s1, const Fr* tw_lo, const Fr* tw_hi, uint32_t h) {
  const uint64_t n = 1ull << k;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t K = i >> (k - s1), jr = i & ((1ull << (k - s1)) - 1);
    const uint64_t e = jr * K;
    st_fp(tab + i, mul(ld_fp_nc(tw_lo + (uint32_t)(e & ((1ull << h) - 1))), ld_fp_nc(tw_hi + (uint32_t)(e >> h))));
  }
}

// ---------------------------------------------------------------------------
// Four-step helpers (single giant NTT sharded across GPUs; no counterpart in
// the reference, SURVEY.md 2.2 "four-step transpose / twiddle")
// ---------------------------------------------------------------------------
// out[b][c][r] = in[b * in_bstride + r * in_rstride + c],  r < rows, c < cols
__global__ void __launch_bounds__(256)
    transpose_kernel(const Fr* in, Fr* out, uint32_t rows, uint32_t cols, uint64_t in_rstride,
                     uint64_t in_bstride, uint64_t out_bstride) {
  __shared__ uint32_t tile[8][32][33];  // limb planes, padded: conflict-free both ways
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const Fr* src = in + (uint64_t)blockIdx.z * in_bstride;
  Fr* dst = out + (uint64_t)blockIdx.z * out_bstride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < rows && c < cols) {
      const Fr v = ld_fp(src + (uint64_t)r * in_rstride + c);
#pragma unroll
      for (int l = 0; l < 8; ++l) tile[l][ty + 8 * i][tx] = v.v[l];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < rows && c < cols) {
      Fr v;
#pragma unroll
      for (int l = 0; l < 8; ++l) v.v[l] = tile[l][tx][ty + 8 * i];
      st_fp(dst + (uint64_t)c * rows + r, v);
    }
  }
}

// Transpose fused with the exchange: rank `rank` holds rows [rank*rows, (rank+1)*rows) of a
// global R x cols matrix (R = G*rows); element (r, c) is stored straight into the buffer of
// the peer h = c / cl that owns column c of the transposed matrix, at its final place
// dst_h[(c % cl) * R + rank*rows + r].  Peer buffers are NVLink-mapped device pointers
// (symmetric memory), so the all-to-all IS these stores: no pack, no collective, no unpack.
__global__ void __launch_bounds__(256)
    transpose_scatter_kernel(const Fr* in, PeerPtrs peers, uint32_t rank, uint32_t rows, uint32_t cols,
                             uint32_t cl, uint64_t R) {
  __shared__ uint32_t tile[8][32][33];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < rows && c < cols) {
      const Fr v = ld_fp(in + (uint64_t)r * cols + c);
#pragma unroll
      for (int l = 0; l < 8; ++l) tile[l][ty + 8 * i][tx] = v.v[l];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < rows && c < cols) {
      Fr v;
#pragma unroll
      for (int l = 0; l < 8; ++l) v.v[l] = tile[l][tx][ty + 8 * i];
      Fr* dst = peers.p[c / cl];
      st_fp(dst + (uint64_t)(c % cl) * R + (uint64_t)rank * rows + r, v);
    }
  }
}

// out[b][a][c] = in[a][b][c]   (a < A, b < B, c < C; C-element runs stay contiguous)
__global__ void permute3_kernel(const Fr* in, Fr* out, uint32_t A, uint32_t B, uint32_t C) {
  const uint64_t total = (uint64_t)A * B * C;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = (uint32_t)(i % C);
    const uint64_t ab = i / C;
    const uint32_t a = (uint32_t)(ab % A), b = (uint32_t)(ab / A);  // i indexes out[b][a][c]
    st_fp(out + i, ld_fp(in + ((uint64_t)a * B + b) * C + c));
  }
}

// a[r][c] *= omega^((row0 + r) * c),  r < nrows, c < ncols; exponent < 2^tw.k
__global__ void twiddle_rows_kernel(Fr* a, uint64_t row0, uint32_t nrows, uint32_t ncols,
                                    const Fr* tw_lo, const Fr* tw_hi, uint32_t h) {
  const uint64_t total = (uint64_t)nrows * ncols;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = row0 + i / ncols, c = i % ncols;
    const uint64_t e = r * c;
    if (e == 0) continue;
    Fr x = ld_fp(a + i);
    x = mul(x, ld_fp_nc(tw_lo + (uint32_t)(e & ((1ull << h) - 1))));
    x = mul(x, ld_fp_nc(tw_hi + (uint32_t)(e >> h)));
    st_fp(a + i, x);
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
static bool omega_has_order(const Fr& omega, uint32_t k) {
  if (k == 0) return omega == Fr::one();
  Fr t = omega;
  for (uint32_t i = 0; i + 1 < k; ++i) t = sqr(t);
  return t == neg(Fr::one());  // omega^(2^(k-1)) = -1  <=>  exact order 2^k
}

static int ntt_plan(uint32_t k, uint32_t* s);

int ntt_get_table(h2b_ctx* ctx, const Fr& omega, uint32_t k, const TwTable** out) {
  for (auto& t : ctx->tw)
    if (t.k == k && t.omega == omega) {
      *out = &t;
      return H2B_OK;
    }
  if (!omega_has_order(omega, k))
    return fail(ctx, H2B_ERR_BAD_OMEGA, "omega is not a primitive 2^log_n-th root of unity");
  TwTable t;
  t.omega = omega;
  t.k = k;
  t.h = (k + 1) / 2;
  const uint32_t nlo = 1u << t.h, nhi = 1u << (k - t.h);
  const uint32_t rt_log = k < 9 ? k : 9;
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&t.d_lo, (size_t)nlo * sizeof(Fr)));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&t.d_hi, (size_t)nhi * sizeof(Fr)));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&t.d_rt, ((size_t)1 << rt_log) * sizeof(Fr)));
  H2B_TRY(launch(ctx, pow_table_kernel, dim3((nlo + 127) / 128), dim3(128), 0, t.d_lo, omega, 0u,
                 nlo));
  H2B_TRY(launch(ctx, pow_table_kernel, dim3((nhi + 127) / 128), dim3(128), 0, t.d_hi, omega, t.h,
                 nhi));
  H2B_TRY(launch(ctx, pow_table_kernel, dim3(((1u << rt_log) + 127) / 128), dim3(128), 0, t.d_rt,
                 omega, k - rt_log, 1u << rt_log));
  H2B_CUDA(ctx, dev_malloc(ctx, (void**)&t.d_rts, ((size_t)2 << rt_log) * sizeof(Fr)));
  H2B_TRY(launch(ctx, shoup_table_kernel, dim3(((1u << rt_log) + 127) / 128), dim3(128), 0, (const Fr*)t.d_rt, t.d_rts,
                 1u << rt_log));
  // single-level tables (later passes: 2^(k-6) entries at most; first pass: folded full table, opt-in)
  if (k > 8) {
    uint32_t s[4];
    ntt_plan(k, s);
    t.mid_log = k - s[0];
    const uint32_t nmid = 1u << t.mid_log;
    if (dev_malloc(ctx, (void**)&t.d_mid, (size_t)nmid * sizeof(Fr)) == cudaSuccess) {
      H2B_TRY(launch(ctx, pow_table_kernel, dim3((nmid + 127) / 128), dim3(128), 0, t.d_mid, omega, k - t.mid_log,
                     nmid));
    } else {
      t.d_mid = nullptr;
      cudaGetLastError();
    }
    // First pass: one table entry per OUTPUT element, streamed with the store's own coalesced pattern (n * 32 B:
    // 512 MiB at k = 24).  One multiplication instead of the two of the two-level tables; the extra HBM read is
    // free next to the arithmetic (a pass moves 64 B per element in the time of ~4 modular multiplications).
    // (A folded table indexed by the exponent was measured slower in round 1: random 32-byte reads.)
    uint32_t out_lo = 18, out_hi = 26;
    if (const char* e = getenv("H2B_NTT_OUT_TABLE")) {
      if (atoi(e) == 0) out_lo = 99;
      else if (atoi(e) > 1) out_lo = (uint32_t)atoi(e);
    }
    if (k >= out_lo && k <= out_hi) {
      const uint64_t nfull = 1ull << k;
      if (dev_malloc(ctx, (void**)&t.d_out, (size_t)nfull * sizeof(Fr)) == cudaSuccess) {
        H2B_TRY(launch(ctx, tw_out_table_kernel, dim3((uint32_t)ctx->sm_count * 16), dim3(256), 0, t.d_out, k, s[0],
                       (const Fr*)t.d_lo, (const Fr*)t.d_hi, t.h));
        t.out_s1 = s[0];
      } else {
        t.d_out = nullptr;
        cudaGetLastError();
      }
    }
  }
  ctx->tw.push_back(t);
  *out = &ctx->tw.back();
  return H2B_OK;
}

void ntt_free_tables(h2b_ctx* ctx) {
  for (auto& t : ctx->tw) {
    cudaFree(t.d_lo);
    cudaFree(t.d_hi);
    cudaFree(t.d_rt);
    if (t.d_rts) cudaFree(t.d_rts);
    if (t.d_out) cudaFree(t.d_out);
    if (t.d_mid) cudaFree(t.d_mid);
  }
  ctx->tw.clear();
}

// Digit widths of the passes, ascending so the widest (fewest columns, best
// coalescing on its strided load) comes last.
static int ntt_plan(uint32_t k, uint32_t* s) {
  if (k <= 8) {
    s[0] = k;
    return 1;
  }
  // fewest passes with digits <= 9; below 12 bits two digits of >= 6 are not possible anyway
  int P = (int)((k + 8) / 9);
  if (const char* e = getenv("H2B_NTT_MAX_DIGIT"))
    if (atoi(e) == 8) P = (int)((k + 7) / 8);
  const uint32_t base = k / P, rem = k % P;
  for (int i = 0; i < P; ++i) s[i] = base + ((uint32_t)i >= (uint32_t)P - rem ? 1u : 0u);
  return P;
}

// the register kernel for digit width s, pass kind and (first pass only) a fused pre-scale
typedef void (*FastKernel)(PassParams);
template <int S>
static FastKernel fast_kernel(int kind, bool pre) {
  switch (kind) {
    case KIND_TWOLEVEL: return pre ? ntt_pass_fast<S, KIND_TWOLEVEL, true> : ntt_pass_fast<S, KIND_TWOLEVEL, false>;
    case KIND_OUT_TABLE: return pre ? ntt_pass_fast<S, KIND_OUT_TABLE, true> : ntt_pass_fast<S, KIND_OUT_TABLE, false>;
    case KIND_MID_TABLE: return ntt_pass_fast<S, KIND_MID_TABLE, false>;
    case KIND_LAST_PEER: return ntt_pass_fast<S, KIND_LAST_PEER, false>;
    default: return pre ? ntt_pass_fast<S, KIND_LAST, true> : ntt_pass_fast<S, KIND_LAST, false>;  // pre: single-pass (sb) only
  }
}
static FastKernel fast_kernel(uint32_t s, int kind, bool pre) {
  switch (s) {
    case 6: return fast_kernel<6>(kind, pre);
    case 7: return fast_kernel<7>(kind, pre);
    case 8: return fast_kernel<8>(kind, pre);
    default: return fast_kernel<9>(kind, pre);
  }
}
static int launch_fast(h2b_ctx* ctx, uint32_t s, int kind, bool pre, dim3 grid, const PassParams& p) {
  return launch(ctx, fast_kernel(s, kind, pre), grid, dim3(256), 65536, p);
}

int ntt_run(h2b_ctx* ctx, const Fr* d_in, Fr* d_out, uint32_t k, const TwTable* tw, uint64_t n_in,
            const Fr* d_pre, uint32_t pre_mod, const Fr* d_post, uint32_t post_mod,
            uint64_t n_out, uint32_t batch, uint64_t in_stride, uint64_t out_stride, const NttScatter* sc) {
  if (batch == 0) return H2B_OK;
  if (k > 28) return fail(ctx, H2B_ERR_ARG, "log_n > 28 (Fr two-adicity)");
  const uint64_t n = 1ull << k;
  uint32_t s[4];
  const int P = ntt_plan(k, s);

  if (!ctx->ntt_attr_done) {
    H2B_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_generic,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (uint32_t sw = 6; sw <= 9; ++sw)
      for (int kind = 0; kind < 5; ++kind)
        for (int pre = 0; pre < 2; ++pre)
          H2B_CUDA(ctx, cudaFuncSetAttribute(fast_kernel(sw, kind, pre != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             65536));
    ctx->ntt_attr_done = true;
  }

  // Scratch holds the intermediate passes of a group of batch members.
  uint64_t kScratchCap = 8ull << 30;
  if (const char* e = getenv("H2B_NTT_SCRATCH_CAP")) {  // tests: force several column groups
    const long long v = atoll(e);
    if (v > 0) kScratchCap = (uint64_t)v;
  }
  uint32_t group = batch;
  if (P > 1) {
    const uint64_t per = n * sizeof(Fr);
    uint64_t g = kScratchCap / per;
    if (g < 1) g = 1;
    if (g < group) group = (uint32_t)g;
    H2B_TRY(ensure_scratch(ctx, (size_t)(per * group)));
  }
  Fr* scratch = reinterpret_cast<Fr*>(ctx->scratch);

  for (uint32_t b0 = 0; b0 < batch; b0 += group) {
    const uint32_t nb = (batch - b0 < group) ? batch - b0 : group;
    uint32_t lm = k;
    // which passes run on the register kernel (the rule of the loop below, evaluated ahead: a pass hands lazy residues
    // to the next one only if both do)
    bool fastv[5] = {false, false, false, false, false};
    {
      const uint32_t ilq = (sc && P >= 2 && s[P - 1] >= 6 && s[P - 1] <= 9 && nb % (1u << (11 - s[P - 1])) == 0 &&
                            getenv("H2B_NTT_NO_IL") == nullptr)
                               ? 1u << (11 - s[P - 1])
                               : 0u;
      uint32_t lmq = k;
      for (int pi = 0; pi < P; ++pi) {
        const bool lastq = pi == P - 1, singleq = P == 1;
        const uint32_t lcq = 11 - s[pi];
        const uint32_t availq = singleq ? 0 : (lastq ? s[0] : lmq - s[pi]);
        bool f = !singleq && s[pi] >= 6 && s[pi] <= 9 && lcq <= availq;
        if (singleq && s[pi] >= 6 && s[pi] <= 9 && nb >= (1u << lcq) && nb % (1u << lcq) == 0 &&
            getenv("H2B_NTT_NO_SB") == nullptr)
          f = true;
        else if (ilq && lastq)
          f = true;
        fastv[pi] = f;
        lmq -= s[pi];
      }
    }
    for (int pi = 0; pi < P; ++pi) {
      PassParams p;
      p.k = k;
      p.lm = lm;
      p.s = s[pi];
      p.first = pi == 0;
      p.last = pi == P - 1;
      p.single = P == 1;
      p.sb = 0;
      p.il_out = p.il_in = 0;
      p.lazy_out = (pi + 1 < P && fastv[pi] && fastv[pi + 1] && getenv("H2B_NTT_NO_LAZY") == nullptr) ? 1u : 0u;
      p.s1 = s[0];
      p.nmid = 0;
      for (int i = 1; i + 1 < P; ++i) p.mid_s[p.nmid++] = s[i];
      p.n_in = p.first ? n_in : n;
      p.n_out = p.last ? n_out : n;
      p.pre = p.first ? d_pre : nullptr;
      p.pre_mod = pre_mod ? pre_mod : 1;
      p.post = p.last ? d_post : nullptr;
      p.post_mod = post_mod ? post_mod : 1;
      p.tw_lo = tw->d_lo;
      p.tw_hi = tw->d_hi;
      p.h = tw->h;
      p.rt = tw->d_rt;
      p.rts = tw->d_rts;
      p.rt_log = k < 9 ? k : 9;
      p.tw1 = nullptr;
      p.tw1_shift = 0;
      p.tw_out = nullptr;
      p.sc_cl = 0;
      p.sc_lo = p.sc_hi = nullptr;
      p.sc_h = 0;
      p.sc_R = p.sc_row0 = 0;
      if (sc && p.last) {
        for (int i = 0; i < 16; ++i) p.sc_peers.p[i] = sc->peers[i];
        p.sc_cl = sc->cl;
        p.sc_R = sc->R;
        p.sc_row0 = sc->row0 + b0;
        if (sc->tw) {
          p.sc_lo = sc->tw->d_lo;
          p.sc_hi = sc->tw->d_hi;
          p.sc_h = sc->tw->h;
        }
      }
      if (!p.last && pi == 0 && tw->d_out && tw->out_s1 == s[0]) {
        p.tw_out = tw->d_out;
      } else if (!p.last && pi > 0 && tw->d_mid && lm <= tw->mid_log) {
        p.tw1 = tw->d_mid;
        p.tw1_shift = tw->mid_log - lm;
      }
      if (p.first) {
        p.in = d_in + (uint64_t)b0 * in_stride;
        p.in_bstride = in_stride;
      } else {
        p.in = scratch;
        p.in_bstride = n;
      }
      if (p.last) {
        p.out = d_out + (uint64_t)b0 * out_stride;
        p.out_bstride = out_stride;
      } else {
        p.out = scratch;
        p.out_bstride = n;
      }
      // peer-scattering row transforms of >= 2 passes: interleave the scratch of the last two passes (see PassParams)
      const uint32_t il = (sc && P >= 2 && s[P - 1] >= 6 && s[P - 1] <= 9 && nb % (1u << (11 - s[P - 1])) == 0 &&
                           getenv("H2B_NTT_NO_IL") == nullptr)
                              ? 1u << (11 - s[P - 1])
                              : 0u;
      if (il && pi == P - 2) p.il_out = il;
      if (il && pi == P - 1) p.il_in = il;
      // columns per tile
      uint32_t lc = 11 - p.s;
      const uint32_t avail = p.single ? 0 : (p.last ? p.s1 : lm - p.s);
      bool fast = !p.single && p.s >= 6 && p.s <= 9 && lc <= avail;
      // a single-pass transform of 64..512 points: the register kernel with batch members as tile columns
      const bool sb = p.single && p.s >= 6 && p.s <= 9 && nb >= (1u << lc) && nb % (1u << lc) == 0 &&
                      getenv("H2B_NTT_NO_SB") == nullptr;
      if (sb) {
        fast = true;
        p.sb = 1;
      } else if (p.il_in) {
        fast = true;  // the tile's columns are batch members: one strided column of the transform per tile
        lc = 0;
      } else if (lc > avail) {
        lc = avail;
      }
      p.lc = lc;
      if (fast != fastv[pi]) return fail(ctx, H2B_ERR_ARG, "internal: NTT pass plan mismatch");  // lazy_out relies on it
      const uint32_t tiles = sb ? nb >> (11 - p.s) : (uint32_t)(n >> (p.s + lc));
      const dim3 grid(tiles, sb ? 1u : p.il_in ? nb / p.il_in : nb);
      const bool prof = ctx->profile && b0 == 0 && pi < 5;
      if (prof) H2B_CUDA(ctx, cudaEventRecord(ctx->pass_ev[pi], ctx->stream));
      if (fast) {
        const int kind = (p.last && p.sc_cl) ? KIND_LAST_PEER : p.last ? KIND_LAST : p.tw_out ? KIND_OUT_TABLE : p.tw1 ? KIND_MID_TABLE : KIND_TWOLEVEL;
        const bool pre = p.first && p.pre;
        H2B_TRY(launch_fast(ctx, p.s, kind, pre, grid, p));
      } else {
        H2B_TRY(launch(ctx, ntt_pass_generic, grid, dim3(256), 65536, p));
      }
      if (prof) {
        H2B_CUDA(ctx, cudaEventRecord(ctx->pass_ev[pi + 1], ctx->stream));
        ctx->last_npass = pi + 1;
      }
      lm -= p.s;
    }
  }
  return H2B_OK;
}

}  // namespace h2b

// ===========================================================================
// C ABI
// ===========================================================================
using namespace h2b;

static const Fr* as_fr(const h2b_fr* p) { return reinterpret_cast<const Fr*>(p); }
static Fr* as_fr(h2b_fr* p) { return reinterpret_cast<Fr*>(p); }

namespace {
// Device view of a caller buffer: the pointer itself for H2B_DEVICE, a staged
// copy for H2B_HOST.
struct Staged {
  h2b_ctx* ctx;
  int which;
  Fr* dev = nullptr;
  int in(const h2b_fr* p, int loc, size_t count, bool copy) {
    if (loc == H2B_DEVICE) {
      dev = const_cast<Fr*>(as_fr(p));
      return H2B_OK;
    }
    H2B_TRY(ensure_stage(ctx, which, count * sizeof(Fr)));
    dev = reinterpret_cast<Fr*>(ctx->stage[which]);
    if (copy) H2B_TRY(copy_h2d_any(ctx, dev, p, count * sizeof(Fr), ctx->stream));
    return H2B_OK;
  }
  int out(h2b_fr* p, int loc, size_t count) {
    if (loc == H2B_DEVICE) return H2B_OK;
    H2B_TRY(copy_d2h_any(ctx, p, dev, count * sizeof(Fr), ctx->stream));
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2B_OK;
  }
};
}  // namespace

// ---------------------------------------------------------------------------
// Host-pointer batches that are too large to stage at once (64 columns of 2^26 extended evaluations are 128 GiB):
// column groups through two staging slots per direction.  Group g + 1 travels host -> device (copy stream) and
// group g - 1 device -> host (its own stream; PCIe is full duplex) while group g is transformed.  `run(din, dout,
// cols)` enqueues the transform of `cols` packed columns on the context's stream.
// ---------------------------------------------------------------------------
static size_t stream_slot_bytes() {
  size_t v = (size_t)1 << 30;
  if (const char* e = getenv("H2B_STREAM_SLOT_BYTES")) {
    const long long x = atoll(e);
    if (x >= 32) v = (size_t)x;
  }
  return v;
}

// columns per group, or 0 if the batch should take the plain (stage everything) path
static uint32_t stream_group(uint32_t ncols, size_t in_count, size_t out_count) {
  if (ncols < 2) return 0;
  const size_t per = std::max(in_count, out_count) * sizeof(Fr);
  const size_t slot = stream_slot_bytes();
  if ((size_t)ncols * per <= slot) return 0;  // fits one slot: nothing to overlap
  size_t g = slot / per;
  if (g < 1) g = 1;
  return (uint32_t)std::min<size_t>(g, (ncols + 1) / 2);  // at least two groups
}

template <class Run>
static int host_batch_streamed(h2b_ctx* ctx, const h2b_fr* in, size_t in_stride, size_t in_count, h2b_fr* out,
                               size_t out_stride, size_t out_count, uint32_t ncols, uint32_t group, Run run) {
  H2B_TRY(ensure_stage(ctx, 0, 2 * (size_t)group * in_count * sizeof(Fr)));
  H2B_TRY(ensure_stage(ctx, 1, 2 * (size_t)group * out_count * sizeof(Fr)));
  if (!ctx->d2h_stream) H2B_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  cudaStream_t st = ctx->stream, up = ctx->copy_stream, down = ctx->d2h_stream;
  cudaEvent_t *e_h2d = &ctx->copy_ev[0], *e_comp = &ctx->copy_ev[2], *e_d2h = &ctx->copy_ev[4];
  Fr* din = reinterpret_cast<Fr*>(ctx->stage[0]);
  Fr* dout = reinterpret_cast<Fr*>(ctx->stage[1]);
  // the staging buffers may still be read by work queued earlier on the main stream
  H2B_CUDA(ctx, cudaEventRecord(ctx->copy_ev[7], st));
  H2B_CUDA(ctx, cudaStreamWaitEvent(up, ctx->copy_ev[7], 0));
  H2B_CUDA(ctx, cudaStreamWaitEvent(down, ctx->copy_ev[7], 0));
  const uint32_t ngroups = (ncols + group - 1) / group;
  auto upload = [&](uint32_t g) -> int {
    const uint32_t slot = g & 1, c0 = g * group, nc = std::min(group, ncols - c0);
    if (g >= 2) H2B_CUDA(ctx, cudaStreamWaitEvent(up, e_comp[slot], 0));  // group g - 2 has left the slot
    Fr* d = din + (size_t)slot * group * in_count;
    for (uint32_t c = 0; c < nc; ++c)
      H2B_TRY(copy_h2d_any(ctx, d + (size_t)c * in_count, in + (size_t)(c0 + c) * in_stride, in_count * sizeof(Fr), up));
    H2B_CUDA(ctx, cudaEventRecord(e_h2d[slot], up));
    return H2B_OK;
  };
  H2B_TRY(upload(0));
  for (uint32_t g = 0; g < ngroups; ++g) {
    const uint32_t slot = g & 1, c0 = g * group, nc = std::min(group, ncols - c0);
    Fr* di = din + (size_t)slot * group * in_count;
    Fr* dq = dout + (size_t)slot * group * out_count;
    H2B_CUDA(ctx, cudaStreamWaitEvent(st, e_h2d[slot], 0));
    if (g >= 2) H2B_CUDA(ctx, cudaStreamWaitEvent(st, e_d2h[slot], 0));  // group g - 2 has been read back
    H2B_TRY(run(di, dq, nc));
    H2B_CUDA(ctx, cudaEventRecord(e_comp[slot], st));
    if (g + 1 < ngroups) H2B_TRY(upload(g + 1));
    H2B_CUDA(ctx, cudaStreamWaitEvent(down, e_comp[slot], 0));
    for (uint32_t c = 0; c < nc; ++c)
      H2B_TRY(copy_d2h_any(ctx, out + (size_t)(c0 + c) * out_stride, dq + (size_t)c * out_count, out_count * sizeof(Fr),
                           down));
    H2B_CUDA(ctx, cudaEventRecord(e_d2h[slot], down));
  }
  H2B_CUDA(ctx, cudaStreamSynchronize(down));
  H2B_CUDA(ctx, cudaStreamSynchronize(st));
  return H2B_OK;
}

extern "C" int h2b_best_fft_batch(h2b_ctx* ctx, h2b_fr* a, int loc, const h2b_fr* omega,
                                  uint32_t log_n, uint32_t ncols, size_t stride) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a || !omega) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (log_n > 28) return fail(ctx, H2B_ERR_ARG, "log_n > 28");
  const size_t n = (size_t)1 << log_n;
  if (ncols == 0) return H2B_OK;
  if (stride < n) return fail(ctx, H2B_ERR_LENGTH, "stride < 2^log_n");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, *as_fr(omega), log_n, &tw));
  if (loc != H2B_DEVICE)
    if (const uint32_t group = stream_group(ncols, n, n))
      return host_batch_streamed(ctx, a, stride, n, a, stride, n, ncols, group, [&](Fr* di, Fr* dq, uint32_t nc) {
        return ntt_run(ctx, di, dq, log_n, tw, n, nullptr, 1, nullptr, 1, n, nc, n, n);
      });
  Staged st{ctx, 0};
  const size_t count = (size_t)(ncols - 1) * stride + n;
  H2B_TRY(st.in(a, loc, count, true));
  H2B_TRY(ntt_run(ctx, st.dev, st.dev, log_n, tw, n, nullptr, 1, nullptr, 1, n, ncols, stride,
                  stride));
  return st.out(a, loc, count);
}

extern "C" int h2b_best_fft(h2b_ctx* ctx, h2b_fr* a, int loc, const h2b_fr* omega,
                            uint32_t log_n) {
  return h2b_best_fft_batch(ctx, a, loc, omega, log_n, 1, (size_t)1 << (log_n > 28 ? 0 : log_n));
}

// ---- EvaluationDomain ------------------------------------------------------
static Fr fr_from_u64_canonical(const uint64_t l[4]) {
  Fr r;
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = (uint32_t)l[i];
    r.v[2 * i + 1] = (uint32_t)(l[i] >> 32);
  }
  return to_mont(r);
}

static Fr fr_pow2k(Fr a, uint32_t times) {
  for (uint32_t i = 0; i < times; ++i) a = sqr(a);
  return a;
}

extern "C" int h2b_domain_new(h2b_ctx* ctx, uint32_t j, uint32_t k, h2b_domain** out) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (j < 1) return fail(ctx, H2B_ERR_ARG, "j < 1");  // j = 1 (quotient degree 0, extended_k = k) is what the reference's own tests build (domain.rs:494, kzg/commitment.rs:374)
  // Fr::ROOT_OF_UNITY (order 2^28) and Fr::ZETA, canonical values
  static const uint64_t kRoot[4] = {0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull,
                                    0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull};
  static const uint64_t kZeta[4] = {0x8b17ea66b99c90ddull, 0x5bfc41088d8daaa7ull,
                                    0xb3c4d79d41a91758ull, 0x0ull};
  const uint32_t S = 28;
  const uint32_t qd = j - 1;  // quotient_poly_degree, domain.rs:41
  if (k > S) return fail(ctx, H2B_ERR_ARG, "k exceeds Fr two-adicity 28");  // before any shift by k
  uint32_t ek = k;
  while (ek <= S && (1ull << ek) < (1ull << k) * qd) ++ek;  // domain.rs:49-52
  if (ek > S) return fail(ctx, H2B_ERR_ARG, "extended_k exceeds Fr two-adicity 28");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));

  h2b_domain* d = new h2b_domain();
  d->ctx = ctx;
  d->j = j;
  d->k = k;
  d->extended_k = ek;
  d->quotient_poly_degree = qd;
  d->extended_omega = fr_pow2k(fr_from_u64_canonical(kRoot), S - ek);  // domain.rs:54-61
  d->omega = fr_pow2k(d->extended_omega, ek - k);                        // domain.rs:70-73
  d->extended_omega_inv = inv(d->extended_omega);
  d->omega_inv = inv(d->omega);
  d->g_coset = fr_from_u64_canonical(kZeta);  // domain.rs:81
  d->g_coset_inv = sqr(d->g_coset);            // domain.rs:82
  Fr two = add(Fr::one(), Fr::one());
  Fr nk = Fr::one(), nek = Fr::one();
  for (uint32_t i = 0; i < k; ++i) nk = mul(nk, two);
  for (uint32_t i = 0; i < ek; ++i) nek = mul(nek, two);
  d->ifft_divisor = inv(nk);
  d->extended_ifft_divisor = inv(nek);
  // t_evaluations, domain.rs:84-124
  {
    Fr orig = fr_pow2k(d->g_coset, 0);
    // zeta^n with n = 2^k
    orig = fr_pow2k(d->g_coset, k);
    const Fr step = fr_pow2k(d->extended_omega, k);
    Fr cur = orig;
    const Fr one = Fr::one();
    do {
      d->t_evaluations.push_back(inv(sub(cur, one)));
      cur = mul(cur, step);
    } while (cur != orig && d->t_evaluations.size() < ((size_t)1 << (ek - k)) + 1);
    if (d->t_evaluations.size() != (size_t)1 << (ek - k)) {
      delete d;
      return fail(ctx, H2B_ERR_ARG, "t_evaluations cycle length mismatch");
    }
  }
  // device tables
  std::vector<Fr> zin = {Fr::one(), d->g_coset, d->g_coset_inv};
  std::vector<Fr> zout = {d->extended_ifft_divisor, mul(d->extended_ifft_divisor, d->g_coset_inv),
                          mul(d->extended_ifft_divisor, d->g_coset)};
  auto up = [&](Fr** dst, const std::vector<Fr>& v) -> int {
    H2B_CUDA(ctx, cudaMalloc((void**)dst, v.size() * sizeof(Fr)));
    H2B_CUDA(ctx, cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(Fr), cudaMemcpyHostToDevice,
                                  ctx->stream));
    H2B_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return H2B_OK;
  };
  int rc = up(&d->d_zeta_in, zin);
  if (rc == H2B_OK) rc = up(&d->d_ext_post, zout);
  if (rc == H2B_OK) rc = up(&d->d_ifft_post, std::vector<Fr>{d->ifft_divisor});
  if (rc == H2B_OK) rc = up(&d->d_t_inv, d->t_evaluations);
  if (rc != H2B_OK) {
    delete d;
    return rc;
  }
  *out = d;
  return H2B_OK;
}

extern "C" void h2b_domain_free(h2b_domain* d) {
  if (!d) return;
  std::lock_guard<std::recursive_mutex> lk(d->ctx->mu);
  cudaFree(d->d_zeta_in);
  cudaFree(d->d_ext_post);
  cudaFree(d->d_ifft_post);
  cudaFree(d->d_t_inv);
  delete d;
}

extern "C" uint32_t h2b_domain_k(const h2b_domain* d) { return d ? d->k : 0; }
extern "C" uint32_t h2b_domain_extended_k(const h2b_domain* d) { return d ? d->extended_k : 0; }
extern "C" size_t h2b_domain_quotient_len(const h2b_domain* d) {
  return d ? ((size_t)1 << d->k) * d->quotient_poly_degree : 0;
}

extern "C" int h2b_domain_constant(const h2b_domain* d, uint32_t which, h2b_fr* out) {
  if (!d || !out) return H2B_ERR_ARG;
  const Fr* src = nullptr;
  switch (which) {
    case 0: src = &d->omega; break;
    case 1: src = &d->omega_inv; break;
    case 2: src = &d->extended_omega; break;
    case 3: src = &d->extended_omega_inv; break;
    case 4: src = &d->g_coset; break;
    case 5: src = &d->g_coset_inv; break;
    case 6: src = &d->ifft_divisor; break;
    case 7: src = &d->extended_ifft_divisor; break;
    default:
      if (which - 8 < d->t_evaluations.size()) src = &d->t_evaluations[which - 8];
  }
  if (!src) return H2B_ERR_ARG;
  memcpy(out, src, sizeof(Fr));
  return H2B_OK;
}

extern "C" int h2b_lagrange_to_coeff_batch(h2b_domain* d, h2b_fr* a, int loc, uint32_t ncols,
                                           size_t stride) {
  if (!d) return H2B_ERR_ARG;
  h2b_ctx* ctx = d->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (ncols == 0) return H2B_OK;
  const size_t n = (size_t)1 << d->k;
  if (stride < n) return fail(ctx, H2B_ERR_LENGTH, "stride < n");  // domain.rs:227
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, d->omega_inv, d->k, &tw));
  if (loc != H2B_DEVICE)
    if (const uint32_t group = stream_group(ncols, n, n))
      return host_batch_streamed(ctx, a, stride, n, a, stride, n, ncols, group, [&](Fr* di, Fr* dq, uint32_t nc) {
        return ntt_run(ctx, di, dq, d->k, tw, n, nullptr, 1, d->d_ifft_post, 1, n, nc, n, n);
      });
  Staged st{ctx, 0};
  const size_t count = (size_t)(ncols - 1) * stride + n;
  H2B_TRY(st.in(a, loc, count, true));
  H2B_TRY(ntt_run(ctx, st.dev, st.dev, d->k, tw, n, nullptr, 1, d->d_ifft_post, 1, n, ncols,
                  stride, stride));
  return st.out(a, loc, count);
}

extern "C" int h2b_lagrange_to_coeff(h2b_domain* d, h2b_fr* a, int loc) {
  return h2b_lagrange_to_coeff_batch(d, a, loc, 1, d ? (size_t)1 << d->k : 0);
}

extern "C" int h2b_coeff_to_extended_batch(h2b_domain* d, const h2b_fr* in, size_t in_stride,
                                           h2b_fr* out, size_t out_stride, int loc,
                                           uint32_t ncols) {
  if (!d) return H2B_ERR_ARG;
  h2b_ctx* ctx = d->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!in || !out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (ncols == 0) return H2B_OK;
  const size_t n = (size_t)1 << d->k, ne = (size_t)1 << d->extended_k;
  if (in_stride < n || out_stride < ne)
    return fail(ctx, H2B_ERR_LENGTH, "stride shorter than the polynomial");  // domain.rs:244
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, d->extended_omega, d->extended_k, &tw));
  if (loc != H2B_DEVICE)
    if (const uint32_t group = stream_group(ncols, n, ne))
      return host_batch_streamed(ctx, in, in_stride, n, out, out_stride, ne, ncols, group,
                                 [&](Fr* di, Fr* dq, uint32_t nc) {
                                   return ntt_run(ctx, di, dq, d->extended_k, tw, n, d->d_zeta_in, 3, nullptr, 1, ne, nc, n,
                                                  ne);
                                 });
  Staged sin{ctx, 0}, sout{ctx, 1};
  const size_t cin = (size_t)(ncols - 1) * in_stride + n;
  const size_t cout = (size_t)(ncols - 1) * out_stride + ne;
  H2B_TRY(sin.in(in, loc, cin, true));
  H2B_TRY(sout.in(out, loc, cout, false));
  H2B_TRY(ntt_run(ctx, sin.dev, sout.dev, d->extended_k, tw, n, d->d_zeta_in, 3, nullptr, 1, ne,
                  ncols, in_stride, out_stride));
  return sout.out(out, loc, cout);
}

extern "C" int h2b_coeff_to_extended(h2b_domain* d, const h2b_fr* in, h2b_fr* out, int loc) {
  if (!d) return H2B_ERR_ARG;
  return h2b_coeff_to_extended_batch(d, in, (size_t)1 << d->k, out, (size_t)1 << d->extended_k,
                                     loc, 1);
}

extern "C" int h2b_extended_to_coeff_batch(h2b_domain* d, const h2b_fr* in, size_t in_stride,
                                           h2b_fr* out, size_t out_stride, int loc,
                                           uint32_t ncols, int divide_by_vanishing) {
  if (!d) return H2B_ERR_ARG;
  h2b_ctx* ctx = d->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  const size_t ne = (size_t)1 << d->extended_k;
  const size_t nq = ((size_t)1 << d->k) * d->quotient_poly_degree;
  if (!in || (!out && nq)) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (ncols == 0 || nq == 0) return H2B_OK;  // j = 1: truncate(0), an empty Vec (domain.rs:299-300)
  if (in_stride < ne || out_stride < nq)
    return fail(ctx, H2B_ERR_LENGTH, "stride shorter than the polynomial");  // domain.rs:282
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, d->extended_omega_inv, d->extended_k, &tw));
  if (loc != H2B_DEVICE)
    if (const uint32_t group = stream_group(ncols, ne, nq))
      return host_batch_streamed(ctx, in, in_stride, ne, out, out_stride, nq, ncols, group,
                                 [&](Fr* di, Fr* dq, uint32_t nc) {
                                   return ntt_run(ctx, di, dq, d->extended_k, tw, ne,
                                                  divide_by_vanishing ? d->d_t_inv : nullptr,
                                                  (uint32_t)d->t_evaluations.size(), d->d_ext_post, 3, nq, nc, ne, nq);
                                 });
  Staged sin{ctx, 0}, sout{ctx, 1};
  const size_t cin = (size_t)(ncols - 1) * in_stride + ne;
  const size_t cout = (size_t)(ncols - 1) * out_stride + nq;
  H2B_TRY(sin.in(in, loc, cin, true));
  H2B_TRY(sout.in(out, loc, cout, false));
  H2B_TRY(ntt_run(ctx, sin.dev, sout.dev, d->extended_k, tw, ne,
                  divide_by_vanishing ? d->d_t_inv : nullptr,
                  (uint32_t)d->t_evaluations.size(), d->d_ext_post, 3, nq, ncols, in_stride,
                  out_stride));
  return sout.out(out, loc, cout);
}

extern "C" int h2b_extended_to_coeff(h2b_domain* d, const h2b_fr* in, h2b_fr* out, int loc,
                                     int divide_by_vanishing) {
  if (!d) return H2B_ERR_ARG;
  return h2b_extended_to_coeff_batch(d, in, (size_t)1 << d->extended_k, out,
                                     h2b_domain_quotient_len(d), loc, 1, divide_by_vanishing);
}

extern "C" int h2b_divide_by_vanishing_poly(h2b_domain* d, h2b_fr* a, int loc) {
  if (!d) return H2B_ERR_ARG;
  h2b_ctx* ctx = d->ctx;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a) return fail(ctx, H2B_ERR_ARG, "null pointer");
  const size_t ne = (size_t)1 << d->extended_k;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  Staged st{ctx, 0};
  H2B_TRY(st.in(a, loc, ne, true));
  const uint32_t blocks = (uint32_t)((ne + 255) / 256 < 148 * 16 ? (ne + 255) / 256 : 148 * 16);
  H2B_TRY(launch(ctx, scale_mod_kernel, dim3(blocks), dim3(256), 0, st.dev, (uint64_t)ne,
                 (const Fr*)d->d_t_inv, (uint32_t)d->t_evaluations.size()));
  return st.out(a, loc, ne);
}

// ---- four-step helpers (device pointers only) --------------------------------
extern "C" int h2b_fr_transpose_batch(h2b_ctx* ctx, const h2b_fr* in, h2b_fr* out, uint32_t rows,
                                      uint32_t cols, size_t in_row_stride, uint32_t nbatch,
                                      size_t in_batch_stride, size_t out_batch_stride) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!in || !out || in == out) return fail(ctx, H2B_ERR_ARG, "null or aliased pointer");
  if (in_row_stride < cols) return fail(ctx, H2B_ERR_LENGTH, "row stride < cols");
  if (rows == 0 || cols == 0 || nbatch == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  return launch(ctx, transpose_kernel, dim3((cols + 31) / 32, (rows + 31) / 32, nbatch), dim3(256), 0,
                as_fr(in), as_fr(out), rows, cols, (uint64_t)in_row_stride, (uint64_t)in_batch_stride,
                (uint64_t)out_batch_stride);
}

extern "C" int h2b_fr_permute3(h2b_ctx* ctx, const h2b_fr* in, h2b_fr* out, uint32_t A, uint32_t B,
                               uint32_t C) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!in || !out || in == out) return fail(ctx, H2B_ERR_ARG, "null or aliased pointer");
  const uint64_t total = (uint64_t)A * B * C;
  if (total == 0) return H2B_OK;
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint64_t want = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
  return launch(ctx, permute3_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0, as_fr(in),
                as_fr(out), A, B, C);
}

extern "C" int h2b_fr_twiddle_rows(h2b_ctx* ctx, h2b_fr* a, const h2b_fr* omega, uint32_t log_n,
                                   uint64_t row0, uint32_t nrows, uint32_t ncols) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!a || !omega) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (log_n > 28) return fail(ctx, H2B_ERR_ARG, "log_n > 28");
  const uint64_t total = (uint64_t)nrows * ncols;
  if (total == 0) return H2B_OK;
  if ((row0 + nrows - 1) * (uint64_t)(ncols - 1) >= (1ull << log_n))
    return fail(ctx, H2B_ERR_LENGTH, "twiddle exponent exceeds 2^log_n");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, *as_fr(omega), log_n, &tw));
  const uint64_t want = (total + 255) / 256, cap = (uint64_t)ctx->sm_count * 16;
  return launch(ctx, twiddle_rows_kernel, dim3((uint32_t)(want < cap ? want : cap)), dim3(256), 0, as_fr(a),
                row0, nrows, ncols, (const Fr*)tw->d_lo, (const Fr*)tw->d_hi, tw->h);
}

extern "C" int h2b_fr_transpose_scatter(h2b_ctx* ctx, const h2b_fr* in, void* const* peer_out,
                                        uint32_t world, uint32_t rank, uint32_t rows_local,
                                        uint32_t cols) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!in || !peer_out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (world == 0 || world > 16 || rank >= world || cols % world)
    return fail(ctx, H2B_ERR_ARG, "bad world / rank / column count");
  if (rows_local == 0 || cols == 0) return H2B_OK;
  PeerPtrs peers;
  for (uint32_t i = 0; i < 16; ++i) peers.p[i] = i < world ? reinterpret_cast<Fr*>(peer_out[i]) : nullptr;
  for (uint32_t i = 0; i < world; ++i)
    if (!peers.p[i] || peers.p[i] == as_fr(in)) return fail(ctx, H2B_ERR_ARG, "null or aliased peer buffer");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  return launch(ctx, transpose_scatter_kernel, dim3((cols + 31) / 32, (rows_local + 31) / 32), dim3(256), 0,
                as_fr(in), peers, rank, rows_local, cols, cols / world, (uint64_t)rows_local * world);
}

// Row transforms of the four-step NTT fused with the distributed transpose that follows them: `nrows` best_fft's of
// 2^log_n points each (rows_dev: nrows x 2^log_n, device, left untouched), output Ko of row r stored -- times
// big_omega^((row0 + r) * Ko) if big_omega != nullptr -- into peer_out[Ko / cl][(Ko % cl) * total_rows + row0 + r],
// cl = 2^log_n / world.  The peers' buffers are NVLink-mapped (symmetric memory): no separate twiddle pass, no
// transpose kernel, no collective.
extern "C" int h2b_best_fft_rows_scatter(h2b_ctx* ctx, const h2b_fr* rows_dev, const h2b_fr* omega, uint32_t log_n,
                                         uint32_t nrows, void* const* peer_out, uint32_t world, uint64_t row0,
                                         uint64_t total_rows, const h2b_fr* big_omega, uint32_t big_log_n) {
  if (!ctx) return H2B_ERR_ARG;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (!rows_dev || !omega || !peer_out) return fail(ctx, H2B_ERR_ARG, "null pointer");
  if (log_n > 28 || big_log_n > 28) return fail(ctx, H2B_ERR_ARG, "log_n > 28");
  const uint64_t n = 1ull << log_n;
  if (world == 0 || world > 16 || (world & (world - 1)) || n % world)
    return fail(ctx, H2B_ERR_ARG, "bad world size for this row length");
  if (nrows == 0) return H2B_OK;
  if (row0 + nrows > total_rows) return fail(ctx, H2B_ERR_LENGTH, "rows outside the matrix");
  if (big_omega && (row0 + nrows - 1) * (n - 1) >= (1ull << big_log_n))
    return fail(ctx, H2B_ERR_LENGTH, "twiddle exponent exceeds 2^big_log_n");
  H2B_CUDA(ctx, cudaSetDevice(ctx->device));
  NttScatter sc;
  for (uint32_t i = 0; i < 16; ++i) sc.peers[i] = i < world ? reinterpret_cast<Fr*>(peer_out[i]) : nullptr;
  for (uint32_t i = 0; i < world; ++i)
    if (!sc.peers[i] || sc.peers[i] == as_fr(rows_dev)) return fail(ctx, H2B_ERR_ARG, "null or aliased peer buffer");
  sc.cl = (uint32_t)(n / world);
  sc.R = total_rows;
  sc.row0 = row0;
  sc.tw = nullptr;
  if (big_omega) H2B_TRY(ntt_get_table(ctx, *as_fr(big_omega), big_log_n, &sc.tw));
  const TwTable* tw;
  H2B_TRY(ntt_get_table(ctx, *as_fr(omega), log_n, &tw));
  return ntt_run(ctx, as_fr(rows_dev), nullptr, log_n, tw, n, nullptr, 1, nullptr, 1, n, nrows, n, n, &sc);
}
