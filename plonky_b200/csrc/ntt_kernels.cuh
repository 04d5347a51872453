// Radix-2 NTT over the Tweedle / BLS12-377 prime fields for sm_100a.
//
// Replaces src/fft.rs of the reference: fft_precompute (:47-59), fft_with_precomputation (:61-80),
// fft_with_precomputation_power_of_2 (:103-156), ifft_with_precomputation_power_of_2 (:82-101),
// plus the coset / zero-pad callers in src/polynomial.rs:135-151,330-380 and
// src/plonk_util.rs:169-190.  Same contract: natural order in, natural order out,
// out[k] = sum_j c_j w^(jk) with w = primitive_root_of_unity(log2 n) (src/field/field.rs:429-435).
//
// Design (B200-first, nothing like the reference's log n passes over a 2n-entry table):
//   n = R_1 * R_2 * ... * R_m, R_i = 2^r_i <= 256, m = ceil(log n / 8) passes (3 for 2^24).
//   Input index  j = j_1 + R_1 j_2 + R_1 R_2 j_3 + ...      (j_1 least significant)
//   Output index k = k_m + R_m k_{m-1} + ...                (k_1 most significant)
//   Pass 1 transforms over j_m (input stride n/R_m) and writes the "work layout" in which the
//   remaining digits are stored most-significant-first (a mixed-radix digit reversal), so that
//   every later pass d = m-1 .. 1 is an IN-PLACE strided transform over j_d preceded by the twiddle
//   w_{N_d}^(j_d k''), N_d = R_d M_d, M_d = R_{d+1}..R_m, k'' = the already transformed low part.
//   After the last pass the data is in natural order: no separate bit-reversal pass ever runs.
//   Each CTA owns a tile of 2^r rows x 8 columns (8 x 32 B = 256 B contiguous per row: full-sector
//   128-bit loads), staged in shared memory as split 16-byte pieces with an XOR swizzle so that
//   row-wise, column-wise and butterfly accesses are all bank-conflict free.  Butterflies run as
//   radix-8 / 4 / 2 register rounds (3 / 2 / 1 layers per shared-memory round trip).
//   Twiddles: sub-transform twiddles from a 128-entry table; inter-pass twiddles w_n^e either on the fly from two small
//   tables (w^lo, w^(hi << lo_bits)) and one multiply, or -- for passes with N_d <= 2^24, built lazily per direction
//   (ensure_direct, ntt_host.cuh) -- from a full table: one load and one product less per element, N_d * 32 bytes more.
//   Fused: zero-padding (rows beyond n_in are never read), coset shift c_j g^j on load, n^-1 folded
//   into the last pass' twiddle table for the inverse, g^-i / pointwise factors on the final store.
//
// Cost model for the roofline: algorithmic bytes = 2 * n * 32; this implementation moves
// m * 2 * n * 32 bytes; arithmetic ~ (n/2) log n + 2 n (m-1) Montgomery products.
#pragma once
#include "ntt_plan.h"
#include "fp.cuh"

namespace plk {

struct NttPassParams {
  const void* in;
  void* out;
  int log_n;
  int r;            // this pass' digit
  int log_m;        // log2 M_d (in-place passes)
  int log_t;        // log2 columns per tile
  int first;        // 1: gathering pass, 0: in-place pass
  int last;         // 1: this pass produces the final natural-order output
  int ndig;         // number of low digits (m-1) for the digit reversal of the first pass
  int digs[kMaxDigits];
  unsigned long long n_in;        // first pass: elements actually present in the input row
  unsigned long long in_stride;   // elements between batch rows
  unsigned long long out_stride;
  const void* wsub;               // w_256^k, k < 128
  const void* tw_lo;              // w_n^e, e < 2^lo_bits
  const void* tw_hi;              // w_n^(e << lo_bits)
  int lo_bits;
  int tw_all;                     // also multiply row 0 (n^-1 folded into tw_hi)
  const void* tw_direct;          // optional full table w_{N_d}^(j_d k'') at [(j_d << log_m) + k''] (small N_d): one load, one product
  const void* pre_lo;             // optional input multiplier s^j (two-level, same lo_bits)
  const void* pre_hi;
  const void* post_lo;            // optional output multiplier s^k (two-level)
  const void* post_hi;
  const void* post_periodic;      // optional output multiplier tbl[k & post_mask]
  unsigned long long post_mask;
  const void* scale;              // optional constant multiplier (single-pass inverse)
  int post_lo_bits;               // lo_bits of the post tables
  int tw_none;                    // in-place pass without inter-pass twiddles (phase B of the distributed transform)
  int post_rowmul;                // post factor exponent = (post_row_base + batch row) * k instead of k
  unsigned long long post_row_base;
  int remap;                      // final store goes to the all-to-all send layout [dest][row][k mod 2^remap_cl_log]
  int remap_cl_log;
  unsigned long long remap_rows;
  // remap == 2: the final store goes STRAIGHT into the receive buffers of the peer GPUs (CUDA IPC mappings over NVLink):
  // value k of local row `row` lands in peer (k >> remap_cl_log) at [remap_row_base + row][k mod 2^remap_cl_log]
  void* peer[kMaxPeers];
  unsigned long long remap_row_base;
};

// Output position of final value k of batch row `row`: natural (k) or packed for the all-to-all of the
// domain-split transform: destination rank (k >> cl) major, then row, then the column inside the block.
__device__ __forceinline__ unsigned long long ntt_out_index(const NttPassParams& p, unsigned long long k, unsigned row,
                                                           unsigned long long out_stride) {
  if (!p.remap) return (unsigned long long)row * out_stride + k;
  const unsigned long long dest = k >> p.remap_cl_log, kl = k & ((1ull << p.remap_cl_log) - 1);
  return ((dest * p.remap_rows + row) << p.remap_cl_log) + kl;
}

// piece 0 of final value k of batch row `row`: local buffer, or a peer's receive buffer (remap == 2)
__device__ __forceinline__ uint4* ntt_out_ptr(const NttPassParams& p, uint4* out, unsigned long long k, unsigned row, int pieces) {
  if (p.remap == 2) {
    const unsigned long long dest = k >> p.remap_cl_log, kl = k & ((1ull << p.remap_cl_log) - 1);
    return reinterpret_cast<uint4*>(p.peer[dest]) + ((((p.remap_row_base + row) << p.remap_cl_log) + kl) * pieces);
  }
  return out + ntt_out_index(p, k, row, p.out_stride) * pieces;
}

__device__ __forceinline__ unsigned bitrev(unsigned x, int bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

template <class F>
__device__ __forceinline__ F lds_fp(const uint4* smem, int elems, int sidx) {
  F r;
#pragma unroll
  for (int pc = 0; pc < F::N / 4; ++pc) {
    uint4 v = smem[pc * elems + sidx];
    r.l[4 * pc] = v.x; r.l[4 * pc + 1] = v.y; r.l[4 * pc + 2] = v.z; r.l[4 * pc + 3] = v.w;
  }
  return r;
}
template <class F>
__device__ __forceinline__ void sts_fp(uint4* smem, int elems, int sidx, const F& a) {
#pragma unroll
  for (int pc = 0; pc < F::N / 4; ++pc)
    smem[pc * elems + sidx] = make_uint4(a.l[4 * pc], a.l[4 * pc + 1], a.l[4 * pc + 2], a.l[4 * pc + 3]);
}
template <class F>
__device__ __forceinline__ F two_level(const void* lo, const void* hi, int lo_bits, unsigned long long e) {
  F a = load_fp<F>(lo, (size_t)(e & ((1ull << lo_bits) - 1)));
  unsigned long long h = e >> lo_bits;
  if (h == 0) return a;                      // hi[0] == 1 for every non-folded table
  return F::mul(a, load_fp<F>(hi, (size_t)h));
}
template <class F>
__device__ __forceinline__ F two_level_always(const void* lo, const void* hi, int lo_bits, unsigned long long e) {
  F a = load_fp<F>(lo, (size_t)(e & ((1ull << lo_bits) - 1)));
  return F::mul(a, load_fp<F>(hi, (size_t)(e >> lo_bits)));
}

// Q layers of decimation-in-time butterflies on 2^Q register-resident elements whose rows are
// base + (e << l0); `low` = base mod 2^l0 selects the twiddles.  LAYER0: l0 == 0 (twiddle 1 skipped).
template <class F, int Q, bool LAYER0>
__device__ __forceinline__ void dit_layers(F (&x)[1 << Q], int l0, int low, const uint4* wsub) {
#pragma unroll
  for (int t = 1; t <= Q; ++t) {
    const int half = 1 << (t - 1);
#pragma unroll
    for (int kk = 0; kk < half; ++kk) {
      const bool trivial = LAYER0 && kk == 0;
      F w;
      if (!trivial) w = load_fp<F>(wsub, (size_t)((low + (kk << l0)) << (kSubLog - (l0 + t))));   // shared-memory copy
#pragma unroll
      for (int blk = 0; blk < (1 << Q); blk += 2 * half) {
        F v = trivial ? x[blk + kk + half] : F::mul(x[blk + kk + half], w);
        F u = x[blk + kk];
        x[blk + kk] = F::add(u, v);
        x[blk + kk + half] = F::sub(u, v);
      }
    }
  }
}

// tile geometry shared by all phases of one CTA
struct TileGeom {
  unsigned long long gbase;   // global element index of (row 0, col 0)
  int rowshift;               // global row stride = 1 << rowshift
  unsigned long long k0;      // in-place: k'' of column 0 ; first: jlow of column 0
};

// GEO = r > 0: the tile geometry is a compile-time constant (2^GEO rows, 2^kGeoLogT columns, kNttThreads threads; GEO = 6, 7, 8
// are instantiated: every pass of transforms of 2^16 points and more) -- all index arithmetic of the load / store loops
// and the rounds folds into constants and per-thread bases (the generic loops spend ~20 % of the kernel's instructions
// on 64-bit shifts, bit reversals and swizzles per 16-byte piece).  GEO = 0: geometry read from the parameters.
constexpr int kGeoLogT = 3, kGeoRMin = 6, kGeoRMax = 8;

template <class F, int Q, bool LAYER0, bool FUSE_TW = false, int GEO = 0>
__device__ __forceinline__ void ntt_round(const NttPassParams& p, uint4* smem, const uint4* wsub, int l0, unsigned long long k0 = 0) {
  const int log_t = GEO ? kGeoLogT : p.log_t, pr = GEO ? GEO : p.r;
  const int T = 1 << log_t, elems = 1 << (pr + log_t);
  const int groups = elems >> Q;
  const int nthreads = GEO ? kNttThreads : (int)blockDim.x;
#pragma unroll 1
  for (int gi = threadIdx.x; gi < groups; gi += nthreads) {
    const int col = gi & (T - 1);
    const int gr = gi >> log_t;
    const int low = gr & ((1 << l0) - 1);
    const int base_row = low | ((gr >> l0) << (l0 + Q));
    F x[1 << Q];
#pragma unroll
    for (int e = 0; e < (1 << Q); ++e) {
      const int row = base_row + (e << l0);
      x[e] = lds_fp<F>(smem, elems, row * T + (col ^ (row & (T - 1))));
    }
    if (FUSE_TW) {
      // inter-pass twiddle from the full table, folded into the first register round (l0 == 0; rows sit at
      // their bit-reversed position): no separate shared-memory pass, no extra barrier
      const size_t kk = (size_t)(k0 + col);
#pragma unroll
      for (int e = 0; e < (1 << Q); ++e) {
        const size_t orow = bitrev((unsigned)(base_row + e), pr);
        // GEO: multiply unconditionally -- the table holds ONE where the exponent is zero (a product by one returns the same
        // canonical value), and the skip is a divergent branch around every product for one row in 256
        if (GEO || p.tw_all || (orow != 0 && kk != 0)) x[e] = F::mul(x[e], load_fp<F>(p.tw_direct, (orow << p.log_m) + kk));
      }
    }
    dit_layers<F, Q, LAYER0>(x, l0, low, wsub);
#pragma unroll
    for (int e = 0; e < (1 << Q); ++e) {
      const int row = base_row + (e << l0);
      sts_fp<F>(smem, elems, row * T + (col ^ (row & (T - 1))), x[e]);
    }
  }
  __syncthreads();
}


// Input-side factors, applied once per element in shared memory before the butterflies (rows sit at
// their bit-reversed position): coset shift s^j (first pass), inter-pass twiddle w_{N_d}^(j_d k'')
// (in-place passes; n^-1 folded into the table on the last inverse pass), constant scale.
template <class F>
__device__ __noinline__ void ntt_pre_factors(const NttPassParams& p, const TileGeom& g, uint4* smem) {
  const int T = 1 << p.log_t, elems = 1 << (p.r + p.log_t);
#pragma unroll 2
  for (int idx = threadIdx.x; idx < elems; idx += blockDim.x) {
    const int col = idx & (T - 1);
    const int srow = idx >> p.log_t;
    const unsigned orow = bitrev((unsigned)srow, p.r);
    const int sidx = srow * T + (col ^ (srow & (T - 1)));
    F fac;
    bool have = false;
    if (p.first) {
      const unsigned long long j = g.gbase + col + ((unsigned long long)orow << g.rowshift);
      if (p.pre_lo && j < p.n_in && j != 0) { fac = two_level<F>(p.pre_lo, p.pre_hi, p.lo_bits, j); have = true; }
    } else {
      const unsigned long long kk = g.k0 + col;
      const unsigned long long ex = ((unsigned long long)orow * kk) << (p.log_n - p.r - p.log_m);
      if (p.tw_none) { }
      else if (p.tw_direct) {
        if (p.tw_all || ex != 0) { fac = load_fp<F>(p.tw_direct, ((size_t)orow << p.log_m) + (size_t)kk); have = true; }
      }
      else if (p.tw_all) { fac = two_level_always<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex); have = true; }
      else if (ex != 0) { fac = two_level<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex); have = true; }
    }
    if (p.scale) {
      F s = load_fp<F>(p.scale, 0);
      fac = have ? F::mul(fac, s) : s;
      have = true;
    }
    if (have) sts_fp<F>(smem, elems, sidx, F::mul(lds_fp<F>(smem, elems, sidx), fac));
  }
  __syncthreads();
}
// Output-side factors on the final natural-order values: s^-k (inverse coset) and / or a periodic table.
template <class F>
__device__ __noinline__ void ntt_post_factors(const NttPassParams& p, const TileGeom& g, uint4* smem) {
  const int T = 1 << p.log_t, elems = 1 << (p.r + p.log_t);
  for (int idx = threadIdx.x; idx < elems; idx += blockDim.x) {
    const int col = idx & (T - 1);
    const unsigned long long row = idx >> p.log_t;
    const int sidx = (int)row * T + (col ^ ((int)row & (T - 1)));
    const unsigned long long k = p.first ? row : (g.k0 + col + (row << p.log_m));
    F x = lds_fp<F>(smem, elems, sidx);
    const unsigned long long pe = p.post_rowmul ? (p.post_row_base + blockIdx.y) * k : k;
    if (p.post_lo && pe != 0) x = F::mul(x, two_level<F>(p.post_lo, p.post_hi, p.post_lo_bits, pe));
    if (p.post_periodic) x = F::mul(x, load_fp<F>(p.post_periodic, (size_t)(k & p.post_mask)));
    sts_fp<F>(smem, elems, sidx, x);
  }
  __syncthreads();
}

// MAXQ = 3: radix-8 register rounds, 2 CTAs per SM (128 registers).  MAXQ = 2: radix-4 rounds, 3 CTAs per SM
// (85 registers): more shared-memory round trips, more warps to hide them behind.
template <class F, int MAXQ, int GEO = 0>
__global__ void __launch_bounds__(kNttThreads, MAXQ == 3 ? 2 : 3) ntt_pass_kernel(NttPassParams p) {
  extern __shared__ uint4 smem[];
  constexpr int PIECES = F::N / 4;
  const int log_t = GEO ? kGeoLogT : p.log_t, pr = GEO ? GEO : p.r;
  const int nthreads = GEO ? kNttThreads : (int)blockDim.x;
  const int T = 1 << log_t, R = 1 << pr, elems = R * T;
  constexpr int kPieceIters = GEO ? ((1 << (GEO + kGeoLogT)) * PIECES) / kNttThreads : 1;   // trip count of the piece loops
  const unsigned long long tile = blockIdx.x;
  const uint4* in = reinterpret_cast<const uint4*>(p.in) + (size_t)blockIdx.y * p.in_stride * PIECES;
  uint4* out = reinterpret_cast<uint4*>(p.out);   // batch row offset is applied by ntt_out_index

  TileGeom g;
  if (p.first) {
    g.k0 = tile << log_t;               // jlow of column 0
    g.gbase = g.k0;
    g.rowshift = p.log_n - pr;          // input stride of j_m
  } else {
    const int tph = p.log_m - log_t;    // log2 tiles per hi block
    const unsigned long long hi = tile >> tph;
    g.k0 = (tile & ((1ull << tph) - 1)) << log_t;
    g.gbase = g.k0 + (hi << (p.log_m + pr));
    g.rowshift = p.log_m;
  }

  // the 128 sub-transform twiddles live in shared memory behind the tile (fixed 29-cycle LDS instead of
  // L1/L2 round trips on the butterflies' critical path)
  uint4* wsub_s = smem + (size_t)PIECES * elems;
  for (int idx = threadIdx.x; idx < (1 << (kSubLog - 1)) * PIECES; idx += nthreads)
    wsub_s[idx] = reinterpret_cast<const uint4*>(p.wsub)[idx];

  // ---- load: 16-byte pieces, columns fastest (256 B contiguous per row), global -> shared directly with
  // cp.async (LDGSTS): all of a thread's 16 pieces are in flight at once and no register staging is
  // needed; rows at or beyond n_in (zero padding of the LDE) are zero-filled without touching memory.
#pragma unroll(kPieceIters)
  for (int idx = threadIdx.x; idx < elems * PIECES; idx += nthreads) {
    const int piece = idx % PIECES;
    const int e = idx / PIECES;
    const int col = e & (T - 1);
    const int row = e >> log_t;
    const unsigned long long gidx = g.gbase + col + ((unsigned long long)row << g.rowshift);
    const int srow = (int)bitrev((unsigned)row, pr);
    uint4* dst = &smem[piece * elems + srow * T + (col ^ (srow & (T - 1)))];
    const bool present = !p.first || gidx < p.n_in;
    const uint4* src = in + (present ? gidx * PIECES + piece : 0);
    const unsigned nbytes = present ? 16u : 0u;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(src), "r"(nbytes) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- input-side factors, then butterflies as radix-8 / 4 / 2 register rounds ----
  const bool fuse_tw = MAXQ == 2 && !p.first && !p.tw_none && !p.scale && p.tw_direct && pr >= 2;
  if (!fuse_tw && ((p.first && p.pre_lo) || (!p.first && !p.tw_none) || p.scale)) ntt_pre_factors<F>(p, g, smem);
  {
    int l0 = 0;
    const int r = pr;
    if (GEO) {
      // radix-4 rounds, one radix-2 round at the end when GEO is odd (the host selects GEO > 0 for MAXQ == 2 only)
      if (fuse_tw) ntt_round<F, 2, true, true, GEO>(p, smem, wsub_s, 0, g.k0);
      else ntt_round<F, 2, true, false, GEO>(p, smem, wsub_s, 0);
#pragma unroll 1
      for (l0 = 2; l0 + 2 <= GEO; l0 += 2) ntt_round<F, 2, false, false, GEO>(p, smem, wsub_s, l0);
      if (GEO & 1) ntt_round<F, 1, false, false, GEO>(p, smem, wsub_s, GEO - 1);
    } else {
    if (MAXQ >= 3 && r >= 3) { ntt_round<F, (MAXQ >= 3 ? 3 : 2), true>(p, smem, wsub_s, 0); l0 = 3; }
    else if (r >= 2) {
      if (fuse_tw) ntt_round<F, 2, true, true>(p, smem, wsub_s, 0, g.k0);
      else ntt_round<F, 2, true>(p, smem, wsub_s, 0);
      l0 = 2;
    }
    else if (r == 1) { ntt_round<F, 1, true>(p, smem, wsub_s, 0); l0 = 1; }
    while (l0 < r) {
      const int q = (r - l0 >= MAXQ) ? MAXQ : (r - l0);
      if (MAXQ >= 3 && q == 3) ntt_round<F, (MAXQ >= 3 ? 3 : 2), false>(p, smem, wsub_s, l0);
      else if (q == 2) ntt_round<F, 2, false>(p, smem, wsub_s, l0);
      else ntt_round<F, 1, false>(p, smem, wsub_s, l0);
      l0 += q;
    }
    }
  }
  if (p.last && (p.post_lo || p.post_periodic)) ntt_post_factors<F>(p, g, smem);

  // ---- store ----
  if (!p.first) {
#pragma unroll(kPieceIters)
    for (int idx = threadIdx.x; idx < elems * PIECES; idx += nthreads) {
      const int piece = idx % PIECES;
      const int e = idx / PIECES;
      const int col = e & (T - 1);
      const int row = e >> log_t;
      const unsigned long long gidx = g.gbase + col + ((unsigned long long)row << g.rowshift);
      uint4* dst = p.last ? ntt_out_ptr(p, out, gidx, blockIdx.y, PIECES)
                          : out + ((unsigned long long)blockIdx.y * p.out_stride + gidx) * PIECES;
      dst[piece] = smem[piece * elems + row * T + (col ^ (row & (T - 1)))];
    }
  } else {
    // column c of the tile becomes a run of R contiguous outputs at rev_digits(jlow) * R
#pragma unroll(kPieceIters < 4 ? kPieceIters : 4)
    for (int idx = threadIdx.x; idx < elems * PIECES; idx += nthreads) {
      const int piece = idx % PIECES;
      const int e = idx / PIECES;
      const int row = e & (R - 1);
      const int col = e >> pr;
      unsigned long long x = g.k0 + col, pos = 0;
      for (int i = 0; i < p.ndig; ++i) {
        pos = (pos << p.digs[i]) | (x & ((1ull << p.digs[i]) - 1));
        x >>= p.digs[i];
      }
      const unsigned long long gidx = (pos << pr) + row;
      uint4* dst = p.last ? ntt_out_ptr(p, out, gidx, blockIdx.y, PIECES)
                          : out + ((unsigned long long)blockIdx.y * p.out_stride + gidx) * PIECES;
      dst[piece] = smem[piece * elems + row * T + (col ^ (row & (T - 1)))];
    }
  }
}

// out[i] = base^(i * stride) for i < count (64-bit exponents), optionally times `scale`
template <class F>
__global__ void pow_table_kernel(F base, unsigned long long stride_log2, unsigned long long count, const F* scale, F* out) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= count) return;
  // base^(i << stride_log2): square the base stride_log2 times first (uniform), then binary powering
  F b = base;
  for (unsigned long long s = 0; s < stride_log2; ++s) b = F::sqr(b);
  F acc = F::one();
  unsigned long long e = i;
  while (e) {
    if (e & 1) acc = F::mul(acc, b);
    b = F::sqr(b);
    e >>= 1;
  }
  if (scale) acc = F::mul(acc, *scale);
  out[i] = acc;
}

// full inter-pass twiddle table of one in-place pass: out[(j << log_m) + k] = w_n^((j k) << sh) (times the
// folded scale when hi is the scaled table)
template <class F>
__global__ void direct_twiddle_kernel(const void* lo, const void* hi, int lo_bits, int log_m, int r, int sh, int always, F* out) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= (1ull << (log_m + r))) return;
  const unsigned long long j = i >> log_m, kk = i & ((1ull << log_m) - 1);
  const unsigned long long ex = (j * kk) << sh;
  out[i] = always ? two_level_always<F>(lo, hi, lo_bits, ex) : two_level<F>(lo, hi, lo_bits, ex);
}

// denominators of divide_by_z_h (src/polynomial.rs:351-361): tbl[i] = 1 / (g^n * w^(n i) - 1), i < period
template <class F>
__global__ void zh_inverse_table_kernel(F gn, F wn, unsigned count, F* out, int* zero_flag) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  F acc = F::one(), b = wn;
  unsigned e = i;
  while (e) {
    if (e & 1) acc = F::mul(acc, b);
    b = F::sqr(b);
    e >>= 1;
  }
  F d = F::sub(F::mul(gn, acc), F::one());
  if (d.is_zero()) { *zero_flag = 1; out[i] = d; return; }
  out[i] = F::inverse(d);
}

}  // namespace plk
