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
//   Twiddles: sub-transform twiddles from a 128-entry table; inter-pass twiddles w_n^e on the fly
//   from two small tables (w^lo, w^(hi << lo_bits)) and one multiply -- never an n-entry table.
//   Fused: zero-padding (rows beyond n_in are never read), coset shift c_j g^j on load, n^-1 folded
//   into the last pass' twiddle table for the inverse, g^-i / pointwise factors on the final store.
//
// Cost model for the roofline: algorithmic bytes = 2 * n * 32; this implementation moves
// m * 2 * n * 32 bytes; arithmetic ~ (n/2) log n + 2 n (m-1) Montgomery products.
#include <mutex>
#include <map>
#include "common.cuh"
#include "fp.cuh"

namespace plk {

constexpr int kSubLog = 8;          // largest sub-transform: 2^8 rows
constexpr int kMaxDigits = 6;       // 6 * 8 = 48 >= largest TWO_ADICITY (47)
constexpr int kTileColsLog = 3;     // 8 columns per tile
constexpr int kNttThreads = 256;

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
  const void* pre_lo;             // optional input multiplier s^j (two-level, same lo_bits)
  const void* pre_hi;
  const void* post_lo;            // optional output multiplier s^k (two-level)
  const void* post_hi;
  const void* post_periodic;      // optional output multiplier tbl[k & post_mask]
  unsigned long long post_mask;
  const void* scale;              // optional constant multiplier (single-pass inverse)
};

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
__device__ __forceinline__ void dit_layers(F (&x)[1 << Q], int l0, int low, const void* wsub) {
#pragma unroll
  for (int t = 1; t <= Q; ++t) {
    const int half = 1 << (t - 1);
#pragma unroll
    for (int kk = 0; kk < half; ++kk) {
      const bool trivial = LAYER0 && kk == 0;
      F w;
      if (!trivial) w = load_fp<F>(wsub, (size_t)((low + (kk << l0)) << (kSubLog - (l0 + t))));
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

template <class F, int Q, bool LAYER0>
__device__ __forceinline__ void ntt_round(const NttPassParams& p, const TileGeom& g, uint4* smem, int l0, bool last_round) {
  const int T = 1 << p.log_t, elems = 1 << (p.r + p.log_t);
  const int groups = elems >> Q;
  for (int gi = threadIdx.x; gi < groups; gi += blockDim.x) {
    const int col = gi & (T - 1);
    const int gr = gi >> p.log_t;
    const int low = gr & ((1 << l0) - 1);
    const int base_row = low | ((gr >> l0) << (l0 + Q));
    F x[1 << Q];
#pragma unroll
    for (int e = 0; e < (1 << Q); ++e) {
      const int row = base_row + (e << l0);
      x[e] = lds_fp<F>(smem, elems, row * T + (col ^ (row & (T - 1))));
    }
    if (LAYER0) {
      // rows sit at their bit-reversed position; undo to find the source index / twiddle
#pragma unroll
      for (int e = 0; e < (1 << Q); ++e) {
        const unsigned orow = bitrev((unsigned)(base_row + e), p.r);
        if (p.first) {
          const unsigned long long j = g.gbase + col + ((unsigned long long)orow << g.rowshift);
          if (p.pre_lo && j < p.n_in) x[e] = F::mul(x[e], two_level<F>(p.pre_lo, p.pre_hi, p.lo_bits, j));
        } else {
          const unsigned long long kk = g.k0 + col;
          const unsigned long long ex = ((unsigned long long)orow * kk) << (p.log_n - p.r - p.log_m);
          if (p.tw_all) x[e] = F::mul(x[e], two_level_always<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex));
          else if (orow != 0 && kk != 0) x[e] = F::mul(x[e], two_level<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex));
        }
        if (p.scale) x[e] = F::mul(x[e], load_fp<F>(p.scale, 0));
      }
    }
    dit_layers<F, Q, LAYER0>(x, l0, low, p.wsub);
    if (last_round && p.last && (p.post_lo || p.post_periodic)) {
#pragma unroll
      for (int e = 0; e < (1 << Q); ++e) {
        const unsigned long long row = base_row + (e << l0);
        // final output index of this element (natural order)
        const unsigned long long k = p.first ? row : (g.k0 + col + (row << p.log_m));
        if (p.post_lo) x[e] = F::mul(x[e], two_level<F>(p.post_lo, p.post_hi, p.lo_bits, k));
        if (p.post_periodic) x[e] = F::mul(x[e], load_fp<F>(p.post_periodic, (size_t)(k & p.post_mask)));
      }
    }
#pragma unroll
    for (int e = 0; e < (1 << Q); ++e) {
      const int row = base_row + (e << l0);
      sts_fp<F>(smem, elems, row * T + (col ^ (row & (T - 1))), x[e]);
    }
  }
  __syncthreads();
}

template <class F>
__global__ void __launch_bounds__(kNttThreads, 2) ntt_pass_kernel(NttPassParams p) {
  extern __shared__ uint4 smem[];
  constexpr int PIECES = F::N / 4;
  const int T = 1 << p.log_t, R = 1 << p.r, elems = R * T;
  const unsigned long long tile = blockIdx.x;
  const uint4* in = reinterpret_cast<const uint4*>(p.in) + (size_t)blockIdx.y * p.in_stride * PIECES;
  uint4* out = reinterpret_cast<uint4*>(p.out) + (size_t)blockIdx.y * p.out_stride * PIECES;

  TileGeom g;
  if (p.first) {
    g.k0 = tile << p.log_t;             // jlow of column 0
    g.gbase = g.k0;
    g.rowshift = p.log_n - p.r;         // input stride of j_m
  } else {
    const int tph = p.log_m - p.log_t;  // log2 tiles per hi block
    const unsigned long long hi = tile >> tph;
    g.k0 = (tile & ((1ull << tph) - 1)) << p.log_t;
    g.gbase = g.k0 + (hi << (p.log_m + p.r));
    g.rowshift = p.log_m;
  }

  // ---- load: 16-byte pieces, columns fastest (256 B contiguous per row) ----
  for (int idx = threadIdx.x; idx < elems * PIECES; idx += blockDim.x) {
    const int piece = idx % PIECES;
    const int e = idx / PIECES;
    const int col = e & (T - 1);
    const int row = e >> p.log_t;
    const unsigned long long gidx = g.gbase + col + ((unsigned long long)row << g.rowshift);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!p.first || gidx < p.n_in) v = in[gidx * PIECES + piece];
    const int srow = (int)bitrev((unsigned)row, p.r);
    smem[piece * elems + srow * T + (col ^ (srow & (T - 1)))] = v;
  }
  __syncthreads();

  // ---- butterflies: radix-8 / 4 / 2 register rounds ----
  {
    int l0 = 0;
    const int r = p.r;
    if (r >= 3) { ntt_round<F, 3, true>(p, g, smem, 0, r == 3); l0 = 3; }
    else if (r == 2) { ntt_round<F, 2, true>(p, g, smem, 0, true); l0 = 2; }
    else if (r == 1) { ntt_round<F, 1, true>(p, g, smem, 0, true); l0 = 1; }
    else {
      // r == 0: a 1-point transform still has to apply the pre/post factors
      ntt_round<F, 0, true>(p, g, smem, 0, true);
    }
    while (l0 < r) {
      const int q = (r - l0 >= 3) ? 3 : (r - l0);
      const bool lr = (l0 + q == r);
      if (q == 3) ntt_round<F, 3, false>(p, g, smem, l0, lr);
      else if (q == 2) ntt_round<F, 2, false>(p, g, smem, l0, lr);
      else ntt_round<F, 1, false>(p, g, smem, l0, lr);
      l0 += q;
    }
  }

  // ---- store ----
  if (!p.first) {
    for (int idx = threadIdx.x; idx < elems * PIECES; idx += blockDim.x) {
      const int piece = idx % PIECES;
      const int e = idx / PIECES;
      const int col = e & (T - 1);
      const int row = e >> p.log_t;
      const unsigned long long gidx = g.gbase + col + ((unsigned long long)row << g.rowshift);
      out[gidx * PIECES + piece] = smem[piece * elems + row * T + (col ^ (row & (T - 1)))];
    }
  } else {
    // column c of the tile becomes a run of R contiguous outputs at rev_digits(jlow) * R
    for (int idx = threadIdx.x; idx < elems * PIECES; idx += blockDim.x) {
      const int piece = idx % PIECES;
      const int e = idx / PIECES;
      const int row = e & (R - 1);
      const int col = e >> p.r;
      unsigned long long x = g.k0 + col, pos = 0;
      for (int i = 0; i < p.ndig; ++i) {
        pos = (pos << p.digs[i]) | (x & ((1ull << p.digs[i]) - 1));
        x >>= p.digs[i];
      }
      const unsigned long long gidx = (pos << p.r) + row;
      out[gidx * PIECES + piece] = smem[piece * elems + row * T + (col ^ (row & (T - 1)))];
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

using namespace plk;

// defined in api.cu: out[i] = in[i]^-1 elementwise on device
void plk_launch_field_inverse(int field, const void* d_in, void* d_out, size_t n, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct CosetTables {
  DevBuf fwd_lo, fwd_hi;    // s^j
  DevBuf inv_lo, inv_hi;    // s^-j
};

struct plk_fft_plan {
  int field = 0;
  int log_n = 0;
  size_t n = 0;
  int device = 0;
  int m = 0;                 // passes
  int dig[kMaxDigits];       // r_1 .. r_m
  int lo_bits = 0;
  size_t elem_bytes = 32;
  DevBuf wsub[2];            // [0] forward, [1] inverse
  DevBuf tw_lo[2], tw_hi[2];
  DevBuf tw_hi_inv_scaled;   // inverse hi table with n^-1 folded in
  DevBuf n_inv;              // one element: n^-1
  std::mutex mu;
  std::map<std::vector<uint32_t>, CosetTables*> cosets;   // keyed by the shift's limbs
  std::map<size_t, DevBuf*> zh_tables;                     // keyed by n_gates
  // scratch for the host-pointer entry points
  DevBuf h_in, h_out;
  ~plk_fft_plan() {
    for (auto& kv : cosets) delete kv.second;
    for (auto& kv : zh_tables) delete kv.second;
  }
};

namespace {

template <class F>
F fp_from_host(const uint32_t* limbs) {
  F r;
  for (int i = 0; i < F::N; ++i) r.l[i] = limbs[i];
  return r;
}

template <class P>
void build_pow_table(const uint32_t* base_limbs, int stride_log2, size_t count, const void* d_scale, DevBuf& buf, cudaStream_t st) {
  typedef Fp<P> F;
  buf.alloc(count * sizeof(F));
  F base = fp_from_host<F>(base_limbs);
  unsigned blocks = (unsigned)((count + 127) / 128);
  pow_table_kernel<F><<<blocks, 128, 0, st>>>(base, (unsigned long long)stride_log2, (unsigned long long)count,
                                               reinterpret_cast<const F*>(d_scale), buf.as<F>());
  PLK_LAUNCHED();
}

template <class P>
void plan_build(plk_fft_plan* pl) {
  typedef Fp<P> F;
  typedef FieldTables<P> Tb;
  cudaStream_t st = thread_stream();
  const int L = pl->log_n;
  pl->elem_bytes = sizeof(F);
  // digits: m = ceil(L / 8) (at least 1), as even as possible
  pl->m = L <= kSubLog ? 1 : (L + kSubLog - 1) / kSubLog;
  for (int i = 0; i < pl->m; ++i) pl->dig[i] = L / pl->m + (i < L % pl->m ? 1 : 0);
  pl->lo_bits = (L + 1) / 2;
  // n^-1
  pl->n_inv.alloc(sizeof(F));
  PLK_CUDA(cudaMemcpyAsync(pl->n_inv.p, Tb::pow2_inv(L), sizeof(F), cudaMemcpyHostToDevice, st));
  for (int inv = 0; inv < 2; ++inv) {
    const uint32_t* w256 = inv ? Tb::root_inv(kSubLog <= P::TWO_ADICITY ? kSubLog : P::TWO_ADICITY)
                               : Tb::root(kSubLog <= P::TWO_ADICITY ? kSubLog : P::TWO_ADICITY);
    build_pow_table<P>(w256, 0, (size_t)1 << (kSubLog - 1), nullptr, pl->wsub[inv], st);
    const uint32_t* wn = inv ? Tb::root_inv(L) : Tb::root(L);
    build_pow_table<P>(wn, 0, (size_t)1 << pl->lo_bits, nullptr, pl->tw_lo[inv], st);
    build_pow_table<P>(wn, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), nullptr, pl->tw_hi[inv], st);
    if (inv) build_pow_table<P>(wn, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), pl->n_inv.p, pl->tw_hi_inv_scaled, st);
  }
  PLK_CUDA(cudaStreamSynchronize(st));
}

struct FusedOps {
  const void* pre_lo = nullptr;
  const void* pre_hi = nullptr;
  const void* post_lo = nullptr;
  const void* post_hi = nullptr;
  const void* post_periodic = nullptr;
  unsigned long long post_mask = 0;
};

// One transform of k rows: d_in (n_in elements per row, stride in_stride) -> d_out (n per row).
// Uses d_out as the work buffer of the in-place passes.
template <class P>
void run_ntt(const plk_fft_plan* pl, const void* d_in, size_t n_in, size_t in_stride, void* d_out, size_t k,
             bool inverse, const FusedOps& ops, cudaStream_t st) {
  typedef Fp<P> F;
  const int L = pl->log_n;
  const int m = pl->m;
  const int inv = inverse ? 1 : 0;
  static bool attr_set[8] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t max_smem = (size_t)(F::N / 4) * 16 * ((size_t)1 << (kSubLog + kTileColsLog));
  if (!attr_set[dev & 7]) {
    PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    attr_set[dev & 7] = true;
  }
  int log_m_acc = 0;   // log2 of M_d for the pass being issued
  for (int pass = 0; pass < m; ++pass) {
    const int d = m - 1 - pass;          // digit index (0-based): pass 0 handles r_m
    NttPassParams p;
    memset(&p, 0, sizeof(p));
    p.log_n = L;
    p.r = pl->dig[d];
    p.first = pass == 0;
    p.last = pass == m - 1;
    p.in = p.first ? d_in : d_out;
    p.out = d_out;
    p.in_stride = p.first ? in_stride : pl->n;
    p.out_stride = pl->n;
    p.n_in = n_in;
    p.wsub = pl->wsub[inv].p;
    p.tw_lo = pl->tw_lo[inv].p;
    p.tw_hi = pl->tw_hi[inv].p;
    p.lo_bits = pl->lo_bits;
    if (p.first) {
      p.log_t = (L - p.r) < kTileColsLog ? (L - p.r) : kTileColsLog;
      p.ndig = m - 1;
      for (int i = 0; i < m - 1; ++i) p.digs[i] = pl->dig[i];
      p.pre_lo = ops.pre_lo;
      p.pre_hi = ops.pre_hi;
      if (inverse && m == 1) p.scale = pl->n_inv.p;
    } else {
      p.log_m = log_m_acc;
      p.log_t = p.log_m < kTileColsLog ? p.log_m : kTileColsLog;
      if (inverse && p.last) { p.tw_all = 1; p.tw_hi = pl->tw_hi_inv_scaled.p; }
    }
    if (p.last) {
      p.post_lo = ops.post_lo;
      p.post_hi = ops.post_hi;
      p.post_periodic = ops.post_periodic;
      p.post_mask = ops.post_mask;
    }
    const size_t tiles = pl->n >> (p.r + p.log_t);
    const size_t smem = (size_t)(F::N / 4) * 16 * ((size_t)1 << (p.r + p.log_t));
    if (tiles > 0x7fffffffull || k > 65535) fail(PLK_EINVAL, "transform grid too large");
    dim3 grid((unsigned)tiles, (unsigned)k);
    ntt_pass_kernel<F><<<grid, kNttThreads, smem, st>>>(p);
    PLK_LAUNCHED();
    log_m_acc += p.r;
  }
}

template <class P>
CosetTables* get_coset(plk_fft_plan* pl, const uint32_t* shift_limbs, cudaStream_t st) {
  typedef Fp<P> F;
  std::vector<uint32_t> key(shift_limbs, shift_limbs + F::N);
  std::lock_guard<std::mutex> lk(pl->mu);
  auto it = pl->cosets.find(key);
  if (it != pl->cosets.end()) return it->second;
  // s^-1 on device (one thread) -> host
  DevBuf tmp(2 * sizeof(F));
  PLK_CUDA(cudaMemcpyAsync(tmp.p, shift_limbs, sizeof(F), cudaMemcpyHostToDevice, st));
  plk_launch_field_inverse(pl->field, tmp.p, (char*)tmp.p + sizeof(F), 1, st);
  uint32_t inv_limbs[F::N];
  PLK_CUDA(cudaMemcpyAsync(inv_limbs, (char*)tmp.p + sizeof(F), sizeof(F), cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaStreamSynchronize(st));
  auto* ct = new CosetTables();
  const int L = pl->log_n;
  build_pow_table<P>(shift_limbs, 0, (size_t)1 << pl->lo_bits, nullptr, ct->fwd_lo, st);
  build_pow_table<P>(shift_limbs, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), nullptr, ct->fwd_hi, st);
  build_pow_table<P>(inv_limbs, 0, (size_t)1 << pl->lo_bits, nullptr, ct->inv_lo, st);
  build_pow_table<P>(inv_limbs, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), nullptr, ct->inv_hi, st);
  PLK_CUDA(cudaStreamSynchronize(st));
  pl->cosets[key] = ct;
  return ct;
}

template <class P>
const uint32_t* default_shift() {
  return FieldTables<P>::consts().gen;   // MULTIPLICATIVE_SUBGROUP_GENERATOR, Montgomery form
}

#define PLK_FIELD_DISPATCH(field, FN, ...)                                         \
  switch (field) {                                                                 \
    case PLK_FIELD_TWEEDLEDEE_BASE: FN<TweedledeeBaseParams>(__VA_ARGS__); break;  \
    case PLK_FIELD_TWEEDLEDUM_BASE: FN<TweedledumBaseParams>(__VA_ARGS__); break;  \
    case PLK_FIELD_BLS12_377_SCALAR: FN<Bls12377ScalarParams>(__VA_ARGS__); break; \
    case PLK_FIELD_BLS12_377_BASE: FN<Bls12377BaseParams>(__VA_ARGS__); break;     \
    default: fail(PLK_EINVAL, "unknown field id");                                 \
  }

int field_two_adicity(int field) {
  switch (field) {
    case PLK_FIELD_TWEEDLEDEE_BASE: return TweedledeeBaseParams::TWO_ADICITY;
    case PLK_FIELD_TWEEDLEDUM_BASE: return TweedledumBaseParams::TWO_ADICITY;
    case PLK_FIELD_BLS12_377_SCALAR: return Bls12377ScalarParams::TWO_ADICITY;
    case PLK_FIELD_BLS12_377_BASE: return Bls12377BaseParams::TWO_ADICITY;
  }
  return -1;
}

// dispatch wrappers (function templates cannot be passed to the macro with differing arity otherwise)
template <class P> void do_run(const plk_fft_plan* pl, const void* d_in, size_t n_in, size_t in_stride, void* d_out,
                               size_t k, bool inverse, const FusedOps* ops, cudaStream_t st) {
  run_ntt<P>(pl, d_in, n_in, in_stride, d_out, k, inverse, *ops, st);
}
template <class P> void do_coset(plk_fft_plan* pl, const uint64_t* shift, bool inverse, FusedOps* ops, cudaStream_t st) {
  const uint32_t* s = shift ? reinterpret_cast<const uint32_t*>(shift) : default_shift<P>();
  CosetTables* ct = get_coset<P>(pl, s, st);
  if (!inverse) { ops->pre_lo = ct->fwd_lo.p; ops->pre_hi = ct->fwd_hi.p; }
  else { ops->post_lo = ct->inv_lo.p; ops->post_hi = ct->inv_hi.p; }
}
template <class P> void do_zh_table(plk_fft_plan* pl, size_t n_gates, FusedOps* ops, cudaStream_t st) {
  typedef Fp<P> F;
  typedef FieldTables<P> Tb;
  std::lock_guard<std::mutex> lk(pl->mu);
  const size_t period = pl->n / n_gates;    // w^n_gates has order size / n_gates
  auto it = pl->zh_tables.find(n_gates);
  if (it == pl->zh_tables.end()) {
    // g^n and w^n via the pow-table kernel (1 entry each)
    DevBuf gw(2 * sizeof(F));
    F g = fp_from_host<F>(default_shift<P>());
    F w = fp_from_host<F>(Tb::root(pl->log_n));
    // base^(1 << log2 n_gates): count = 2 gives entries {1, base^stride}; take entry 1
    DevBuf t1(2 * sizeof(F)), t2(2 * sizeof(F));
    const int lg = log2_floor(n_gates);
    pow_table_kernel<F><<<1, 32, 0, st>>>(g, (unsigned long long)lg, 2ull, nullptr, t1.as<F>());
    PLK_LAUNCHED();
    pow_table_kernel<F><<<1, 32, 0, st>>>(w, (unsigned long long)lg, 2ull, nullptr, t2.as<F>());
    PLK_LAUNCHED();
    F hg[2], hw[2];
    PLK_CUDA(cudaMemcpyAsync(hg, t1.p, 2 * sizeof(F), cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(hw, t2.p, 2 * sizeof(F), cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    auto* buf = new DevBuf(period * sizeof(F));
    DevBuf flag(sizeof(int));
    PLK_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    zh_inverse_table_kernel<F><<<(unsigned)((period + 63) / 64), 64, 0, st>>>(hg[1], hw[1], (unsigned)period, buf->as<F>(), flag.as<int>());
    PLK_LAUNCHED();
    int hz = 0;
    PLK_CUDA(cudaMemcpyAsync(&hz, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    if (hz) { delete buf; fail(PLK_EZERO, "No inverse"); }   // field.rs:267
    it = pl->zh_tables.emplace(n_gates, buf).first;
  }
  ops->post_periodic = it->second->p;
  ops->post_mask = period - 1;
}

// host-pointer front end: copy in, run, copy out on the calling thread's stream
void host_transform(plk_fft_plan* pl, const uint64_t* in, size_t n_in, size_t k, bool inverse, const FusedOps& ops,
                    uint64_t* out) {
  cudaStream_t st = thread_stream();
  const size_t eb = pl->elem_bytes;
  void* d_in = thread_scratch(0, n_in * k * eb);
  void* d_out = thread_scratch(1, pl->n * k * eb);
  PLK_CUDA(cudaMemcpyAsync(d_in, in, n_in * k * eb, cudaMemcpyHostToDevice, st));
  PLK_FIELD_DISPATCH(pl->field, do_run, pl, d_in, n_in, n_in, d_out, k, inverse, &ops, st);
  PLK_CUDA(cudaMemcpyAsync(out, d_out, pl->n * k * eb, cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaStreamSynchronize(st));
}

void check_plan(const plk_fft_plan* p) {
  if (!p) fail(PLK_EINVAL, "NULL plan");
}

}  // namespace

extern "C" {

int plk_fft_precompute(int field, size_t degree, plk_fft_plan** out) {
  return guarded([&] {
    if (!out) fail(PLK_EINVAL, "NULL out");
    *out = nullptr;
    const int ta = field_two_adicity(field);
    if (ta < 0) fail(PLK_EINVAL, "unknown field id");
    if (degree == 0) degree = 1;
    const int L = log2_ceil(degree);         // fft.rs:48
    if (L > ta) fail(PLK_ETOOBIG, "log2(size) exceeds TWO_ADICITY");   // field.rs:430 assert
    auto* pl = new plk_fft_plan();
    pl->field = field;
    pl->log_n = L;
    pl->n = (size_t)1 << L;
    cudaGetDevice(&pl->device);
    try {
      PLK_FIELD_DISPATCH(field, plan_build, pl);
    } catch (...) {
      delete pl;
      throw;
    }
    *out = pl;
  });
}

size_t plk_fft_size(const plk_fft_plan* p) { return p ? p->n : 0; }

void plk_fft_free(plk_fft_plan* p) { delete p; }

int plk_fft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");                     // util.rs:16-19
    if (n != p->n) fail(PLK_ESIZE, "Number of coefficients does not match size of subgroup in precomputation");
    host_transform(const_cast<plk_fft_plan*>(p), in, n, 1, false, FusedOps(), out);
  });
}

int plk_ifft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");
    if (n != p->n) fail(PLK_ESIZE, "Number of points does not match size of subgroup in precomputation");
    host_transform(const_cast<plk_fft_plan*>(p), in, n, 1, true, FusedOps(), out);
  });
}

int plk_fft(const plk_fft_plan* p, const uint64_t* in, size_t n_in, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if ((!in && n_in) || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n) fail(PLK_ESIZE, "more coefficients than the precomputation size");
    // fft.rs:61-80 pads to the next power of two of n_in, which must be the plan size
    if (((size_t)1 << log2_ceil(n_in ? n_in : 1)) != p->n) fail(PLK_ESIZE, "padded length does not match precomputation size");
    host_transform(const_cast<plk_fft_plan*>(p), in, n_in, 1, false, FusedOps(), out);
  });
}

int plk_fft_batch(const plk_fft_plan* p, const uint64_t* in, size_t n_in, size_t k, int inverse, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad row length");
    if (inverse && n_in != p->n) fail(PLK_ESIZE, "inverse transform needs full rows");
    if (k == 0) return;
    host_transform(const_cast<plk_fft_plan*>(p), in, n_in, k, inverse != 0, FusedOps(), out);
  });
}

int plk_coset_lde(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, const uint64_t* shift, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!coeffs || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad coefficient count");
    FusedOps ops;
    auto* pl = const_cast<plk_fft_plan*>(p);
    PLK_FIELD_DISPATCH(pl->field, do_coset, pl, shift, false, &ops, thread_stream());
    host_transform(pl, coeffs, n_in, 1, false, ops, out);
  });
}

int plk_coset_ifft(const plk_fft_plan* p, const uint64_t* evals, const uint64_t* shift, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!evals || !out) fail(PLK_EINVAL, "NULL buffer");
    FusedOps ops;
    auto* pl = const_cast<plk_fft_plan*>(p);
    PLK_FIELD_DISPATCH(pl->field, do_coset, pl, shift, true, &ops, thread_stream());
    host_transform(pl, evals, pl->n, 1, true, ops, out);
  });
}

int plk_divide_by_z_h(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, size_t n_gates, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!coeffs || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad coefficient count");
    if (!is_pow2(n_gates) || n_gates > p->n) fail(PLK_EINVAL, "n_gates must be a power of two dividing the plan size");
    auto* pl = const_cast<plk_fft_plan*>(p);
    cudaStream_t st = thread_stream();
    FusedOps fwd, bwd;
    PLK_FIELD_DISPATCH(pl->field, do_coset, pl, nullptr, false, &fwd, st);
    PLK_FIELD_DISPATCH(pl->field, do_zh_table, pl, n_gates, &fwd, st);
    PLK_FIELD_DISPATCH(pl->field, do_coset, pl, nullptr, true, &bwd, st);
    const size_t eb = pl->elem_bytes;
    void* d_in = thread_scratch(0, n_in * eb);
    void* d_mid = thread_scratch(1, pl->n * eb);
    void* d_out = thread_scratch(2, pl->n * eb);
    PLK_CUDA(cudaMemcpyAsync(d_in, coeffs, n_in * eb, cudaMemcpyHostToDevice, st));
    PLK_FIELD_DISPATCH(pl->field, do_run, pl, d_in, n_in, n_in, d_mid, 1, false, &fwd, st);
    PLK_FIELD_DISPATCH(pl->field, do_run, pl, d_mid, pl->n, pl->n, d_out, 1, true, &bwd, st);
    PLK_CUDA(cudaMemcpyAsync(out, d_out, pl->n * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

int plk_fft_dev(const plk_fft_plan* p, const void* d_in, size_t n_in, size_t k, unsigned flags, void* d_out, void* stream) {
  return guarded([&] {
    check_plan(p);
    if (!d_in || !d_out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad row length");
    const bool inverse = (flags & PLK_FFT_INVERSE) != 0;
    if (inverse && n_in != p->n) fail(PLK_ESIZE, "inverse transform needs full rows");
    if (k == 0) return;
    auto* pl = const_cast<plk_fft_plan*>(p);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FusedOps ops;
    if (flags & PLK_FFT_COSET) { PLK_FIELD_DISPATCH(pl->field, do_coset, pl, nullptr, inverse, &ops, st); }
    PLK_FIELD_DISPATCH(pl->field, do_run, pl, d_in, n_in, n_in, d_out, k, inverse, &ops, st);
  });
}

}  // extern "C"
