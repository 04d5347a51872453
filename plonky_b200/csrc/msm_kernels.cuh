// Multi-scalar multiplication for sm_100a: sum_i s_i * P_i over Tweedledee / Tweedledum / BLS12-377 G1.
//
// Replaces src/curve/curve_msm.rs of the reference: msm_precompute (:27-52), msm_execute (:63-100),
// msm_execute_parallel (:102-157), msm_parallel (:54-61), to_digits (:159-180), and through them
// pedersen_hash (src/plonk_util.rs:193-198).  The reference is a Yao-style method: per-generator
// precomputed powers [(2^w)^j] G_i and ONE shared set of 2^w buckets, per-bucket batched-affine
// sums on rayon workers, then a sequential running sum.  The result is the unique group element
// sum_i s_i P_i, so any window size / coordinate system gives the same normalised point (F3, F5
// in SURVEY.md); the caller's `w` is kept only for interface fidelity.
//
// Device pipeline (per execute, all on one stream, no host round trip):
//   1. msm_count_kernel     scalars Montgomery -> canonical (one product by 1, = to_canonical_u64_vec),
//                           signed c-bit window recoding (digits in [-2^(c-1), 2^(c-1)]),
//                           histogram of |digit| over 2^(c-1) shared buckets (all windows share the
//                           buckets because the table holds [2^(c j)] P_i -- the reference's own trick)
//   2. msm_scan_kernel      exclusive scans: bucket offsets and task offsets (ceil(count / S) tasks)
//   3. msm_scatter_kernel   counting-sort scatter of (window, point) ids + sign into bucket order
//   4. msm_accumulate_kernel one thread per task: <= S mixed additions XYZZ += affine, points gathered
//                           from the table with 128-bit loads, next point prefetched
//   5. msm_bucket_sum_kernel one thread per bucket: sum of its task partials
//   6. msm_range_kernel     running sums over ranges of buckets + [lo] * range sum
//   7. msm_final_kernel     tree reduction of the range partials, to_affine
// Exceptional cases (identity, P == Q, P == -Q) are handled in every addition like the reference
// (src/curve/curve_adds.rs:12-33, src/curve/curve_summations.rs:107-141).
//
// Roofline accounting: algorithmic bytes = n * (32 + 2 * 8L); the table walk reads
// nwin * 2 * 8L bytes per term instead of one point (the reference streams its table the same way).
#pragma once
#include <stdlib.h>
#include "msm_plan.h"
#include "ec.cuh"

namespace plk {

// signed window recoding of a canonical scalar; calls f(window, bucket_index, negative)
// `base`: first bucket of the vector's bucket set (merged batch), 0 otherwise
template <class SF, class Fn>
__device__ __forceinline__ void for_each_digit(const SF& canon, const MsmGeom& g, unsigned base, Fn&& f) {
  if (g.c == 16 && !g.variable && g.nwin <= 2 * SF::N) {
    // 16-bit windows are the two halves of each limb: fully unrolled, no dynamic limb indexing (the generic loop
    // below costs ~1500 instructions per scalar, most of them address arithmetic around canon.l[limb])
    unsigned carry = 0;
#pragma unroll
    for (int j = 0; j < 2 * SF::N; ++j) {
      if (j < g.nwin) {
        unsigned raw = ((canon.l[j >> 1] >> ((j & 1) * 16)) & 0xffffu) + carry;
        carry = 0;
        if (raw > 0x8000u) {
          carry = 1;
          if (raw != 0x10000u) f(j, base + (0x10000u - raw - 1), true);
        } else if (raw != 0) {
          f(j, base + (raw - 1), false);
        }
      }
    }
    return;
  }
  unsigned carry = 0;
  const unsigned full = 1u << g.c, halfw = 1u << (g.c - 1), mask = full - 1;
  for (int j = 0; j < g.nwin; ++j) {
    const int bit = j * g.c;
    const int limb = bit >> 5, off = bit & 31;
    unsigned long long two = 0;
    if (limb < SF::N) two = canon.l[limb];
    if (limb + 1 < SF::N) two |= (unsigned long long)canon.l[limb + 1] << 32;
    unsigned raw = ((unsigned)(two >> off) & mask) + carry;
    carry = 0;
    if (raw > halfw) {
      // digit = raw - 2^c (negative or, when raw == 2^c, zero), carry one into the next window
      carry = 1;
      if (raw != full) f(j, base + (g.variable ? (unsigned)j * g.nbw : 0u) + (full - raw - 1), true);
    } else if (raw != 0) {
      f(j, base + (g.variable ? (unsigned)j * g.nbw : 0u) + (raw - 1), false);
    }
  }
}

// term i of a (possibly merged) execute: vector v = i / n_pts, point index i - v * n_pts, bucket set base v * nbw
__device__ __forceinline__ void term_of(const MsmGeom& g, unsigned long long i, unsigned long long& idx, unsigned& base) {
  idx = i;
  base = 0;
  if (g.batch > 1) {
    const unsigned v = (unsigned)(i / g.n_pts);
    idx = i - (unsigned long long)v * g.n_pts;
    base = v * g.nbw;
  }
}

template <class C>
__global__ void msm_count_kernel(const uint4* __restrict__ scalars, MsmGeom g, unsigned* __restrict__ counts) {
  typedef Fp<typename C::Scalar> SF;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  SF s = SF::to_canonical(load_fp<SF>(scalars, i));     // curve_msm.rs:164 to_canonical_u64_vec
  unsigned long long idx;
  unsigned base;
  term_of(g, i, idx, base);
  for_each_digit(s, g, base, [&](int, unsigned b, bool) { atomicAdd(&counts[b], 1u); });
}

// single CTA: offsets[b] = exclusive sum of counts, task_off[b] = exclusive sum of ceil(count / S);
// offsets[nb] / task_off[nb] = totals; cursors[b] = offsets[b] (scatter positions).  1024 entries per
// step: warp-shuffle scans, one shared-memory hop for the 32 warp totals.  (One contiguous strip per thread + a single
// block scan was measured slower: 0.106 vs 0.036 ms at 2^15 buckets -- strided, latency-bound loads from one CTA.)
static __global__ void msm_scan_kernel(const unsigned* __restrict__ counts, unsigned nb, unsigned task, unsigned* __restrict__ offsets,
                                       unsigned* __restrict__ task_off, unsigned* __restrict__ cursors) {
  __shared__ unsigned w_a[32], w_b[32];
  __shared__ unsigned carry_a, carry_b;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
  __syncthreads();
  for (unsigned base = 0; base < nb; base += 1024) {
    const unsigned i = base + threadIdx.x;
    const unsigned ca = i < nb ? counts[i] : 0;
    const unsigned cb = (ca + task - 1) / task;
    unsigned sa = ca, sb = cb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned ta = __shfl_up_sync(0xffffffffu, sa, d), tb = __shfl_up_sync(0xffffffffu, sb, d);
      if (lane >= (unsigned)d) { sa += ta; sb += tb; }
    }
    if (lane == 31) { w_a[warp] = sa; w_b[warp] = sb; }
    __syncthreads();
    if (warp == 0) {
      unsigned ta = w_a[lane], tb = w_b[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned ua = __shfl_up_sync(0xffffffffu, ta, d), ub = __shfl_up_sync(0xffffffffu, tb, d);
        if (lane >= (unsigned)d) { ta += ua; tb += ub; }
      }
      w_a[lane] = ta;       // inclusive scan of the warp totals
      w_b[lane] = tb;
    }
    __syncthreads();
    const unsigned pa = carry_a + (warp ? w_a[warp - 1] : 0), pb = carry_b + (warp ? w_b[warp - 1] : 0);
    if (i < nb) {
      offsets[i] = pa + sa - ca;
      cursors[i] = pa + sa - ca;
      task_off[i] = pb + sb - cb;
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry_a += w_a[31]; carry_b += w_b[31]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { offsets[nb] = carry_a; task_off[nb] = carry_b; }
}

template <class C>
__global__ void msm_scatter_kernel(const uint4* __restrict__ scalars, MsmGeom g, unsigned* __restrict__ cursors,
                                   unsigned* __restrict__ sorted) {
  typedef Fp<typename C::Scalar> SF;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  SF s = SF::to_canonical(load_fp<SF>(scalars, i));
  unsigned long long idx;
  unsigned base;
  term_of(g, i, idx, base);
  for_each_digit(s, g, base, [&](int j, unsigned b, bool negative) {
    unsigned pos = atomicAdd(&cursors[b], 1u);
    // entry = table slot (window-major: j * n_pts + idx) with the sign in bit 31 (n_pts * nwin < 2^31 checked on host)
    sorted[pos] = (unsigned)(g.variable ? i : (unsigned long long)j * g.n_pts + idx) | (negative ? 0x80000000u : 0u);
  });
}

// ---- counting sort without global atomics (fixed-base geometry, nb * 4 B fits in shared memory) ----
// The L2 atomic unit serves ~1e11 atomics/s: 2 x 16.7 M global atomics were 0.45 ms of a 3.7 ms execute.  Here CTA k
// owns the scalars [k * chunk, (k + 1) * chunk): msm_hist_kernel histograms its digits in shared memory and stores the
// row cta_hist[k][.]; msm_colsum_kernel turns each column into exclusive prefixes over the CTAs (+ the bucket totals);
// after the scan msm_scatter_smem_kernel re-derives the same digits and takes positions from shared-memory cursors
// initialised to offsets[b] + cta_hist[k][b].  Entries of one bucket end up grouped by CTA; their order inside a
// group depends on the atomics' arrival order, which only permutes the terms of a sum.
template <class C>
__global__ void __launch_bounds__(1024) msm_hist_kernel(const uint4* __restrict__ scalars, MsmGeom g, unsigned chunk,
                                                        unsigned* __restrict__ cta_hist) {
  typedef Fp<typename C::Scalar> SF;
  extern __shared__ unsigned sh_bins[];
  for (unsigned b = threadIdx.x; b < g.nb; b += blockDim.x) sh_bins[b] = 0;
  __syncthreads();
  const unsigned long long lo = (unsigned long long)blockIdx.x * chunk;
  unsigned long long hi = lo + chunk;
  if (hi > g.n) hi = g.n;
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    SF s = SF::to_canonical(load_fp<SF>(scalars, i));     // curve_msm.rs:164 to_canonical_u64_vec
    unsigned long long idx;
    unsigned base;
    term_of(g, i, idx, base);
    for_each_digit(s, g, base, [&](int, unsigned b, bool) { atomicAdd(&sh_bins[b], 1u); });
  }
  __syncthreads();
  unsigned* row = cta_hist + (size_t)blockIdx.x * g.nb;
  for (unsigned b = threadIdx.x; b < g.nb; b += blockDim.x) row[b] = sh_bins[b];
}
// one thread per bucket: cta_hist[k][b] <- sum_{k' < k} cta_hist[k'][b], counts[b] <- column total
static __global__ void msm_colsum_kernel(unsigned* __restrict__ cta_hist, unsigned nb, unsigned rows, unsigned* __restrict__ counts) {
  const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  unsigned run = 0;
  // batches of 16 rows: the loads of a batch are independent (a load-store-load chain over `rows` L2 round trips was 90 us)
  for (unsigned k0 = 0; k0 < rows; k0 += 16) {
    unsigned v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (k0 + u < rows) ? cta_hist[(size_t)(k0 + u) * nb + b] : 0u;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (k0 + u < rows) cta_hist[(size_t)(k0 + u) * nb + b] = run;
      run += v[u];
    }
  }
  counts[b] = run;
}
template <class C>
__global__ void __launch_bounds__(1024) msm_scatter_smem_kernel(const uint4* __restrict__ scalars, MsmGeom g, unsigned chunk,
                                                                const unsigned* __restrict__ cta_hist, const unsigned* __restrict__ offsets,
                                                                unsigned* __restrict__ sorted, unsigned b_lo, unsigned b_hi) {
  typedef Fp<typename C::Scalar> SF;
  extern __shared__ unsigned sh_bins[];
  const unsigned* row = cta_hist + (size_t)blockIdx.x * g.nb;
  for (unsigned b = threadIdx.x; b < g.nb; b += blockDim.x) sh_bins[b] = offsets[b] + row[b];
  __syncthreads();
  const unsigned long long lo = (unsigned long long)blockIdx.x * chunk;
  unsigned long long hi = lo + chunk;
  if (hi > g.n) hi = g.n;
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    SF s = SF::to_canonical(load_fp<SF>(scalars, i));
    unsigned long long idx;
    unsigned base;
    term_of(g, i, idx, base);
    for_each_digit(s, g, base, [&](int j, unsigned b, bool negative) {
      if (b < b_lo || b >= b_hi) return;         // another pass of the range-partitioned scatter (see execute_one)
      const unsigned pos = atomicAdd(&sh_bins[b], 1u);
      sorted[pos] = (unsigned)((unsigned long long)j * g.n_pts + idx) | (negative ? 0x80000000u : 0u);
    });
  }
}

template <class C>
__device__ __forceinline__ Affine<C> load_affine(const void* table, size_t slot) {
  typedef Fp<typename C::Base> F;
  Affine<C> a;
  a.x = load_fp<F>(table, 2 * slot);
  a.y = load_fp<F>(table, 2 * slot + 1);
  return a;
}
template <class C>
__device__ __forceinline__ void store_xyzz(void* base, size_t idx, const XYZZ<C>& p) {
  typedef Fp<typename C::Base> F;
  store_fp<F>(base, 4 * idx, p.x);
  store_fp<F>(base, 4 * idx + 1, p.y);
  store_fp<F>(base, 4 * idx + 2, p.zz);
  store_fp<F>(base, 4 * idx + 3, p.zzz);
}
template <class C>
__device__ __forceinline__ XYZZ<C> load_xyzz(const void* base, size_t idx) {
  typedef Fp<typename C::Base> F;
  XYZZ<C> p;
  p.x = load_fp<F>(base, 4 * idx);
  p.y = load_fp<F>(base, 4 * idx + 1);
  p.zz = load_fp<F>(base, 4 * idx + 2);
  p.zzz = load_fp<F>(base, 4 * idx + 3);
  return p;
}

}  // namespace plk
#include "msm_affine.cuh"
namespace plk {

// one thread per task: task t of bucket b adds entries [offsets[b] + k S, min(offsets[b+1], .. + S))
template <class C, int COMPACT = 0>
__global__ void __launch_bounds__(kAccThreads, (Fp<typename C::Base>::N <= 8 ? 4 : 1)) msm_accumulate_kernel(const void* __restrict__ table, const unsigned* __restrict__ sorted,
                                                                     const unsigned* __restrict__ offsets,
                                                                     const unsigned* __restrict__ task_off, unsigned b_lo, unsigned b_hi,
                                                                     unsigned task, void* __restrict__ partials) {
  typedef Fp<typename C::Base> F;
  // the tasks of buckets [b_lo, b_hi): the whole bucket set in one launch, or one part of the overlapped pipeline.
  // Grid-stride: a part's grid is sized for an even share of the tasks (its exact count is only known on the device);
  // a skewed digit distribution makes some threads take a second task instead of making the launch carry thousands
  // of CTAs that exit at once (each of those still waits for a register-file slot behind the working CTAs).
  const unsigned tend = task_off[b_hi];
  for (unsigned t = task_off[b_lo] + blockIdx.x * blockDim.x + threadIdx.x; t < tend; t += gridDim.x * blockDim.x) {
    // bucket of task t: last b with task_off[b] <= t
    unsigned lo = b_lo, hi = b_hi;
    while (hi - lo > 1) {
      unsigned mid = (lo + hi) >> 1;
      if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const unsigned b = lo;
    const unsigned start = offsets[b] + (t - task_off[b]) * task;
    unsigned end = offsets[b + 1];
    if (end > start + task) end = start + task;
    XYZZ<C> acc = XYZZ<C>::identity();
    unsigned e = sorted[start];
    Affine<C> next = load_affine<C>(table, e & 0x7fffffffu);
    for (unsigned k = start; k < end; ++k) {
      Affine<C> p = next;
      const bool negative = (e >> 31) != 0;
      if (k + 1 < end) {
        e = sorted[k + 1];
        next = load_affine<C>(table, e & 0x7fffffffu);
      }
      if (negative) p.y = F::neg(p.y);
      acc = COMPACT == 3 ? XYZZ<C>::template madd_compact<2>(acc, p)
          : COMPACT == 2 ? XYZZ<C>::template madd_compact<1>(acc, p)
          : COMPACT == 1 ? XYZZ<C>::template madd_compact<0>(acc, p) : XYZZ<C>::madd(acc, p);
    }
    store_xyzz<C>(partials, t, acc);
  }
}

// The three tail kernels below run one QUAD (4 lanes) per logical work item, see QuadXYZZ in ec.cuh.
// Two lanes per bucket: lane h sums the task partials t0 + h, t0 + h + 2, ..., one full-warp shuffle exchanges the two
// halves and both lanes add them (lane 0 stores).  No lane leaves before the shuffle, so the full mask is legal and
// the exchange is 32 plain SHFLs.  Half the dependent chain and twice the warps of one thread per bucket.
// (Measured alternatives at 2^15 buckets x 16 partials: one thread per bucket 0.20 ms; four lanes with quad-mask
// shuffles 0.27 ms; one QuadXYZZ quad per bucket 0.27 ms.)
template <class C>
__global__ void __launch_bounds__(128) msm_bucket_sum_kernel(const void* __restrict__ partials, const unsigned* __restrict__ task_off, unsigned b_lo,
                                                            unsigned b_hi, void* __restrict__ buckets, unsigned* __restrict__ big_list) {
  typedef Fp<typename C::Base> F;
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned b = b_lo + (gid >> 1), h = gid & 1;
  const bool valid = b < b_hi;
  const unsigned t0 = valid ? task_off[b] : 0u, t1 = valid ? task_off[b + 1] : 0u;
  const bool big = t1 - t0 > kBigBucket;            // skewed digit distribution: leave it to msm_big_bucket_kernel
  XYZZ<C> acc = XYZZ<C>::identity();
  if (!big)
    for (unsigned t = t0 + h; t < t1; t += 2) acc = XYZZ<C>::add(acc, load_xyzz<C>(partials, t));
  XYZZ<C> other;
#pragma unroll
  for (int k = 0; k < F::N; ++k) {
    other.x.l[k] = __shfl_xor_sync(0xffffffffu, acc.x.l[k], 1);
    other.y.l[k] = __shfl_xor_sync(0xffffffffu, acc.y.l[k], 1);
    other.zz.l[k] = __shfl_xor_sync(0xffffffffu, acc.zz.l[k], 1);
    other.zzz.l[k] = __shfl_xor_sync(0xffffffffu, acc.zzz.l[k], 1);
  }
  if (!valid || h != 0) return;
  if (big) { big_list[1 + atomicAdd(&big_list[0], 1u)] = b; return; }
  store_xyzz<C>(buckets, b, XYZZ<C>::add(acc, other));
}
// Buckets holding very many entries (all scalars equal, a short top window, ...): one CTA per bucket, strided
// partial sums per thread + shared-memory tree, so that the worst case stays a log-depth reduction.
template <class C>
__global__ void __launch_bounds__(256) msm_big_bucket_kernel(const void* __restrict__ partials, const unsigned* __restrict__ task_off,
                                                             void* __restrict__ buckets, const unsigned* __restrict__ big_list) {
  extern __shared__ uint4 sm[];
  const unsigned nbig = big_list[0];
  for (unsigned i = blockIdx.x; i < nbig; i += gridDim.x) {
    const unsigned b = big_list[1 + i];
    const unsigned t0 = task_off[b], t1 = task_off[b + 1];
    XYZZ<C> acc = XYZZ<C>::identity();
    for (unsigned t = t0 + threadIdx.x; t < t1; t += blockDim.x) acc = XYZZ<C>::add(acc, load_xyzz<C>(partials, t));
    store_xyzz<C>(sm, threadIdx.x, acc);
    __syncthreads();
    for (unsigned d = blockDim.x >> 1; d > 0; d >>= 1) {
      if (threadIdx.x < d) {
        XYZZ<C> x = load_xyzz<C>(sm, threadIdx.x), y = load_xyzz<C>(sm, threadIdx.x + d);
        store_xyzz<C>(sm, threadIdx.x, XYZZ<C>::add(x, y));
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz<C>(buckets, b, load_xyzz<C>(sm, 0));
    __syncthreads();
  }
}

// bucket index b carries weight (b + 1).  Range [lo, lo + R): sum_b (b + 1) B_b =
//   sum_b (b - lo + 1) B_b  (running sum, curve_msm.rs:149-154)  +  lo * sum_b B_b
template <class C>
__global__ void __launch_bounds__(kQuadThreads) msm_range_kernel(const void* __restrict__ buckets, unsigned r_lo, unsigned nb, unsigned nbw,
                                                                void* __restrict__ range_out) {
  typedef QuadXYZZ<C> Q;
  __shared__ uint4 quad_sm[(kQuadThreads / 4) * Q::kSmemVec];
  typename Q::Ctx qc = Q::make_ctx(quad_sm);
  const unsigned r = r_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 2);   // ranges r_lo .. ceil(nb / R) - 1
  const unsigned lo = r * kRangeSize;
  if (lo >= nb) return;
  unsigned hi = lo + kRangeSize;
  if (hi > nb) hi = nb;
  XYZZ<C> run = XYZZ<C>::identity(), sum = XYZZ<C>::identity();
  for (unsigned b = hi; b-- > lo;) {
    run = Q::add(run, load_xyzz<C>(buckets, b), qc);
    sum = Q::add(sum, run, qc);
  }
  const unsigned wlo = lo & (nbw - 1);            // index inside its window's bucket set (ranges never straddle windows)
  if (wlo != 0 && !run.is_identity()) sum = Q::add(sum, Q::mul_u64(run, wlo, qc), qc);
  if (qc.ql == 0) store_xyzz<C>(range_out, r, sum);
}

// out[i] = sum of in[i * chunk .. min(count, (i + 1) * chunk)), one quad per output
template <class C>
__global__ void __launch_bounds__(kQuadThreads) msm_sum_chunks_kernel(const void* __restrict__ in, unsigned count, unsigned chunk,
                                                                     void* __restrict__ out) {
  typedef QuadXYZZ<C> Q;
  __shared__ uint4 quad_sm[(kQuadThreads / 4) * Q::kSmemVec];
  typename Q::Ctx qc = Q::make_ctx(quad_sm);
  const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const unsigned lo = i * chunk;
  if (lo >= count) return;
  unsigned hi = lo + chunk;
  if (hi > count) hi = count;
  XYZZ<C> acc = load_xyzz<C>(in, lo);
  for (unsigned j = lo + 1; j < hi; ++j) acc = Q::add(acc, load_xyzz<C>(in, j), qc);
  if (qc.ql == 0) store_xyzz<C>(out, i, acc);
}

// single CTA of (blockDim / 4) quads: sum `count` XYZZ points; optionally normalise to (x, y, z = 1 | zero flag)
template <class C>
__global__ void __launch_bounds__(4 * kFinalQuadsMax) msm_final_kernel(const void* __restrict__ in, unsigned count, void* __restrict__ out_xyzz,
                                                                      uint32_t* __restrict__ out_xyz, unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  typedef QuadXYZZ<C> Q;
  extern __shared__ uint4 sm[];
  __shared__ uint4 quad_sm[kFinalQuadsMax * Q::kSmemVec];
  typename Q::Ctx qc = Q::make_ctx(quad_sm);
  const unsigned q = threadIdx.x >> 2, nq = blockDim.x >> 2;
  const int ql = qc.ql;
  // block b (merged batch: one block per vector) sums in[b * count .. (b + 1) * count) into the b-th output
  const size_t first = (size_t)blockIdx.x * count;
  if (out_xyzz) out_xyzz = static_cast<char*>(out_xyzz) + (size_t)blockIdx.x * 4 * sizeof(F);
  if (out_xyz) { out_xyz += (size_t)blockIdx.x * 3 * F::N; out_zero += blockIdx.x; }
  XYZZ<C> acc = XYZZ<C>::identity();
  for (unsigned i = q; i < count; i += nq) acc = Q::add(acc, load_xyzz<C>(in, first + i), qc);
  if (ql == 0) store_xyzz<C>(sm, q, acc);
  __syncthreads();
  for (unsigned d = nq >> 1; d > 0; d >>= 1) {
    XYZZ<C> s;
    if (q < d) s = Q::add(load_xyzz<C>(sm, q), load_xyzz<C>(sm, q + d), qc);
    __syncthreads();
    if (q < d && ql == 0) store_xyzz<C>(sm, q, s);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    XYZZ<C> total = load_xyzz<C>(sm, 0);
    if (out_xyzz) store_xyzz<C>(out_xyzz, 0, total);
    if (out_xyz) {
      Affine<C> a = XYZZ<C>::to_affine_gcd(total);
      const bool z = total.is_identity();
      F one = z ? F::zero() : F::one();
      for (int i = 0; i < F::N; ++i) {
        out_xyz[i] = a.x.l[i];
        out_xyz[F::N + i] = a.y.l[i];
        out_xyz[2 * F::N + i] = one.l[i];
      }
      *out_zero = z ? 1 : 0;
    }
  }
}

// variable base: result = sum_j 2^(c j) W_j (Horner from the top window: c doublings per step), one quad
template <class C>
__global__ void msm_window_combine_kernel(const void* __restrict__ window_sums, int nwin, int c, void* __restrict__ out_xyzz,
                                          uint32_t* __restrict__ out_xyz, unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  typedef QuadXYZZ<C> Q;
  __shared__ uint4 quad_sm[Q::kSmemVec];          // launched with one quad
  typename Q::Ctx qc = Q::make_ctx(quad_sm);
  XYZZ<C> acc = load_xyzz<C>(window_sums, nwin - 1);
  for (int j = nwin - 2; j >= 0; --j) {
    for (int k = 0; k < c; ++k) acc = Q::dbl(acc, qc);
    acc = Q::add(acc, load_xyzz<C>(window_sums, j), qc);
  }
  if (threadIdx.x == 0) {
    if (out_xyzz) store_xyzz<C>(out_xyzz, 0, acc);
    if (out_xyz) {
      Affine<C> a = XYZZ<C>::to_affine_gcd(acc);
      const bool z = acc.is_identity();
      F one = z ? F::zero() : F::one();
      for (int i = 0; i < F::N; ++i) {
        out_xyz[i] = a.x.l[i];
        out_xyz[F::N + i] = a.y.l[i];
        out_xyz[2 * F::N + i] = one.l[i];
      }
      *out_zero = z ? 1 : 0;
    }
  }
}

// ---- table construction: [2^(c j)] P_i for j < nwin (curve_msm.rs:40-52 with our own window) ----
// One thread per generator.  Forward sweep: c doublings per window in XYZZ; the un-normalised (X, Y) go to the
// table slot, (ZZ, ZZZ, running product of ZZ*ZZZ) to a scratch strip.  One inversion of the total product
// (Montgomery's trick, the batch_to_affine of the reference, curve.rs:216-232), then a backward sweep
// normalises every power with 7 products instead of one ~380-product inversion each.
template <class C>
__global__ void __launch_bounds__(128) msm_table_kernel(const void* __restrict__ points, unsigned long long n, int c, int nwin,
                                                        void* __restrict__ table, void* __restrict__ scratch) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p = load_affine<C>(points, i);
  store_fp<F>(table, 2 * i, p.x);            // window 0 is the point itself
  store_fp<F>(table, 2 * i + 1, p.y);
  if (p.is_identity()) {                     // every power of the identity is the identity
    for (int j = 1; j < nwin; ++j) {
      store_fp<F>(table, 2 * ((size_t)j * n + i), p.x);
      store_fp<F>(table, 2 * ((size_t)j * n + i) + 1, p.y);
    }
    return;
  }
  XYZZ<C> q = XYZZ<C>::from_affine(p);
  F prod = F::one();
  for (int j = 1; j < nwin; ++j) {
    for (int k = 0; k < c; ++k) q = XYZZ<C>::dbl(q);
    const size_t slot = (size_t)j * n + i;
    // A power can be the identity when the generator has 2-power order (BLS12-377 G1 has an even cofactor: e.g. (-1, 0)
    // on y^2 = x^3 + 1): it and every later power are stored as the identity and stay out of the batch product.
    const bool dead = q.is_identity();
    store_fp<F>(table, 2 * slot, dead ? F::zero() : q.x);
    store_fp<F>(table, 2 * slot + 1, dead ? F::zero() : q.y);
    store_fp<F>(scratch, 3 * slot, q.zz);                 // zz == 0 marks an identity power for the backward sweep
    store_fp<F>(scratch, 3 * slot + 1, q.zzz);
    store_fp<F>(scratch, 3 * slot + 2, prod);             // product of the ZZ*ZZZ of the live powers before this one
    if (!dead) prod = F::mul(prod, F::mul(q.zz, q.zzz));
  }
  F inv = F::inverse(prod);
  for (int j = nwin - 1; j >= 1; --j) {
    const size_t slot = (size_t)j * n + i;
    const F zz = load_fp<F>(scratch, 3 * slot), zzz = load_fp<F>(scratch, 3 * slot + 1), before = load_fp<F>(scratch, 3 * slot + 2);
    if (zz.is_zero()) continue;                           // identity power: (0, 0) is already in the table
    const F inv_j = F::mul(inv, before);                // 1 / (zz * zzz)
    inv = F::mul(inv, F::mul(zz, zzz));
    const F x = F::mul(load_fp<F>(table, 2 * slot), F::mul(inv_j, zzz));      // X / ZZ
    const F y = F::mul(load_fp<F>(table, 2 * slot + 1), F::mul(inv_j, zz));   // Y / ZZZ
    store_fp<F>(table, 2 * slot, x);
    store_fp<F>(table, 2 * slot + 1, y);
  }
}

// projective (x, y, z, zero) / affine (x, y, zero) host layouts -> device affine with identity = (0, 0)
template <class C>
__global__ void msm_import_points_kernel(const void* __restrict__ in, const unsigned char* __restrict__ zero, unsigned long long n,
                                         int projective, void* __restrict__ out) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> a;
  bool z = zero && zero[i];
  if (projective) {
    F x = load_fp<F>(in, 3 * i), y = load_fp<F>(in, 3 * i + 1), zz = load_fp<F>(in, 3 * i + 2);
    if (z || zz.is_zero()) a = Affine<C>::identity();
    else if (zz == F::one()) { a.x = x; a.y = y; }
    else { F zi = F::inverse(zz); a.x = F::mul(x, zi); a.y = F::mul(y, zi); }   // curve.rs:206-214
  } else {
    a.x = load_fp<F>(in, 2 * i);
    a.y = load_fp<F>(in, 2 * i + 1);
    if (z) a = Affine<C>::identity();
  }
  store_fp<F>(out, 2 * i, a.x);
  store_fp<F>(out, 2 * i + 1, a.y);
}

// P_i = [splitmix64(seed + i)] G (synthetic generators; same recipe as oracle/ref_port.cpp gen_points_t)
__device__ __forceinline__ unsigned long long splitmix_hash(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class C>
__global__ void __launch_bounds__(128) points_generate_kernel(Affine<C> gen, unsigned long long seed, unsigned long long n, void* __restrict__ out) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = splitmix_hash(seed + i);
  XYZZ<C> acc = XYZZ<C>::identity();
  const XYZZ<C> g = XYZZ<C>::from_affine(gen);
  for (int b = 63; b >= 0; --b) {
    acc = XYZZ<C>::dbl(acc);
    if ((k >> b) & 1) acc = XYZZ<C>::madd(acc, gen);
  }
  (void)g;
  Affine<C> a = XYZZ<C>::to_affine(acc);
  store_fp<F>(out, 2 * i, a.x);
  store_fp<F>(out, 2 * i + 1, a.y);
}

// normalise n projective points (batch_to_affine, curve.rs:216-232): one thread per point
template <class C>
__global__ void batch_to_affine_kernel(const void* __restrict__ in, const unsigned char* __restrict__ zero, unsigned long long n,
                                       void* __restrict__ out_xy, unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = load_fp<F>(in, 3 * i), y = load_fp<F>(in, 3 * i + 1), z = load_fp<F>(in, 3 * i + 2);
  const bool isz = (zero && zero[i]);
  if (isz) { x = F::zero(); y = F::zero(); }
  else { F zi = F::inverse(z); x = F::mul(x, zi); y = F::mul(y, zi); }
  store_fp<F>(out_xy, 2 * i, x);
  store_fp<F>(out_xy, 2 * i + 1, y);
  out_zero[i] = isz ? 1 : 0;
}


// ---- affine (multi-)summation, src/curve/curve_summations.rs:18-158 --------------------------------
// One CTA per list: every thread folds a strided slice of the list into an XYZZ accumulator with mixed
// additions, then a shared-memory tree adds the per-thread sums; the result is normalised like every
// other point this library returns.  (The reference batches affine additions with Montgomery's trick
// because a CPU inversion is cheap relative to 11 multiplications; on the device the mixed XYZZ add at
// 8M + 2S needs no inversion at all.)
template <class C>
__global__ void affine_multisum_kernel(const void* __restrict__ points, const unsigned char* __restrict__ zero,
                                       const unsigned long long* __restrict__ offsets, uint32_t* __restrict__ out_xyz,
                                       unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  extern __shared__ uint4 sm[];
  const unsigned long long lo = offsets[blockIdx.x], hi = offsets[blockIdx.x + 1];
  XYZZ<C> acc = XYZZ<C>::identity();
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (zero && zero[i]) continue;
    acc = XYZZ<C>::madd(acc, load_affine<C>(points, i));
  }
  store_xyzz<C>(sm, threadIdx.x, acc);
  __syncthreads();
  for (unsigned d = blockDim.x >> 1; d > 0; d >>= 1) {
    if (threadIdx.x < d) {
      XYZZ<C> a = load_xyzz<C>(sm, threadIdx.x), b = load_xyzz<C>(sm, threadIdx.x + d);
      store_xyzz<C>(sm, threadIdx.x, XYZZ<C>::add(a, b));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    XYZZ<C> total = load_xyzz<C>(sm, 0);
    Affine<C> a = XYZZ<C>::to_affine(total);
    const bool z = total.is_identity();
    F one = z ? F::zero() : F::one();
    uint32_t* o = out_xyz + (size_t)blockIdx.x * 3 * F::N;
    for (int i = 0; i < F::N; ++i) { o[i] = a.x.l[i]; o[F::N + i] = a.y.l[i]; o[2 * F::N + i] = one.l[i]; }
    out_zero[blockIdx.x] = z ? 1 : 0;
  }
}

// ---- n independent scalar multiplications, src/curve/curve_multiplication.rs:20-70 ------------------
// (CurveScalar * ProjectivePoint: blinding terms of src/poly_commit.rs:44 and the naive oracle of the
// reference's tests.)  One thread per pair, 4-bit fixed windows over the canonical scalar.
template <class C>
__global__ void __launch_bounds__(64) curve_mul_kernel(const void* __restrict__ points_xy, const void* __restrict__ scalars,
                                                       unsigned long long n, uint32_t* __restrict__ out_xyz,
                                                       unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine<C> p = load_affine<C>(points_xy, i);
  const SF s = SF::to_canonical(load_fp<SF>(scalars, i));
  XYZZ<C> acc = XYZZ<C>::identity();
  for (int bit = SF::N * 32 - 1; bit >= 0; --bit) {
    acc = XYZZ<C>::dbl(acc);
    if ((s.l[bit >> 5] >> (bit & 31)) & 1) acc = XYZZ<C>::madd(acc, p);
  }
  Affine<C> a = XYZZ<C>::to_affine(acc);
  const bool z = acc.is_identity();
  F one = z ? F::zero() : F::one();
  uint32_t* o = out_xyz + (size_t)i * 3 * F::N;
  for (int k = 0; k < F::N; ++k) { o[k] = a.x.l[k]; o[F::N + k] = a.y.l[k]; o[2 * F::N + k] = one.l[k]; }
  out_zero[i] = z ? 1 : 0;
}

// ---- blinded commitments, src/poly_commit.rs:32-66 -----------------------------------------------------
// coeffs_to_commitment: pedersen_hash(coeffs) + [blinding_factor] * blinding_point, then batch_to_affine over the k
// commitments.  One thread per commitment: the MSM results arrive normalised (msm_xyz: x, y, z per commitment), the
// blinding term is a plain double-and-add (one point, k scalars), one inversion each for the final affine form.
template <class C>
__global__ void __launch_bounds__(64) commit_blind_kernel(const uint32_t* __restrict__ msm_xyz, const unsigned char* __restrict__ msm_zero,
                                                          const void* __restrict__ blinding, Affine<C> h, unsigned long long k,
                                                          void* __restrict__ out_xy, unsigned char* __restrict__ out_zero) {
  typedef Fp<typename C::Base> F;
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= k) return;
  Affine<C> m = Affine<C>::identity();
  if (!msm_zero[i]) {
    const uint32_t* p = msm_xyz + (size_t)i * 3 * F::N;
    for (int j = 0; j < F::N; ++j) { m.x.l[j] = p[j]; m.y.l[j] = p[F::N + j]; }
  }
  XYZZ<C> acc = XYZZ<C>::identity();
  if (blinding) {
    const SF s = SF::to_canonical(load_fp<SF>(blinding, i));
    for (int bit = SF::N * 32 - 1; bit >= 0; --bit) {
      acc = XYZZ<C>::dbl(acc);
      if ((s.l[bit >> 5] >> (bit & 31)) & 1) acc = XYZZ<C>::madd(acc, h);
    }
  }
  acc = XYZZ<C>::madd(acc, m);
  const Affine<C> a = XYZZ<C>::to_affine(acc);
  store_fp<F>(out_xy, 2 * i, a.x);
  store_fp<F>(out_xy, 2 * i + 1, a.y);
  out_zero[i] = acc.is_identity() ? 1 : 0;
}

// ---- host-side launch templates ----

template <class C>
void table_build(plk_msm_table* t, const void* d_points, cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  t->point_bytes = 2 * sizeof(F);
  if (t->g.variable) {                       // msm_parallel: no powers are precomputed, the "table" is the point set
    t->table.alloc((size_t)t->n * t->point_bytes);
    if (t->n) PLK_CUDA(cudaMemcpyAsync(t->table.p, d_points, (size_t)t->n * t->point_bytes, cudaMemcpyDeviceToDevice, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    return;
  }
  t->table.alloc((size_t)t->g.nwin * t->n * t->point_bytes);
  if (t->n == 0) return;
  unsigned blocks = (unsigned)((t->n + 127) / 128);
  DevBuf scratch((size_t)t->g.nwin * t->n * 3 * sizeof(F), st);  // (ZZ, ZZZ, prefix product) per power; freed after the build
  msm_table_kernel<C><<<blocks, 128, 0, st>>>(d_points, t->n, t->g.c, t->g.nwin, t->table.p, scratch.p);
  PLK_LAUNCHED();
  PLK_CUDA(cudaStreamSynchronize(st));
}

template <class C>
void import_points(const void* d_raw, const unsigned char* d_zero, size_t n, int projective, void* d_out, cudaStream_t st) {
  if (n == 0) return;
  unsigned blocks = (unsigned)((n + 127) / 128);
  msm_import_points_kernel<C><<<blocks, 128, 0, st>>>(d_raw, d_zero, n, projective, d_out);
  PLK_LAUNCHED();
}

// bucket offsets + task offsets (+ the per-round offsets of the batched-affine tree)
static inline void launch_scan(plk_msm_scratch* s, const MsmGeom& g, cudaStream_t st) {
  if (g.affine_rounds == 0) {
    msm_scan_kernel<<<1, 1024, 0, st>>>(s->counts.as<unsigned>(), g.nb, g.task, s->offsets.as<unsigned>(), s->task_off.as<unsigned>(),
                                        s->cursors.as<unsigned>());
  } else {
    AffineRoundOffsets ro;
    for (int r = 0; r < kAffineRoundsMax; ++r) ro.off[r] = r < g.affine_rounds ? s->aff_off[r].as<unsigned>() : nullptr;
    if (g.affine_rounds == 1)
      msm_scan_rounds_kernel<1><<<1, 1024, 0, st>>>(s->counts.as<unsigned>(), g.nb, g.task, s->offsets.as<unsigned>(), ro, s->task_off.as<unsigned>(), s->cursors.as<unsigned>());
    else if (g.affine_rounds == 2)
      msm_scan_rounds_kernel<2><<<1, 1024, 0, st>>>(s->counts.as<unsigned>(), g.nb, g.task, s->offsets.as<unsigned>(), ro, s->task_off.as<unsigned>(), s->cursors.as<unsigned>());
    else
      msm_scan_rounds_kernel<3><<<1, 1024, 0, st>>>(s->counts.as<unsigned>(), g.nb, g.task, s->offsets.as<unsigned>(), ro, s->task_off.as<unsigned>(), s->cursors.as<unsigned>());
  }
  PLK_LAUNCHED();
}

// the whole pipeline for one scalar vector; writes either the normalised point or the XYZZ partial
template <class C>
void execute_one(plk_msm_table* t, plk_msm_scratch* s, const void* d_scalars, void* d_out_xyz, void* d_out_zero, void* d_partial,
                 cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  const MsmGeom g = s->g;                 // the table's geometry, or a merged batch of it (run_batch)
  const size_t xyzz = 4 * sizeof(F);
  if (g.n == 0) {
    // empty sum = identity (curve_msm.rs: y stays ProjectivePoint::ZERO)
    if (d_partial) PLK_CUDA(cudaMemsetAsync(d_partial, 0, xyzz, st));
    if (d_out_xyz) {
      PLK_CUDA(cudaMemsetAsync(d_out_xyz, 0, 3 * sizeof(F), st));
      PLK_CUDA(cudaMemsetAsync(d_out_zero, 1, 1, st));
    }
    return;
  }
  PLK_CUDA(cudaMemsetAsync(s->big_list.p, 0, 4, st));
  s->timer.begin(st);
  if (s->sort_rows) {
    // shared-memory histograms per CTA, no global atomics (see msm_hist_kernel)
    const unsigned rows = s->sort_rows, chunk = (unsigned)((g.n + rows - 1) / rows);
    const size_t smem = (size_t)g.nb * 4;
    {
      // the opt-in to > 48 KiB of dynamic shared memory is per device (a process may drive several GPUs)
      static std::mutex attr_mu;
      static bool attr_done[64] = {};
      int dev = 0;
      PLK_CUDA(cudaGetDevice(&dev));
      std::lock_guard<std::mutex> lk(attr_mu);
      if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        PLK_CUDA(cudaFuncSetAttribute(msm_hist_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortMaxBins * 4));
        PLK_CUDA(cudaFuncSetAttribute(msm_scatter_smem_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortMaxBins * 4));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
      }
    }
    msm_hist_kernel<C><<<rows, 1024, smem, st>>>(reinterpret_cast<const uint4*>(d_scalars), g, chunk, s->cta_hist.as<unsigned>());
    PLK_LAUNCHED();
    msm_colsum_kernel<<<(g.nb + 255) / 256, 256, 0, st>>>(s->cta_hist.as<unsigned>(), g.nb, rows, s->counts.as<unsigned>());
    PLK_LAUNCHED();
    s->timer.mark(st);
    launch_scan(s, g, st);
    s->timer.mark(st);
    // The scatter writes 4-byte entries all over `sorted`.  While the array fits in L2 (67 MB at 2^20 terms) the partial
    // sectors merge there; beyond it they are evicted half-filled and fetched again (BLS12-377 2^22 on one GPU, 268 MB:
    // 2.1 ms for 4x the entries that take 0.17 ms).  Then the scatter runs once per bucket range of <= 64 MiB of output --
    // the digits are recomputed per pass, which is cheap next to the write traffic.
    // PLK_MSM_SCATTER_PASS_KB: output bytes per pass (0 = always one pass; a small value forces the passes on small inputs: tests)
    static const long long pass_kb = getenv("PLK_MSM_SCATTER_PASS_KB") ? atoll(getenv("PLK_MSM_SCATTER_PASS_KB")) : -1;
    const size_t pass_bytes = pass_kb >= 0 ? (size_t)pass_kb << 10 : (size_t)64 << 20;
    const size_t threshold = pass_kb >= 0 ? pass_bytes : (size_t)96 << 20;
    const size_t sorted_bytes = (size_t)g.n * g.nwin * 4;
    unsigned passes = (pass_bytes && sorted_bytes > threshold) ? (unsigned)((sorted_bytes + pass_bytes - 1) / pass_bytes) : 1u;
    if (passes > 16) passes = 16;
    for (unsigned ps = 0; ps < passes; ++ps) {
      const unsigned b_lo = (unsigned)((unsigned long long)g.nb * ps / passes), b_hi = (unsigned)((unsigned long long)g.nb * (ps + 1) / passes);
      msm_scatter_smem_kernel<C><<<rows, 1024, smem, st>>>(reinterpret_cast<const uint4*>(d_scalars), g, chunk, s->cta_hist.as<unsigned>(),
                                                           s->offsets.as<unsigned>(), s->sorted.as<unsigned>(), b_lo, b_hi);
      PLK_LAUNCHED();
    }
    s->timer.mark(st);
  } else {
    PLK_CUDA(cudaMemsetAsync(s->counts.p, 0, (size_t)g.nb * 4, st));
    const unsigned sblocks = (unsigned)((g.n + 255) / 256);
    msm_count_kernel<C><<<sblocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_scalars), g, s->counts.as<unsigned>());
    PLK_LAUNCHED();
    s->timer.mark(st);
    launch_scan(s, g, st);
    s->timer.mark(st);
    msm_scatter_kernel<C><<<sblocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_scalars), g, s->cursors.as<unsigned>(),
                                                   s->sorted.as<unsigned>());
    PLK_LAUNCHED();
    s->timer.mark(st);
  }
  const unsigned ablocks = (unsigned)((s->max_tasks + kAccThreads - 1) / kAccThreads);
  if (g.affine_rounds > 0) {
    // batched-affine tree rounds (msm_affine.cuh), then the XYZZ task kernel over the shortened lists
    const size_t entries = (size_t)g.n * g.nwin;
    const char* pte = getenv("PLK_MSM_AFF_PER_THREAD");          // tuning knob
    const int per_thread_env = pte ? atoi(pte) : 0;
    const unsigned* in_off = s->offsets.as<unsigned>();
    for (int r = 0; r < g.affine_rounds; ++r) {
      const size_t bound = (entries >> (r + 1)) + g.nb + 1;                 // output slots of this round, upper bound
      // ~2 waves of 3 resident CTAs per SM, between 16 and 64 slots per thread (fewer slots: the shared inversion
      // amortises worse; more: too few CTAs to cover each other's inversion stall)
      size_t per = bound / ((size_t)148 * 3 * kAffThreads * 2);
      if (per < 16) per = 16;
      if (per > 64) per = 64;
      if (per_thread_env >= 1 && per_thread_env <= 4096) per = per_thread_env;
      const unsigned ctas = (unsigned)((bound + kAffThreads * per - 1) / (kAffThreads * per));
      unsigned* out_off = s->aff_off[r].as<unsigned>();
      if (r == 0)
        msm_affine_round_kernel<C, true><<<ctas, kAffThreads, 0, st>>>(t->table.p, s->sorted.as<unsigned>(), nullptr, in_off, out_off, g.nb,
                                                                       (unsigned)per, s->aff_prefix.p, s->aff[0].p);
      else
        msm_affine_round_kernel<C, false><<<ctas, kAffThreads, 0, st>>>(nullptr, nullptr, s->aff[(r - 1) & 1].p, in_off, out_off, g.nb,
                                                                        (unsigned)per, s->aff_prefix.p, s->aff[r & 1].p);
      PLK_LAUNCHED();
      in_off = out_off;
    }
    msm_accumulate_points_kernel<C><<<ablocks, kAccThreads, 0, st>>>(s->aff[(g.affine_rounds - 1) & 1].p, in_off, s->task_off.as<unsigned>(), g.nb,
                                                                     g.task, s->partials.p);
    PLK_LAUNCHED();
  } else {
    static const int compact = getenv("PLK_MSM_MADD_COMPACT") ? atoi(getenv("PLK_MSM_MADD_COMPACT")) : kMaddCompactDefault;   // see ec.cuh
    auto accumulate = [&](cudaStream_t q, unsigned b_lo, unsigned b_hi, unsigned ablocks) {
      if (compact == 3)
        msm_accumulate_kernel<C, 3><<<ablocks, kAccThreads, 0, q>>>(t->table.p, s->sorted.as<unsigned>(), s->offsets.as<unsigned>(),
                                                                    s->task_off.as<unsigned>(), b_lo, b_hi, g.task, s->partials.p);
      else if (compact == 2)
        msm_accumulate_kernel<C, 2><<<ablocks, kAccThreads, 0, q>>>(t->table.p, s->sorted.as<unsigned>(), s->offsets.as<unsigned>(),
                                                                    s->task_off.as<unsigned>(), b_lo, b_hi, g.task, s->partials.p);
      else if (compact == 1)
        msm_accumulate_kernel<C, 1><<<ablocks, kAccThreads, 0, q>>>(t->table.p, s->sorted.as<unsigned>(), s->offsets.as<unsigned>(),
                                                                    s->task_off.as<unsigned>(), b_lo, b_hi, g.task, s->partials.p);
      else
        msm_accumulate_kernel<C, 0><<<ablocks, kAccThreads, 0, q>>>(t->table.p, s->sorted.as<unsigned>(), s->offsets.as<unsigned>(),
                                                                    s->task_off.as<unsigned>(), b_lo, b_hi, g.task, s->partials.p);
      PLK_LAUNCHED();
    };
    const int parts = s->parts_for(g, t->temporary);
    if (parts > 1) {
      // Overlapped pipeline (opt-in, PLK_MSM_OVERLAP_PARTS; see parts_for for why it is not the default).  The bucket set
      // is cut into `parts` equal ranges; the accumulate kernels of the ranges run on streams of descending priority
      // (PLK_MSM_OVERLAP_MODE=1: back to back on the caller's stream), and as soon as one range is accumulated its
      // reduction tail -- bucket_sum, big buckets, running sums, chunk sums -- runs on the highest-priority stream
      // underneath the accumulation of the following ranges; only the last range's tail and the final tree stay exposed.
      // PLK_MSM_OVERLAP_TRACE=1 prints when each part's accumulation and tail finished (profiles/r2_msm_overlap_trace.txt).
      s->ensure_streams(parts);
      const unsigned per = g.nb / parts, rper = per / kRangeSize;
      const unsigned part_blocks = ablocks / parts + ablocks / (32 * parts) + 1;      // an even share + 3 %; the kernel strides over any excess
      PLK_CUDA(cudaMemsetAsync(s->big_list.p, 0, ((size_t)g.nb + 1 + kPartsMax) * 4, st));
      PLK_CUDA(cudaEventRecord(s->fork_ev, st));
      static const bool trace = getenv("PLK_MSM_OVERLAP_TRACE") != nullptr;
      static cudaEvent_t tr0, tr_acc[kPartsMax], tr_tail[kPartsMax];
      static bool tr_made = false;
      if (trace && !tr_made) {
        cudaEventCreate(&tr0);
        for (int i = 0; i < kPartsMax; ++i) { cudaEventCreate(&tr_acc[i]); cudaEventCreate(&tr_tail[i]); }
        tr_made = true;
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        fprintf(stderr, "[trace] stream priority range least=%d greatest=%d\n", least, greatest);
      }
      if (trace) cudaEventRecord(tr0, st);
      unsigned out_per = 0;
      const void* last = nullptr;
      for (int k = 0; k < parts; ++k) {
        const unsigned b_lo = k * per, b_hi = b_lo + per;
        static const int same_stream = getenv("PLK_MSM_OVERLAP_MODE") ? atoi(getenv("PLK_MSM_OVERLAP_MODE")) : 0;
        cudaStream_t q = same_stream ? st : s->part_st[k];
        if (!same_stream) PLK_CUDA(cudaStreamWaitEvent(q, s->fork_ev, 0));
        accumulate(q, b_lo, b_hi, part_blocks);
        PLK_CUDA(cudaEventRecord(s->acc_ev[k], q));
        if (trace) cudaEventRecord(tr_acc[k], q);
        cudaStream_t ts = s->tail_st;
        PLK_CUDA(cudaStreamWaitEvent(ts, s->acc_ev[k], 0));
        unsigned* list = s->big_list.as<unsigned>() + b_lo + k;            // per - 1 slots + the count: one list per part
        msm_bucket_sum_kernel<C><<<(2 * per + 127) / 128, 128, 0, ts>>>(s->partials.p, s->task_off.as<unsigned>(), b_lo, b_hi, s->buckets.p, list);
        PLK_LAUNCHED();
        // every tail CTA must fit the hole ONE retiring accumulate CTA leaves (128 threads x 128 registers): a larger CTA
        // is never placed while the accumulation of the next part keeps refilling the register file, and the whole
        // tail chain waits behind it (measured: 256-thread big-bucket CTAs serialised the tails after the accumulation)
        msm_big_bucket_kernel<C><<<16, 128, 128 * xyzz, ts>>>(s->partials.p, s->task_off.as<unsigned>(), s->buckets.p, list);
        PLK_LAUNCHED();
        msm_range_kernel<C><<<(4 * rper + kQuadThreads - 1) / kQuadThreads, kQuadThreads, 0, ts>>>(s->buckets.p, k * rper, b_hi, g.nbw, s->ranges.p);
        PLK_LAUNCHED();
        // chunk sums of this part's ranges down to <= 16 values, ping-pong in the two small buffers; every part does the
        // same number of passes, so the last pass of all parts lands contiguously in one buffer
        const char* in = static_cast<const char*>(s->ranges.p) + (size_t)k * rper * xyzz;
        unsigned cnt = rper;
        int pass = 0;
        while (cnt > 16) {
          const unsigned chunk = (cnt + 15) / 16 < 8 ? (cnt + 15) / 16 : 8;
          const unsigned nout = (cnt + chunk - 1) / chunk;
          char* dst = static_cast<char*>(s->chunks[pass & 1].p) + (size_t)k * nout * xyzz;
          msm_sum_chunks_kernel<C><<<(4 * nout + kQuadThreads - 1) / kQuadThreads, kQuadThreads, 0, ts>>>(in, cnt, chunk, dst);
          PLK_LAUNCHED();
          in = dst;
          cnt = nout;
          ++pass;
        }
        out_per = cnt;
        last = pass ? s->chunks[(pass - 1) & 1].p : s->ranges.p;
        if (trace) cudaEventRecord(tr_tail[k], ts);
      }
      if (trace) {
        cudaStreamSynchronize(s->tail_st);
        for (int k = 0; k < parts; ++k) {
          float a = 0, b = 0;
          cudaEventElapsedTime(&a, tr0, tr_acc[k]);
          cudaEventElapsedTime(&b, tr0, tr_tail[k]);
          fprintf(stderr, "[trace] part %d: accumulate done at %.3f ms, tail done at %.3f ms\n", k, a, b);
        }
      }
      PLK_CUDA(cudaEventRecord(s->tail_ev, s->tail_st));
      PLK_CUDA(cudaStreamWaitEvent(st, s->tail_ev, 0));
      const unsigned fcount = out_per * parts;
      unsigned fquads = 1;
      while (fquads < fcount && fquads < 64) fquads <<= 1;
      msm_final_kernel<C><<<1, 4 * fquads, fquads * xyzz, st>>>(last, fcount, d_partial, reinterpret_cast<uint32_t*>(d_out_xyz),
                                                              reinterpret_cast<unsigned char*>(d_out_zero));
      PLK_LAUNCHED();
      return;
    }
    accumulate(st, 0, g.nb, ablocks);
  }
  s->timer.mark(st);
  msm_bucket_sum_kernel<C><<<(2 * g.nb + 127) / 128, 128, 0, st>>>(s->partials.p, s->task_off.as<unsigned>(), 0, g.nb, s->buckets.p,
                                                               s->big_list.as<unsigned>());
  PLK_LAUNCHED();
  msm_big_bucket_kernel<C><<<64, 256, 256 * xyzz, st>>>(s->partials.p, s->task_off.as<unsigned>(), s->buckets.p, s->big_list.as<unsigned>());
  PLK_LAUNCHED();
  s->timer.mark(st);
  const unsigned nranges = (g.nb + kRangeSize - 1) / kRangeSize;
  msm_range_kernel<C><<<(4 * nranges + kQuadThreads - 1) / kQuadThreads, kQuadThreads, 0, st>>>(s->buckets.p, 0, g.nb, g.nbw, s->ranges.p);
  PLK_LAUNCHED();
  s->timer.mark(st);
  const void* fin = s->ranges.p;
  unsigned fcount = nranges;
  auto chunk_pass = [&](unsigned chunk) {
    const unsigned nout = (fcount + chunk - 1) / chunk;
    void* dst = (fin == s->partials.p) ? s->buckets.p : s->partials.p;     // ping-pong between consumed buffers
    msm_sum_chunks_kernel<C><<<(4 * nout + kQuadThreads - 1) / kQuadThreads, kQuadThreads, 0, st>>>(fin, fcount, chunk, dst);
    PLK_LAUNCHED();
    fin = dst;
    fcount = nout;
  };
  if (g.variable) {
    // per-window sums (chunks never straddle a window: both sizes are powers of two), then Horner
    unsigned per_window = g.nbw / kRangeSize;
    while (per_window > 1) {
      const unsigned chunk = per_window < 8 ? per_window : 8;
      chunk_pass(chunk);
      per_window /= chunk;
    }
    msm_window_combine_kernel<C><<<1, 4, 0, st>>>(fin, g.nwin, g.c, d_partial, reinterpret_cast<uint32_t*>(d_out_xyz),
                                                  reinterpret_cast<unsigned char*>(d_out_zero));
    PLK_LAUNCHED();
    s->timer.mark(st);
    return;
  }
  if (g.batch > 1) {
    // merged batch: chunk sums inside every vector's bucket set (sizes are powers of two: chunks never straddle two
    // sets) down to <= 64 values per vector, then one final block per vector
    unsigned per_set = g.nbw / kRangeSize;
    while (per_set > 64) {
      chunk_pass(8);
      per_set /= 8;
    }
    unsigned bq = 1;
    while (bq < per_set && bq < 64) bq <<= 1;
    msm_final_kernel<C><<<g.batch, 4 * bq, bq * xyzz, st>>>(fin, per_set, nullptr, reinterpret_cast<uint32_t*>(d_out_xyz),
                                                          reinterpret_cast<unsigned char*>(d_out_zero));
    PLK_LAUNCHED();
    s->timer.mark(st);
    return;
  }
  // fixed base: chunks of 8 -> <= 64 quads -> tree -> normalise
  while (fcount > 64) chunk_pass((fcount + 63) / 64 < 8 ? (fcount + 63) / 64 : 8);
  unsigned fquads = 1;
  while (fquads < fcount && fquads < 64) fquads <<= 1;
  msm_final_kernel<C><<<1, 4 * fquads, fquads * xyzz, st>>>(fin, fcount, d_partial, reinterpret_cast<uint32_t*>(d_out_xyz),
                                                          reinterpret_cast<unsigned char*>(d_out_zero));
  PLK_LAUNCHED();
  s->timer.mark(st);
}

template <class C>
void combine_partials(const void* d_partials, size_t count, void* d_out_xyz, void* d_out_zero, cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  unsigned fquads = 1;
  while (fquads < count && fquads < 8) fquads <<= 1;
  msm_final_kernel<C><<<1, 4 * fquads, fquads * 4 * sizeof(F), st>>>(d_partials, (unsigned)count, nullptr,
                                                                     reinterpret_cast<uint32_t*>(d_out_xyz),
                                                                     reinterpret_cast<unsigned char*>(d_out_zero));
  PLK_LAUNCHED();
}

template <class C>
void generate_points(uint64_t seed, size_t n, void* d_out, cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  if (n == 0) return;
  Affine<C> g;
  for (int i = 0; i < F::N; ++i) { g.x.l[i] = CurveTables<C>::gen_x()[i]; g.y.l[i] = CurveTables<C>::gen_y()[i]; }
  points_generate_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, seed, n, d_out);
  PLK_LAUNCHED();
}

template <class C>
void to_affine_batch(const void* d_in, const unsigned char* d_zero, size_t n, void* d_out, unsigned char* d_out_zero, cudaStream_t st) {
  if (n == 0) return;
  batch_to_affine_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_in, d_zero, n, d_out, d_out_zero);
  PLK_LAUNCHED();
}


template <class C>
void multisum(const void* d_points, const unsigned char* d_zero, const unsigned long long* d_offsets, size_t lists, void* d_out_xyz,
              unsigned char* d_out_zero, cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  if (lists == 0) return;
  const unsigned threads = 128;
  affine_multisum_kernel<C><<<(unsigned)lists, threads, threads * 4 * sizeof(F), st>>>(d_points, d_zero, d_offsets,
                                                                                   reinterpret_cast<uint32_t*>(d_out_xyz), d_out_zero);
  PLK_LAUNCHED();
}
template <class C>
void curve_mul(const void* d_points_xy, const void* d_scalars, size_t n, void* d_out_xyz, unsigned char* d_out_zero, cudaStream_t st) {
  if (n == 0) return;
  curve_mul_kernel<C><<<(unsigned)((n + 63) / 64), 64, 0, st>>>(d_points_xy, d_scalars, n, reinterpret_cast<uint32_t*>(d_out_xyz), d_out_zero);
  PLK_LAUNCHED();
}

template <class C>
void commit_blind(const void* d_msm_xyz, const unsigned char* d_msm_zero, const void* d_blinding, const uint64_t* h_xy, bool h_zero, size_t k,
                  void* d_out_xy, unsigned char* d_out_zero, cudaStream_t st) {
  typedef Fp<typename C::Base> F;
  if (k == 0) return;
  Affine<C> h = Affine<C>::identity();
  if (!h_zero && h_xy) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(h_xy);
    for (int i = 0; i < F::N; ++i) { h.x.l[i] = p[i]; h.y.l[i] = p[F::N + i]; }
  }
  commit_blind_kernel<C><<<(unsigned)((k + 63) / 64), 64, 0, st>>>(reinterpret_cast<const uint32_t*>(d_msm_xyz), d_msm_zero, d_blinding, h, k,
                                                                 d_out_xy, d_out_zero);
  PLK_LAUNCHED();
}

template <class C>
const MsmOps* make_msm_ops() {
  static const MsmOps ops = {&table_build<C>, &import_points<C>, &execute_one<C>, &combine_partials<C>, &generate_points<C>,
                             &to_affine_batch<C>, &multisum<C>, &curve_mul<C>, &commit_blind<C>};
  return &ops;
}
}  // namespace plk
