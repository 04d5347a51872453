// Batched-affine bucket accumulation for the fixed-base MSM (sm_100a).
//
// The reference sums every bucket with affine additions whose divisions share ONE inversion
// (affine_multisummation_batch_inversion, src/curve/curve_summations.rs:70-158: pair up neighbours, collect
// x1 - x2 or 2 y1 (:86-92), batch-invert (:96), add / double per pair (:124-140), carry the odd tail (:146-148),
// recurse on the halves).  This is the same tree on the device, applied to the bucket-sorted entry list:
//   round r:  every bucket's list of m points becomes ceil(m / 2) points: (P_0 + P_1), (P_2 + P_3), ..., [P_(m-1)]
//   an affine addition with a shared inversion costs 5 products + 1 squaring (1 product for the running product,
//   2 to peel the pair's inverse off the batch inverse, 1 for lambda, 1 squaring for x3, 1 product for y3) against
//   8 + 2 for the mixed XYZZ addition; after kAffineRoundsMax rounds the (4x or 8x shorter) lists go to the XYZZ
//   task kernel, whose partials then feed the unchanged bucket-sum / running-sum tails.
// Work decomposition: the OUTPUT slots of a round are numbered globally (bucket-major); thread t of a CTA owns
// `per_thread` consecutive slots wherever the bucket borders fall, so skewed digit distributions (all scalars
// equal, a short top window) stay balanced.  One inversion per CTA: every thread multiplies up the denominators of
// its slots (forward sweep, running products parked in a global strip), the 256 thread products are combined by
// warp-shuffle prefix / suffix scans, ONE lane inverts the CTA total with the binary extended GCD (the reference's own
// inversion, src/bigint/bigint_inverse.rs:6-55 -- integer add / shift work on the ALU pipe, not on the IMAD pipe the
// products saturate) while the other CTAs of the SM keep the multipliers busy, and the backward sweep peels one
// inverse per slot.  P == Q (doubling: denominator 2 y, :88-90), P == -Q (identity), identity operands and odd
// tails are classified per pair exactly like the reference does; they contribute 1 to the batch product.
#pragma once
#include "ec.cuh"

namespace plk {

constexpr int kAffThreads = 128;     // small CTAs: one lane per CTA inverts while the others wait; 4 resident CTAs per SM cover for it
constexpr int kAffineRoundsMax = 3;

struct AffineRoundOffsets {
  unsigned* off[kAffineRoundsMax];      // off[r][b]: exclusive scan of count_(r+1)[b] = ceil(count_r[b] / 2)
};

// One pass over the bucket histogram: offsets of the sorted entries, offsets of every affine round's output and the
// task offsets of the XYZZ kernel that consumes the last round.  Single CTA, 1024 buckets per step (see msm_scan_kernel).
template <int R>
__global__ void msm_scan_rounds_kernel(const unsigned* __restrict__ counts, unsigned nb, unsigned task, unsigned* __restrict__ offsets,
                                       AffineRoundOffsets ro, unsigned* __restrict__ task_off, unsigned* __restrict__ cursors) {
  constexpr int NQ = R + 2;
  __shared__ unsigned w[NQ][32];
  __shared__ unsigned carry[NQ];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ) carry[threadIdx.x] = 0;
  __syncthreads();
  for (unsigned base = 0; base < nb; base += 1024) {
    const unsigned i = base + threadIdx.x;
    unsigned v[NQ], s[NQ];
    v[0] = i < nb ? counts[i] : 0;
#pragma unroll
    for (int r = 1; r <= R; ++r) v[r] = (v[r - 1] + 1) >> 1;
    v[R + 1] = (v[R] + task - 1) / task;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      s[q] = v[q];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, s[q], d);
        if (lane >= (unsigned)d) s[q] += t;
      }
      if (lane == 31) w[q][warp] = s[q];
    }
    __syncthreads();
    if (warp < NQ) {
      unsigned t = w[warp][lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= (unsigned)d) t += u;
      }
      w[warp][lane] = t;
    }
    __syncthreads();
    if (i < nb) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const unsigned ex = carry[q] + (warp ? w[q][warp - 1] : 0) + s[q] - v[q];
        if (q == 0) { offsets[i] = ex; cursors[i] = ex; }
        else if (q <= R) ro.off[q - 1][i] = ex;
        else task_off[i] = ex;
      }
    }
    __syncthreads();
    if (threadIdx.x < NQ) carry[threadIdx.x] += w[threadIdx.x][31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    offsets[nb] = carry[0];
#pragma unroll
    for (int r = 1; r <= R; ++r) ro.off[r - 1][nb] = carry[r];
    task_off[nb] = carry[R + 1];
  }
}

// what one pair of a round turns into
enum AffCase : int { kAffNormal = 0, kAffDouble = 1, kAffCopy1 = 2, kAffCopy2 = 3, kAffZero = 4 };

template <class C>
__device__ __forceinline__ int aff_classify(const Affine<C>& p1, const Affine<C>& p2) {
  if (p1.is_identity()) return kAffCopy2;                    // curve_summations.rs:107-113 (zero operands)
  if (p2.is_identity()) return kAffCopy1;
  if (p1.x != p2.x) return kAffNormal;
  if (p1.y == p2.y && !p1.y.is_zero()) return kAffDouble;    // :88-90, :126-133
  return kAffZero;                                           // P == -Q (:134-136), or a point of order two doubled
}

// Product of the 256 thread values `mine` (non-zero) -> every thread receives 1 / mine.  Warp-shuffle prefix and
// suffix scans, the 8 warp totals meet in shared memory, lane 0 of warp 0 inverts the CTA total.
template <class F>
__device__ __forceinline__ F cta_batch_inverse(const F& mine, uint4* sm /* (2 * 8 + 1) elements of F */) {
  constexpr int V = F::N / 4;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kAffThreads / 32;
  // inclusive prefix / suffix products inside the warp
  F pre = mine, suf = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    F a, b;
#pragma unroll
    for (int k = 0; k < F::N; ++k) {
      a.l[k] = __shfl_up_sync(0xffffffffu, pre.l[k], d);
      b.l[k] = __shfl_down_sync(0xffffffffu, suf.l[k], d);
    }
    if (lane >= (unsigned)d) pre = fp_mul_call<F>(pre, a);
    if (lane + d < 32) suf = fp_mul_call<F>(suf, b);
  }
  // exclusive versions: value of the neighbouring lane (ONE at the warp border)
  F pre_ex, suf_ex;
#pragma unroll
  for (int k = 0; k < F::N; ++k) {
    pre_ex.l[k] = __shfl_up_sync(0xffffffffu, pre.l[k], 1);
    suf_ex.l[k] = __shfl_down_sync(0xffffffffu, suf.l[k], 1);
  }
  if (lane == 0) pre_ex = F::one();
  if (lane == 31) suf_ex = F::one();
  if (lane == 31) store_fp<F>(sm, warp, pre);                // warp total
  __syncthreads();
  if (threadIdx.x == 0) {
    // exclusive prefix / suffix products of the warp totals, then ONE inversion
    F tot[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) tot[i] = load_fp<F>(sm, i);
    F acc = F::one();
#pragma unroll
    for (int i = 0; i < NW; ++i) { store_fp<F>(sm, i, acc); acc = fp_mul_call<F>(acc, tot[i]); }            // sm[i] = prod_{j < i}
    const F inv_total = F::inverse_gcd(acc);
    F back = inv_total;
#pragma unroll
    for (int i = NW - 1; i >= 0; --i) { store_fp<F>(sm, NW + i, back); back = fp_mul_call<F>(back, tot[i]); }   // sm[NW + i] = inv_total * prod_{j > i}
  }
  __syncthreads();
  // 1 / mine = inv_total * (product of everything else) = pre_ex * suf_ex * Wpre[warp] * (inv_total * Wsuf[warp])
  F r = fp_mul_call<F>(pre_ex, suf_ex);
  r = fp_mul_call<F>(r, load_fp<F>(sm, warp));
  r = fp_mul_call<F>(r, load_fp<F>(sm, NW + warp));
  (void)V;
  return r;
}

template <class C, bool FIRST>
__device__ __forceinline__ Affine<C> aff_load_point(const void* __restrict__ table, const unsigned* __restrict__ sorted, const void* __restrict__ pts,
                                                    unsigned idx) {
  typedef Fp<typename C::Base> F;
  if (FIRST) {
    const unsigned e = sorted[idx];
    Affine<C> p = load_affine<C>(table, e & 0x7fffffffu);
    if (e >> 31) p.y = F::neg(p.y);
    return p;
  }
  return load_affine<C>(pts, idx);
}
template <class C, bool FIRST>
__device__ __forceinline__ Fp<typename C::Base> aff_load_x(const void* __restrict__ table, const unsigned* __restrict__ sorted,
                                                           const void* __restrict__ pts, unsigned idx) {
  typedef Fp<typename C::Base> F;
  if (FIRST) return load_fp<F>(table, 2 * (size_t)(sorted[idx] & 0x7fffffffu));
  return load_fp<F>(pts, 2 * (size_t)idx);
}

// One round of the pairwise tree.  in_off / out_off: bucket offsets of the input list and of this round's output.
template <class C, bool FIRST>
__global__ void __launch_bounds__(kAffThreads, 4) msm_affine_round_kernel(const void* __restrict__ table, const unsigned* __restrict__ sorted,
                                                                       const void* __restrict__ in_pts, const unsigned* __restrict__ in_off,
                                                                       const unsigned* __restrict__ out_off, unsigned nb, unsigned per_thread,
                                                                       void* __restrict__ prefix, void* __restrict__ out_pts) {
  typedef Fp<typename C::Base> F;
  __shared__ uint4 inv_sm[(2 * (kAffThreads / 32)) * (F::N / 4)];
  const unsigned total = out_off[nb];
  const unsigned cta_base = blockIdx.x * (kAffThreads * per_thread);
  if (cta_base >= total) return;                             // uniform per CTA
  unsigned s0 = cta_base + threadIdx.x * per_thread;
  unsigned s1 = s0 + per_thread;
  if (s0 > total) s0 = total;
  if (s1 > total) s1 = total;
  // bucket of slot s0: last b with out_off[b] <= s0
  unsigned b = 0;
  if (s0 < s1) {
    unsigned lo = 0, hi = nb;
    while (hi - lo > 1) {
      const unsigned mid = (lo + hi) >> 1;
      if (out_off[mid] <= s0) lo = mid; else hi = mid;
    }
    b = lo;
  }
  unsigned ob = out_off[b], oe = out_off[b + 1], ib = in_off[b], ie = in_off[b + 1];

  // ---- forward sweep: running product of the denominators, parked per slot.  The x coordinates of the next slot are
  // requested before the current product is issued (the chain run *= d is serial; the gathers are not). ----
  F run = F::one();
  unsigned i1 = 0;
  bool has2 = false;
  F nx1 = F::zero(), nx2 = F::zero();
  if (s0 < s1) {
    i1 = ib + 2 * (s0 - ob);
    has2 = i1 + 1 < ie;
    if (has2) { nx1 = aff_load_x<C, FIRST>(table, sorted, in_pts, i1); nx2 = aff_load_x<C, FIRST>(table, sorted, in_pts, i1 + 1); }
  }
  for (unsigned s = s0; s < s1; ++s) {
    const unsigned ci1 = i1;
    const bool chas2 = has2;
    const F x1 = nx1, x2 = nx2;
    if (s + 1 < s1) {
      while (s + 1 >= oe) { ++b; ob = oe; oe = out_off[b + 1]; ib = ie; ie = in_off[b + 1]; }
      i1 = ib + 2 * (s + 1 - ob);
      has2 = i1 + 1 < ie;
      if (has2) { nx1 = aff_load_x<C, FIRST>(table, sorted, in_pts, i1); nx2 = aff_load_x<C, FIRST>(table, sorted, in_pts, i1 + 1); }
    }
    store_fp<F>(prefix, s, run);
    if (chas2) {
      F d = F::sub(x2, x1);
      if (d.is_zero() || x1.is_zero() || x2.is_zero()) {     // rare: equal x, or an operand that may be the identity
        const Affine<C> p1 = aff_load_point<C, FIRST>(table, sorted, in_pts, ci1), p2 = aff_load_point<C, FIRST>(table, sorted, in_pts, ci1 + 1);
        const int kind = aff_classify<C>(p1, p2);
        if (kind == kAffDouble) d = F::dbl(p1.y);
        else if (kind != kAffNormal) d = F::one();
      }
      run = F::mul(run, d);
    }
  }
  // ---- one inversion per CTA ----
  F inv = cta_batch_inverse<F>(run, inv_sm);
  // ---- backward sweep ----
  for (unsigned s = s1; s-- > s0;) {
    while (s < ob) { --b; oe = ob; ob = out_off[b]; ie = ib; ib = in_off[b]; }
    const unsigned i1 = ib + 2 * (s - ob);
    const Affine<C> p1 = aff_load_point<C, FIRST>(table, sorted, in_pts, i1);
    Affine<C> o = p1;                                         // odd tail: carried (curve_summations.rs:146-148)
    if (i1 + 1 < ie) {
      const Affine<C> p2 = aff_load_point<C, FIRST>(table, sorted, in_pts, i1 + 1);
      const int kind = aff_classify<C>(p1, p2);
      if (kind == kAffCopy2) o = p2;
      else if (kind == kAffZero) o = Affine<C>::identity();
      else if (kind != kAffCopy1) {
        const F d = kind == kAffNormal ? F::sub(p2.x, p1.x) : F::dbl(p1.y);
        const F dinv = F::mul(inv, load_fp<F>(prefix, s));    // 1 / d
        inv = F::mul(inv, d);
        F num;
        if (kind == kAffNormal) num = F::sub(p2.y, p1.y);
        else { const F xx = F::sqr(p1.x); num = F::add(F::dbl(xx), xx); }     // 3 x^2 (a = 0 on every supported curve)
        const F lam = F::mul(num, dinv);
        o.x = F::sub(F::sub(F::sqr(lam), p1.x), p2.x);
        o.y = F::sub(F::mul(lam, F::sub(p1.x, o.x)), p1.y);
      }
    }
    store_fp<F>(out_pts, 2 * (size_t)s, o.x);
    store_fp<F>(out_pts, 2 * (size_t)s + 1, o.y);
  }
}

// XYZZ task kernel over a list of affine POINTS (the output of the last affine round) instead of table slots
template <class C>
__global__ void __launch_bounds__(kAccThreads) msm_accumulate_points_kernel(const void* __restrict__ pts, const unsigned* __restrict__ offsets,
                                                                            const unsigned* __restrict__ task_off, unsigned nb, unsigned task,
                                                                            void* __restrict__ partials) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = task_off[nb];
  if (t >= total) return;
  unsigned lo = 0, hi = nb;
  while (hi - lo > 1) {
    unsigned mid = (lo + hi) >> 1;
    if (task_off[mid] <= t) lo = mid; else hi = mid;
  }
  const unsigned b = lo;
  const unsigned start = offsets[b] + (t - task_off[b]) * task;
  unsigned end = offsets[b + 1];
  if (end > start + task) end = start + task;
  XYZZ<C> acc = XYZZ<C>::identity();
  Affine<C> next = load_affine<C>(pts, start);
  for (unsigned k = start; k < end; ++k) {
    const Affine<C> p = next;
    if (k + 1 < end) next = load_affine<C>(pts, k + 1);
    acc = XYZZ<C>::madd(acc, p);
  }
  store_xyzz<C>(partials, t, acc);
}

}  // namespace plk
