// Halo inner-product-argument rounds on device (src/halo.rs:63-124 of the reference).
//
// The reference keeps halo_a, halo_b (scalar-field vectors) and halo_g (points) on the host and, per round j:
//   L_j = msm_parallel(a_lo, G_hi, 8) + blinding + [<a_lo, b_hi>] U'            (halo.rs:87-89)
//   R_j = msm_parallel(a_hi, G_lo, 8) + blinding + [<a_hi, b_lo>] U'            (halo.rs:91-93)
//   a' = u^-1 a_hi + u a_lo,  b' = u^-1 b_lo + u b_hi                           (halo.rs:117-118)
//   G'_i = msm_parallel([u^-1, u], [G_lo_i, G_hi_i], 4)                         (halo.rs:119-123)
// Here the three vectors stay in HBM for all log2(n) rounds; only the two points, the two inner products and the
// challenge cross PCIe.  The challenger (Rescue sponge) and the blinding / U' terms stay with the caller.
//
// Kernels: ipa_inner_kernel (+ _final) both inner products in one pass; ipa_fold_scalars_kernel; ipa_fold_points_kernel
// (u^-1 P + u Q by Shamir's trick -- every thread shares the two scalars, so the bit pattern is warp-uniform -- and
// one Fermat inversion to keep G affine for the next round's bucket MSM).  The MSMs are msm_variable_dev.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace plk {

constexpr int kIpaThreads = 256;

template <class SF>
__device__ __forceinline__ void ipa_block_reduce2(SF& s1, SF& s2, uint4* sm) {
  store_fp<SF>(sm, 2 * threadIdx.x, s1);
  store_fp<SF>(sm, 2 * threadIdx.x + 1, s2);
  __syncthreads();
  for (unsigned d = blockDim.x >> 1; d > 0; d >>= 1) {
    if (threadIdx.x < d) {
      s1 = SF::add(load_fp<SF>(sm, 2 * threadIdx.x), load_fp<SF>(sm, 2 * (threadIdx.x + d)));
      s2 = SF::add(load_fp<SF>(sm, 2 * threadIdx.x + 1), load_fp<SF>(sm, 2 * (threadIdx.x + d) + 1));
      store_fp<SF>(sm, 2 * threadIdx.x, s1);
      store_fp<SF>(sm, 2 * threadIdx.x + 1, s2);
    }
    __syncthreads();
  }
}
// partials[2 k] = sum over CTA k's slice of a[i] b[half + i], partials[2 k + 1] = ... a[half + i] b[i]   (field.rs:214-221)
template <class C>
__global__ void __launch_bounds__(kIpaThreads) ipa_inner_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                                unsigned long long half, void* __restrict__ partials) {
  typedef Fp<typename C::Scalar> SF;
  __shared__ uint4 sm[2 * kIpaThreads * (SF::N / 4)];
  SF s1 = SF::zero(), s2 = SF::zero();
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < half;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    s1 = SF::add(s1, SF::mul(load_fp<SF>(a, i), load_fp<SF>(b, half + i)));
    s2 = SF::add(s2, SF::mul(load_fp<SF>(a, half + i), load_fp<SF>(b, i)));
  }
  ipa_block_reduce2<SF>(s1, s2, sm);
  if (threadIdx.x == 0) {
    store_fp<SF>(partials, 2 * blockIdx.x, s1);
    store_fp<SF>(partials, 2 * blockIdx.x + 1, s2);
  }
}
template <class C>
__global__ void __launch_bounds__(kIpaThreads) ipa_inner_final_kernel(const void* __restrict__ partials, unsigned count,
                                                                      void* __restrict__ out2) {
  typedef Fp<typename C::Scalar> SF;
  __shared__ uint4 sm[2 * kIpaThreads * (SF::N / 4)];
  SF s1 = SF::zero(), s2 = SF::zero();
  for (unsigned i = threadIdx.x; i < count; i += blockDim.x) {
    s1 = SF::add(s1, load_fp<SF>(partials, 2 * i));
    s2 = SF::add(s2, load_fp<SF>(partials, 2 * i + 1));
  }
  ipa_block_reduce2<SF>(s1, s2, sm);
  if (threadIdx.x == 0) {
    store_fp<SF>(out2, 0, s1);
    store_fp<SF>(out2, 1, s2);
  }
}
// in place on the low halves: a[i] <- u^-1 a[half + i] + u a[i], b[i] <- u^-1 b[i] + u b[half + i];  uu = (u, u^-1)
template <class C>
__global__ void ipa_fold_scalars_kernel(void* __restrict__ a, void* __restrict__ b, unsigned long long half, const void* __restrict__ uu) {
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= half) return;
  const SF u = load_fp<SF>(uu, 0), u_inv = load_fp<SF>(uu, 1);
  const SF a_lo = load_fp<SF>(a, i), a_hi = load_fp<SF>(a, half + i), b_lo = load_fp<SF>(b, i), b_hi = load_fp<SF>(b, half + i);
  store_fp<SF>(a, i, SF::add(SF::mul(u_inv, a_hi), SF::mul(u, a_lo)));
  store_fp<SF>(b, i, SF::add(SF::mul(u_inv, b_lo), SF::mul(u, b_hi)));
}
// in place on the low half: G[i] <- u^-1 G[i] + u G[half + i], normalised to affine (identity = (0, 0))
template <class C>
__global__ void __launch_bounds__(128) ipa_fold_points_kernel(void* __restrict__ g, unsigned long long half, const void* __restrict__ uu) {
  typedef Fp<typename C::Base> F;
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= half) return;
  const SF u = SF::to_canonical(load_fp<SF>(uu, 0)), u_inv = SF::to_canonical(load_fp<SF>(uu, 1));
  Affine<C> lo, hi;
  lo.x = load_fp<F>(g, 2 * i);
  lo.y = load_fp<F>(g, 2 * i + 1);
  hi.x = load_fp<F>(g, 2 * (half + i));
  hi.y = load_fp<F>(g, 2 * (half + i) + 1);
  const XYZZ<C> both = XYZZ<C>::madd(XYZZ<C>::from_affine(lo), hi);        // G_lo + G_hi
  XYZZ<C> acc = XYZZ<C>::identity();
  for (int bit = C::Scalar::BITS - 1; bit >= 0; --bit) {
    acc = XYZZ<C>::dbl(acc);
    const unsigned sel = ((u_inv.l[bit >> 5] >> (bit & 31)) & 1u) | (((u.l[bit >> 5] >> (bit & 31)) & 1u) << 1);   // warp-uniform
    if (sel == 1) acc = XYZZ<C>::madd(acc, lo);
    else if (sel == 2) acc = XYZZ<C>::madd(acc, hi);
    else if (sel == 3) acc = XYZZ<C>::add(acc, both);
  }
  const Affine<C> r = XYZZ<C>::to_affine(acc);
  store_fp<F>(g, 2 * i, r.x);
  store_fp<F>(g, 2 * i + 1, r.y);
}

// ---- table mode: G is never folded ------------------------------------------------------------------------------
// With the fixed-base table of the ORIGINAL generators at hand (the prover holds pedersen_g_msm_precomputation,
// src/plonk.rs:64-69), folding G costs more than it saves: G^(j)[k] = sum over i == k (mod m) of c_i G_i with
// c_i = prod over the rounds so far of (bit of i selected by that round ? u : u^-1), so
//   <a_lo, G^(j)_hi> = sum over i with bit (m/2) set    of (a[i & (m/2 - 1)]       c_i) G_i
//   <a_hi, G^(j)_lo> = sum over i with bit (m/2) clear  of (a[m/2 + (i & (m/2 - 1))] c_i) G_i
// are two fixed-base MSMs over all n generators (half of the scalars are zero and cost nothing after the digit
// pass), and the final halo_g is the MSM of the c_i.  Same group elements as halo.rs:87-123, no per-round chain of
// 255 doublings per generator.
// sl_sr: 2 * n0 scalars (row 0 = L, row 1 = R); m = current length of a
template <class C>
__global__ void ipa_expand_scalars_kernel(const void* __restrict__ a, const void* __restrict__ coef, unsigned long long n0,
                                          unsigned long long m, void* __restrict__ sl_sr) {
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n0) return;
  const unsigned long long half = m >> 1, k = i & (half - 1);
  const bool hi = (i & half) != 0;
  const SF c = load_fp<SF>(coef, i);
  const SF v = SF::mul(load_fp<SF>(a, hi ? k : half + k), c);
  store_fp<SF>(sl_sr, i, hi ? v : SF::zero());
  store_fp<SF>(sl_sr, n0 + i, hi ? SF::zero() : v);
}
template <class C>
__global__ void ipa_init_coef_kernel(void* __restrict__ coef, unsigned long long n0) {
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n0) store_fp<SF>(coef, i, SF::one());
}
// fold of length m -> m/2: c_i *= (bit m/2 of i) ? u : u^-1
template <class C>
__global__ void ipa_update_coef_kernel(void* __restrict__ coef, unsigned long long n0, unsigned long long m, const void* __restrict__ uu) {
  typedef Fp<typename C::Scalar> SF;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n0) return;
  const SF f = load_fp<SF>(uu, (i & (m >> 1)) ? 0 : 1);
  store_fp<SF>(coef, i, SF::mul(load_fp<SF>(coef, i), f));
}

struct IpaOps {
  void (*inner_products)(const void* d_a, const void* d_b, size_t half, void* d_partials, void* d_out2, cudaStream_t st);
  void (*fold)(void* d_a, void* d_b, void* d_g, size_t half, const void* d_uu, cudaStream_t st);   // d_g may be NULL (table mode)
  void (*init_coef)(void* d_coef, size_t n0, cudaStream_t st);
  void (*expand_scalars)(const void* d_a, const void* d_coef, size_t n0, size_t m, void* d_sl_sr, cudaStream_t st);
  void (*update_coef)(void* d_coef, size_t n0, size_t m, const void* d_uu, cudaStream_t st);
};
constexpr unsigned kIpaMaxBlocks = 296;

template <class C>
void ipa_inner_products(const void* d_a, const void* d_b, size_t half, void* d_partials, void* d_out2, cudaStream_t st) {
  unsigned blocks = (unsigned)((half + kIpaThreads - 1) / kIpaThreads);
  if (blocks > kIpaMaxBlocks) blocks = kIpaMaxBlocks;
  ipa_inner_kernel<C><<<blocks, kIpaThreads, 0, st>>>(d_a, d_b, half, d_partials);
  PLK_LAUNCHED();
  ipa_inner_final_kernel<C><<<1, kIpaThreads, 0, st>>>(d_partials, blocks, d_out2);
  PLK_LAUNCHED();
}
template <class C>
void ipa_fold(void* d_a, void* d_b, void* d_g, size_t half, const void* d_uu, cudaStream_t st) {
  ipa_fold_scalars_kernel<C><<<(unsigned)((half + 127) / 128), 128, 0, st>>>(d_a, d_b, half, d_uu);
  PLK_LAUNCHED();
  if (!d_g) return;
  ipa_fold_points_kernel<C><<<(unsigned)((half + 127) / 128), 128, 0, st>>>(d_g, half, d_uu);
  PLK_LAUNCHED();
}
template <class C>
void ipa_init_coef(void* d_coef, size_t n0, cudaStream_t st) {
  ipa_init_coef_kernel<C><<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(d_coef, n0);
  PLK_LAUNCHED();
}
template <class C>
void ipa_expand_scalars(const void* d_a, const void* d_coef, size_t n0, size_t m, void* d_sl_sr, cudaStream_t st) {
  ipa_expand_scalars_kernel<C><<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(d_a, d_coef, n0, m, d_sl_sr);
  PLK_LAUNCHED();
}
template <class C>
void ipa_update_coef(void* d_coef, size_t n0, size_t m, const void* d_uu, cudaStream_t st) {
  ipa_update_coef_kernel<C><<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(d_coef, n0, m, d_uu);
  PLK_LAUNCHED();
}
template <class C>
const IpaOps* make_ipa_ops() {
  static const IpaOps ops = {&ipa_inner_products<C>, &ipa_fold<C>, &ipa_init_coef<C>, &ipa_expand_scalars<C>, &ipa_update_coef<C>};
  return &ops;
}
}  // namespace plk
