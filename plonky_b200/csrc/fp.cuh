// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces (value-for-value) src/field/monty.rs:38-177 and the inlined twins
// src/field/bls12_377_base.rs:58-98, src/field/bls12_377_scalar.rs:53-93 of the reference: every
// operation returns the unique fully reduced Montgomery representative in [0, p), R = 2^(32*N).
//
// Blackwell has no 64x64 multiplier: a "4x64-bit" field element is 8 x u32 here (12 for the
// 377-bit base field).  The product uses two interleaved carry chains of mad.lo.cc/madc.hi.cc pairs
// (ptxas fuses each pair into one IMAD.WIDE.U32.X, 64 lanes/clk/SM on the FMA pipe) over "even" and
// "odd" column accumulators, with the Montgomery reduction interleaved limb by limb (CIOS).  All
// four moduli are 1 mod 2^32, so the quotient digit is just -t0 (no multiply), and the Tweedle
// moduli 2^254 + c have three zero limbs whose products are skipped at compile time.
//
// The helpers are __host__ __device__: on the host the PTX carry flag is emulated so that the SAME
// source can be unit-tested without a GPU (tests/test_fp_host.py); the product path never runs it
// on the host.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#endif
#include "field_constants.cuh"

#define PLK_HD __host__ __device__ __forceinline__
// Cold paths (inversion, full group additions in the reduction tails) are real function calls: inlining
// dozens of 150-instruction Montgomery products per call site buys nothing there and costs minutes of ptxas.
#define PLK_HD_NOINLINE __host__ __device__ __noinline__

namespace plk {
// -p^-1 mod 2^32 is 0xffffffff for all four moduli (each is 1 mod 2^32).  It is deliberately read
// from constant memory instead of being an immediate: when ptxas can see that the quotient digit is
// just -t0 it rewrites the following mad.lo.cc/madc.hi.cc chain algebraically and no longer fuses
// the pairs into IMAD.WIDE.U32.X (measured: 189 vs 140 FMA-pipe instructions per product).
#ifdef __CUDACC__
static __constant__ uint32_t kMontMu32 = 0xffffffffu;
#endif
// Two measured alternatives for the reduction rows, both off (tools/bench_mul.cu on B200, products/s, Tweedle | 377-bit):
//   baseline 7.34e10 | 2.96e10;  PLK_MOD_LIMB_ONE_AS_ADD (modulus limb 0 == 1 as two additions instead of IMAD + IMAD.HI)
//   7.39e10 | 3.05e10, squarings 8.33 -> 8.23e10;  + PLK_MU_AS_NEG (quotient digit as an opaque negation) 7.36e10 | 3.08e10.
// ptxas re-balances by itself: the additions it is handed come back as IMAD.X / IMAD.MOV on the FMA pipe (16 -> 14 narrow
// IMADs per product instead of 16 -> 0), so the gain stays inside the noise for the 8-limb fields.
#ifndef PLK_MU_AS_NEG
#define PLK_MU_AS_NEG 0
#endif
#ifndef PLK_MOD_LIMB_ONE_AS_ADD
#define PLK_MOD_LIMB_ONE_AS_ADD 0
#endif
namespace ptx {
#ifdef __CUDA_ARCH__
PLK_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PLK_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLK_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLK_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PLK_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code register (unit tests only).
static thread_local uint32_t CF = 0;
PLK_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
PLK_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
PLK_HD uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
PLK_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; CF = (uint32_t)(t >> 63); return (uint32_t)t; }
PLK_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF; CF = (uint32_t)(t >> 63); return (uint32_t)t; }
PLK_HD uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
PLK_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
PLK_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
PLK_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + CF; CF = (uint32_t)(t >> 32); return (uint32_t)t; }
PLK_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c + CF; }
#endif
}  // namespace ptx

template <class P>
struct Fp {
  static constexpr int N = P::LIMBS;
  typedef P Params;
  uint32_t l[N];

  PLK_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = 0;
    return r;
  }
  PLK_HD static Fp one() {   // R mod p (monty.rs:36 ONE = R)
    Fp r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = P::one(i);
    return r;
  }
  PLK_HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = P::r2(i);
    return r;
  }
  PLK_HD static Fp modulus() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = P::mod(i);
    return r;
  }
  PLK_HD bool is_zero() const {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc |= l[i];
    return acc == 0;
  }
  PLK_HD bool operator==(const Fp& o) const {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc |= l[i] ^ o.l[i];
    return acc == 0;
  }
  PLK_HD bool operator!=(const Fp& o) const { return !(*this == o); }

  // r = (a >= p) ? a - p : a, for a < 2p that fits N limbs (true for every modulus here: 2p < 2^(32N)).
  PLK_HD static Fp reduce_once(const Fp& a) {
    Fp t;
    t.l[0] = ptx::sub_cc(a.l[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; ++i) t.l[i] = ptx::subc_cc(a.l[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);   // 0 or 0xffffffff
    Fp r;
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = borrow ? a.l[i] : t.l[i];
    return r;
  }
  // monty.rs:38-46
  PLK_HD static Fp add(const Fp& a, const Fp& b) {
    Fp s;
    s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    s.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);
    return reduce_once(s);
  }
  // monty.rs:48-56 (value: a - b mod p)
  PLK_HD static Fp sub(const Fp& a, const Fp& b) {
    Fp d;
    d.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) d.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx::subc(0, 0);   // all ones when a < b
    Fp r;
    r.l[0] = ptx::add_cc(d.l[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) r.l[i] = ptx::addc_cc(d.l[i], P::mod(i) & borrow);
    r.l[N - 1] = ptx::addc(d.l[N - 1], P::mod(N - 1) & borrow);
    return r;
  }
  // monty.rs:58-64
  PLK_HD static Fp neg(const Fp& a) {
    Fp r;
    r.l[0] = ptx::sub_cc(P::mod(0), a.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) r.l[i] = ptx::subc_cc(P::mod(i), a.l[i]);
    r.l[N - 1] = ptx::subc(P::mod(N - 1), a.l[N - 1]);
    bool z = a.is_zero();
#pragma unroll
    for (int i = 0; i < N; ++i) r.l[i] = z ? 0u : r.l[i];
    return r;
  }
  PLK_HD static Fp dbl(const Fp& a) { return add(a, a); }

  // One carry chain: acc[0..N-1] += (x[start], x[start+2], ...) * y, product j on columns (j, j+1).
  // CIN: the chain continues a pending carry (CC.CF); COUT: the carry-out is added to acc[N].
  template <bool CIN, bool COUT, int LEN>
  PLK_HD static void mad_chain(uint32_t (&acc)[LEN], const uint32_t (&x)[N], int start, uint32_t y) {
    acc[0] = CIN ? ptx::madc_lo_cc(x[start], y, acc[0]) : ptx::mad_lo_cc(x[start], y, acc[0]);
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      if (j > 0) acc[j] = ptx::madc_lo_cc(x[start + j], y, acc[j]);
      if (j + 2 < N || COUT) acc[j + 1] = ptx::madc_hi_cc(x[start + j], y, acc[j + 1]);
      else acc[j + 1] = ptx::madc_hi(x[start + j], y, acc[j + 1]);
    }
    if (COUT) acc[N] = ptx::addc(acc[N], 0);
  }
  // Same with x = the modulus: limbs are compile-time constants, zero limbs only propagate the carry.
  template <bool COUT, int LEN>
  PLK_HD static void mad_chain_mod(uint32_t (&acc)[LEN], int start, uint32_t y) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const bool last = !(j + 2 < N || COUT);
      if (PLK_MOD_LIMB_ONE_AS_ADD && P::mod(start + j) == 1) {
        // limb == 1: the 64-bit product is just y -- two additions on the ALU pipe instead of an IMAD / IMAD.HI pair
        acc[j] = (j == 0) ? ptx::add_cc(acc[j], y) : ptx::addc_cc(acc[j], y);
        acc[j + 1] = last ? ptx::addc(acc[j + 1], 0) : ptx::addc_cc(acc[j + 1], 0);
      } else if (P::mod(start + j) != 0) {
        acc[j] = (j == 0) ? ptx::mad_lo_cc(P::mod(start + j), y, acc[j]) : ptx::madc_lo_cc(P::mod(start + j), y, acc[j]);
        acc[j + 1] = last ? ptx::madc_hi(P::mod(start + j), y, acc[j + 1]) : ptx::madc_hi_cc(P::mod(start + j), y, acc[j + 1]);
      } else {
        acc[j] = (j == 0) ? ptx::add_cc(acc[j], 0) : ptx::addc_cc(acc[j], 0);
        acc[j + 1] = last ? ptx::addc(acc[j + 1], 0) : ptx::addc_cc(acc[j + 1], 0);
      }
    }
    if (COUT) acc[N] = ptx::addc(acc[N], 0);
  }

  // Montgomery product a*b*R^-1 mod p, interleaved limb by limb (CIOS) like monty.rs:66-107.
  // Running value T = E + 2^32*O + pend, where E collects the 64-bit products that start on even
  // columns and O those that start on odd columns (two independent carry chains, no carry
  // hand-off between neighbouring products).  After each reduction step T is divided by 2^32: O
  // becomes the new E, E >> 64 the new O, and the single limb E[1] stays "pending" on column 0;
  // it is folded in with one add.cc whose carry enters the next odd chain (column 1).
  // Bounds: T < 2p < 2^(32N) at step boundaries, T < 2^(32(N+1)) inside a step, hence O < 2^(32N)
  // always (no carry out of the odd chains) and E[N] is only transiently non-zero.
  PLK_HD static Fp mul(const Fp& a, const Fp& b) {
    uint32_t E[N + 1], O[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { E[i] = 0; O[i] = 0; }
    E[N] = 0;
    uint32_t pend = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i == 0) {
        mad_chain<false, false, N>(O, a.l, 1, b.l[i]);
      } else {
        E[0] = ptx::add_cc(E[0], pend);
        mad_chain<true, false, N>(O, a.l, 1, b.l[i]);
      }
      mad_chain<false, true, N + 1>(E, a.l, 0, b.l[i]);
      static_assert(P::MU32 == 0xffffffffu, "kMontMu32 assumes p == 1 mod 2^32");
#if defined(__CUDA_ARCH__) && PLK_MU_AS_NEG
      uint32_t m;
      asm volatile("sub.u32 %0, 0, %1;" : "=r"(m) : "r"(E[0]));
#elif defined(__CUDA_ARCH__)
      uint32_t m = E[0] * kMontMu32;        // quotient digit
#else
      uint32_t m = E[0] * P::MU32;
#endif
      mad_chain_mod<true, N + 1>(E, 0, m);   // E[0] becomes 0
      mad_chain_mod<false, N>(O, 1, m);
      // T /= 2^32
      pend = E[1];
      uint32_t nE[N + 1], nO[N];
#pragma unroll
      for (int k = 0; k < N; ++k) nE[k] = O[k];
      nE[N] = 0;
#pragma unroll
      for (int k = 0; k + 2 <= N; ++k) nO[k] = E[k + 2];
      nO[N - 1] = 0;
#pragma unroll
      for (int k = 0; k < N; ++k) { E[k] = nE[k]; O[k] = nO[k]; }
      E[N] = 0;
    }
    Fp t;
    t.l[0] = ptx::add_cc(E[0], pend);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) t.l[k] = ptx::addc_cc(E[k], O[k - 1]);
    t.l[N - 1] = ptx::addc(E[N - 1], O[N - 2]);
    return reduce_once(t);
  }
  // Row i of the upper-triangular squaring: like mad_chain, but operand limbs t = start + j below `from` are
  // compile-time zeros (no product; a CIN chain still has to carry through them).
  template <bool CIN, bool COUT, int LEN>
  PLK_HD static void mad_chain_from(uint32_t (&acc)[LEN], const uint32_t (&x)[N], int start, int from, uint32_t y) {
    bool started = CIN;
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const bool last = !(j + 2 < N || COUT);
      if (start + j < from) {
        if (started) {
          acc[j] = ptx::addc_cc(acc[j], 0);
          acc[j + 1] = last ? ptx::addc(acc[j + 1], 0) : ptx::addc_cc(acc[j + 1], 0);
        }
      } else {
        acc[j] = started ? ptx::madc_lo_cc(x[start + j], y, acc[j]) : ptx::mad_lo_cc(x[start + j], y, acc[j]);
        acc[j + 1] = last ? ptx::madc_hi(x[start + j], y, acc[j + 1]) : ptx::madc_hi_cc(x[start + j], y, acc[j + 1]);
        started = true;
      }
    }
    if (COUT && started) acc[N] = ptx::addc(acc[N], 0);
  }
  // monty.rs:109-160 (same value as mul(a, a)).  PLK_SQR_DEDICATED: upper-triangular rows -- row i multiplies a_i with
  // a_i and with the DOUBLED limbs above it (36 instead of 64 limb products for N = 8), same interleaved reduction.
  // Doubling limbs i+1.. as a number: d_t = (a_t << 1) | (a_(t-1) >> 31), except that limb i + 1 takes no bit from a_i.
#ifndef PLK_SQR_DEDICATED
#define PLK_SQR_DEDICATED 1
#endif
  PLK_HD static Fp sqr(const Fp& a) {
#if PLK_SQR_DEDICATED
    static_assert(N % 2 == 0, "even limb count");
    uint32_t d[N];                      // 2a, limb-wise (2a < 2^(32N) for every modulus here)
    d[0] = a.l[0] << 1;
#pragma unroll
    for (int t = 1; t < N; ++t) d[t] = (a.l[t] << 1) | (a.l[t - 1] >> 31);
    uint32_t E[N + 1], O[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { E[i] = 0; O[i] = 0; }
    E[N] = 0;
    uint32_t pend = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      uint32_t x[N];                    // operand limbs of row i: 0 below i, a_i, then the doubled limbs above
#pragma unroll
      for (int t = 0; t < N; ++t) x[t] = t < i ? 0u : (t == i ? a.l[i] : (t == i + 1 ? (d[t] & 0xfffffffeu) : d[t]));
      if (i == 0) {
        mad_chain_from<false, false, N>(O, x, 1, i, a.l[i]);
      } else {
        E[0] = ptx::add_cc(E[0], pend);
        mad_chain_from<true, false, N>(O, x, 1, i, a.l[i]);
      }
      mad_chain_from<false, true, N + 1>(E, x, 0, i, a.l[i]);
#ifdef __CUDA_ARCH__
      uint32_t m = E[0] * kMontMu32;
#else
      uint32_t m = E[0] * P::MU32;
#endif
      mad_chain_mod<true, N + 1>(E, 0, m);
      mad_chain_mod<false, N>(O, 1, m);
      pend = E[1];
      uint32_t nE[N + 1], nO[N];
#pragma unroll
      for (int k = 0; k < N; ++k) nE[k] = O[k];
      nE[N] = 0;
#pragma unroll
      for (int k = 0; k + 2 <= N; ++k) nO[k] = E[k + 2];
      nO[N - 1] = 0;
#pragma unroll
      for (int k = 0; k < N; ++k) { E[k] = nE[k]; O[k] = nO[k]; }
      E[N] = 0;
    }
    Fp t;
    t.l[0] = ptx::add_cc(E[0], pend);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) t.l[k] = ptx::addc_cc(E[k], O[k - 1]);
    t.l[N - 1] = ptx::addc(E[N - 1], O[N - 2]);
    return reduce_once(t);
#else
    return mul(a, a);
#endif
  }

  PLK_HD Fp operator+(const Fp& o) const { return add(*this, o); }
  PLK_HD Fp operator-(const Fp& o) const { return sub(*this, o); }
  PLK_HD Fp operator*(const Fp& o) const { return mul(*this, o); }

  // Montgomery form -> canonical (monty.rs:174-177 "to_monty": multiply by 1)
  PLK_HD static Fp to_canonical(const Fp& a) {
    Fp o = zero();
    o.l[0] = 1;
    return mul(a, o);
  }
  // canonical -> Montgomery form (monty.rs:169-172 "from_monty": multiply by R^2)
  PLK_HD static Fp from_canonical(const Fp& c) { return mul(c, r2()); }

  // a^e for a little-endian exponent of nbits bits (field.rs:309-330, same value)
  PLK_HD_NOINLINE static Fp pow(const Fp& a, const uint32_t* e, int nbits) {
    Fp cur = a, prod = one();
    for (int i = 0; i < nbits; ++i) {
      if ((e[i >> 5] >> (i & 31)) & 1) prod = mul(prod, cur);
      cur = sqr(cur);
    }
    return prod;
  }
  // a^-1 by Fermat (the reference uses a binary GCD, bigint_inverse.rs:6-55 + monty.rs:162-167; the
  // inverse is unique, so the representative is identical).  a must be non-zero.
  PLK_HD_NOINLINE static Fp inverse(const Fp& a) {
    uint32_t e[N];
    // e = p - 2  (p is odd and p0 >= 3 for every modulus here? p0 = 1 -> borrow) -- do it generally
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      uint64_t t = (uint64_t)P::mod(i) - (i == 0 ? 2u : 0u) - borrow;
      e[i] = (uint32_t)t;
      borrow = (uint32_t)(t >> 63);
    }
    Fp cur = a, prod = one();
    for (int i = 0; i < P::BITS; ++i) {
      if ((e[i >> 5] >> (i & 31)) & 1) prod = mul(prod, cur);
      cur = sqr(cur);
    }
    return prod;
  }

  // a^-1 by the binary extended GCD -- the reference's own algorithm (bigint_inverse.rs:6-55 followed by the
  // product with R^3, monty.rs:162-167).  Data-dependent control flow: used where ONE thread normalises a
  // result (the Fermat ladder above costs ~380 dependent products there); a must be non-zero.
  PLK_HD_NOINLINE static Fp inverse_gcd(const Fp& a) {
    if (a.is_zero()) return a;
    uint32_t u[N], v[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { u[i] = a.l[i]; v[i] = P::mod(i); b[i] = 0; c[i] = 0; }
    b[0] = 1;
    // (every loop below has a compile-time trip count and is unrolled, so u, v, b, c stay in registers)
    auto is_one = [](const uint32_t (&x)[N]) {
      uint32_t acc = x[0] ^ 1u;
#pragma unroll
      for (int i = 1; i < N; ++i) acc |= x[i];
      return acc == 0;
    };
    auto shr = [](uint32_t (&x)[N], int k) {           // 1 <= k <= 31
#pragma unroll
      for (int i = 0; i < N - 1; ++i) x[i] = (x[i] >> k) | (x[i + 1] << (32 - k));
      x[N - 1] >>= k;
    };
    // x <- x / 2^k mod p for 1 <= k <= 31: add the multiple m p that clears the low k bits (p == 1 mod 2^32, so
    // m = -x mod 2^k), then shift.  x < p and m < 2^31 keep x + m p below 2^(32 N + 31): one extra limb.
    auto div_pow2 = [&shr](uint32_t (&x)[N], int k) {
      const uint32_t m = (0u - x[0]) & ((1u << k) - 1u);
      uint64_t cy = 0;
      uint32_t top;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        cy += (uint64_t)m * P::mod(i) + x[i];
        x[i] = (uint32_t)cy;
        cy >>= 32;
      }
      top = (uint32_t)cy;
      shr(x, k);
      x[N - 1] |= top << (32 - k);
    };
    auto less = [](const uint32_t (&x)[N], const uint32_t (&y)[N]) {
      bool lt = false;
#pragma unroll
      for (int i = 0; i < N; ++i) lt = (x[i] < y[i]) || (x[i] == y[i] && lt);      // most significant limb decides last
      return lt;
    };
    auto add_p = [](uint32_t (&x)[N]) {
      uint64_t cy = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) { cy += (uint64_t)x[i] + P::mod(i); x[i] = (uint32_t)cy; cy >>= 32; }
    };
    auto sub = [](uint32_t (&x)[N], const uint32_t (&y)[N]) {
      uint64_t bw = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) { uint64_t t = (uint64_t)x[i] - y[i] - bw; x[i] = (uint32_t)t; bw = (t >> 63) & 1; }
    };
    auto ctz31 = [](uint32_t w) { int k = 0; while (k < 31 && !((w >> k) & 1)) ++k; return k; };   // w != 0 in practice; capped at 31
    while (!is_one(u) && !is_one(v)) {
      // strip the trailing zero bits of u (resp. v) in one go, dividing b (resp. c) by the same power of two
      while (!(u[0] & 1)) { const int k = u[0] ? ctz31(u[0]) : 31; shr(u, k); div_pow2(b, k); }
      while (!(v[0] & 1)) { const int k = v[0] ? ctz31(v[0]) : 31; shr(v, k); div_pow2(c, k); }
      if (less(u, v)) { sub(v, u); if (less(c, b)) add_p(c); sub(c, b); }
      else { sub(u, v); if (less(b, c)) add_p(b); sub(b, c); }
    }
    Fp r, r3;
#pragma unroll
    for (int i = 0; i < N; ++i) { r.l[i] = is_one(u) ? b[i] : c[i]; r3.l[i] = P::r3(i); }
    return mul(r, r3);
  }
};

// 128-bit vectorised global/shared access of one element (32 B -> 2 x uint4, 48 B -> 3 x uint4)
template <class F>
__device__ __forceinline__ F load_fp(const void* base, size_t idx) {
  F r;
#ifdef __CUDACC__
  const uint4* p = reinterpret_cast<const uint4*>(base) + idx * (F::N / 4);
#pragma unroll
  for (int i = 0; i < F::N / 4; ++i) {
    uint4 v = p[i];
    r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
  }
#endif
  return r;
}
template <class F>
__device__ __forceinline__ void store_fp(void* base, size_t idx, const F& a) {
#ifdef __CUDACC__
  uint4* p = reinterpret_cast<uint4*>(base) + idx * (F::N / 4);
#pragma unroll
  for (int i = 0; i < F::N / 4; ++i) p[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
#endif
}

}  // namespace plk
