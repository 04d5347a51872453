// Short-Weierstrass group law for the MSM kernels (Tweedledee, Tweedledum, BLS12-377 G1; a = 0).
//
// Replaces the VALUES computed by src/curve/curve_adds.rs:5-128, src/curve/curve.rs:206-260 and the
// affine adds of src/curve/curve_summations.rs:107-141.  The reference works in homogeneous
// projective coordinates (x/z, y/z); a sum of points is a unique group element, so any coordinate
// system is admissible as long as the final point is normalised to the same affine (x, y).  The
// device accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): a mixed add costs
// 8M + 2S instead of the reference's 11 multiplications, and the identity needs no flag (ZZ = 0).
//
// Exceptional cases are handled exactly like the reference does (curve_adds.rs:12-17, :66-75):
// identity operands, P == Q (doubling) and P == -Q (identity).
//
// Affine points are stored as (x, y) in Montgomery form; the identity (AffinePoint::ZERO,
// curve.rs:81-85, `zero: true`) is encoded as x = y = 0, which is not on any supported curve (b != 0).
#pragma once
#include "fp.cuh"

namespace plk {

// The Montgomery product as a real function call.  The latency-bound reduction tails run ~1 warp per scheduler through
// straight-line code: with every product inlined one point addition is ~50 KB of instructions executed once per
// pass, and the kernels run at instruction-fetch speed (~9 cycles per instruction measured).  Calling one 3.5 KB
// body keeps the working set inside the instruction cache.  (The throughput-bound accumulate kernel keeps the
// inlined product: its many warps share every fetched line.)
template <class F>
PLK_HD_NOINLINE F fp_mul_call(const F a, const F b) { return F::mul(a, b); }
template <class F>
PLK_HD_NOINLINE F fp_sqr_call(const F a) { return F::sqr(a); }

template <class C>
struct Affine {
  typedef Fp<typename C::Base> F;
  F x, y;
  PLK_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  PLK_HD static Affine identity() { Affine a; a.x = F::zero(); a.y = F::zero(); return a; }
  PLK_HD static Affine neg(const Affine& p) { Affine r; r.x = p.x; r.y = F::neg(p.y); return r; }
};

template <class C>
struct XYZZ {
  typedef Fp<typename C::Base> F;
  F x, y, zz, zzz;

  PLK_HD bool is_identity() const { return zz.is_zero(); }
  PLK_HD static XYZZ identity() {
    XYZZ r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r;
  }
  PLK_HD static XYZZ from_affine(const Affine<C>& p) {
    if (p.is_identity()) return identity();
    XYZZ r; r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one(); return r;
  }
  // 2*(x1, y1) for a non-identity affine point with y1 != 0 (always true in an odd-order group)
  PLK_HD_NOINLINE static XYZZ dbl_affine(const Affine<C>& p) {
    if (p.is_identity() || p.y.is_zero()) return identity();
    F u = F::dbl(p.y);
    F v = F::sqr(u);
    F w = F::mul(u, v);
    F s = F::mul(p.x, v);
    F xx = F::sqr(p.x);
    F m = F::add(F::dbl(xx), xx);           // 3 x^2 (+ a*1, a = 0)
    XYZZ r;
    r.x = F::sub(F::sqr(m), F::dbl(s));
    r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
  }
  PLK_HD_NOINLINE static XYZZ dbl(const XYZZ& p) {
    if (p.is_identity() || p.y.is_zero()) return identity();
    F u = F::dbl(p.y);
    F v = fp_sqr_call<F>(u);
    F w = fp_mul_call<F>(u, v);
    F s = fp_mul_call<F>(p.x, v);
    F xx = fp_sqr_call<F>(p.x);
    F m = F::add(F::dbl(xx), xx);           // a = 0
    XYZZ r;
    r.x = F::sub(fp_sqr_call<F>(m), F::dbl(s));
    r.y = F::sub(fp_mul_call<F>(m, F::sub(s, r.x)), fp_mul_call<F>(w, p.y));
    r.zz = fp_mul_call<F>(v, p.zz);
    r.zzz = fp_mul_call<F>(w, p.zzz);
    return r;
  }
  // acc + q, q affine (value of curve_adds.rs:50-90)
  PLK_HD static XYZZ madd(const XYZZ& a, const Affine<C>& q) {
    if (q.is_identity()) return a;
    if (a.is_identity()) return from_affine(q);
    F u2 = F::mul(q.x, a.zz);
    F s2 = F::mul(q.y, a.zzz);
    F p = F::sub(u2, a.x);
    F r = F::sub(s2, a.y);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl_affine(q);
      return identity();
    }
    F pp = F::sqr(p);
    F ppp = F::mul(p, pp);
    F qq = F::mul(a.x, pp);
    XYZZ o;
    o.x = F::sub(F::sub(F::sqr(r), ppp), F::dbl(qq));
    o.y = F::sub(F::mul(r, F::sub(qq, o.x)), F::mul(a.y, ppp));
    o.zz = F::mul(a.zz, pp);
    o.zzz = F::mul(a.zzz, ppp);
    return o;
  }
  // madd with six of its ten products as calls to ONE shared body each (squaring, product): the fully inlined addition is
  // ~2500 instructions = 40 KB per loop iteration of the accumulate kernel, more than the 32 KB instruction cache level that
  // backs the SM (ncu: no_instruction 0.94 stalls per issue); this variant keeps the loop at ~28 KB.  Same values.
  // MODE 0: four products inline + six calls; 1: all ten as calls; 2: six inline (also P*PP and X1*PP) + four calls -- the
  // shipped one.  Measured accumulate times at 2^20 Tweedledee terms: fully inlined 2.519 ms, MODE 0 2.495, MODE 1 2.550,
  // MODE 2 2.456; seven or eight products inline are back at 2.483-2.486 (the loop outgrows the instruction cache again).
  template <int MODE = 0>
  PLK_HD static XYZZ madd_compact(const XYZZ& a, const Affine<C>& q) {
    constexpr bool ALL = MODE == 1, SIX = MODE == 2;
    if (q.is_identity()) return a;
    if (a.is_identity()) return from_affine(q);
    F u2 = ALL ? fp_mul_call<F>(q.x, a.zz) : F::mul(q.x, a.zz);
    F s2 = ALL ? fp_mul_call<F>(q.y, a.zzz) : F::mul(q.y, a.zzz);
    F p = F::sub(u2, a.x);
    F r = F::sub(s2, a.y);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl_affine(q);
      return identity();
    }
    F pp = fp_sqr_call<F>(p);
    F ppp = SIX ? F::mul(p, pp) : fp_mul_call<F>(p, pp);
    F qq = SIX ? F::mul(a.x, pp) : fp_mul_call<F>(a.x, pp);
    XYZZ o;
    o.x = F::sub(F::sub(fp_sqr_call<F>(r), ppp), F::dbl(qq));
    o.y = ALL ? F::sub(fp_mul_call<F>(r, F::sub(qq, o.x)), fp_mul_call<F>(a.y, ppp)) : F::sub(F::mul(r, F::sub(qq, o.x)), F::mul(a.y, ppp));
    o.zz = fp_mul_call<F>(a.zz, pp);
    o.zzz = fp_mul_call<F>(a.zzz, ppp);
    return o;
  }
  // a + b (value of curve_adds.rs:5-48)
  PLK_HD_NOINLINE static XYZZ add(const XYZZ& a, const XYZZ& b) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    F u1 = fp_mul_call<F>(a.x, b.zz);
    F u2 = fp_mul_call<F>(b.x, a.zz);
    F s1 = fp_mul_call<F>(a.y, b.zzz);
    F s2 = fp_mul_call<F>(b.y, a.zzz);
    F p = F::sub(u2, u1);
    F r = F::sub(s2, s1);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl(a);
      return identity();
    }
    F pp = fp_sqr_call<F>(p);
    F ppp = fp_mul_call<F>(p, pp);
    F qq = fp_mul_call<F>(u1, pp);
    XYZZ o;
    o.x = F::sub(F::sub(fp_sqr_call<F>(r), ppp), F::dbl(qq));
    o.y = F::sub(fp_mul_call<F>(r, F::sub(qq, o.x)), fp_mul_call<F>(s1, ppp));
    o.zz = fp_mul_call<F>(fp_mul_call<F>(a.zz, b.zz), pp);
    o.zzz = fp_mul_call<F>(fp_mul_call<F>(a.zzz, b.zzz), ppp);
    return o;
  }
  PLK_HD static XYZZ neg(const XYZZ& a) { XYZZ r = a; r.y = F::neg(a.y); return r; }
  // to_affine (curve.rs:206-214): one inversion of ZZ*ZZZ
  PLK_HD_NOINLINE static Affine<C> to_affine(const XYZZ& a) {
    if (a.is_identity()) return Affine<C>::identity();
    F j = F::inverse(F::mul(a.zz, a.zzz));
    Affine<C> r;
    r.x = F::mul(a.x, F::mul(j, a.zzz));   // X / ZZ
    r.y = F::mul(a.y, F::mul(j, a.zz));    // Y / ZZZ
    return r;
  }
  // same, with the binary-GCD inverse: for the single thread that normalises a final result
  PLK_HD_NOINLINE static Affine<C> to_affine_gcd(const XYZZ& a) {
    if (a.is_identity()) return Affine<C>::identity();
    F j = F::inverse_gcd(F::mul(a.zz, a.zzz));
    Affine<C> r;
    r.x = F::mul(a.x, F::mul(j, a.zzz));
    r.y = F::mul(a.y, F::mul(j, a.zz));
    return r;
  }
  // [k] p for a small scalar k (double-and-add, MSB first)
  PLK_HD_NOINLINE static XYZZ mul_u64(const XYZZ& p, uint64_t k) {
    XYZZ acc = identity();
    int top = 63;
    while (top >= 0 && !((k >> top) & 1)) --top;      // skip the leading zero bits
    for (int i = top; i >= 0; --i) {
      if (i != top) acc = dbl(acc);
      if ((k >> i) & 1) acc = add(acc, p);
    }
    return acc;
  }
};

#ifdef __CUDACC__
// ---- quad-cooperative group law ---------------------------------------------------------------------
// The MSM's reduction tails (running sums, final tree) are chains of a few dozen DEPENDENT point additions
// executed by few threads: they are bound by the latency of one thread's 14 sequential Montgomery products per
// addition, not by throughput.  Here the four lanes of a quad hold identical copies of the operands and each
// computes a different product of the same dependency level; an addition becomes 4 product rounds instead of 14
// products, a doubling 3 rounds instead of 9.  Every lane returns the full result.  All lanes of a quad must call
// with identical point arguments.
// The products are exchanged through a quad-private strip of shared memory (one 16/48-byte store, one __syncwarp,
// four loads per round).  Warp shuffles were the first implementation: with a quad mask that differs between the
// lanes of a warp each __shfl_sync compiles to a MATCH/WARPSYNC/BSSY sequence (~7 instructions), 128 of them per
// addition -- more instructions than the four Montgomery products themselves.
template <class C>
struct QuadXYZZ {
  typedef Fp<typename C::Base> F;
  typedef XYZZ<C> P;
  static constexpr int kVec = F::N / 4;                 // uint4 per field element
  static constexpr int kSmemVec = 2 * 4 * kVec;         // per quad: double-buffered, 4 lanes

  struct Ctx {
    uint4* qs;          // this quad's strip of shared memory (kSmemVec uint4)
    unsigned qmask;     // the quad's 4-bit lane mask
    int ql;             // lane within the quad
    int phase;          // which half of the strip the next exchange uses
  };
  // `pool` holds kSmemVec uint4 per quad of the CTA
  static __device__ __forceinline__ Ctx make_ctx(uint4* pool) {
    Ctx c;
    c.qs = pool + (threadIdx.x >> 2) * kSmemVec;
    c.qmask = 0xFu << (threadIdx.x & 28);
    c.ql = threadIdx.x & 3;
    c.phase = 0;
    return c;
  }
  // Lane ql's operand out of four, WITHOUT branches: written as ternaries on ql the compiler emits divergent
  // branches around blocks of register moves (26 % of the range kernel's stall samples were branch_resolving, plus
  // the BSSY/BSYNC/BRA instructions themselves); bit masks cannot be turned back into control flow.
  static __device__ __forceinline__ F sel(int ql, const F& a0, const F& a1, const F& a2, const F& a3) {
    const uint32_t m0 = 0u - (uint32_t)(ql == 0), m1 = 0u - (uint32_t)(ql == 1), m2 = 0u - (uint32_t)(ql == 2), m3 = 0u - (uint32_t)(ql == 3);
    F r;
#pragma unroll
    for (int k = 0; k < F::N; ++k) r.l[k] = (a0.l[k] & m0) | (a1.l[k] & m1) | (a2.l[k] & m2) | (a3.l[k] & m3);
    return r;
  }
  // two alternatives: lanes 0 and 2 take `even`, lanes 1 and 3 take `odd`
  static __device__ __forceinline__ F sel2(int ql, const F& even, const F& odd) {
    const uint32_t mo = 0u - (uint32_t)(ql & 1);
    F r;
#pragma unroll
    for (int k = 0; k < F::N; ++k) r.l[k] = (even.l[k] & ~mo) | (odd.l[k] & mo);
    return r;
  }
  // lane 0 takes x0, lane 1 x1, lanes 2 and 3 take x23
  static __device__ __forceinline__ F sel3(int ql, const F& x0, const F& x1, const F& x23) {
    const uint32_t m0 = 0u - (uint32_t)(ql == 0), m1 = 0u - (uint32_t)(ql == 1), m23 = 0u - (uint32_t)(ql >= 2);
    F r;
#pragma unroll
    for (int k = 0; k < F::N; ++k) r.l[k] = (x0.l[k] & m0) | (x1.l[k] & m1) | (x23.l[k] & m23);
    return r;
  }
  // Every lane contributes `mine` and receives all four.  Double buffering: a lane can only reach the exchange after
  // next (which reuses this half) once all four lanes have passed the next exchange's barrier, i.e. finished reading.
  static __device__ __forceinline__ void gather(Ctx& c, const F& mine, F (&all)[4]) {
    uint4* buf = c.qs + c.phase * 4 * kVec;
    store_fp<F>(buf, c.ql, mine);
    __syncwarp(c.qmask);
#pragma unroll
    for (int i = 0; i < 4; ++i) all[i] = load_fp<F>(buf, i);
    c.phase ^= 1;
  }
  static __device__ __forceinline__ P dbl(const P& a, Ctx& c) {
    if (a.is_identity() || a.y.is_zero()) return P::identity();
    const int ql = c.ql;
    F g[4];
    const F u = F::dbl(a.y);
    { const F t = sel2(ql, u, a.x); gather(c, fp_mul_call<F>(t, t), g); }          // v = u^2 | xx = x^2
    const F v = g[0], xx = g[1];
    const F m = F::add(F::dbl(xx), xx);                                                    // 3 x^2 (a = 0)
    gather(c, fp_mul_call<F>(sel(ql, u, a.x, m, v), sel(ql, v, v, m, a.zz)), g);             // w | s | m^2 | zz3
    const F w = g[0], s = g[1], mm = g[2];
    P r;
    r.zz = g[3];
    r.x = F::sub(mm, F::dbl(s));
    { const F d = F::sub(s, r.x); gather(c, fp_mul_call<F>(sel3(ql, w, w, m), sel3(ql, a.y, a.zzz, d)), g); }   // w y | zzz3 | m (s - x3)
    r.zzz = g[1];
    r.y = F::sub(g[2], g[0]);
    return r;
  }
  static __device__ __forceinline__ P add(const P& a, const P& b, Ctx& c) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    const int ql = c.ql;
    F g[4];
    gather(c, fp_mul_call<F>(sel(ql, a.x, b.x, a.y, b.y), sel(ql, b.zz, a.zz, b.zzz, a.zzz)), g);   // u1 | u2 | s1 | s2
    const F u1 = g[0], s1 = g[2];
    const F p = F::sub(g[1], g[0]);
    const F r = F::sub(g[3], g[2]);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl(a, c);
      return P::identity();
    }
    gather(c, fp_mul_call<F>(sel(ql, p, r, a.zz, a.zzz), sel(ql, p, r, b.zz, b.zzz)), g);            // pp | rr | zz1 zz2 | zzz1 zzz2
    const F pp = g[0], rr = g[1], zz12 = g[2], zzz12 = g[3];
    gather(c, fp_mul_call<F>(sel3(ql, p, u1, zz12), pp), g);                                    // ppp | q | zz3
    const F ppp = g[0], qq = g[1];
    P o;
    o.zz = g[2];
    o.x = F::sub(F::sub(rr, ppp), F::dbl(qq));
    { const F d = F::sub(qq, o.x); gather(c, fp_mul_call<F>(sel3(ql, r, s1, zzz12), sel3(ql, d, ppp, ppp)), g); }   // r (q - x3) | s1 ppp | zzz3
    o.y = F::sub(g[0], g[1]);
    o.zzz = g[2];
    return o;
  }
  static __device__ P mul_u64(const P& p, uint64_t k, Ctx& c) {
    P acc = P::identity();
    int top = 63;
    while (top >= 0 && !((k >> top) & 1)) --top;
    for (int i = top; i >= 0; --i) {
      if (i != top) acc = dbl(acc, c);
      if ((k >> i) & 1) acc = add(acc, p, c);
    }
    return acc;
  }
};
#endif

}  // namespace plk
