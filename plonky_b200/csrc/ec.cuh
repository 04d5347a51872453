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
    F v = F::sqr(u);
    F w = F::mul(u, v);
    F s = F::mul(p.x, v);
    F xx = F::sqr(p.x);
    F m = F::add(F::dbl(xx), xx);           // a = 0
    XYZZ r;
    r.x = F::sub(F::sqr(m), F::dbl(s));
    r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
    r.zz = F::mul(v, p.zz);
    r.zzz = F::mul(w, p.zzz);
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
  // a + b (value of curve_adds.rs:5-48)
  PLK_HD_NOINLINE static XYZZ add(const XYZZ& a, const XYZZ& b) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    F u1 = F::mul(a.x, b.zz);
    F u2 = F::mul(b.x, a.zz);
    F s1 = F::mul(a.y, b.zzz);
    F s2 = F::mul(b.y, a.zzz);
    F p = F::sub(u2, u1);
    F r = F::sub(s2, s1);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl(a);
      return identity();
    }
    F pp = F::sqr(p);
    F ppp = F::mul(p, pp);
    F qq = F::mul(u1, pp);
    XYZZ o;
    o.x = F::sub(F::sub(F::sqr(r), ppp), F::dbl(qq));
    o.y = F::sub(F::mul(r, F::sub(qq, o.x)), F::mul(s1, ppp));
    o.zz = F::mul(F::mul(a.zz, b.zz), pp);
    o.zzz = F::mul(F::mul(a.zzz, b.zzz), ppp);
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
// The MSM's reduction tails (bucket sums, running sums, final tree) are chains of a few dozen DEPENDENT
// point additions executed by few threads: they are bound by the latency of one thread's 14 sequential
// Montgomery products per addition, not by throughput.  Here the four lanes of a quad hold identical
// copies of the operands and each computes a different product of the same dependency level; the
// products are exchanged with shuffles.  An addition becomes 4 product rounds instead of 14 products,
// a doubling 3 rounds instead of 9.  Every lane returns the full result.  All lanes of a quad must call
// with identical point arguments; `ql` = lane within the quad, `qmask` = the quad's 4-bit lane mask.
template <class C>
struct QuadXYZZ {
  typedef Fp<typename C::Base> F;
  typedef XYZZ<C> P;

  static __device__ __forceinline__ F sel(int ql, const F& a0, const F& a1, const F& a2, const F& a3) {
    F r;
#pragma unroll
    for (int k = 0; k < F::N; ++k) r.l[k] = ql == 0 ? a0.l[k] : (ql == 1 ? a1.l[k] : (ql == 2 ? a2.l[k] : a3.l[k]));
    return r;
  }
  static __device__ __forceinline__ void gather(const F& mine, F (&all)[4], unsigned qmask) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < F::N; ++k) all[i].l[k] = __shfl_sync(qmask, mine.l[k], i, 4);
  }
  static __device__ __noinline__ P dbl(const P& a, int ql, unsigned qmask) {
    if (a.is_identity() || a.y.is_zero()) return P::identity();
    F g[4];
    const F u = F::dbl(a.y);
    gather(F::mul(sel(ql, u, a.x, u, a.x), sel(ql, u, a.x, u, a.x)), g, qmask);          // v = u^2 | xx = x^2
    const F v = g[0], xx = g[1];
    const F m = F::add(F::dbl(xx), xx);                                                    // 3 x^2 (a = 0)
    gather(F::mul(sel(ql, u, a.x, m, v), sel(ql, v, v, m, a.zz)), g, qmask);             // w | s | m^2 | zz3
    const F w = g[0], s = g[1], mm = g[2];
    P r;
    r.zz = g[3];
    r.x = F::sub(mm, F::dbl(s));
    gather(F::mul(sel(ql, w, w, m, m), sel(ql, a.y, a.zzz, F::sub(s, r.x), F::sub(s, r.x))), g, qmask);   // w y | zzz3 | m (s - x3)
    r.zzz = g[1];
    r.y = F::sub(g[2], g[0]);
    return r;
  }
  static __device__ __noinline__ P add(const P& a, const P& b, int ql, unsigned qmask) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    F g[4];
    gather(F::mul(sel(ql, a.x, b.x, a.y, b.y), sel(ql, b.zz, a.zz, b.zzz, a.zzz)), g, qmask);   // u1 | u2 | s1 | s2
    const F u1 = g[0], s1 = g[2];
    const F p = F::sub(g[1], g[0]);
    const F r = F::sub(g[3], g[2]);
    if (p.is_zero()) {
      if (r.is_zero()) return dbl(a, ql, qmask);
      return P::identity();
    }
    gather(F::mul(sel(ql, p, r, a.zz, a.zzz), sel(ql, p, r, b.zz, b.zzz)), g, qmask);            // pp | rr | zz1 zz2 | zzz1 zzz2
    const F pp = g[0], rr = g[1], zz12 = g[2], zzz12 = g[3];
    gather(F::mul(sel(ql, p, u1, zz12, zz12), pp), g, qmask);                                    // ppp | q | zz3
    const F ppp = g[0], qq = g[1];
    P o;
    o.zz = g[2];
    o.x = F::sub(F::sub(rr, ppp), F::dbl(qq));
    gather(F::mul(sel(ql, r, s1, zzz12, zzz12), sel(ql, F::sub(qq, o.x), ppp, ppp, ppp)), g, qmask);   // r (q - x3) | s1 ppp | zzz3
    o.y = F::sub(g[0], g[1]);
    o.zzz = g[2];
    return o;
  }
  static __device__ P mul_u64(const P& p, uint64_t k, int ql, unsigned qmask) {
    P acc = P::identity();
    int top = 63;
    while (top >= 0 && !((k >> top) & 1)) --top;
    for (int i = top; i >= 0; --i) {
      if (i != top) acc = dbl(acc, ql, qmask);
      if ((k >> i) & 1) acc = add(acc, p, ql, qmask);
    }
    return acc;
  }
};
#endif

}  // namespace plk
