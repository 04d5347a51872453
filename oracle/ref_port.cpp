// CPU restatement ("port") of plonky's MSM / NTT hot path -- TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (plonky_b200/, libplonky_b200.so) never links or calls it.
//
// The reference (0xPolygonZero/plonky) is Rust and cannot be compiled in this image (no
// cargo/rustc, nightly-only crate), so this file restates its ALGORITHMS in C++17 + OpenMP
// (OpenMP stands in for rayon with the same work partitioning).  Each function cites the
// reference file:line it follows; paths are relative to the reference root.  The restatement is
// pinned against the reference's own known-answer vectors and against the big-integer oracle
// (oracle/plonky_oracle.py) by tests/test_oracle_*.py.
//
// Build: see oracle/Makefile  (g++ -O3 -march=native -fopenmp, mirroring .cargo/config:2).
#include <stdint.h>
#include <string.h>
#include <omp.h>
#include <vector>
#include <algorithm>
#include "ref_constants.h"

typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------
// src/bigint/bigint_arithmetic.rs
// ------------------------------------------------------------------------------------------
template <int N> struct Big { uint64_t l[N]; };

template <int N> static inline int big_cmp(const Big<N>& a, const Big<N>& b) {   // :11-22
  for (int i = N - 1; i >= 0; --i) {
    if (a.l[i] < b.l[i]) return -1;
    if (a.l[i] > b.l[i]) return 1;
  }
  return 0;
}
template <int N> static inline bool big_eq(const Big<N>& a, const Big<N>& b) { return big_cmp(a, b) == 0; }
template <int N> static inline Big<N> big_add(const Big<N>& a, const Big<N>& b) {   // :25-38 add_no_overflow
  Big<N> s; unsigned carry = 0;
  for (int i = 0; i < N; ++i) {
    u128 t = (u128)a.l[i] + b.l[i] + carry;
    s.l[i] = (uint64_t)t; carry = (unsigned)(t >> 64);
  }
  return s;
}
template <int N> static inline Big<N> big_sub(const Big<N>& a, const Big<N>& b) {   // :42-56
  Big<N> d; unsigned borrow = 0;
  for (int i = 0; i < N; ++i) {
    u128 t = (u128)a.l[i] - b.l[i] - borrow;
    d.l[i] = (uint64_t)t; borrow = (unsigned)((t >> 64) & 1);
  }
  return d;
}
template <int N> static inline Big<N> big_mul2(const Big<N>& x) {   // :70-80
  Big<N> r; r.l[0] = x.l[0] << 1;
  for (int i = 1; i < N; ++i) r.l[i] = (x.l[i] << 1) | (x.l[i - 1] >> 63);
  return r;
}
template <int N> static inline Big<N> big_div2(const Big<N>& x) {   // :84-91
  Big<N> r;
  for (int i = 0; i < N - 1; ++i) r.l[i] = (x.l[i] >> 1) | (x.l[i + 1] << 63);
  r.l[N - 1] = x.l[N - 1] >> 1;
  return r;
}
template <int N> static inline Big<N> big_from(const uint64_t* p) { Big<N> r; memcpy(r.l, p, 8 * N); return r; }
template <int N> static inline Big<N> big_small(uint64_t v) { Big<N> r; memset(r.l, 0, 8 * N); r.l[0] = v; return r; }

// src/bigint/bigint_inverse.rs:6-55 -- binary extended GCD on raw limbs
template <int N> static Big<N> big_inverse(const Big<N>& a, const Big<N>& order) {
  const Big<N> one = big_small<N>(1);
  Big<N> u = a, v = order, b = one, c = big_small<N>(0);
  while (!big_eq(u, one) && !big_eq(v, one)) {
    while ((u.l[0] & 1) == 0) {
      u = big_div2(u);
      if (b.l[0] & 1) b = big_add(b, order);
      b = big_div2(b);
    }
    while ((v.l[0] & 1) == 0) {
      v = big_div2(v);
      if (c.l[0] & 1) c = big_add(c, order);
      c = big_div2(c);
    }
    if (big_cmp(u, v) < 0) {
      v = big_sub(v, u);
      if (big_cmp(c, b) < 0) c = big_add(c, order);
      c = big_sub(c, b);
    } else {
      u = big_sub(u, v);
      if (big_cmp(b, c) < 0) b = big_add(b, order);
      b = big_sub(b, c);
    }
  }
  return big_eq(u, one) ? b : c;
}

// ------------------------------------------------------------------------------------------
// src/field/monty.rs (4-limb) and the inlined twins in bls12_377_{base,scalar}.rs
// ------------------------------------------------------------------------------------------
template <class F> struct Fe {
  static constexpr int N = F::N;
  Big<F::N> v;   // Montgomery form, reduced

  static Big<N> order() { return big_from<N>(F::ORDER); }
  static Fe zero() { Fe r; r.v = big_small<N>(0); return r; }
  static Fe one() { Fe r; r.v = big_from<N>(F::R); return r; }
  static Fe from_limbs(const uint64_t* p) { Fe r; r.v = big_from<N>(p); return r; }
  bool is_zero() const { for (int i = 0; i < N; ++i) if (v.l[i]) return false; return true; }
  bool operator==(const Fe& o) const { return big_eq(v, o.v); }
  bool operator!=(const Fe& o) const { return !big_eq(v, o.v); }

  Fe operator+(const Fe& o) const {            // monty.rs:38-46
    Fe r; r.v = big_add(v, o.v);
    // BLS12-377 base has 7 spare bits, Tweedle fields 1: the widened sum never overflows N limbs.
    if (big_cmp(r.v, order()) >= 0) r.v = big_sub(r.v, order());
    return r;
  }
  Fe neg() const {                             // monty.rs:58-64
    if (is_zero()) return *this;
    Fe r; r.v = big_sub(order(), v); return r;
  }
  Fe operator-(const Fe& o) const {            // monty.rs:48-56
    Fe r;
    if (big_cmp(v, o.v) < 0) r.v = big_add(v, o.neg().v); else r.v = big_sub(v, o.v);
    return r;
  }
  // Interleaved Montgomery product, monty.rs:66-107 (bls12_377_base.rs:58-98 for N = 6):
  // for each limb a[i]: c += a[i]*b; q = mu*c[0]; c += q*p; drop the (now zero) low limb.
  static Big<N> mont_mul(const Big<N>& a, const Big<N>& b) {
    uint64_t c[N + 2];
    memset(c, 0, sizeof(c));
    for (int i = 0; i < N; ++i) {
      uint64_t carry = 0;
      for (int j = 0; j < N; ++j) {
        u128 t = (u128)a.l[i] * b.l[j] + c[j] + carry;
        c[j] = (uint64_t)t; carry = (uint64_t)(t >> 64);
      }
      u128 t = (u128)c[N] + carry;
      c[N] = (uint64_t)t; c[N + 1] = (uint64_t)(t >> 64);
      uint64_t q = F::MU * c[0];
      carry = 0;
      for (int j = 0; j < N; ++j) {
        u128 t2 = (u128)q * F::ORDER[j] + c[j] + carry;
        c[j] = (uint64_t)t2; carry = (uint64_t)(t2 >> 64);
      }
      t = (u128)c[N] + carry;
      c[N] = (uint64_t)t; c[N + 1] += (uint64_t)(t >> 64);
      for (int j = 0; j <= N; ++j) c[j] = c[j + 1];   // shift right one limb (c[0] == 0)
      c[N + 1] = 0;
    }
    Big<N> r; memcpy(r.l, c, 8 * N);
    if (c[N] != 0 || big_cmp(r, order()) >= 0) r = big_sub(r, order());   // final conditional subtraction
    return r;
  }
  Fe operator*(const Fe& o) const { Fe r; r.v = mont_mul(v, o.v); return r; }
  // monty.rs:109-160 has a dedicated squaring for the Tweedle fields; its VALUE is a*a*R^-1 mod p,
  // the unique reduced representative, so the product routine restates it exactly.
  Fe square() const { return *this * *this; }
  Fe cube() const { return square() * *this; }
  Fe dbl() const {
    if (F::SHIFT_DOUBLE) {                      // bls12_377_base.rs:229-237
      Fe r; r.v = big_mul2(v);
      if (big_cmp(r.v, order()) >= 0) r.v = big_sub(r.v, order());
      return r;
    }
    return *this * from_limbs(F::TWO);         // field.rs:181-183
  }
  Fe triple() const {
    if (F::SHIFT_DOUBLE) {                      // bls12_377_base.rs:239-253
      Big<N> s = big_add(big_mul2(v), v);
      Big<N> x2 = big_from<N>(F::ORDER_X2);
      Fe r;
      if (big_cmp(s, order()) < 0) r.v = s;
      else if (big_cmp(s, x2) < 0) r.v = big_sub(s, order());
      else r.v = big_sub(s, x2);
      return r;
    }
    return *this * from_limbs(F::THREE);       // field.rs:186-188
  }
  // monty.rs:162-167: binary GCD on the Montgomery limbs, then times R^3
  Fe inverse() const { Fe r; r.v = mont_mul(big_inverse(v, order()), big_from<N>(F::R3)); return r; }
  static Fe from_canonical(const Big<N>& c) { Fe r; r.v = mont_mul(c, big_from<N>(F::R2)); return r; }   // monty.rs:169-172
  Big<N> to_canonical() const { return mont_mul(v, big_small<N>(1)); }                               // monty.rs:174-177
  static Fe from_u64(uint64_t x) { return from_canonical(big_small<N>(x)); }
  // field.rs:309-330 exp with a canonical exponent given as limbs
  Fe exp_big(const Big<N>& e) const {
    Fe cur = *this, prod = one();
    int top = -1;
    for (int i = 64 * N - 1; i >= 0; --i) if ((e.l[i / 64] >> (i % 64)) & 1) { top = i; break; }
    for (int i = 0; i <= top; ++i) {
      if ((e.l[i / 64] >> (i % 64)) & 1) prod = prod * cur;
      cur = cur.square();
    }
    return prod;
  }
  // field.rs:429-435
  static Fe primitive_root_of_unity(int n_power) {
    Fe base_root = from_limbs(F::GENERATOR).exp_big(from_limbs(F::T).to_canonical());
    Big<N> e = big_small<N>(0);
    int sh = F::TWO_ADICITY - n_power;
    e.l[sh / 64] = 1ull << (sh % 64);
    return base_root.exp_big(e);
  }
};

// field.rs:251-278 Montgomery's trick; returns false on a zero input (the reference panics, :267)
template <class F> static bool batch_inverse(const std::vector<Fe<F>>& x, std::vector<Fe<F>>& out) {
  size_t n = x.size();
  out.clear();
  if (n == 0) return true;
  std::vector<Fe<F>> a(n);
  a[0] = x[0];
  for (size_t i = 1; i < n; ++i) a[i] = a[i - 1] * x[i];
  if (a[n - 1].is_zero()) return false;
  std::vector<Fe<F>> a_inv(n);
  a_inv[n - 1] = a[n - 1].inverse();
  for (size_t i = n - 1; i-- > 0;) a_inv[i] = x[i + 1] * a_inv[i + 1];
  out.resize(n);
  out[0] = a_inv[0];
  for (size_t i = 1; i < n; ++i) out[i] = a[i - 1] * a_inv[i];
  return true;
}

// ------------------------------------------------------------------------------------------
// src/curve/curve.rs, curve_adds.rs
// ------------------------------------------------------------------------------------------
template <class C> struct Aff { Fe<typename C::Base> x, y; bool zero; };
template <class C> struct Proj { Fe<typename C::Base> x, y, z; bool zero; };

template <class C> static Aff<C> aff_zero() { Aff<C> r; r.x = r.y = Fe<typename C::Base>::zero(); r.zero = true; return r; }
template <class C> static Proj<C> proj_zero() { Proj<C> r; r.x = r.y = r.z = Fe<typename C::Base>::zero(); r.zero = true; return r; }
template <class C> static Proj<C> to_proj(const Aff<C>& a) { Proj<C> r; r.x = a.x; r.y = a.y; r.z = Fe<typename C::Base>::one(); r.zero = a.zero; return r; }
template <class C> static bool aff_eq(const Aff<C>& a, const Aff<C>& b) {      // curve.rs:153-170
  if (a.zero || b.zero) return a.zero == b.zero;
  return a.x == b.x && a.y == b.y;
}
template <class C> static Aff<C> aff_neg(const Aff<C>& a) { Aff<C> r = a; r.y = a.y.neg(); return r; }

template <class C> static Proj<C> proj_double(const Proj<C>& p) {               // curve.rs:234-260
  typedef Fe<typename C::Base> F;
  if (p.zero) return proj_zero<C>();
  F xx = p.x.square(), zz = p.z.square();
  F w = xx.triple();
  if (!C::A_IS_ZERO) w = w + F::from_limbs(C::A) * zz;
  F s = p.y.dbl() * p.z;
  F r = p.y * s;
  F rr = r.square();
  F b = (p.x + r).square() - (xx + rr);
  F h = w.square() - b.dbl();
  Proj<C> o; o.x = h * s; o.y = w * (b - h) - rr.dbl(); o.z = s.cube(); o.zero = false;
  return o;
}
template <class C> static Proj<C> proj_add(const Proj<C>& a, const Proj<C>& b) {  // curve_adds.rs:5-48
  typedef Fe<typename C::Base> F;
  if (a.zero) return b;
  if (b.zero) return a;
  F x1z2 = a.x * b.z, y1z2 = a.y * b.z, x2z1 = b.x * a.z, y2z1 = b.y * a.z;
  if (x1z2 == x2z1) {
    if (y1z2 == y2z1) return proj_double(a);
    if (y1z2 == y2z1.neg()) return proj_zero<C>();
  }
  F z1z2 = a.z * b.z;
  F u = y2z1 - y1z2, uu = u.square();
  F v = x2z1 - x1z2, vv = v.square(), vvv = v * vv;
  F r = vv * x1z2;
  F aa = uu * z1z2 - vvv - r.dbl();
  Proj<C> o; o.x = v * aa; o.y = u * (r - aa) - vvv * y1z2; o.z = vvv * z1z2; o.zero = false;
  return o;
}
template <class C> static Proj<C> proj_add_aff(const Proj<C>& a, const Aff<C>& b) {  // curve_adds.rs:50-90
  typedef Fe<typename C::Base> F;
  if (a.zero) return to_proj(b);
  if (b.zero) return a;
  F x2z1 = b.x * a.z, y2z1 = b.y * a.z;
  if (a.x == x2z1) {
    if (a.y == y2z1) return proj_double(a);
    if (a.y == y2z1.neg()) return proj_zero<C>();
  }
  F u = y2z1 - a.y, uu = u.square();
  F v = x2z1 - a.x, vv = v.square(), vvv = v * vv;
  F r = vv * a.x;
  F aa = uu * a.z - vvv - r.dbl();
  Proj<C> o; o.x = v * aa; o.y = u * (r - aa) - vvv * a.y; o.z = vvv * a.z; o.zero = false;
  return o;
}
template <class C> static Proj<C> aff_add_aff(const Aff<C>& a, const Aff<C>& b) {   // curve_adds.rs:92-128
  typedef Fe<typename C::Base> F;
  if (a.zero) return to_proj(b);
  if (b.zero) return to_proj(a);
  if (a.x == b.x) {
    if (a.y == b.y) return proj_double(to_proj(a));
    if (a.y == b.y.neg()) return proj_zero<C>();
  }
  F u = b.y - a.y, uu = u.square();
  F v = b.x - a.x, vv = v.square(), vvv = v * vv;
  F r = vv * a.x;
  F aa = uu - vvv - r.dbl();
  Proj<C> o; o.x = v * aa; o.y = u * (r - aa) - vvv * a.y; o.z = vvv; o.zero = false;
  return o;
}
template <class C> static Aff<C> to_affine(const Proj<C>& p) {                      // curve.rs:206-214
  if (p.zero) return aff_zero<C>();
  Fe<typename C::Base> zi = p.z.inverse();
  Aff<C> r; r.x = p.x * zi; r.y = p.y * zi; r.zero = false; return r;
}
template <class C> static std::vector<Aff<C>> batch_to_affine(const std::vector<Proj<C>>& ps) {   // curve.rs:216-232
  typedef Fe<typename C::Base> F;
  std::vector<F> nz; std::vector<size_t> idx(ps.size());
  for (size_t i = 0; i < ps.size(); ++i) {            // field.rs:223-249 (_opt: zeros skipped)
    if (!ps[i].z.is_zero()) { idx[i] = nz.size(); nz.push_back(ps[i].z); } else idx[i] = (size_t)-1;
  }
  std::vector<F> inv; batch_inverse(nz, inv);
  std::vector<Aff<C>> out(ps.size());
  for (size_t i = 0; i < ps.size(); ++i) {
    if (ps[i].zero) out[i] = aff_zero<C>();
    else { out[i].x = ps[i].x * inv[idx[i]]; out[i].y = ps[i].y * inv[idx[i]]; out[i].zero = false; }
  }
  return out;
}

// ------------------------------------------------------------------------------------------
// src/curve/curve_summations.rs
// ------------------------------------------------------------------------------------------
template <class C> static Proj<C> affine_summation_pairwise(const std::vector<Aff<C>>& pts) {   // :50-58
  std::vector<Proj<C>> red;
  for (size_t i = 0; i < pts.size(); i += 2) {
    if (i + 1 < pts.size()) red.push_back(aff_add_aff(pts[i], pts[i + 1])); else red.push_back(to_proj(pts[i]));
  }
  Proj<C> s = proj_zero<C>();
  for (auto& r : red) s = proj_add(s, r);
  return s;
}
template <class C> static std::vector<Proj<C>> multisum_best(std::vector<std::vector<Aff<C>>>& sums);

template <class C> static std::vector<Proj<C>> multisum_batch_inversion(std::vector<std::vector<Aff<C>>>& sums) {   // :70-158
  typedef Fe<typename C::Base> F;
  std::vector<F> to_inv;
  for (auto& s : sums) {
    size_t n = s.size(), end = n == 0 ? 0 : n - 1;
    for (size_t i = 0; i < end; i += 2) {
      const Aff<C>&p1 = s[i], &p2 = s[i + 1];
      if (p1.zero || p2.zero || aff_eq(p1, aff_neg(p2))) {
      } else if (aff_eq(p1, p2)) to_inv.push_back(p1.y.dbl());
      else to_inv.push_back(p1.x - p2.x);
    }
  }
  std::vector<F> inv; batch_inverse(to_inv, inv);
  std::vector<std::vector<Aff<C>>> all(sums.size());
  size_t k = 0;
  for (size_t si = 0; si < sums.size(); ++si) {
    auto& s = sums[si];
    size_t n = s.size(), end = n == 0 ? 0 : n - 1;
    auto& red = all[si];
    red.reserve((n + 1) / 2);
    for (size_t i = 0; i < end; i += 2) {
      const Aff<C>&p1 = s[i], &p2 = s[i + 1];
      Aff<C> sum;
      if (p1.zero) sum = p2;
      else if (p2.zero) sum = p1;
      else if (aff_eq(p1, aff_neg(p2))) sum = aff_zero<C>();
      else {
        F iv = inv[k++];
        F q;
        if (aff_eq(p1, p2)) {
          F num = p1.x.square().triple();
          if (!C::A_IS_ZERO) num = num + F::from_limbs(C::A);
          q = num * iv;
          sum.x = q.square() - p1.x.dbl();
        } else {
          q = (p1.y - p2.y) * iv;
          sum.x = q.square() - p1.x - p2.x;
        }
        sum.y = q * (p1.x - sum.x) - p1.y;
        sum.zero = false;
      }
      red.push_back(sum);
    }
    if (n % 2 == 1) red.push_back(s[n - 1]);
    std::vector<Aff<C>>().swap(s);
  }
  return multisum_best(all);
}
template <class C> static std::vector<Proj<C>> multisum_best(std::vector<std::vector<Aff<C>>>& sums) {   // :24-35
  size_t pairs = 0;
  for (auto& s : sums) pairs += s.size() / 2;
  if (pairs < 70) {
    std::vector<Proj<C>> out;
    for (auto& s : sums) out.push_back(affine_summation_pairwise(s));
    return out;
  }
  return multisum_batch_inversion(sums);
}

// ------------------------------------------------------------------------------------------
// src/curve/curve_msm.rs
// ------------------------------------------------------------------------------------------
template <class C> static std::vector<unsigned> to_digits(const Fe<typename C::Scalar>& x, int w) {   // :159-180
  const int bits = C::Scalar::BITS;
  int nd = (bits + w - 1) / w;
  auto c = x.to_canonical();
  std::vector<unsigned> d(nd);
  for (int i = 0; i < nd; ++i) {
    unsigned digit = 0;
    int hi = std::min((i + 1) * w, bits);
    for (int j = hi - 1; j >= i * w; --j) digit = (digit << 1) | (unsigned)((c.l[j / 64] >> (j % 64)) & 1);
    d[i] = digit;
  }
  return d;
}
template <class C> struct MsmPre { std::vector<std::vector<Aff<C>>> powers; int w; };

template <class C> static MsmPre<C>* msm_precompute(const std::vector<Proj<C>>& gens, int w) {   // :27-52
  auto* pre = new MsmPre<C>();
  pre->w = w;
  pre->powers.resize(gens.size());
  int digits = (C::Scalar::BITS + w - 1) / w;
#pragma omp parallel for schedule(dynamic, 16)
  for (long i = 0; i < (long)gens.size(); ++i) {
    std::vector<Proj<C>> pw; pw.reserve(digits);
    pw.push_back(gens[i]);
    for (int j = 1; j < digits; ++j) {
      Proj<C> t = pw[j - 1];
      for (int k = 0; k < w; ++k) t = proj_double(t);
      pw.push_back(t);
    }
    pre->powers[i] = batch_to_affine(pw);
  }
  return pre;
}
template <class C> static void build_occurrences(const MsmPre<C>& pre, const std::vector<Fe<typename C::Scalar>>& sc,
                                                 std::vector<std::vector<std::pair<uint32_t, uint32_t>>>& occ) {   // :117-126
  occ.assign((size_t)1 << pre.w, {});
  for (size_t i = 0; i < sc.size(); ++i) {
    auto d = to_digits<C>(sc[i], pre.w);
    for (size_t j = 0; j < d.size(); ++j) occ[d[j]].push_back({(uint32_t)i, (uint32_t)j});
  }
}
template <class C> static Proj<C> msm_execute(const MsmPre<C>& pre, const std::vector<Fe<typename C::Scalar>>& sc) {   // :63-100
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> occ;
  build_occurrences(pre, sc, occ);
  Proj<C> y = proj_zero<C>(), u = proj_zero<C>();
  for (size_t d = occ.size() - 1; d >= 1; --d) {
    for (auto& ij : occ[d]) u = proj_add_aff(u, pre.powers[ij.first][ij.second]);
    y = proj_add(y, u);
  }
  return y;
}
template <class C> static Proj<C> msm_execute_parallel(const MsmPre<C>& pre, const std::vector<Fe<typename C::Scalar>>& sc) {   // :102-157
  const size_t CHUNK = 80;   // DIGITS_PER_CHUNK, :14
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> occ;
  build_occurrences(pre, sc, occ);          // single-threaded scatter, as in the reference
  size_t base = occ.size();
  std::vector<Proj<C>> acc(base);
  long nchunks = (long)((base + CHUNK - 1) / CHUNK);
#pragma omp parallel for schedule(dynamic, 1)
  for (long ch = 0; ch < nchunks; ++ch) {
    size_t lo = ch * CHUNK, hi = std::min(base, lo + CHUNK);
    std::vector<std::vector<Aff<C>>> sums(hi - lo);
    for (size_t d = lo; d < hi; ++d) {
      sums[d - lo].reserve(occ[d].size());
      for (auto& ij : occ[d]) sums[d - lo].push_back(pre.powers[ij.first][ij.second]);
    }
    auto res = multisum_best(sums);
    for (size_t d = lo; d < hi; ++d) acc[d] = res[d - lo];
  }
  Proj<C> y = proj_zero<C>(), u = proj_zero<C>();
  for (size_t d = base - 1; d >= 1; --d) { u = proj_add(u, acc[d]); y = proj_add(y, u); }   // :149-154
  return y;
}

// src/curve/curve_multiplication.rs:20-85 -- single scalar * point, 4-bit Yao
template <class C> static Proj<C> scalar_mul(const Fe<typename C::Scalar>& s, const Proj<C>& p) {
  const int WB = 4, BASE = 16;
  int nd = (C::Scalar::BITS + WB - 1) / WB;
  std::vector<Proj<C>> pw; pw.push_back(p);
  for (int i = 1; i < nd; ++i) { Proj<C> t = pw[i - 1]; for (int j = 0; j < WB; ++j) t = proj_double(t); pw.push_back(t); }
  auto powers = batch_to_affine(pw);
  auto c = s.to_canonical();
  std::vector<unsigned> digits;
  for (int l = 0; l < C::Scalar::N; ++l) for (int j = 0; j < 16; ++j) digits.push_back((unsigned)((c.l[l] >> (4 * j)) & 15));
  Proj<C> y = proj_zero<C>(), u = proj_zero<C>();
  for (int j = BASE - 1; j >= 1; --j) {
    std::vector<std::vector<Aff<C>>> one(1);
    for (int i = 0; i < nd && i < (int)digits.size(); ++i) if ((int)digits[i] == j) one[0].push_back(powers[i]);
    u = proj_add(u, multisum_batch_inversion(one)[0]);
    y = proj_add(y, u);
  }
  return y;
}

// ------------------------------------------------------------------------------------------
// src/fft.rs
// ------------------------------------------------------------------------------------------
static inline size_t reverse_bits(size_t n, int bits) {   // :18-26
  size_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((n >> i) & 1) << (bits - 1 - i);
  return r;
}
static inline int log2_strict(size_t n) {   // util.rs:16-19 (caller turns -1 into the panic/error)
  if (n == 0 || (n & (n - 1))) return -1;
  int k = 0; while (((size_t)1 << k) < n) ++k; return k;
}
template <class T> static std::vector<T> reverse_index_bits(const std::vector<T>& a) {   // :8-16
  int k = log2_strict(a.size());
  std::vector<T> r(a.size());
  for (size_t i = 0; i < a.size(); ++i) r[i] = a[reverse_bits(i, k)];
  return r;
}
template <class F> struct FftPre { std::vector<std::vector<Fe<F>>> subgroups_rev; };

template <class F> static FftPre<F>* fft_precompute(size_t degree) {   // :47-59
  int pw = 0; while (((size_t)1 << pw) < degree) ++pw;    // log2_ceil
  auto* pre = new FftPre<F>();
  for (int i = 0; i <= pw; ++i) {
    Fe<F> g = Fe<F>::primitive_root_of_unity(i);
    std::vector<Fe<F>> sub((size_t)1 << i);
    Fe<F> cur = Fe<F>::one();
    for (size_t k = 0; k < sub.size(); ++k) { sub[k] = cur; cur = cur * g; }    // field.rs:292-300
    pre->subgroups_rev.push_back(reverse_index_bits(sub));
  }
  return pre;
}
template <class F> static std::vector<Fe<F>> fft_pow2(const std::vector<Fe<F>>& coeffs, const FftPre<F>& pre) {   // :103-156
  size_t n = coeffs.size(), half = n >> 1;
  int pw = log2_strict(n);
  std::vector<Fe<F>> ev = reverse_index_bits(coeffs);
  const long CH = 2000;                                   // par_chunks(2000), :130
  for (int i = 1; i <= pw; ++i) {
    size_t ppp = (size_t)1 << i, pairs = (size_t)1 << (i - 1);
    std::vector<Fe<F>> nw(n);                              // a fresh Vec per layer, as the reference
    long nch = (long)((half + CH - 1) / CH);
    const auto& tw = pre.subgroups_rev[i];
#pragma omp parallel for schedule(dynamic, 1)
    for (long c = 0; c < nch; ++c) {
      size_t lo = (size_t)c * CH, hi = std::min(half, lo + (size_t)CH);
      for (size_t pair = lo; pair < hi; ++pair) {
        size_t poly = pair / pairs, within = pair % pairs;
        size_t c0 = poly * ppp + within, c1 = c0 + pairs;
        Fe<F> prod = tw[within * 2] * ev[c1];
        nw[2 * pair] = ev[c0] + prod;
        nw[2 * pair + 1] = ev[c0] - prod;
      }
    }
    ev.swap(nw);
  }
  return reverse_index_bits(ev);
}
template <class F> static std::vector<Fe<F>> ifft_pow2(const std::vector<Fe<F>>& pts, const FftPre<F>& pre) {   // :82-101
  size_t n = pts.size();
  Fe<F> n_inv = Fe<F>::from_u64((uint64_t)n).inverse();
  auto r = fft_pow2(pts, pre);
  r[0] = r[0] * n_inv;
  if (n > 1) r[n / 2] = r[n / 2] * n_inv;
  for (size_t i = 1; i < n / 2; ++i) {
    size_t j = n - i;
    Fe<F> ri = r[j] * n_inv, rj = r[i] * n_inv;
    r[i] = ri; r[j] = rj;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// C ABI for ctypes (tests / bench cpu_baseline).  ids as in include/plonky_b200.h.
// ------------------------------------------------------------------------------------------
enum { OP_ADD = 0, OP_SUB, OP_MUL, OP_SQUARE, OP_NEG, OP_INVERSE, OP_TO_CANONICAL, OP_FROM_CANONICAL, OP_DOUBLE, OP_TRIPLE };

template <class F> static int field_op_t(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  const int N = F::N;
  for (size_t i = 0; i < n; ++i) {
    Fe<F> x = Fe<F>::from_limbs(a + i * N), y = b ? Fe<F>::from_limbs(b + i * N) : Fe<F>::zero(), r;
    switch (op) {
      case OP_ADD: r = x + y; break;
      case OP_SUB: r = x - y; break;
      case OP_MUL: r = x * y; break;
      case OP_SQUARE: r = x.square(); break;
      case OP_NEG: r = x.neg(); break;
      case OP_INVERSE: if (x.is_zero()) return -2; r = x.inverse(); break;
      case OP_TO_CANONICAL: r.v = x.to_canonical(); break;
      case OP_FROM_CANONICAL: r = Fe<F>::from_canonical(x.v); break;
      case OP_DOUBLE: r = x.dbl(); break;
      case OP_TRIPLE: r = x.triple(); break;
      default: return -1;
    }
    memcpy(out + i * N, r.v.l, 8 * N);
  }
  return 0;
}
template <class F> static int fft_t(const uint64_t* in, uint64_t* out, size_t n, int inverse) {
  if (log2_strict(n) < 0) return -1;
  std::vector<Fe<F>> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = Fe<F>::from_limbs(in + i * F::N);
  FftPre<F>* pre = fft_precompute<F>(n);
  auto r = inverse ? ifft_pow2(v, *pre) : fft_pow2(v, *pre);
  delete pre;
  for (size_t i = 0; i < n; ++i) memcpy(out + i * F::N, r[i].v.l, 8 * F::N);
  return 0;
}
template <class C> static std::vector<Aff<C>> load_affine(const uint64_t* xy, const uint8_t* zero, size_t n) {
  const int N = C::Base::N;
  std::vector<Aff<C>> p(n);
  for (size_t i = 0; i < n; ++i) {
    p[i].x = Fe<typename C::Base>::from_limbs(xy + (2 * i) * N);
    p[i].y = Fe<typename C::Base>::from_limbs(xy + (2 * i + 1) * N);
    p[i].zero = zero ? zero[i] != 0 : false;
    if (p[i].zero) p[i] = aff_zero<C>();
  }
  return p;
}
template <class C> static void store_affine(const Aff<C>& a, uint64_t* xy, uint8_t* zero) {
  const int N = C::Base::N;
  memcpy(xy, a.x.v.l, 8 * N); memcpy(xy + N, a.y.v.l, 8 * N); *zero = a.zero ? 1 : 0;
}
static uint64_t splitmix_hash(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

extern "C" {

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }

int ref_field_op(int fid, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  switch (fid) {
    case 0: return field_op_t<TweedledeeBase>(op, a, b, out, n);
    case 1: return field_op_t<TweedledumBase>(op, a, b, out, n);
    case 2: return field_op_t<Bls12377Scalar>(op, a, b, out, n);
    case 3: return field_op_t<Bls12377Base>(op, a, b, out, n);
  }
  return -1;
}
}  // extern "C"
template <class F> static int batch_inverse_t(const uint64_t* in, uint64_t* out, size_t n) {
  std::vector<Fe<F>> x(n), r;
  for (size_t i = 0; i < n; ++i) x[i] = Fe<F>::from_limbs(in + i * F::N);
  if (!batch_inverse(x, r)) return -2;
  for (size_t i = 0; i < n; ++i) memcpy(out + i * F::N, r[i].v.l, 8 * F::N);
  return 0;
}
template <class F> static int root_t(int k, uint64_t* out) { auto r = Fe<F>::primitive_root_of_unity(k); memcpy(out, r.v.l, 8 * F::N); return 0; }
extern "C" {
int ref_batch_inverse(int fid, const uint64_t* in, uint64_t* out, size_t n) {
  switch (fid) {
    case 0: return batch_inverse_t<TweedledeeBase>(in, out, n);
    case 1: return batch_inverse_t<TweedledumBase>(in, out, n);
    case 2: return batch_inverse_t<Bls12377Scalar>(in, out, n);
    case 3: return batch_inverse_t<Bls12377Base>(in, out, n);
  }
  return -1;
}
int ref_div2(int nlimbs, const uint64_t* in, uint64_t* out) {
  if (nlimbs == 4) { auto r = big_div2(big_from<4>(in)); memcpy(out, r.l, 32); return 0; }
  if (nlimbs == 6) { auto r = big_div2(big_from<6>(in)); memcpy(out, r.l, 48); return 0; }
  return -1;
}
uint64_t ref_reverse_bits(uint64_t n, int bits) { return reverse_bits(n, bits); }

int ref_primitive_root_of_unity(int fid, int n_power, uint64_t* out) {
  switch (fid) {
    case 0: return root_t<TweedledeeBase>(n_power, out);
    case 1: return root_t<TweedledumBase>(n_power, out);
    case 2: return root_t<Bls12377Scalar>(n_power, out);
    case 3: return root_t<Bls12377Base>(n_power, out);
  }
  return -1;
}
// forward (inverse=0) / inverse (1) power-of-two transform, src/fft.rs:103 / :82
int ref_fft(int fid, const uint64_t* in, uint64_t* out, size_t n, int inverse) {
  switch (fid) {
    case 0: return fft_t<TweedledeeBase>(in, out, n, inverse);
    case 1: return fft_t<TweedledumBase>(in, out, n, inverse);
    case 2: return fft_t<Bls12377Scalar>(in, out, n, inverse);
    case 3: return fft_t<Bls12377Base>(in, out, n, inverse);
  }
  return -1;
}
}  // extern "C"

// FFT with a persistent precomputation (the timed configuration: table built once, benches/fft.rs:20-33)
struct RefFftHandle { int fid; size_t n; void* pre; };
template <class F> static int fft_run_t(RefFftHandle* h, const uint64_t* in, uint64_t* out, int inverse) {
  std::vector<Fe<F>> v(h->n);
  memcpy((void*)v.data(), in, h->n * 8 * F::N);
  auto r = inverse ? ifft_pow2(v, *(FftPre<F>*)h->pre) : fft_pow2(v, *(FftPre<F>*)h->pre);
  memcpy(out, (void*)r.data(), h->n * 8 * F::N);
  return 0;
}
extern "C" {
void* ref_fft_precompute(int fid, size_t n) {
  if (log2_strict(n) < 0) return nullptr;
  auto* h = new RefFftHandle{fid, n, nullptr};
  switch (fid) {
    case 0: h->pre = fft_precompute<TweedledeeBase>(n); break;
    case 1: h->pre = fft_precompute<TweedledumBase>(n); break;
    case 2: h->pre = fft_precompute<Bls12377Scalar>(n); break;
    case 3: h->pre = fft_precompute<Bls12377Base>(n); break;
    default: delete h; return nullptr;
  }
  return h;
}
int ref_fft_run(void* hv, const uint64_t* in, uint64_t* out, int inverse) {
  auto* h = (RefFftHandle*)hv;
  switch (h->fid) {
    case 0: return fft_run_t<TweedledeeBase>(h, in, out, inverse);
    case 1: return fft_run_t<TweedledumBase>(h, in, out, inverse);
    case 2: return fft_run_t<Bls12377Scalar>(h, in, out, inverse);
    case 3: return fft_run_t<Bls12377Base>(h, in, out, inverse);
  }
  return -1;
}
void ref_fft_free(void* hv) {
  auto* h = (RefFftHandle*)hv;
  if (!h) return;
  switch (h->fid) {
    case 0: delete (FftPre<TweedledeeBase>*)h->pre; break;
    case 1: delete (FftPre<TweedledumBase>*)h->pre; break;
    case 2: delete (FftPre<Bls12377Scalar>*)h->pre; break;
    case 3: delete (FftPre<Bls12377Base>*)h->pre; break;
  }
  delete h;
}
}  // extern "C"

// ---- curve entry points ------------------------------------------------------------------
struct RefMsmHandle { int cid; size_t n; void* pre; };

template <class C> static void* msm_pre_t(const uint64_t* xy, const uint8_t* zero, size_t n, int w) {
  auto aff = load_affine<C>(xy, zero, n);
  std::vector<Proj<C>> g(n);
  for (size_t i = 0; i < n; ++i) g[i] = to_proj(aff[i]);
  return msm_precompute<C>(g, w);
}
template <class C> static int msm_exec_t(RefMsmHandle* h, const uint64_t* scalars, size_t n, int parallel, uint64_t* out_xy, uint8_t* out_zero) {
  auto* pre = (MsmPre<C>*)h->pre;
  if (n != pre->powers.size()) return -3;   // assert_eq!, curve_msm.rs:67,106
  typedef Fe<typename C::Scalar> S;
  std::vector<S> sc(n);
  for (size_t i = 0; i < n; ++i) sc[i] = S::from_limbs(scalars + i * S::N);
  Proj<C> r = parallel ? msm_execute_parallel(*pre, sc) : msm_execute(*pre, sc);
  store_affine(to_affine(r), out_xy, out_zero);
  return 0;
}
template <class C> static int curve_mul_t(const uint64_t* xy, uint8_t zero, const uint64_t* scalar, uint64_t* out_xy, uint8_t* out_zero) {
  auto p = load_affine<C>(xy, &zero, 1);
  auto r = scalar_mul<C>(Fe<typename C::Scalar>::from_limbs(scalar), to_proj(p[0]));
  store_affine(to_affine(r), out_xy, out_zero);
  return 0;
}
// mode 0: pairwise, 1: batch inversion, 2: best (curve_summations.rs:18-158)
template <class C> static int affine_sum_t(const uint64_t* xy, const uint8_t* zero, size_t n, int mode, uint64_t* out_xy, uint8_t* out_zero) {
  std::vector<std::vector<Aff<C>>> one(1);
  one[0] = load_affine<C>(xy, zero, n);
  Proj<C> r;
  if (mode == 0) r = affine_summation_pairwise(one[0]);
  else if (mode == 1) r = multisum_batch_inversion(one)[0];
  else r = multisum_best(one)[0];
  store_affine(to_affine(r), out_xy, out_zero);
  return 0;
}
// P_i = [k_i] G with k_i = splitmix_hash(seed + i) -- the synthetic point set shared with the GPU generator
template <class C> static int gen_points_t(uint64_t seed, size_t n, uint64_t* out_xy) {
  Aff<C> g; g.x = Fe<typename C::Base>::from_limbs(C::GX); g.y = Fe<typename C::Base>::from_limbs(C::GY); g.zero = false;
  // 64 precomputed doublings of G (affine), then per point a double-and-add over the set bits
  std::vector<Proj<C>> dbl(64); dbl[0] = to_proj(g);
  for (int i = 1; i < 64; ++i) dbl[i] = proj_double(dbl[i - 1]);
  auto tab = batch_to_affine(dbl);
  const long B = 1024;
#pragma omp parallel for schedule(dynamic, 1)
  for (long b0 = 0; b0 < (long)n; b0 += B) {
    long b1 = std::min((long)n, b0 + B);
    std::vector<Proj<C>> acc(b1 - b0);
    for (long i = b0; i < b1; ++i) {
      uint64_t k = splitmix_hash(seed + (uint64_t)i);
      Proj<C> a = proj_zero<C>();
      for (int j = 0; j < 64; ++j) if ((k >> j) & 1) a = proj_add_aff(a, tab[j]);
      acc[i - b0] = a;
    }
    auto aff = batch_to_affine(acc);
    for (long i = b0; i < b1; ++i) { uint8_t z; store_affine(aff[i - b0], out_xy + 2 * i * C::Base::N, &z); }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// src/hash_to_curve.rs:13-76 -- blake_field / blake_hash_base_field_to_curve / blake_hash_usize_to_curve, the
// derivation of pedersen_g (src/circuit_builder.rs:1127).  BLAKE3 is the blake3 crate (Cargo.toml: "0.3.3", not
// vendored): restated from the published compression function for the one case the path needs (a single block
// of at most 64 bytes in, at most 64 bytes of extended output), like oracle/plonky_oracle.py, and pinned by the
// same golden vectors (tests/golden/blake_hash_to_curve.json, made with the independent `blake3` package).
// ------------------------------------------------------------------------------------------
static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void blake3_one_block(const uint8_t* data, size_t len, uint8_t out[64]) {
  static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
  uint8_t block[64] = {0};
  memcpy(block, data, len);
  uint32_t m[16], v[16];
  for (int i = 0; i < 16; ++i) m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
  for (int i = 0; i < 8; ++i) v[i] = IV[i];
  for (int i = 0; i < 4; ++i) v[8 + i] = IV[i];
  v[12] = 0; v[13] = 0; v[14] = (uint32_t)len; v[15] = 1 | 2 | 8;       // CHUNK_START | CHUNK_END | ROOT, counter 0
  auto g = [&](int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    v[a] = v[a] + v[b] + mx; v[d] = rotr32(v[d] ^ v[a], 16);
    v[c] = v[c] + v[d];      v[b] = rotr32(v[b] ^ v[c], 12);
    v[a] = v[a] + v[b] + my; v[d] = rotr32(v[d] ^ v[a], 8);
    v[c] = v[c] + v[d];      v[b] = rotr32(v[b] ^ v[c], 7);
  };
  for (int rnd = 0; rnd < 7; ++rnd) {
    g(0, 4, 8, 12, m[0], m[1]); g(1, 5, 9, 13, m[2], m[3]); g(2, 6, 10, 14, m[4], m[5]); g(3, 7, 11, 15, m[6], m[7]);
    g(0, 5, 10, 15, m[8], m[9]); g(1, 6, 11, 12, m[10], m[11]); g(2, 7, 8, 13, m[12], m[13]); g(3, 4, 9, 14, m[14], m[15]);
    if (rnd < 6) { uint32_t t[16]; for (int i = 0; i < 16; ++i) t[i] = m[PERM[i]]; memcpy(m, t, sizeof(m)); }
  }
  for (int i = 0; i < 8; ++i) { v[i] ^= v[i + 8]; v[i + 8] ^= IV[i]; }
  for (int i = 0; i < 16; ++i) for (int b = 0; b < 4; ++b) out[4 * i + b] = (uint8_t)(v[i] >> (8 * b));
}
// Field::square_root (field.rs:440-473) with is_quadratic_residue (:377-392, Euler's criterion): the reference's
// Tonelli-Shanks loop, so the same one of the two roots comes out.
template <class F> static bool fe_sqrt(const Fe<F>& a, Fe<F>* out) {
  typedef Fe<F> E;
  constexpr int N = F::N;
  if (a.is_zero()) { *out = a; return true; }
  const Big<N> pm1 = big_sub(E::order(), big_small<N>(1));
  if (a.exp_big(big_div2(pm1)) != E::one()) return false;
  const Big<N> t = E::from_limbs(F::T).to_canonical();
  E z = E::from_limbs(F::GENERATOR).exp_big(t);
  E w = a.exp_big(big_div2(big_sub(t, big_small<N>(1))));
  E x = w * a, b = x * w;
  int v = F::TWO_ADICITY;
  while (b != E::one()) {
    int k = 0;
    E b2k = b;
    while (b2k != E::one()) { b2k = b2k.square(); ++k; }
    w = z;
    for (int i = 0; i < v - k - 1; ++i) w = w.square();
    z = w.square(); b = b * z; x = x * w; v = k;
  }
  *out = x;
  return true;
}
template <class C> static Aff<C> blake_hash_base_field_to_curve(const Fe<typename C::Base>& seed) {
  typedef typename C::Base F;
  typedef Fe<F> E;
  constexpr int N = F::N, NB = 8 * N;
  const Big<N> sc = seed.to_canonical();
  uint8_t msg[NB + 2];
  for (int i = 0; i < NB; ++i) msg[i] = (uint8_t)(sc.l[i / 8] >> (8 * (i % 8)));
  for (int it = 0; it < 256; ++it) {                         // hash_to_curve.rs:59-75
    E x; bool y_neg = false;
    for (int j = 0; j < 256; ++j) {                          // blake_field, :13-51
      msg[NB] = (uint8_t)it; msg[NB + 1] = (uint8_t)j;
      uint8_t h[64];
      blake3_one_block(msg, NB + 2, h);
      h[NB - 1] >>= (8 * NB - F::BITS);
      Big<N> c;
      for (int l = 0; l < N; ++l) { c.l[l] = 0; for (int b = 0; b < 8; ++b) c.l[l] |= (uint64_t)h[8 * l + b] << (8 * b); }
      if (big_cmp(c, E::order()) < 0) { x = E::from_canonical(c); y_neg = h[NB] & 1; break; }
    }
    E cand = x.cube() + E::from_limbs(C::B);
    if (!C::A_IS_ZERO) cand = cand + E::from_limbs(C::A) * x;
    E y;
    if (fe_sqrt<F>(cand, &y)) {
      Aff<C> r; r.x = x; r.y = y_neg ? y.neg() : y; r.zero = false;
      return r;
    }
  }
  return aff_zero<C>();
}
template <class C> static int blake_points_t(uint64_t seed_start, size_t n, uint64_t* out_xy) {
  typedef Fe<typename C::Base> E;
#pragma omp parallel for schedule(dynamic, 64)
  for (long i = 0; i < (long)n; ++i) {
    uint8_t z;
    store_affine(blake_hash_base_field_to_curve<C>(E::from_u64(seed_start + (uint64_t)i)), out_xy + 2 * i * C::Base::N, &z);
  }
  return 0;
}
extern "C" int ref_blake_hash_usize_to_curve(int cid, uint64_t seed_start, size_t n, uint64_t* out_xy) {
  switch (cid) {
    case 0: return blake_points_t<Tweedledee>(seed_start, n, out_xy);
    case 1: return blake_points_t<Tweedledum>(seed_start, n, out_xy);
    case 2: return blake_points_t<Bls12377>(seed_start, n, out_xy);
  }
  return -1;
}

extern "C" {
void* ref_msm_precompute(int cid, const uint64_t* xy, const uint8_t* zero, size_t n, int w) {
  if (w < 1 || w > 24) return nullptr;
  auto* h = new RefMsmHandle{cid, n, nullptr};
  switch (cid) {
    case 0: h->pre = msm_pre_t<Tweedledee>(xy, zero, n, w); break;
    case 1: h->pre = msm_pre_t<Tweedledum>(xy, zero, n, w); break;
    case 2: h->pre = msm_pre_t<Bls12377>(xy, zero, n, w); break;
    default: delete h; return nullptr;
  }
  return h;
}
int ref_msm_execute(void* hv, const uint64_t* scalars, size_t n, int parallel, uint64_t* out_xy, uint8_t* out_zero) {
  auto* h = (RefMsmHandle*)hv;
  switch (h->cid) {
    case 0: return msm_exec_t<Tweedledee>(h, scalars, n, parallel, out_xy, out_zero);
    case 1: return msm_exec_t<Tweedledum>(h, scalars, n, parallel, out_xy, out_zero);
    case 2: return msm_exec_t<Bls12377>(h, scalars, n, parallel, out_xy, out_zero);
  }
  return -1;
}
void ref_msm_free(void* hv) {
  auto* h = (RefMsmHandle*)hv;
  if (!h) return;
  switch (h->cid) {
    case 0: delete (MsmPre<Tweedledee>*)h->pre; break;
    case 1: delete (MsmPre<Tweedledum>*)h->pre; break;
    case 2: delete (MsmPre<Bls12377>*)h->pre; break;
  }
  delete h;
}
int ref_curve_mul(int cid, const uint64_t* xy, uint8_t zero, const uint64_t* scalar, uint64_t* out_xy, uint8_t* out_zero) {
  switch (cid) {
    case 0: return curve_mul_t<Tweedledee>(xy, zero, scalar, out_xy, out_zero);
    case 1: return curve_mul_t<Tweedledum>(xy, zero, scalar, out_xy, out_zero);
    case 2: return curve_mul_t<Bls12377>(xy, zero, scalar, out_xy, out_zero);
  }
  return -1;
}
int ref_affine_sum(int cid, const uint64_t* xy, const uint8_t* zero, size_t n, int mode, uint64_t* out_xy, uint8_t* out_zero) {
  switch (cid) {
    case 0: return affine_sum_t<Tweedledee>(xy, zero, n, mode, out_xy, out_zero);
    case 1: return affine_sum_t<Tweedledum>(xy, zero, n, mode, out_xy, out_zero);
    case 2: return affine_sum_t<Bls12377>(xy, zero, n, mode, out_xy, out_zero);
  }
  return -1;
}
int ref_gen_points(int cid, uint64_t seed, size_t n, uint64_t* out_xy) {
  switch (cid) {
    case 0: return gen_points_t<Tweedledee>(seed, n, out_xy);
    case 1: return gen_points_t<Tweedledum>(seed, n, out_xy);
    case 2: return gen_points_t<Bls12377>(seed, n, out_xy);
  }
  return -1;
}
}  // extern "C"
template <class C> static int digits_t(const uint64_t* s, int w, uint32_t* out) {
  auto d = to_digits<C>(Fe<typename C::Scalar>::from_limbs(s), w);
  for (size_t i = 0; i < d.size(); ++i) out[i] = d[i];
  return (int)d.size();
}
extern "C" int ref_to_digits(int cid, const uint64_t* scalar, int w, uint32_t* out) {
  switch (cid) {
    case 0: return digits_t<Tweedledee>(scalar, w, out);
    case 1: return digits_t<Tweedledum>(scalar, w, out);
    case 2: return digits_t<Bls12377>(scalar, w, out);
  }
  return -1;
}

// ------------------------------------------------------------------------------------------
// src/halo.rs:63-124 -- one round of the Halo inner-product argument (the parts that touch the L1 kernels).
// The challenger (Rescue sponge), the blinding terms [l_j] H and the U' terms are the caller's: they are not
// data parallel.  msm_parallel(scalars, generators, w) = msm_precompute + msm_execute_parallel (curve_msm.rs:54-61).
// ------------------------------------------------------------------------------------------
template <class C> static Proj<C> msm_parallel(const std::vector<Fe<typename C::Scalar>>& sc, const std::vector<Proj<C>>& gens, int w) {
  MsmPre<C>* pre = msm_precompute<C>(gens, w);
  Proj<C> r = msm_execute_parallel(*pre, sc);
  delete pre;
  return r;
}
// halo.rs:87-93: <a_lo, G_hi>, <a_hi, G_lo> (window_size = 8, :78) and the inner products <a_lo, b_hi>, <a_hi, b_lo>
template <class C> static int ipa_round_lr_t(const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero, size_t n,
                                             uint64_t* out_lr_xy, uint8_t* out_lr_zero, uint64_t* out_ip) {
  typedef Fe<typename C::Scalar> S;
  const size_t mid = n / 2;
  auto g = load_affine<C>(g_xy, g_zero, n);
  std::vector<S> av(n), bv(n);
  for (size_t i = 0; i < n; ++i) { av[i] = S::from_limbs(a + i * S::N); bv[i] = S::from_limbs(b + i * S::N); }
  std::vector<Proj<C>> g_lo(mid), g_hi(mid);
  for (size_t i = 0; i < mid; ++i) { g_lo[i] = to_proj(g[i]); g_hi[i] = to_proj(g[mid + i]); }
  std::vector<S> a_lo(av.begin(), av.begin() + mid), a_hi(av.begin() + mid, av.end());
  Proj<C> l = msm_parallel<C>(a_lo, g_hi, 8), r = msm_parallel<C>(a_hi, g_lo, 8);
  store_affine(to_affine(l), out_lr_xy, out_lr_zero);
  store_affine(to_affine(r), out_lr_xy + 2 * C::Base::N, out_lr_zero + 1);
  S ip_l = S::zero(), ip_r = S::zero();           // Field::inner_product, field.rs:214-221
  for (size_t i = 0; i < mid; ++i) { ip_l = ip_l + av[i] * bv[mid + i]; ip_r = ip_r + av[mid + i] * bv[i]; }
  memcpy(out_ip, ip_l.v.l, 8 * S::N);
  memcpy(out_ip + S::N, ip_r.v.l, 8 * S::N);
  return 0;
}
// halo.rs:117-123: a' = u^-1 a_hi + u a_lo, b' = u^-1 b_lo + u b_hi, G'_i = msm_parallel([u^-1, u], [G_lo_i, G_hi_i], 4)
template <class C> static int ipa_fold_t(const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero, size_t n,
                                         const uint64_t* u_limbs, const uint64_t* u_inv_limbs, uint64_t* out_a, uint64_t* out_b,
                                         uint64_t* out_g_xy, uint8_t* out_g_zero) {
  typedef Fe<typename C::Scalar> S;
  const size_t mid = n / 2;
  const S u = S::from_limbs(u_limbs), u_inv = S::from_limbs(u_inv_limbs);
  auto g = load_affine<C>(g_xy, g_zero, n);
  for (size_t i = 0; i < mid; ++i) {
    S a_lo = S::from_limbs(a + i * S::N), a_hi = S::from_limbs(a + (mid + i) * S::N);
    S b_lo = S::from_limbs(b + i * S::N), b_hi = S::from_limbs(b + (mid + i) * S::N);
    S na = u_inv * a_hi + u * a_lo, nb = u_inv * b_lo + u * b_hi;
    memcpy(out_a + i * S::N, na.v.l, 8 * S::N);
    memcpy(out_b + i * S::N, nb.v.l, 8 * S::N);
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (long i = 0; i < (long)mid; ++i) {
    std::vector<S> sc = {u_inv, u};
    std::vector<Proj<C>> pts = {to_proj(g[i]), to_proj(g[mid + i])};
    Proj<C> r = msm_parallel<C>(sc, pts, 4);
    store_affine(to_affine(r), out_g_xy + 2 * i * C::Base::N, out_g_zero + i);
  }
  return 0;
}
extern "C" {
int ref_ipa_round_lr(int cid, const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero, size_t n,
                     uint64_t* out_lr_xy, uint8_t* out_lr_zero, uint64_t* out_ip) {
  if (n < 2 || (n & (n - 1))) return -2;          // log2_strict(degree), halo.rs:62
  switch (cid) {
    case 0: return ipa_round_lr_t<Tweedledee>(a, b, g_xy, g_zero, n, out_lr_xy, out_lr_zero, out_ip);
    case 1: return ipa_round_lr_t<Tweedledum>(a, b, g_xy, g_zero, n, out_lr_xy, out_lr_zero, out_ip);
    case 2: return ipa_round_lr_t<Bls12377>(a, b, g_xy, g_zero, n, out_lr_xy, out_lr_zero, out_ip);
  }
  return -1;
}
int ref_ipa_fold(int cid, const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero, size_t n,
                 const uint64_t* u, const uint64_t* u_inv, uint64_t* out_a, uint64_t* out_b, uint64_t* out_g_xy, uint8_t* out_g_zero) {
  if (n < 2 || (n & (n - 1))) return -2;
  switch (cid) {
    case 0: return ipa_fold_t<Tweedledee>(a, b, g_xy, g_zero, n, u, u_inv, out_a, out_b, out_g_xy, out_g_zero);
    case 1: return ipa_fold_t<Tweedledum>(a, b, g_xy, g_zero, n, u, u_inv, out_a, out_b, out_g_xy, out_g_zero);
    case 2: return ipa_fold_t<Bls12377>(a, b, g_xy, g_zero, n, u, u_inv, out_a, out_b, out_g_xy, out_g_zero);
  }
  return -1;
}
}  // extern "C"
