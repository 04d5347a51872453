// Host-side harness for the __host__ __device__ arithmetic in plonky_b200/csrc/{fp,ec}.cuh.
// Compiled with g++ (no CUDA): the PTX carry flag is emulated, so the SAME source that the kernels
// inline is checked against the big-integer oracle without a GPU (tests/test_fp_host.py).
// Test infrastructure only; limbs cross this boundary as u64 little-endian (== pairs of u32).
#include <stdint.h>
#include <string.h>
#include <stddef.h>
#include "../../plonky_b200/csrc/ec.cuh"

using namespace plk;

template <class P> static int field_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  typedef Fp<P> F;
  const int W = F::N / 2;
  for (size_t i = 0; i < n; ++i) {
    F x, y = F::zero(), r;
    memcpy(x.l, a + i * W, 4 * F::N);
    if (b) memcpy(y.l, b + i * W, 4 * F::N);
    switch (op) {
      case 0: r = F::add(x, y); break;
      case 1: r = F::sub(x, y); break;
      case 2: r = F::mul(x, y); break;
      case 3: r = F::sqr(x); break;
      case 4: r = F::neg(x); break;
      case 5: r = F::inverse(x); break;
      case 6: r = F::to_canonical(x); break;
      case 7: r = F::from_canonical(x); break;
      case 8: r = F::dbl(x); break;
      case 9: r = F::inverse_gcd(x); break;
      default: return -1;
    }
    memcpy(out + i * W, r.l, 4 * F::N);
  }
  return 0;
}

// op 0: madd chain  out = sum_i pts[i] (XYZZ accumulator + affine), normalised
// op 1: tree of XYZZ adds          op 2: double each point then sum (exercises dbl)
template <class C> static int curve_sum(int op, const uint64_t* xy, size_t n, uint64_t* out) {
  typedef Fp<typename C::Base> F;
  const int W = F::N / 2;
  XYZZ<C> acc = XYZZ<C>::identity();
  for (size_t i = 0; i < n; ++i) {
    Affine<C> p;
    memcpy(p.x.l, xy + (2 * i) * W, 4 * F::N);
    memcpy(p.y.l, xy + (2 * i + 1) * W, 4 * F::N);
    if (op == 0) acc = XYZZ<C>::madd(acc, p);
    else if (op == 1) acc = XYZZ<C>::add(XYZZ<C>::from_affine(p), acc);
    else acc = XYZZ<C>::add(acc, XYZZ<C>::dbl(XYZZ<C>::from_affine(p)));
  }
  Affine<C> r = XYZZ<C>::to_affine(acc);
  memcpy(out, r.x.l, 4 * F::N);
  memcpy(out + W, r.y.l, 4 * F::N);
  return 0;
}
template <class C> static int curve_mul64(const uint64_t* xy, uint64_t k, uint64_t* out) {
  typedef Fp<typename C::Base> F;
  const int W = F::N / 2;
  Affine<C> p;
  memcpy(p.x.l, xy, 4 * F::N);
  memcpy(p.y.l, xy + W, 4 * F::N);
  Affine<C> r = XYZZ<C>::to_affine(XYZZ<C>::mul_u64(XYZZ<C>::from_affine(p), k));
  memcpy(out, r.x.l, 4 * F::N);
  memcpy(out + W, r.y.l, 4 * F::N);
  return 0;
}

extern "C" int host_field_op(int fid, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  switch (fid) {
    case 0: return field_op<TweedledeeBaseParams>(op, a, b, out, n);
    case 1: return field_op<TweedledumBaseParams>(op, a, b, out, n);
    case 2: return field_op<Bls12377ScalarParams>(op, a, b, out, n);
    case 3: return field_op<Bls12377BaseParams>(op, a, b, out, n);
  }
  return -1;
}
extern "C" int host_curve_sum(int cid, int op, const uint64_t* xy, size_t n, uint64_t* out) {
  switch (cid) {
    case 0: return curve_sum<TweedledeeParams>(op, xy, n, out);
    case 1: return curve_sum<TweedledumParams>(op, xy, n, out);
    case 2: return curve_sum<Bls12377Params>(op, xy, n, out);
  }
  return -1;
}
extern "C" int host_curve_mul64(int cid, const uint64_t* xy, uint64_t k, uint64_t* out) {
  switch (cid) {
    case 0: return curve_mul64<TweedledeeParams>(xy, k, out);
    case 1: return curve_mul64<TweedledumParams>(xy, k, out);
    case 2: return curve_mul64<Bls12377Params>(xy, k, out);
  }
  return -1;
}
