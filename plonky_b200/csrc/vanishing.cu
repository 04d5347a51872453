// Pointwise evaluation of Plonk's vanishing polynomial on the 8n-point LDE domain.
//
// Replaces the par_iter body of Circuit::vanishing_poly (src/plonk.rs:393-452): for every point x = w_8n^i
//   constraint terms   evaluate_all_constraints (src/gates/mod.rs:46-125): sum over the ten gates of
//                      prefix_filter(constants) * evaluate_unfiltered(constants, local, right, below wires)
//                      (src/gates/{curve_add,curve_dbl,curve_endo,base_4_sum,public_input,buffer,constant,arithmetic,
//                      rescue_a,rescue_b}.rs), index-wise
//   L_1(x) (Z(x) - 1)  eval_l_1 (src/plonk_util.rs:14-24)
//   Z(x) f'(x) - g'(x) Z(g x)   the permutation argument over the NUM_ROUTED_WIRES = 6 routed wires
//   reduce_with_powers(terms, alpha) (src/plonk_util.rs:27-33)
// The wire / constant / sigma / Z evaluations are the outputs of the 8n LDE transforms and already live in HBM; the
// result feeds the 8n inverse transform (Polynomial::from_evaluations, plonk.rs:455) and divide_by_z_h without leaving the
// device (plk_vanishing_poly chains them).
//
// Exact field arithmetic, so algebraically equal rearrangements give the same reduced limbs:
//   * reduce_with_powers is linear: result = z1 + alpha * shift + sum_gates filter_g * sum_k c_(g,k) alpha^(k+2)
//     -- one running accumulator instead of a unified constraint vector;
//   * the ten prefix filters share their common prefixes (a 15-product tree instead of 41 products);
//   * double() / triple() / quadruple() are additions (the reference multiplies by TWO / THREE / FOUR, field.rs:181-197);
//   * L_1(x) = (x^n - 1) / (n (x - 1)) with x^n = w_8^(i mod 8); the 8n divisions are one table per (field, n) built with
//     Montgomery's trick and cached (the reference recomputes a division per point).
// One thread per point; rows are read with coalesced 32-byte loads (adjacent threads, adjacent points).
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"
#include "fp.cuh"
#include "ntt_plan.h"

namespace plk {

constexpr int kVanishDefaultMB = 4;      // measured at 2^19 points: 2.64 ms (1), 2.49 ms (3), 2.44 ms (4)
constexpr int kVanWires = 9, kVanRouted = 6, kVanConsts = 6, kVanGridWidth = 65;
// layout of the small constants buffer (elements of F)
constexpr int kcK = 0, kcAlpha = 6, kcBeta = 7, kcGamma = 8, kcZetaM1 = 9, kcA = 10, kcMds = 11, kcApow = 27, kcN = 37, kcCount = 38;

struct VanishArgs {
  const void* wires;     // kVanWires rows of m elements
  const void* consts;    // kVanConsts rows
  const void* sigma;     // kVanRouted rows
  const void* z;         // m
  const void* subgroup;  // m: x_i = w_8n^i
  const void* l1;        // m: L_1(x_i)
  const void* small;     // kcCount elements
  unsigned long long m;  // 8 n
  void* out;
};

template <class F>
static __device__ __noinline__ F vmul(const F a, const F b) { return F::mul(a, b); }
template <class F>
static __device__ __forceinline__ F from_small(unsigned v) {
  F c = F::zero();
  c.l[0] = v;
  return F::from_canonical(c);
}

// in: k_is[6], alpha, beta, gamma, zeta, a (11 elements, Montgomery) -> the small constants buffer
template <class P>
__global__ void vanish_setup_kernel(const Fp<P>* __restrict__ in, unsigned long long degree, Fp<P>* __restrict__ out) {
  typedef Fp<P> F;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < 9; ++i) out[i] = in[i];
  out[kcZetaM1] = F::sub(in[9], F::one());
  out[kcA] = in[10];
  // mds.rs:63-77: Cauchy matrix 1 / (x_r - y_c), x_r = 4 + r, y_c = c
  F inv[8];
  for (unsigned v = 1; v <= 7; ++v) inv[v] = F::inverse(from_small<F>(v));
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[kcMds + 4 * r + c] = inv[4 + r - c];
  F ap = F::one();
  for (int k = 0; k < 10; ++k) { out[kcApow + k] = ap; ap = F::mul(ap, in[6]); }
  F nn = F::zero();
  nn.l[0] = (uint32_t)degree;
  nn.l[1] = (uint32_t)(degree >> 32);
  out[kcN] = F::from_canonical(nn);
}

// L_1(x_i) for the m = 8n points x_i = w^i of the field's OWN 8n-th root of unity w (plonk_util.rs:14-24): the cached table
// must not depend on the subgroup array a caller hands in.  Strips of 16 points per thread share one inversion; a thread's
// first point w^(16 t) comes from binary powering, the others from successive products.
template <class P>
__global__ void __launch_bounds__(128) vanish_l1_kernel(Fp<P> w, unsigned long long m, const Fp<P>* __restrict__ small, void* __restrict__ out) {
  typedef Fp<P> F;
  constexpr int S = 16;
  const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long lo = t * S;
  if (lo >= m) return;
  const F one = F::one(), n = small[kcN];
  F x0 = one, b = w;                                  // x0 = w^lo
  for (unsigned long long e = lo; e; e >>= 1) {
    if (e & 1) x0 = vmul<F>(x0, b);
    b = vmul<F>(b, b);
  }
  F w8 = w;                                           // w^(m/8): the primitive 8th root, x^n = w8^(i mod 8)
  for (unsigned long long e = m >> 3; e > 1; e >>= 1) w8 = vmul<F>(w8, w8);
  F pre[S];
  F run = one, x = x0;
  for (int k = 0; k < S; ++k) {
    pre[k] = run;
    if (lo + k < m) {
      const F d = (x == one) ? one : vmul<F>(n, F::sub(x, one));
      run = vmul<F>(run, d);
      x = vmul<F>(x, w);
    }
  }
  F inv = F::inverse(run);
  // x^n for the last point of the strip, then walk both down: x /= w is x * w^-1 -- avoided by recomputing from x0 instead
  F xs[S];
  x = x0;
  for (int k = 0; k < S; ++k) { xs[k] = x; x = vmul<F>(x, w); }
  for (int k = S - 1; k >= 0; --k) {
    const unsigned long long i = lo + k;
    if (i >= m) continue;
    F r;
    if (xs[k] == one) r = one;
    else {
      const F dinv = vmul<F>(inv, pre[k]);
      inv = vmul<F>(inv, vmul<F>(n, F::sub(xs[k], one)));
      F xn = one;                                     // w8^(i mod 8)
      for (unsigned j = 0; j < (unsigned)(i & 7); ++j) xn = vmul<F>(xn, w8);
      r = vmul<F>(F::sub(xn, one), dinv);
    }
    store_fp<F>(out, i, r);
  }
}

// MB = resident CTAs per SM the register allocation is held to (1: unconstrained, 185 registers; 3: 168; 4: 128 with a few
// hundred bytes of spills): measured variants, selected by PLK_VANISH_MB
template <class P, int MB>
__global__ void __launch_bounds__(128, MB) vanishing_points_kernel(VanishArgs a) {
  typedef Fp<P> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long m = a.m;
  if (i >= m) return;
  unsigned long long ir = i + 8, ib = i + 8ull * kVanGridWidth;        // a shift of 1 in the n-subgroup = 8 in the 8n-subgroup
  if (ir >= m) ir -= m;
  ib %= m;
  const F* K = reinterpret_cast<const F*>(a.small);
  const F one = F::one();
  auto W = [&](int j) { return load_fp<F>(a.wires, (unsigned long long)j * m + i); };     // local wire j
  auto R = [&](int j) { return load_fp<F>(a.wires, (unsigned long long)j * m + ir); };    // right
  auto B = [&](int j) { return load_fp<F>(a.wires, (unsigned long long)j * m + ib); };    // below
  auto Cn = [&](int j) { return load_fp<F>(a.consts, (unsigned long long)j * m + i); };
  auto AP = [&](int k) { return K[kcApow + k]; };                                         // alpha^k
  auto mul = [](const F& x, const F& y) { return vmul<F>(x, y); };
  auto dbl = [](const F& x) { return F::add(x, x); };

  // ---- L_1(x) (Z(x) - 1) and the permutation term (plonk.rs:424-437) ----
  const F x = load_fp<F>(a.subgroup, i);
  const F z_x = load_fp<F>(a.z, i), z_gz = load_fp<F>(a.z, ir);
  F acc = mul(load_fp<F>(a.l1, i), F::sub(z_x, one));
  {
    F fp = one, gp = one;
    const F beta = K[kcBeta], gamma = K[kcGamma];
    const F bx = mul(beta, x);
#pragma unroll 1
    for (int j = 0; j < kVanRouted; ++j) {
      const F w = W(j);
      const F s_sigma = load_fp<F>(a.sigma, (unsigned long long)j * m + i);
      fp = mul(fp, F::add(F::add(w, mul(K[kcK + j], bx)), gamma));         // beta * (k_i * x) == k_i * (beta * x)
      gp = mul(gp, F::add(F::add(w, mul(beta, s_sigma)), gamma));
    }
    const F shift = F::sub(mul(fp, z_x), mul(gp, z_gz));
    acc = F::add(acc, mul(shift, AP(1)));
  }

  // ---- prefix filters as a tree over the constants c0..c5 (gates/mod.rs:281-293; prefixes in the file header of mod.rs).
  // Every filter is formed right before its gate from freshly loaded constants (L1 hits) so that only the accumulator and one
  // shared prefix stay live across the gate bodies. ----
  auto gate = [&](const F& filter, const F& weighted) { acc = F::add(acc, mul(filter, weighted)); };
  auto notc = [&](int j) { return F::sub(one, Cn(j)); };

  // RescueStepAGate, prefix 00 (rescue_a.rs:37-64): constraints (root_i^5 - in_i, const_i + sum_j mds_ij root_j - out_i) interleaved
  {
    F roots[4];
    for (int k = 0; k < 4; ++k) roots[k] = W(4 + k);
    F h = F::zero();
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const F r2 = mul(roots[k], roots[k]);
      const F r5 = mul(mul(r2, r2), roots[k]);
      h = F::add(h, mul(F::sub(r5, W(k)), AP(2 + 2 * k)));
      F o = Cn(2 + k);
      for (int j = 0; j < 4; ++j) o = F::add(o, mul(K[kcMds + 4 * k + j], roots[j]));
      h = F::add(h, mul(F::sub(o, R(k)), AP(3 + 2 * k)));
    }
    gate(mul(notc(0), notc(1)), h);
  }
  // RescueStepBGate, prefix 01 (rescue_b.rs:32-56)
  {
    F exps[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const F w = W(k);
      const F w2 = mul(w, w);
      exps[k] = mul(mul(w2, w2), w);
    }
    F h = F::zero();
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      F o = Cn(2 + k);
      for (int j = 0; j < 4; ++j) o = F::add(o, mul(K[kcMds + 4 * k + j], exps[j]));
      h = F::add(h, mul(F::sub(o, R(k)), AP(2 + k)));
    }
    gate(mul(notc(0), Cn(1)), h);
  }
  // CurveEndoGate, prefix 11 (curve_endo.rs:37-85)
  {
    const F x1 = W(0), y1 = W(1), b0 = W(6), b1 = W(7), inv = W(8);
    const F x3 = R(0), y3 = R(1);
    const F mult = F::add(mul(K[kcZetaM1], b1), one);
    const F sgn = F::sub(dbl(b0), one);
    const F x2 = mul(mult, W(4)), y2 = mul(sgn, W(5));
    const F lam = mul(F::sub(y1, y2), inv);
    F h = mul(F::sub(F::sub(F::sub(mul(lam, lam), x1), x2), x3), AP(2));
    h = F::add(h, mul(F::sub(F::sub(mul(lam, F::sub(x1, x3)), y1), y3), AP(3)));
    const F su = W(2), ss = W(3);
    h = F::add(h, mul(F::sub(B(2), F::add(F::add(dbl(dbl(su)), dbl(b1)), b0)), AP(4)));
    h = F::add(h, mul(F::sub(B(3), F::add(dbl(ss), mul(sgn, mult))), AP(5)));
    h = F::add(h, mul(mul(b0, F::sub(b0, one)), AP(6)));
    h = F::add(h, mul(mul(b1, F::sub(b1, one)), AP(7)));
    h = F::add(h, mul(F::sub(mul(inv, F::sub(x1, x2)), one), AP(8)));
    gate(mul(Cn(0), Cn(1)), h);
  }
  const F p10 = mul(Cn(0), notc(1));
  {
  const F p100 = mul(p10, notc(2));
  // Base4SumGate, prefix 1000 (base_4_sum.rs:36-62)
  {
    F sum = W(0);
    F h = F::zero();
    const F two = dbl(one), three = F::add(two, one);
#pragma unroll 1
    for (int k = 0; k < 7; ++k) {
      const F l = W(2 + k);
      sum = F::add(dbl(dbl(sum)), l);
      const F prod = mul(mul(l, F::sub(l, one)), mul(F::sub(l, two), F::sub(l, three)));
      h = F::add(h, mul(prod, AP(3 + k)));
    }
    h = F::add(h, mul(F::sub(sum, W(1)), AP(2)));
    gate(mul(p100, notc(3)), h);
  }
  // ArithmeticGate, prefix 1001 (arithmetic.rs:35-49)
  {
    const F out = F::sub(F::add(mul(mul(Cn(4), W(0)), W(1)), mul(Cn(5), W(2))), W(3));
    gate(mul(p100, Cn(3)), mul(out, AP(2)));
  }
  }
  const F p101 = mul(p10, Cn(2));
  {
  const F p1010 = mul(p101, notc(3));
  // CurveAddGate, prefix 10101 (curve_add.rs:38-81)
  {
    const F x1 = W(0), y1 = W(1), x2 = W(4), y2 = W(5), bit = W(6), inv = W(7), lam = W(8);
    const F x4 = R(0), y4 = R(1);
    const F x3 = F::sub(F::sub(mul(lam, lam), x1), x2);
    const F y3 = F::sub(mul(lam, F::sub(x1, x4)), y1);
    const F nb = F::sub(one, bit);
    F h = mul(F::sub(mul(F::sub(y1, y2), inv), lam), AP(2));
    h = F::add(h, mul(F::sub(F::add(mul(bit, x3), mul(nb, x1)), x4), AP(3)));
    h = F::add(h, mul(F::sub(F::add(mul(bit, y3), mul(nb, y1)), y4), AP(4)));
    h = F::add(h, mul(F::sub(W(3), F::add(dbl(W(2)), bit)), AP(5)));
    h = F::add(h, mul(mul(bit, nb), AP(6)));
    h = F::add(h, mul(F::sub(mul(inv, F::sub(x1, x2)), one), AP(7)));
    gate(mul(p1010, Cn(4)), h);
  }
  // PublicInputGate, prefix 101001 (public_input.rs:32-43); BufferGate 101000 has no constraints
  {
    F h = F::zero();
    for (int k = 0; k < 3; ++k) h = F::add(h, mul(F::sub(W(kVanRouted + k), R(k)), AP(2 + k)));
    gate(mul(mul(p1010, notc(4)), Cn(5)), h);
  }
  }
  const F p1011 = mul(p101, Cn(3));
  // CurveDblGate, prefix 10111 (curve_dbl.rs:31-60)
  {
    const F xo = W(0), yo = W(1), xn = W(2), yn = W(3), inv = W(4), lam = W(5);
    const F xx = mul(xo, xo);
    const F num = F::add(F::add(dbl(xx), xx), K[kcA]);
    F h = mul(F::sub(mul(num, inv), lam), AP(2));
    h = F::add(h, mul(F::sub(F::sub(mul(lam, lam), dbl(xo)), xn), AP(3)));
    h = F::add(h, mul(F::sub(F::sub(mul(lam, F::sub(xo, xn)), yo), yn), AP(4)));
    h = F::add(h, mul(F::sub(mul(dbl(yo), inv), one), AP(5)));
    gate(mul(p1011, Cn(4)), h);
  }
  // ConstantGate, prefix 10110 (constant.rs:28-37)
  gate(mul(p1011, notc(4)), mul(F::sub(Cn(5), W(0)), AP(2)));

  store_fp<F>(a.out, i, acc);
}

// L_1 tables, one per (device, field, degree)
static std::mutex g_l1_mu;
static std::map<std::tuple<int, int, unsigned long long>, DevBuf*> g_l1_tables;

template <class P>
void vanishing_points_run(int field, unsigned long long degree, const void* d_wires, const void* d_consts, const void* d_sigma, const void* d_z,
                          const void* d_subgroup, const void* d_params /* 11 elements */, void* d_out, cudaStream_t st) {
  typedef Fp<P> F;
  const unsigned long long m = 8 * degree;
  if (log2_floor(m) > P::TWO_ADICITY) fail(PLK_ETOOBIG, "log2(8 * degree) exceeds TWO_ADICITY");      // field.rs:430
  int dev = 0;
  PLK_CUDA(cudaGetDevice(&dev));
  DevBuf small(kcCount * sizeof(F), st);
  vanish_setup_kernel<P><<<1, 1, 0, st>>>(reinterpret_cast<const F*>(d_params), degree, small.as<F>());
  PLK_LAUNCHED();
  const void* l1 = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_l1_mu);
    auto key = std::make_tuple(dev, field, degree);
    auto it = g_l1_tables.find(key);
    if (it == g_l1_tables.end()) {
      auto* buf = new DevBuf(m * sizeof(F));
      const unsigned long long threads = (m + 15) / 16;
      F w;
      const uint32_t* wl = FieldTables<P>::root(log2_floor(m));      // primitive_root_of_unity(log2(8n)), field.rs:429-435
      for (int i = 0; i < F::N; ++i) w.l[i] = wl[i];
      vanish_l1_kernel<P><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(w, m, small.as<F>(), buf->p);
      PLK_LAUNCHED();
      PLK_CUDA(cudaStreamSynchronize(st));          // the table outlives this stream's ordering
      it = g_l1_tables.emplace(key, buf).first;
    }
    l1 = it->second->p;
  }
  VanishArgs a;
  a.wires = d_wires; a.consts = d_consts; a.sigma = d_sigma; a.z = d_z; a.subgroup = d_subgroup; a.l1 = l1; a.small = small.p; a.m = m; a.out = d_out;
  static const int mb = getenv("PLK_VANISH_MB") ? atoi(getenv("PLK_VANISH_MB")) : kVanishDefaultMB;
  const unsigned blocks = (unsigned)((m + 127) / 128);
  if (mb >= 4) vanishing_points_kernel<P, 4><<<blocks, 128, 0, st>>>(a);
  else if (mb == 3) vanishing_points_kernel<P, 3><<<blocks, 128, 0, st>>>(a);
  else vanishing_points_kernel<P, 1><<<blocks, 128, 0, st>>>(a);
  PLK_LAUNCHED();
}

static void vanishing_dispatch(int field, unsigned long long degree, const void* d_wires, const void* d_consts, const void* d_sigma, const void* d_z,
                               const void* d_subgroup, const void* d_params, void* d_out, cudaStream_t st) {
  switch (field) {
    case PLK_FIELD_TWEEDLEDEE_BASE: vanishing_points_run<TweedledeeBaseParams>(field, degree, d_wires, d_consts, d_sigma, d_z, d_subgroup, d_params, d_out, st); break;
    case PLK_FIELD_TWEEDLEDUM_BASE: vanishing_points_run<TweedledumBaseParams>(field, degree, d_wires, d_consts, d_sigma, d_z, d_subgroup, d_params, d_out, st); break;
    case PLK_FIELD_BLS12_377_SCALAR: vanishing_points_run<Bls12377ScalarParams>(field, degree, d_wires, d_consts, d_sigma, d_z, d_subgroup, d_params, d_out, st); break;
    default: fail(PLK_EINVAL, "vanishing_poly: unsupported field id (4-limb scalar fields only)");
  }
}

}  // namespace plk

using namespace plk;

extern "C" {

int plk_vanishing_points_dev(int field, size_t degree, const void* d_wires_8n, const void* d_constants_8n, const void* d_sigma_8n,
                             const void* d_z_8n, const void* d_subgroup_8n, const void* d_params, void* d_out_8n, void* stream) {
  return guarded([&] {
    if (!is_pow2(degree)) fail(PLK_ENOTPOW2, "Not a power of two");
    if (!d_wires_8n || !d_constants_8n || !d_sigma_8n || !d_z_8n || !d_subgroup_8n || !d_params || !d_out_8n) fail(PLK_EINVAL, "NULL buffer");
    vanishing_dispatch(field, degree, d_wires_8n, d_constants_8n, d_sigma_8n, d_z_8n, d_subgroup_8n, d_params, d_out_8n,
                       reinterpret_cast<cudaStream_t>(stream));
  });
}

int plk_vanishing_points(int field, size_t degree, const uint64_t* wires_8n, const uint64_t* constants_8n, const uint64_t* sigma_8n,
                         const uint64_t* z_8n, const uint64_t* subgroup_8n, const uint64_t* k_is, const uint64_t* alpha, const uint64_t* beta,
                         const uint64_t* gamma, const uint64_t* inner_zeta, const uint64_t* inner_a, uint64_t* out_8n) {
  return guarded([&] {
    if (plk_field_limbs(field) != 4) fail(PLK_EINVAL, "vanishing_poly: unsupported field id (4-limb scalar fields only)");
    if (!is_pow2(degree)) fail(PLK_ENOTPOW2, "Not a power of two");
    if (!wires_8n || !constants_8n || !sigma_8n || !z_8n || !subgroup_8n || !k_is || !alpha || !beta || !gamma || !inner_zeta || !inner_a || !out_8n)
      fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t m = 8 * degree, eb = 32;
    DevBuf d_w(kVanWires * m * eb, st), d_c(kVanConsts * m * eb, st), d_s(kVanRouted * m * eb, st), d_z(m * eb, st), d_x(m * eb, st), d_p(11 * eb, st),
        d_o(m * eb, st);
    PLK_CUDA(cudaMemcpyAsync(d_w.p, wires_8n, kVanWires * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_c.p, constants_8n, kVanConsts * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_s.p, sigma_8n, kVanRouted * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_z.p, z_8n, m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_x.p, subgroup_8n, m * eb, cudaMemcpyHostToDevice, st));
    uint64_t params[11 * 4];
    memcpy(params, k_is, 6 * eb);
    memcpy(params + 24, alpha, eb);
    memcpy(params + 28, beta, eb);
    memcpy(params + 32, gamma, eb);
    memcpy(params + 36, inner_zeta, eb);
    memcpy(params + 40, inner_a, eb);
    PLK_CUDA(cudaMemcpyAsync(d_p.p, params, sizeof(params), cudaMemcpyHostToDevice, st));
    vanishing_dispatch(field, degree, d_w.p, d_c.p, d_s.p, d_z.p, d_x.p, d_p.p, d_o.p, st);
    PLK_CUDA(cudaMemcpyAsync(out_8n, d_o.p, m * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
