// Host-side templates of the NTT: plan construction and pass scheduling for one field.
#pragma once
#include <stdlib.h>
#include "ntt_kernels.cuh"
#include "ntt_tma.cuh"

// defined in api.cu: out[i] = in[i]^-1 elementwise on device
void plk_launch_field_inverse(int field, const void* d_in, void* d_out, size_t n, cudaStream_t st);

namespace plk {

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
  pl->pow2_inv.alloc((size_t)(P::TWO_ADICITY + 1) * sizeof(F));     // 2^-k for every k (contiguous host table)
  PLK_CUDA(cudaMemcpyAsync(pl->pow2_inv.p, Tb::pow2_inv(0), (size_t)(P::TWO_ADICITY + 1) * sizeof(F), cudaMemcpyHostToDevice, st));
  for (int inv = 0; inv < 2; ++inv) {
    const uint32_t* w256 = inv ? Tb::root_inv(kSubLog <= P::TWO_ADICITY ? kSubLog : P::TWO_ADICITY)
                               : Tb::root(kSubLog <= P::TWO_ADICITY ? kSubLog : P::TWO_ADICITY);
    build_pow_table<P>(w256, 0, (size_t)1 << (kSubLog - 1), nullptr, pl->wsub[inv], st);
    const uint32_t* wn = inv ? Tb::root_inv(L) : Tb::root(L);
    build_pow_table<P>(wn, 0, (size_t)1 << pl->lo_bits, nullptr, pl->tw_lo[inv], st);
    build_pow_table<P>(wn, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), nullptr, pl->tw_hi[inv], st);
    if (inv) build_pow_table<P>(wn, pl->lo_bits, (size_t)1 << (L - pl->lo_bits), pl->n_inv.p, pl->tw_hi_inv_scaled, st);
  }
  // the full inter-pass twiddle tables of the in-place passes (ensure_direct) are built on first use, per direction
  PLK_CUDA(cudaStreamSynchronize(st));
}

// Full twiddle table w_{N_d}^(j_d k'') of the in-place pass over digit d in direction v (0 forward, 1 inverse, 2 inverse with
// n^-1 folded in: the last pass), built on first use when N_d <= 2^kDirectLog (plk_fft_set_direct_log / PLK_NTT_DIRECT_LOG lower the cap,
// 0 or PLK_NTT_NO_DIRECT=1 disable it: the pass then forms each twiddle on the fly from two small tables and one more product).
// A plan that only ever runs forward never allocates the inverse tables; a plan that only lends its small tables to the
// domain-split transform allocates none.
template <class P>
const void* ensure_direct(const plk_fft_plan* plc, int v, int d, int log_m_acc, cudaStream_t st) {
  typedef Fp<P> F;
  plk_fft_plan* pl = const_cast<plk_fft_plan*>(plc);
  const int r = pl->dig[d], log_nd = r + log_m_acc, sh = pl->log_n - log_nd;
  if (log_nd > pl->direct_log) return nullptr;
  std::lock_guard<std::mutex> lk(pl->mu);
  if (!pl->direct[v][d].p) {
    const size_t cnt = (size_t)1 << log_nd;
    pl->direct[v][d].alloc(cnt * sizeof(F));
    const int inv = v ? 1 : 0;
    direct_twiddle_kernel<F><<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(pl->tw_lo[inv].p, v == 2 ? pl->tw_hi_inv_scaled.p : pl->tw_hi[inv].p,
                                                                            pl->lo_bits, log_m_acc, r, sh, v == 2 ? 1 : 0,
                                                                            pl->direct[v][d].template as<F>());
    PLK_LAUNCHED();
    PLK_CUDA(cudaStreamSynchronize(st));       // other streams may use the table as soon as the lock is released
  }
  return pl->direct[v][d].p;
}

// One transform of k rows: d_in (n_in elements per row, stride in_stride) -> d_out (n per row).
// Uses d_out as the work buffer of the in-place passes.
template <class P>
void run_ntt(const plk_fft_plan* pl, const void* d_in, size_t n_in, size_t in_stride, void* d_out, size_t k,
             bool inverse, const FusedOps& ops, cudaStream_t st) {
  typedef Fp<P> F;
  const int L = pl->log_n;
  const int m = pl->m;
  const int inv = inverse ? 1 : 0;
  static std::mutex attr_mu;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t wsub_bytes = sizeof(F) << (kSubLog - 1);
  const size_t max_smem = (size_t)(F::N / 4) * 16 * ((size_t)1 << (kSubLog + kTileColsLog)) + wsub_bytes;
  // measurement knobs (defaults are the tuned values): PLK_NTT_TILE_LOG = log2 columns per tile, PLK_NTT_THREADS
  static const int tile_log = getenv("PLK_NTT_TILE_LOG") ? atoi(getenv("PLK_NTT_TILE_LOG")) : kTileColsLog;
  static const int nthreads = getenv("PLK_NTT_THREADS") ? atoi(getenv("PLK_NTT_THREADS")) : kNttThreads;
  // radix-4 register rounds at 3 CTAs / SM measured 6 % faster than radix-8 at 2 CTAs / SM (PLK_NTT_RADIX4=0 selects the latter)
  static const int radix4 = getenv("PLK_NTT_RADIX4") ? atoi(getenv("PLK_NTT_RADIX4")) : 1;
  {
    std::lock_guard<std::mutex> lk(attr_mu);                   // the opt-in to > 48 KiB of dynamic shared memory is per device
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      if constexpr (F::N == 8) {   // compile-time geometry only for the 8-limb fields (the 12-limb instantiations triple that unit's build time)
        PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 2, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
      }
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  int log_m_acc = 0;   // log2 of M_d for the pass being issued
  // per-pass CUDA events only when profiling is on, and then under the plan's timer lock (concurrent callers share the plan)
  PhaseTimer& timer = const_cast<plk_fft_plan*>(pl)->timer;
  std::unique_lock<std::mutex> timer_lock;
  if (g_profiling.load(std::memory_order_relaxed)) timer_lock = std::unique_lock<std::mutex>(const_cast<plk_fft_plan*>(pl)->timer_mu);
  timer.begin(st);
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
      p.log_t = (L - p.r) < tile_log ? (L - p.r) : tile_log;
      p.ndig = m - 1;
      for (int i = 0; i < m - 1; ++i) p.digs[i] = pl->dig[i];
      p.pre_lo = ops.pre_lo;
      p.pre_hi = ops.pre_hi;
      if (inverse && m == 1) p.scale = pl->n_inv.p;
    } else {
      p.log_m = log_m_acc;
      p.log_t = p.log_m < tile_log ? p.log_m : tile_log;
      if (inverse && p.last) { p.tw_all = 1; p.tw_hi = pl->tw_hi_inv_scaled.p; }
      const int v = (inverse && p.last) ? 2 : inv;
      p.tw_direct = ensure_direct<P>(pl, v, d, log_m_acc, st);
    }
    if (p.last) {
      p.post_lo = ops.post_lo;
      p.post_hi = ops.post_hi;
      p.post_periodic = ops.post_periodic;
      p.post_mask = ops.post_mask;
      p.post_rowmul = ops.post_rowmul;
      p.post_row_base = ops.post_row_base;
      p.remap = ops.remap;
      p.remap_cl_log = ops.remap_cl_log;
      p.remap_rows = ops.remap_rows;
      p.remap_row_base = ops.remap_row_base;
      for (int i = 0; i < kMaxPeers; ++i) p.peer[i] = ops.peer[i];
      if (ops.final_out) p.out = ops.final_out;
    }
    // the post tables may belong to the (larger) plan of a distributed transform: their own lo_bits
    p.post_lo_bits = (p.last && ops.post_lo_bits >= 0) ? ops.post_lo_bits : p.lo_bits;
    if (launch_tma_pass<F>(p, pl->n, k, st)) {        // TMA-staged Stockham pass (ntt_tma.cuh) when the geometry qualifies
      timer.mark(st);
      log_m_acc += p.r;
      continue;
    }
    const size_t tiles = pl->n >> (p.r + p.log_t);
    const size_t smem = (size_t)(F::N / 4) * 16 * ((size_t)1 << (p.r + p.log_t)) + wsub_bytes;
    if (tiles > 0x7fffffffull || k > 65535) fail(PLK_EINVAL, "transform grid too large");
    dim3 grid((unsigned)tiles, (unsigned)k);
    // compile-time tile geometry (ntt_kernels.cuh, GEO = 1) whenever the pass has it; PLK_NTT_GEO=0 keeps the generic kernel
    static const int geo = getenv("PLK_NTT_GEO") ? atoi(getenv("PLK_NTT_GEO")) : 1;
    const bool fixed_geo = F::N == 8 && radix4 && geo && p.log_t == kGeoLogT && nthreads == kNttThreads && p.r >= kGeoRMin && p.r <= kGeoRMax;
    if (fixed_geo) {
      if constexpr (F::N == 8) {
        if (p.r == 8) ntt_pass_kernel<F, 2, 8><<<grid, nthreads, smem, st>>>(p);
        else if (p.r == 7) ntt_pass_kernel<F, 2, 7><<<grid, nthreads, smem, st>>>(p);
        else ntt_pass_kernel<F, 2, 6><<<grid, nthreads, smem, st>>>(p);
      }
    }
    else if (radix4) ntt_pass_kernel<F, 2><<<grid, nthreads, smem, st>>>(p);
    else ntt_pass_kernel<F, 3><<<grid, nthreads, smem, st>>>(p);
    PLK_LAUNCHED();
    timer.mark(st);
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


template <class P>
void do_final_pass(const plk_fft_plan* pl, void* d_buf, int r, int log_cols, bool inverse, const void* d_scale, cudaStream_t st) {
  typedef Fp<P> F;
  NttPassParams p;
  memset(&p, 0, sizeof(p));
  p.log_n = r + log_cols;
  p.r = r;
  p.first = 0;
  p.last = 1;
  p.tw_none = 1;
  p.in = d_buf;
  p.out = d_buf;
  p.in_stride = p.out_stride = (unsigned long long)1 << (r + log_cols);
  p.log_m = log_cols;
  p.log_t = log_cols < kTileColsLog ? log_cols : kTileColsLog;
  p.wsub = pl->wsub[inverse ? 1 : 0].p;
  p.lo_bits = pl->lo_bits;
  p.post_lo_bits = pl->lo_bits;
  p.scale = d_scale;
  if (launch_tma_pass<F>(p, (size_t)1 << (r + log_cols), 1, st)) return;
  const size_t wsub_bytes = sizeof(F) << (kSubLog - 1);
  const size_t smem = (size_t)(F::N / 4) * 16 * ((size_t)1 << (p.r + p.log_t)) + wsub_bytes;
  PLK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<F, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((size_t)(F::N / 4) * 16 * ((size_t)1 << (kSubLog + kTileColsLog)) + wsub_bytes)));
  const size_t tiles = ((size_t)1 << (r + log_cols)) >> (p.r + p.log_t);
  ntt_pass_kernel<F, 2><<<dim3((unsigned)tiles, 1), kNttThreads, smem, st>>>(p);
  PLK_LAUNCHED();
}

template <class P>
const NttOps* make_ntt_ops() {
  static const NttOps ops = {&plan_build<P>, &do_run<P>, &do_coset<P>, &do_zh_table<P>, &do_final_pass<P>};
  return &ops;
}
}  // namespace plk
