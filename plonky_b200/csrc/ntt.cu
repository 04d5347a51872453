// C ABI of the NTT (include/plonky_b200.h); kernels live in ntt_kernels.cuh, one translation unit per field.
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include "ntt_plan.h"
#include "field_constants.cuh"

using namespace plk;

void plk_launch_field_mul(int field, const void* d_a, const void* d_b, void* d_out, size_t n, cudaStream_t st);   // api.cu

namespace plk {
const NttOps* ntt_ops_tweedledee_base();
const NttOps* ntt_ops_tweedledum_base();
const NttOps* ntt_ops_bls12_377_scalar();
const NttOps* ntt_ops_bls12_377_base();
}

namespace {

const NttOps* ops_for(int field) {
  switch (field) {
    case PLK_FIELD_TWEEDLEDEE_BASE: return ntt_ops_tweedledee_base();
    case PLK_FIELD_TWEEDLEDUM_BASE: return ntt_ops_tweedledum_base();
    case PLK_FIELD_BLS12_377_SCALAR: return ntt_ops_bls12_377_scalar();
    case PLK_FIELD_BLS12_377_BASE: return ntt_ops_bls12_377_base();
  }
  fail(PLK_EINVAL, "unknown field id");
}

int field_two_adicity(int field) {
  switch (field) {
    case PLK_FIELD_TWEEDLEDEE_BASE: return TweedledeeBaseParams::TWO_ADICITY;
    case PLK_FIELD_TWEEDLEDUM_BASE: return TweedledumBaseParams::TWO_ADICITY;
    case PLK_FIELD_BLS12_377_SCALAR: return Bls12377ScalarParams::TWO_ADICITY;
    case PLK_FIELD_BLS12_377_BASE: return Bls12377BaseParams::TWO_ADICITY;
  }
  return -1;
}

// host-pointer front end: copy in, run, copy out on the calling thread's stream
void host_transform(plk_fft_plan* pl, const uint64_t* in, size_t n_in, size_t k, bool inverse, const FusedOps& ops,
                    uint64_t* out) {
  cudaStream_t st = thread_stream();
  const size_t eb = pl->elem_bytes;
  void* d_in = thread_scratch(0, n_in * k * eb);
  void* d_out = thread_scratch(1, pl->n * k * eb);
  PLK_CUDA(cudaMemcpyAsync(d_in, in, n_in * k * eb, cudaMemcpyHostToDevice, st));
  ops_for(pl->field)->run(pl, d_in, n_in, n_in, d_out, k, inverse, &ops, st);
  PLK_CUDA(cudaMemcpyAsync(out, d_out, pl->n * k * eb, cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaStreamSynchronize(st));
}

void check_plan(const plk_fft_plan* p) {
  if (!p) fail(PLK_EINVAL, "NULL plan");
}

// w^k for k < n on the device (cyclic_subgroup_known_order, src/field/field.rs:292-300), cached in the plan
const void* plan_subgroup(plk_fft_plan* pl, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(pl->sub_mu);        // not pl->mu: the transform below takes that one (lazy twiddle tables)
  if (!pl->subgroup.p) {
    const size_t eb = pl->elem_bytes;
    DevBuf e1(pl->n * eb, st);
    PLK_CUDA(cudaMemsetAsync(e1.p, 0, pl->n * eb, st));
    // ONE in Montgomery form = 2^0 of the plan's table of inverse powers of two
    if (pl->n > 1) PLK_CUDA(cudaMemcpyAsync((char*)e1.p + eb, pl->pow2_inv.p, eb, cudaMemcpyDeviceToDevice, st));
    else PLK_CUDA(cudaMemcpyAsync(e1.p, pl->pow2_inv.p, eb, cudaMemcpyDeviceToDevice, st));
    pl->subgroup.alloc(pl->n * eb);
    FusedOps ops;
    ops_for(pl->field)->run(pl, e1.p, pl->n, pl->n, pl->subgroup.p, 1, false, &ops, st);
    PLK_CUDA(cudaStreamSynchronize(st));
  }
  return pl->subgroup.p;
}

}  // namespace

extern "C" {

int plk_fft_precompute(int field, size_t degree, plk_fft_plan** out) {
  return guarded([&] {
    if (!out) fail(PLK_EINVAL, "NULL out");
    *out = nullptr;
    const int ta = field_two_adicity(field);
    if (ta < 0) fail(PLK_EINVAL, "unknown field id");
    if (degree == 0) degree = 1;
    const int L = log2_ceil(degree);         // fft.rs:48
    if (L > ta) fail(PLK_ETOOBIG, "log2(size) exceeds TWO_ADICITY");   // field.rs:430 assert
    auto* pl = new plk_fft_plan();
    pl->field = field;
    pl->log_n = L;
    pl->n = (size_t)1 << L;
    cudaGetDevice(&pl->device);
    if (const char* e = getenv("PLK_NTT_DIRECT_LOG")) pl->direct_log = atoi(e);
    if (getenv("PLK_NTT_NO_DIRECT")) pl->direct_log = 0;
    try {
      ops_for(field)->plan_build(pl);
    } catch (...) {
      delete pl;
      throw;
    }
    *out = pl;
  });
}

// Circuit::vanishing_poly (src/plonk.rs:375-456) end to end on the device: pad_to_8n + FFT of the Z coefficients, the
// pointwise evaluation (vanishing.cu), Polynomial::from_evaluations (the 8n inverse transform).
int plk_vanishing_poly(const plk_fft_plan* pc, size_t degree, const uint64_t* wires_8n, const uint64_t* constants_8n, const uint64_t* sigma_8n,
                       const uint64_t* plonk_z_coeffs, const uint64_t* k_is, const uint64_t* alpha, const uint64_t* beta, const uint64_t* gamma,
                       const uint64_t* inner_zeta, const uint64_t* inner_a, uint64_t* out_coeffs_8n) {
  return guarded([&] {
    check_plan(pc);
    auto* pl = const_cast<plk_fft_plan*>(pc);
    if (!is_pow2(degree)) fail(PLK_ENOTPOW2, "Not a power of two");
    if (pl->n != 8 * degree) fail(PLK_ESIZE, "the precomputation must have size 8 * degree (fft_precomputation_8n)");
    if (pl->elem_bytes != 32) fail(PLK_EINVAL, "vanishing_poly: unsupported field id (4-limb scalar fields only)");
    if (!wires_8n || !constants_8n || !sigma_8n || !plonk_z_coeffs || !k_is || !alpha || !beta || !gamma || !inner_zeta || !inner_a || !out_coeffs_8n)
      fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t m = pl->n, eb = 32;
    DevBuf d_w(9 * m * eb, st), d_c(6 * m * eb, st), d_s(6 * m * eb, st), d_zc(degree * eb, st), d_z8(m * eb, st), d_p(11 * eb, st), d_pts(m * eb, st),
        d_out(m * eb, st);
    PLK_CUDA(cudaMemcpyAsync(d_w.p, wires_8n, 9 * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_c.p, constants_8n, 6 * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_s.p, sigma_8n, 6 * m * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_zc.p, plonk_z_coeffs, degree * eb, cudaMemcpyHostToDevice, st));
    uint64_t params[11 * 4];
    memcpy(params, k_is, 6 * eb);
    memcpy(params + 24, alpha, eb);
    memcpy(params + 28, beta, eb);
    memcpy(params + 32, gamma, eb);
    memcpy(params + 36, inner_zeta, eb);
    memcpy(params + 40, inner_a, eb);
    PLK_CUDA(cudaMemcpyAsync(d_p.p, params, sizeof(params), cudaMemcpyHostToDevice, st));
    const void* sub = plan_subgroup(pl, st);
    FusedOps ops;
    ops_for(pl->field)->run(pl, d_zc.p, degree, degree, d_z8.p, 1, false, &ops, st);                 // plonk.rs:388-391
    const int rc = plk_vanishing_points_dev(pl->field, degree, d_w.p, d_c.p, d_s.p, d_z8.p, sub, d_p.p, d_pts.p, st);
    if (rc != PLK_OK) fail(rc, plk_last_error_message());
    ops_for(pl->field)->run(pl, d_pts.p, m, m, d_out.p, 1, true, &ops, st);                          // plonk.rs:455
    PLK_CUDA(cudaMemcpyAsync(out_coeffs_8n, d_out.p, m * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
// the plan's subgroup w^k, k < size (self.subgroup_n / subgroup_8n of the reference's Circuit, plonk.rs:47-51)
int plk_fft_subgroup(const plk_fft_plan* pc, uint64_t* out) {
  return guarded([&] {
    check_plan(pc);
    if (!out) fail(PLK_EINVAL, "NULL buffer");
    auto* pl = const_cast<plk_fft_plan*>(pc);
    cudaStream_t st = thread_stream();
    const void* sub = plan_subgroup(pl, st);
    PLK_CUDA(cudaMemcpyAsync(out, sub, pl->n * pl->elem_bytes, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

int plk_fft_set_direct_log(plk_fft_plan* p, int log2_entries) {
  return guarded([&] {
    check_plan(p);
    if (log2_entries < 0 || log2_entries > 30) fail(PLK_EINVAL, "direct table cap out of range");
    std::lock_guard<std::mutex> lk(p->mu);
    p->direct_log = log2_entries;
  });
}
size_t plk_fft_size(const plk_fft_plan* p) { return p ? p->n : 0; }
int plk_fft_num_passes(const plk_fft_plan* p) { return p ? p->m : 0; }
int plk_fft_last_pass_ms(const plk_fft_plan* pc, float* out_ms, int cap) {
  int n = 0;
  int rc = guarded([&] {
    auto* p = const_cast<plk_fft_plan*>(pc);
    if (!p || !out_ms) fail(PLK_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> lk(p->timer_mu);
    n = p->timer.read(out_ms, cap);
  });
  return rc == PLK_OK ? n : -rc;
}

void plk_fft_free(plk_fft_plan* p) { delete p; }

int plk_fft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");                     // util.rs:16-19
    if (n != p->n) fail(PLK_ESIZE, "Number of coefficients does not match size of subgroup in precomputation");
    host_transform(const_cast<plk_fft_plan*>(p), in, n, 1, false, FusedOps(), out);
  });
}

int plk_ifft_pow2(const plk_fft_plan* p, const uint64_t* in, uint64_t* out, size_t n) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");
    if (n != p->n) fail(PLK_ESIZE, "Number of points does not match size of subgroup in precomputation");
    host_transform(const_cast<plk_fft_plan*>(p), in, n, 1, true, FusedOps(), out);
  });
}

int plk_fft(const plk_fft_plan* p, const uint64_t* in, size_t n_in, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if ((!in && n_in) || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n) fail(PLK_ESIZE, "more coefficients than the precomputation size");
    // fft.rs:61-80 pads to the next power of two of n_in, which must be the plan size
    if (((size_t)1 << log2_ceil(n_in ? n_in : 1)) != p->n) fail(PLK_ESIZE, "padded length does not match precomputation size");
    host_transform(const_cast<plk_fft_plan*>(p), in, n_in, 1, false, FusedOps(), out);
  });
}

int plk_fft_batch(const plk_fft_plan* p, const uint64_t* in, size_t n_in, size_t k, int inverse, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad row length");
    if (inverse && n_in != p->n) fail(PLK_ESIZE, "inverse transform needs full rows");
    if (k == 0) return;
    host_transform(const_cast<plk_fft_plan*>(p), in, n_in, k, inverse != 0, FusedOps(), out);
  });
}

int plk_coset_lde(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, const uint64_t* shift, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!coeffs || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad coefficient count");
    FusedOps ops;
    auto* pl = const_cast<plk_fft_plan*>(p);
    ops_for(pl->field)->coset(pl, shift, false, &ops, thread_stream());
    host_transform(pl, coeffs, n_in, 1, false, ops, out);
  });
}

int plk_coset_ifft(const plk_fft_plan* p, const uint64_t* evals, const uint64_t* shift, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!evals || !out) fail(PLK_EINVAL, "NULL buffer");
    FusedOps ops;
    auto* pl = const_cast<plk_fft_plan*>(p);
    ops_for(pl->field)->coset(pl, shift, true, &ops, thread_stream());
    host_transform(pl, evals, pl->n, 1, true, ops, out);
  });
}

int plk_divide_by_z_h(const plk_fft_plan* p, const uint64_t* coeffs, size_t n_in, size_t n_gates, uint64_t* out) {
  return guarded([&] {
    check_plan(p);
    if (!coeffs || !out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad coefficient count");
    if (!is_pow2(n_gates) || n_gates > p->n) fail(PLK_EINVAL, "n_gates must be a power of two dividing the plan size");
    auto* pl = const_cast<plk_fft_plan*>(p);
    cudaStream_t st = thread_stream();
    FusedOps fwd, bwd;
    ops_for(pl->field)->coset(pl, nullptr, false, &fwd, st);
    ops_for(pl->field)->zh_table(pl, n_gates, &fwd, st);
    ops_for(pl->field)->coset(pl, nullptr, true, &bwd, st);
    const size_t eb = pl->elem_bytes;
    void* d_in = thread_scratch(0, n_in * eb);
    void* d_mid = thread_scratch(1, pl->n * eb);
    void* d_out = thread_scratch(2, pl->n * eb);
    PLK_CUDA(cudaMemcpyAsync(d_in, coeffs, n_in * eb, cudaMemcpyHostToDevice, st));
    ops_for(pl->field)->run(pl, d_in, n_in, n_in, d_mid, 1, false, &fwd, st);
    ops_for(pl->field)->run(pl, d_mid, pl->n, pl->n, d_out, 1, true, &bwd, st);
    PLK_CUDA(cudaMemcpyAsync(out, d_out, pl->n * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

/* Polynomial::mul (src/polynomial.rs:209-227): both operands zero-padded to deg a + deg b + 1, FFT of the size
 * 2^log2_ceil(that), pointwise product, IFFT -- here two fused-padding transforms, one product kernel and the inverse
 * transform without leaving the device.  The reference rebuilds the FftPrecomputation on every call (:217); the plans
 * are cached per (field, log size, device) instead. */
int plk_poly_mul(int field, const uint64_t* a, size_t na, const uint64_t* b, size_t nb, uint64_t* out, size_t out_cap, size_t* out_len) {
  return guarded([&] {
    const int L = plk_field_limbs(field);
    if (!L) fail(PLK_EINVAL, "unknown field id");
    if ((na && !a) || (nb && !b) || !out || !out_len) fail(PLK_EINVAL, "NULL buffer");
    auto degree_plus_one = [L](const uint64_t* p, size_t n) {          // Polynomial::degree after trim (:105-118)
      while (n > 0) {
        bool zero = true;
        for (int j = 0; j < L; ++j) zero = zero && p[(n - 1) * L + j] == 0;
        if (!zero) break;
        --n;
      }
      return n;
    };
    const size_t la = degree_plus_one(a, na), lb = degree_plus_one(b, nb);
    if (la == 0 || lb == 0) {                                          // is_zero: Self::zero(1)
      if (out_cap < 1) fail(PLK_ESIZE, "output buffer too small");
      for (int j = 0; j < L; ++j) out[j] = 0;
      *out_len = 1;
      return;
    }
    const size_t size = la + lb - 1;                                   // a_deg + b_deg + 1
    const int ta = field_two_adicity(field);
    const int lg = log2_ceil(size);
    if (lg > ta) fail(PLK_ETOOBIG, "log2(size) exceeds TWO_ADICITY");
    const size_t n = (size_t)1 << lg;
    if (out_cap < n) fail(PLK_ESIZE, "output buffer too small (needs 2^log2_ceil(deg a + deg b + 1) elements)");
    // plan cache
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, plk_fft_plan*> cache;
    int dev = 0;
    PLK_CUDA(cudaGetDevice(&dev));
    // plans up to 2^20 are cached for the life of the process (a few MiB each: direct twiddle tables are built lazily and
    // only in the directions used); larger products build their plan for this call and free it afterwards
    plk_fft_plan* pl = nullptr;
    std::unique_ptr<plk_fft_plan> transient;
    if (lg <= 20) {
      std::lock_guard<std::mutex> lk(mu);
      auto key = std::make_tuple(field, lg, dev);
      auto it = cache.find(key);
      if (it == cache.end()) {
        plk_fft_plan* fresh = nullptr;
        const int rc = plk_fft_precompute(field, n, &fresh);
        if (rc != PLK_OK) fail(rc, plk_last_error_message());
        it = cache.emplace(key, fresh).first;
      }
      pl = it->second;
    } else {
      plk_fft_plan* fresh = nullptr;
      const int rc = plk_fft_precompute(field, n, &fresh);
      if (rc != PLK_OK) fail(rc, plk_last_error_message());
      transient.reset(fresh);
      pl = fresh;
    }
    cudaStream_t st = thread_stream();
    const size_t eb = pl->elem_bytes;
    char* d_in = reinterpret_cast<char*>(thread_scratch(0, (la + lb) * eb));
    char* d_ev = reinterpret_cast<char*>(thread_scratch(1, 2 * n * eb));
    char* d_out = reinterpret_cast<char*>(thread_scratch(2, n * eb));
    PLK_CUDA(cudaMemcpyAsync(d_in, a, la * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_in + la * eb, b, lb * eb, cudaMemcpyHostToDevice, st));
    const FusedOps none;
    ops_for(field)->run(pl, d_in, la, la, d_ev, 1, false, &none, st);
    ops_for(field)->run(pl, d_in + la * eb, lb, lb, d_ev + n * eb, 1, false, &none, st);
    plk_launch_field_mul(field, d_ev, d_ev + n * eb, d_ev, n, st);
    ops_for(field)->run(pl, d_ev, n, n, d_out, 1, true, &none, st);
    PLK_CUDA(cudaMemcpyAsync(out, d_out, n * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    *out_len = n;
  });
}

int plk_fft_dev(const plk_fft_plan* p, const void* d_in, size_t n_in, size_t k, unsigned flags, void* d_out, void* stream) {
  return guarded([&] {
    check_plan(p);
    if (!d_in || !d_out) fail(PLK_EINVAL, "NULL buffer");
    if (n_in > p->n || n_in == 0) fail(PLK_ESIZE, "bad row length");
    const bool inverse = (flags & PLK_FFT_INVERSE) != 0;
    if (inverse && n_in != p->n) fail(PLK_ESIZE, "inverse transform needs full rows");
    if (k == 0) return;
    auto* pl = const_cast<plk_fft_plan*>(p);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FusedOps ops;
    if (flags & PLK_FFT_COSET) ops_for(pl->field)->coset(pl, nullptr, inverse, &ops, st);
    ops_for(pl->field)->run(pl, d_in, n_in, n_in, d_out, k, inverse, &ops, st);
  });
}

/* ---- domain-split transform over `world` ranks ---------------------------------
 * N = R1 * M.  Rank r holds the rows j_1 in [row_base, row_base + rows) of the matrix x[j_1 + R1 j'] (rows
 * of length M).  Phase A: size-M transforms of the local rows, times w_N^(j_1 k'), written in the
 * all-to-all send layout [dest][row][k' mod (M / world)].  After the exchange rank s holds
 * Z[j_1][kl] for all j_1 and its M / world columns; phase B transforms over j_1 in place and leaves
 * X[k' + M k_1] at [k_1][kl], k' = s M / world + kl. */
int plk_fft_dist_phase_a(const plk_fft_plan* plan_m, const plk_fft_plan* plan_n, const void* d_in, size_t rows, size_t row_base,
                         unsigned world, unsigned flags, void* d_work, void* d_send, void* stream) {
  return guarded([&] {
    check_plan(plan_m);
    check_plan(plan_n);
    if (!d_in || !d_work || !d_send || rows == 0 || world == 0) fail(PLK_EINVAL, "bad arguments");
    if (plan_m->field != plan_n->field || plan_m->log_n > plan_n->log_n) fail(PLK_EINVAL, "plans do not match");
    if (!is_pow2(world) || world > plan_m->n) fail(PLK_EINVAL, "world must be a power of two <= M");
    if (rows > 65535) fail(PLK_EINVAL, "too many local rows");
    const bool inverse = (flags & PLK_FFT_INVERSE) != 0;
    FusedOps ops;
    ops.post_lo = plan_n->tw_lo[inverse ? 1 : 0].p;
    ops.post_hi = plan_n->tw_hi[inverse ? 1 : 0].p;
    ops.post_lo_bits = plan_n->lo_bits;
    ops.post_rowmul = 1;
    ops.post_row_base = row_base;
    ops.final_out = d_send;
    ops.remap = 1;
    ops.remap_cl_log = plan_m->log_n - log2_floor(world);
    ops.remap_rows = rows;
    ops_for(plan_m->field)->run(plan_m, d_in, plan_m->n, plan_m->n, d_work, rows, inverse, &ops, reinterpret_cast<cudaStream_t>(stream));
  });
}
// Phase A with the exchange folded into the kernel: the last pass stores every value straight into the receive buffer
// of the GPU that owns its column block (peer_recv[d] = that GPU's buffer, mapped here with plk_ipc_open; this rank's own
// buffer for d == rank), so no send buffer and no all-to-all exist.  The caller orders "all peers finished phase A" before
// phase B (one tiny all-reduce on the stream) and alternates between two receive buffers per transform.
int plk_fft_dist_phase_a_p2p(const plk_fft_plan* plan_m, const plk_fft_plan* plan_n, const void* d_in, size_t rows, size_t row_base,
                             unsigned world, unsigned flags, void* d_work, void* const* peer_recv, void* stream) {
  return guarded([&] {
    check_plan(plan_m);
    check_plan(plan_n);
    if (!d_in || !d_work || !peer_recv || rows == 0 || world == 0) fail(PLK_EINVAL, "bad arguments");
    if (plan_m->field != plan_n->field || plan_m->log_n > plan_n->log_n) fail(PLK_EINVAL, "plans do not match");
    if (!is_pow2(world) || world > plan_m->n || world > (unsigned)kMaxPeers) fail(PLK_EINVAL, "world must be a power of two <= min(M, 8)");
    if (rows > 65535) fail(PLK_EINVAL, "too many local rows");
    const bool inverse = (flags & PLK_FFT_INVERSE) != 0;
    FusedOps ops;
    ops.post_lo = plan_n->tw_lo[inverse ? 1 : 0].p;
    ops.post_hi = plan_n->tw_hi[inverse ? 1 : 0].p;
    ops.post_lo_bits = plan_n->lo_bits;
    ops.post_rowmul = 1;
    ops.post_row_base = row_base;
    ops.remap = 2;
    ops.remap_cl_log = plan_m->log_n - log2_floor(world);
    ops.remap_rows = rows;
    ops.remap_row_base = row_base;
    for (unsigned d = 0; d < world; ++d) {
      if (!peer_recv[d]) fail(PLK_EINVAL, "NULL peer buffer");
      ops.peer[d] = peer_recv[d];
    }
    ops_for(plan_m->field)->run(plan_m, d_in, plan_m->n, plan_m->n, d_work, rows, inverse, &ops, reinterpret_cast<cudaStream_t>(stream));
  });
}
// Device buffers that other processes of the node can map (cudaIpc*): the receive buffers of the domain-split transform.
int plk_ipc_alloc(size_t bytes, void** d_ptr, uint8_t handle[64]) {
  return guarded([&] {
    if (!d_ptr || !handle || bytes == 0) fail(PLK_EINVAL, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PLK_CUDA(cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; throw ::plk::CudaError{e, "cudaIpcGetMemHandle", __FILE__, __LINE__}; }
    memcpy(handle, &h, 64);
  });
}
int plk_ipc_open(const uint8_t handle[64], void** d_ptr) {
  return guarded([&] {
    if (!d_ptr || !handle) fail(PLK_EINVAL, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    PLK_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}
int plk_ipc_close(void* d_ptr) { return guarded([&] { if (d_ptr) PLK_CUDA(cudaIpcCloseMemHandle(d_ptr)); }); }
int plk_ipc_free(void* d_ptr) { return guarded([&] { if (d_ptr) PLK_CUDA(cudaFree(d_ptr)); }); }
int plk_copy_dev(void* d_dst, const void* d_src, size_t bytes, void* stream) {
  return guarded([&] {
    if (bytes && (!d_dst || !d_src)) fail(PLK_EINVAL, "NULL buffer");
    PLK_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  });
}
int plk_fft_dist_phase_b(const plk_fft_plan* plan_n, void* d_recv, unsigned log_r1, unsigned log_cols, unsigned flags, void* stream) {
  return guarded([&] {
    check_plan(plan_n);
    if (!d_recv || log_r1 > 8) fail(PLK_EINVAL, "bad arguments");
    const bool inverse = (flags & PLK_FFT_INVERSE) != 0;
    // phase A already divided by M (inverse); the remaining factor is R1^-1
    const void* scale = inverse ? (const char*)plan_n->pow2_inv.p + (size_t)log_r1 * plan_n->elem_bytes : nullptr;
    ops_for(plan_n->field)->final_pass(plan_n, d_recv, (int)log_r1, (int)log_cols, inverse, scale, reinterpret_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
