// C ABI of the device-resident Halo IPA rounds (include/plonky_b200.h, section "Halo inner-product argument").
#include <memory>
#include <string.h>
#include "ipa_kernels.cuh"
#include "msm_plan.h"

using namespace plk;

namespace plk {
const IpaOps* ipa_ops_tweedledee();
const IpaOps* ipa_ops_tweedledum();
const IpaOps* ipa_ops_bls12_377();
}

struct plk_ipa_state {
  int curve = 0;
  size_t n = 0;            // current length of halo_a / halo_b / halo_g (halves every fold)
  size_t L = 4;            // u64 limbs of a base-field element
  DevBuf a, b, g;          // n scalars (32 B), n scalars, n affine points (2 * L * 8 B, identity = (0, 0))
  DevBuf partials, io;     // inner-product partials; io: 2 points xyz | 2 flags (8 B each) | 2 inner products | u, u^-1
  // table mode (plk_ipa_new_with_table): G is never folded, see ipa_kernels.cuh
  plk_msm_table* table = nullptr;   // borrowed: fixed-base table of the n0 original generators
  size_t n0 = 0;                    // original length
  DevBuf coef, sl_sr;               // c_i (n0 scalars); expanded scalars of the L and R MSMs (2 * n0)
};

namespace {
const IpaOps* ipa_ops_for(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return ipa_ops_tweedledee();
    case PLK_CURVE_TWEEDLEDUM: return ipa_ops_tweedledum();
    case PLK_CURVE_BLS12_377: return ipa_ops_bls12_377();
  }
  fail(PLK_EINVAL, "unknown curve id");
}
// identity points arrive as zero flags: force their coordinates to (0, 0), the device encoding
__global__ void ipa_apply_zero_flags(uint4* g, const unsigned char* zero, unsigned long long n, unsigned vec_per_point) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n || !zero[i]) return;
  for (unsigned k = 0; k < vec_per_point; ++k) g[i * vec_per_point + k] = make_uint4(0, 0, 0, 0);
}
__global__ void ipa_zero_flags_of(const uint4* g, unsigned char* zero, unsigned long long n, unsigned vec_per_point) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned acc = 0;
  for (unsigned k = 0; k < vec_per_point; ++k) { const uint4 v = g[i * vec_per_point + k]; acc |= v.x | v.y | v.z | v.w; }
  zero[i] = acc == 0 ? 1 : 0;
}
}  // namespace

extern "C" {

int plk_ipa_new(int curve, const uint64_t* a, const uint64_t* b, const uint64_t* g_xy, const uint8_t* g_zero, size_t n,
                plk_ipa_state** out) {
  return guarded([&] {
    if (!out) fail(PLK_EINVAL, "NULL out");
    *out = nullptr;
    const int bf = plk_curve_base_field(curve);
    if (bf < 0) fail(PLK_EINVAL, "unknown curve id");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");            // log2_strict(degree), halo.rs:62
    if (!a || !b || !g_xy) fail(PLK_EINVAL, "NULL buffer");
    std::unique_ptr<plk_ipa_state> s(new plk_ipa_state());
    s->curve = curve;
    s->n = n;
    s->L = (size_t)plk_field_limbs(bf);
    cudaStream_t st = thread_stream();
    const size_t pb = 2 * s->L * 8;
    s->a.alloc(n * 32);
    s->b.alloc(n * 32);
    s->g.alloc(n * pb);
    s->partials.alloc(2 * kIpaMaxBlocks * 32);
    s->io.alloc(1024);
    PLK_CUDA(cudaMemcpyAsync(s->a.p, a, n * 32, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(s->b.p, b, n * 32, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(s->g.p, g_xy, n * pb, cudaMemcpyHostToDevice, st));
    if (g_zero) {
      DevBuf dz(n, st);
      PLK_CUDA(cudaMemcpyAsync(dz.p, g_zero, n, cudaMemcpyHostToDevice, st));
      ipa_apply_zero_flags<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->g.as<uint4>(), dz.as<unsigned char>(), n, (unsigned)(pb / 16));
      PLK_LAUNCHED();
    }
    PLK_CUDA(cudaStreamSynchronize(st));
    *out = s.release();
  });
}
int plk_ipa_new_with_table(const plk_msm_table* table, const uint64_t* a, const uint64_t* b, size_t n, plk_ipa_state** out) {
  return guarded([&] {
    if (!out) fail(PLK_EINVAL, "NULL out");
    *out = nullptr;
    if (!table) fail(PLK_EINVAL, "NULL table");
    if (table->g.variable) fail(PLK_EINVAL, "a fixed-base table is required");
    if (n != table->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");
    if (!is_pow2(n)) fail(PLK_ENOTPOW2, "Not a power of two");            // log2_strict(degree), halo.rs:62
    if (!a || !b) fail(PLK_EINVAL, "NULL buffer");
    std::unique_ptr<plk_ipa_state> s(new plk_ipa_state());
    s->curve = table->curve;
    s->n = s->n0 = n;
    s->L = (size_t)plk_field_limbs(plk_curve_base_field(table->curve));
    s->table = const_cast<plk_msm_table*>(table);
    cudaStream_t st = thread_stream();
    s->a.alloc(n * 32);
    s->b.alloc(n * 32);
    s->coef.alloc(n * 32);
    s->sl_sr.alloc(2 * n * 32);
    s->partials.alloc(2 * kIpaMaxBlocks * 32);
    s->io.alloc(1024);
    PLK_CUDA(cudaMemcpyAsync(s->a.p, a, n * 32, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(s->b.p, b, n * 32, cudaMemcpyHostToDevice, st));
    ipa_ops_for(s->curve)->init_coef(s->coef.p, n, st);
    PLK_CUDA(cudaStreamSynchronize(st));
    *out = s.release();
  });
}
size_t plk_ipa_len(const plk_ipa_state* s) { return s ? s->n : 0; }
void plk_ipa_free(plk_ipa_state* s) { delete s; }

int plk_ipa_round_lr(plk_ipa_state* s, uint64_t* out_l_xyz, uint8_t* out_l_zero, uint64_t* out_r_xyz, uint8_t* out_r_zero,
                     uint64_t* out_ip_l, uint64_t* out_ip_r) {
  return guarded([&] {
    if (!s) fail(PLK_EINVAL, "NULL state");
    if (!out_l_xyz || !out_l_zero || !out_r_xyz || !out_r_zero || !out_ip_l || !out_ip_r) fail(PLK_EINVAL, "NULL buffer");
    if (s->n < 2) fail(PLK_EINVAL, "no round left: the vectors have length 1");
    cudaStream_t st = thread_stream();
    const size_t half = s->n / 2, pb = 2 * s->L * 8, xyz = 3 * s->L * 8;
    char* io = s->io.as<char>();
    char* d_l = io;
    char* d_r = io + xyz;
    char* d_flags = io + 2 * xyz;             // 2 x 8 bytes
    char* d_ip = io + 2 * xyz + 16;           // 2 x 32 bytes
    const char* a = s->a.as<char>();
    const char* g = s->g.as<char>();
    if (s->table) {
      // both MSMs against the fixed-base table of the original generators, forked over the table's side streams
      ipa_ops_for(s->curve)->expand_scalars(s->a.p, s->coef.p, s->n0, s->n, s->sl_sr.p, st);
      msm_execute_batch_on(s->table, s->sl_sr.p, 2, d_l, d_flags, st);                 // d_l, d_r adjacent; flags at +0, +1
    } else {
      msm_variable_dev(s->curve, g + half * pb, a, half, d_l, d_flags, st);            // <a_lo, G_hi>  (halo.rs:87)
      msm_variable_dev(s->curve, g, a + half * 32, half, d_r, d_flags + 8, st);        // <a_hi, G_lo>  (halo.rs:91)
    }
    ipa_ops_for(s->curve)->inner_products(s->a.p, s->b.p, half, s->partials.p, d_ip, st);
    // the io block is contiguous (L | R | flags | inner products): one device-to-host copy per round
    const size_t io_bytes = 2 * xyz + 16 + 64;
    uint8_t host_io[2 * 3 * 6 * 8 + 16 + 64];
    PLK_CUDA(cudaMemcpyAsync(host_io, io, io_bytes, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    memcpy(out_l_xyz, host_io, xyz);
    memcpy(out_r_xyz, host_io + xyz, xyz);
    const uint8_t* flags = host_io + 2 * xyz;
    memcpy(out_ip_l, host_io + 2 * xyz + 16, 32);
    memcpy(out_ip_r, host_io + 2 * xyz + 48, 32);
    *out_l_zero = flags[0];
    *out_r_zero = s->table ? flags[1] : flags[8];
  });
}

int plk_ipa_fold(plk_ipa_state* s, const uint64_t* u, const uint64_t* u_inv) {
  return guarded([&] {
    if (!s || !u || !u_inv) fail(PLK_EINVAL, "NULL argument");
    if (s->n < 2) fail(PLK_EINVAL, "no round left: the vectors have length 1");
    cudaStream_t st = thread_stream();
    const size_t half = s->n / 2;
    char* d_uu = s->io.as<char>() + 512;
    uint64_t host_uu[8];
    memcpy(host_uu, u, 32);
    memcpy(host_uu + 4, u_inv, 32);
    PLK_CUDA(cudaMemcpyAsync(d_uu, host_uu, 64, cudaMemcpyHostToDevice, st));
    ipa_ops_for(s->curve)->fold(s->a.p, s->b.p, s->table ? nullptr : s->g.p, half, d_uu, st);
    if (s->table) ipa_ops_for(s->curve)->update_coef(s->coef.p, s->n0, s->n, d_uu, st);
    PLK_CUDA(cudaStreamSynchronize(st));     // u / u_inv are the caller's stack memory
    s->n = half;
  });
}

int plk_ipa_read(const plk_ipa_state* s, uint64_t* a, uint64_t* b, uint64_t* g_xy, uint8_t* g_zero) {
  return guarded([&] {
    if (!s) fail(PLK_EINVAL, "NULL state");
    cudaStream_t st = thread_stream();
    const size_t pb = 2 * s->L * 8;
    if (a) PLK_CUDA(cudaMemcpyAsync(a, s->a.p, s->n * 32, cudaMemcpyDeviceToHost, st));
    if (b) PLK_CUDA(cudaMemcpyAsync(b, s->b.p, s->n * 32, cudaMemcpyDeviceToHost, st));
    if (s->table && (g_xy || g_zero)) {
      // table mode: the folded generators exist only as coefficients; halo_g = sum_i c_i G_i once the length is 1
      if (s->n != 1) fail(PLK_EINVAL, "table mode materialises halo_g only at length 1");
      if (!g_xy || !g_zero) fail(PLK_EINVAL, "g_xy and g_zero are read together");
      char* io = s->io.as<char>();
      msm_execute_batch_on(s->table, s->coef.p, 1, io, io + 2 * 3 * s->L * 8, st);
      PLK_CUDA(cudaMemcpyAsync(g_xy, io, pb, cudaMemcpyDeviceToHost, st));      // normalised: (x, y) are the affine coordinates
      PLK_CUDA(cudaMemcpyAsync(g_zero, io + 2 * 3 * s->L * 8, 1, cudaMemcpyDeviceToHost, st));
      PLK_CUDA(cudaStreamSynchronize(st));
      return;
    }
    if (g_xy) PLK_CUDA(cudaMemcpyAsync(g_xy, s->g.p, s->n * pb, cudaMemcpyDeviceToHost, st));
    if (g_zero) {
      DevBuf dz(s->n, st);
      ipa_zero_flags_of<<<(unsigned)((s->n + 255) / 256), 256, 0, st>>>(s->g.as<uint4>(), dz.as<unsigned char>(), s->n, (unsigned)(pb / 16));
      PLK_LAUNCHED();
      PLK_CUDA(cudaMemcpyAsync(g_zero, dz.p, s->n, cudaMemcpyDeviceToHost, st));
      PLK_CUDA(cudaStreamSynchronize(st));
      return;
    }
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
