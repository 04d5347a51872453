// Process-wide plumbing of the C ABI (status strings, error capture, per-thread streams and scratch)
// and the small elementwise entry points: plk_field_op, plk_batch_inverse.
#include <mutex>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "fp.cuh"

namespace plk {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_profiling{0};

static thread_local std::string t_last_error;
void set_last_error(const std::string& s) { t_last_error = s; }

struct ThreadCtx {
  cudaStream_t stream = nullptr;
  int device = -1;
  DevBuf scratch[8];
  ~ThreadCtx() {
    // the CUDA context may already be gone at thread/process teardown: never throw from here
    for (auto& b : scratch) { if (b.p) { cudaFree(b.p); b.p = nullptr; } }
    if (stream) cudaStreamDestroy(stream);
  }
};
static thread_local ThreadCtx t_ctx;

static void ensure_ctx() {
  int dev = 0;
  PLK_CUDA(cudaGetDevice(&dev));
  if (t_ctx.stream && t_ctx.device == dev) return;
  if (t_ctx.stream) {
    for (auto& b : t_ctx.scratch) b.release();
    cudaStreamDestroy(t_ctx.stream);
    t_ctx.stream = nullptr;
  }
  PLK_CUDA(cudaStreamCreateWithFlags(&t_ctx.stream, cudaStreamNonBlocking));
  t_ctx.device = dev;
}
// The library's OWN stream-ordered memory pool, one per device (never the device's default pool, which other users of the
// process -- e.g. PyTorch with the cudaMallocAsync backend -- share): freed temporaries stay cached up to 1 GiB per device.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};
cudaMemPool_t async_pool() {
  int dev = 0;
  PLK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) fail(PLK_EINVAL, "device index out of range");
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    PLK_CUDA(cudaMemPoolCreate(&g_pools[dev], &props));
    unsigned long long keep = 1ull << 30;
    cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return g_pools[dev];
}
cudaStream_t thread_stream() {
  ensure_ctx();
  return t_ctx.stream;
}
void* thread_scratch(int slot, size_t bytes) {
  ensure_ctx();
  if (slot < 0 || slot >= 8) fail(PLK_EINVAL, "bad scratch slot");
  DevBuf& b = t_ctx.scratch[slot];
  if (bytes > b.bytes) {
    // growing: nothing on this thread's stream may still use the old block
    if (b.p) PLK_CUDA(cudaStreamSynchronize(t_ctx.stream));
    b.alloc(bytes + bytes / 8);
  }
  return b.p;
}

// ---- elementwise kernels ----------------------------------------------------------------------
template <class P>
__global__ void field_op_kernel(int op, const void* a, const void* b, void* out, unsigned long long n) {
  typedef Fp<P> F;
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = load_fp<F>(a, i), y = b ? load_fp<F>(b, i) : F::zero(), r;
  switch (op) {
    case 0: r = F::add(x, y); break;
    case 1: r = F::sub(x, y); break;
    case 2: r = F::mul(x, y); break;
    case 3: r = F::sqr(x); break;
    case 4: r = F::neg(x); break;
    case 5: r = F::inverse(x); break;
    case 6: r = F::to_canonical(x); break;
    case 7: r = F::from_canonical(x); break;
    case 9: r = F::inverse_gcd(x); break;
    default: r = F::dbl(x); break;
  }
  store_fp<F>(out, i, r);
}

// Montgomery's trick (field.rs:251-278) per thread over a strip of `strip` consecutive elements:
// prefix products, one inversion, back substitution.  zero_flag is set if any input is zero.
template <class P>
__global__ void batch_inverse_kernel(const void* in, void* out, unsigned long long n, unsigned strip, int* zero_flag) {
  typedef Fp<P> F;
  unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long lo = t * strip;
  if (lo >= n) return;
  unsigned long long hi = lo + strip < n ? lo + strip : n;
  F acc = F::one();
  for (unsigned long long i = lo; i < hi; ++i) {
    F x = load_fp<F>(in, i);
    if (x.is_zero()) { *zero_flag = 1; return; }
    store_fp<F>(out, i, acc);          // prefix product of the elements before i
    acc = F::mul(acc, x);
  }
  F inv = F::inverse(acc);
  for (unsigned long long i = hi; i-- > lo;) {
    F x = load_fp<F>(in, i);
    F pre = load_fp<F>(out, i);
    store_fp<F>(out, i, F::mul(inv, pre));
    inv = F::mul(inv, x);
  }
}

// Throughput probe: every thread runs 4 independent chains of dependent Montgomery products (the ILP of a mixed
// point addition).  bench.py quotes the MSM / NTT kernels against this measured products/s ceiling next to the HBM
// roofline (tools/bench_mul.cu is the standalone version).
template <class P>
__global__ void __launch_bounds__(128) mul_throughput_kernel(int iters, uint32_t* out) {
  typedef Fp<P> F;
  F a[4], b = F::one();
  for (int k = 0; k < F::N - 1; ++k) {
    b.l[k] ^= threadIdx.x * 2654435761u + k;
    for (int c = 0; c < 4; ++c) a[c].l[k] = 0x9e3779b9u * (k + 1) + c + blockIdx.x;
  }
  for (int c = 0; c < 4; ++c) a[c].l[F::N - 1] = 1;          // operands stay below p
  b.l[F::N - 1] = 2;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = F::mul(a[c], b);
  }
  uint32_t acc = 0;
  for (int c = 0; c < 4; ++c) for (int k = 0; k < F::N; ++k) acc ^= a[c].l[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <class P>
void measure_mul_throughput(double* products_per_s, cudaStream_t st) {
  int dev = 0, sms = 0;
  PLK_CUDA(cudaGetDevice(&dev));
  PLK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 4, threads = 128, iters = 1000;
  DevBuf out((size_t)blocks * threads * 4, st);
  cudaEvent_t e0, e1;
  PLK_CUDA(cudaEventCreate(&e0));
  PLK_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {              // rep 0 is the warm-up
    PLK_CUDA(cudaEventRecord(e0, st));
    mul_throughput_kernel<P><<<blocks, threads, 0, st>>>(iters, out.as<uint32_t>());
    PLK_LAUNCHED();
    PLK_CUDA(cudaEventRecord(e1, st));
    PLK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PLK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *products_per_s = (double)blocks * threads * iters * 4 / (best * 1e-3);
}

template <class P>
void launch_field_op(int op, const void* a, const void* b, void* out, size_t n, cudaStream_t st) {
  if (n == 0) return;
  field_op_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(op, a, b, out, n);
  PLK_LAUNCHED();
}
template <class P>
void launch_batch_inverse(const void* in, void* out, size_t n, int* d_flag, cudaStream_t st) {
  if (n == 0) return;
  const unsigned strip = n >= (1u << 16) ? 16 : 4;
  const size_t threads = (n + strip - 1) / strip;
  batch_inverse_kernel<P><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(in, out, n, strip, d_flag);
  PLK_LAUNCHED();
}

// ---- permutation_polynomial (src/plonk_util.rs:233-262) ---------------------------------------------------------------
// Z(g^0) = 1, Z(g^i) = Z(g^(i-1)) * prod_j (w_j + beta k_j x + gamma) / prod_j (w_j + beta sigma_j + gamma) over the routed
// wires j of gate i - 1.  Device: one thread per gate for numerator and denominator, Montgomery's trick for the n
// divisions, then an exclusive prefix PRODUCT over the ratios (field multiplication is associative and exact, so the
// scan order cannot change a value).
struct PermArgs {
  const void* subgroup;     // n elements
  const void* wires;        // wire (i, j) at i * wire_stride + j
  const void* sigma;        // sigma_j at j * sigma_row + sigma_stride * i
  const void* kis;          // routed shifts k_j
  const void* beta_gamma;   // beta, gamma
  unsigned long long n, wire_stride, sigma_row, sigma_stride;
  unsigned routed;
};
template <class P>
__global__ void perm_ratio_terms_kernel(PermArgs a, void* __restrict__ num, void* __restrict__ den) {
  typedef Fp<P> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i + 1 >= a.n) return;                         // gates 0 .. n - 2 feed Z(g^1) .. Z(g^(n-1))
  const F beta = load_fp<F>(a.beta_gamma, 0), gamma = load_fp<F>(a.beta_gamma, 1);
  const F bx = F::mul(beta, load_fp<F>(a.subgroup, i));
  F nu = F::one(), de = F::one();
  for (unsigned j = 0; j < a.routed; ++j) {
    const F w = F::add(load_fp<F>(a.wires, i * a.wire_stride + j), gamma);
    nu = F::mul(nu, F::add(w, F::mul(load_fp<F>(a.kis, j), bx)));                                        // w + beta k_j x + gamma
    de = F::mul(de, F::add(w, F::mul(beta, load_fp<F>(a.sigma, j * a.sigma_row + a.sigma_stride * i))));   // w + beta sigma + gamma
  }
  store_fp<F>(num, i, nu);
  store_fp<F>(den, i, de);
}
constexpr int kScanThreads = 256, kScanPerThread = 4, kScanTile = kScanThreads * kScanPerThread;
// inclusive product scan of one tile of r = num * den_inv (computed on the fly); tile totals to `totals`
template <class P>
__global__ void __launch_bounds__(kScanThreads) perm_tile_scan_kernel(const void* __restrict__ num, const void* __restrict__ den_inv,
                                                                      unsigned long long m, void* __restrict__ scanned, void* __restrict__ totals) {
  typedef Fp<P> F;
  __shared__ uint4 sm[kScanThreads * (F::N / 4)];
  const unsigned long long base = (unsigned long long)blockIdx.x * kScanTile + (unsigned long long)threadIdx.x * kScanPerThread;
  F v[kScanPerThread];
  F run = F::one();
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    const unsigned long long i = base + e;
    const F r = i < m ? F::mul(load_fp<F>(num, i), load_fp<F>(den_inv, i)) : F::one();
    run = F::mul(run, r);
    v[e] = run;
  }
  // block-wide inclusive scan of the per-thread products (Hillis-Steele in shared memory)
  store_fp<F>(sm, threadIdx.x, run);
  __syncthreads();
  F acc = run;
  for (unsigned d = 1; d < kScanThreads; d <<= 1) {
    F other = F::one();
    if (threadIdx.x >= d) other = load_fp<F>(sm, threadIdx.x - d);
    __syncthreads();
    if (threadIdx.x >= d) acc = F::mul(acc, other);
    store_fp<F>(sm, threadIdx.x, acc);
    __syncthreads();
  }
  const F before = threadIdx.x ? load_fp<F>(sm, threadIdx.x - 1) : F::one();
#pragma unroll
  for (int e = 0; e < kScanPerThread; ++e) {
    const unsigned long long i = base + e;
    if (i < m) store_fp<F>(scanned, i, F::mul(before, v[e]));
  }
  if (threadIdx.x == kScanThreads - 1) store_fp<F>(totals, blockIdx.x, acc);
}
// single CTA: exclusive product scan of the tile totals, in place (tiles <= 4096 here: n <= 2^22)
template <class P>
__global__ void __launch_bounds__(1024) perm_totals_scan_kernel(void* __restrict__ totals, unsigned tiles) {
  typedef Fp<P> F;
  __shared__ uint4 sm[1024 * (F::N / 4)];
  const unsigned per = (tiles + 1023) / 1024;
  const unsigned lo = threadIdx.x * per;
  F run = F::one();
  for (unsigned i = lo; i < lo + per && i < tiles; ++i) run = F::mul(run, load_fp<F>(totals, i));
  store_fp<F>(sm, threadIdx.x, run);
  __syncthreads();
  F acc = run;
  for (unsigned d = 1; d < 1024; d <<= 1) {
    F other = F::one();
    if (threadIdx.x >= d) other = load_fp<F>(sm, threadIdx.x - d);
    __syncthreads();
    if (threadIdx.x >= d) acc = F::mul(acc, other);
    store_fp<F>(sm, threadIdx.x, acc);
    __syncthreads();
  }
  F before = threadIdx.x ? load_fp<F>(sm, threadIdx.x - 1) : F::one();
  for (unsigned i = lo; i < lo + per && i < tiles; ++i) {
    const F t = load_fp<F>(totals, i);
    store_fp<F>(totals, i, before);
    before = F::mul(before, t);
  }
}
// out[0] = 1, out[i + 1] = tile prefix * scanned[i]
template <class P>
__global__ void perm_finish_kernel(const void* __restrict__ scanned, const void* __restrict__ totals, unsigned long long m, void* __restrict__ out) {
  typedef Fp<P> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i == 0) store_fp<F>(out, 0, F::one());
  if (i >= m) return;
  store_fp<F>(out, i + 1, F::mul(load_fp<F>(totals, i / kScanTile), load_fp<F>(scanned, i)));
}
template <class P>
void launch_permutation(const PermArgs& a, void* d_num, void* d_den, void* d_inv, void* d_totals, int* d_flag, void* d_out, cudaStream_t st) {
  typedef Fp<P> F;
  const size_t m = a.n - 1;                        // ratios
  if (m == 0) {
    perm_finish_kernel<P><<<1, 32, 0, st>>>(d_num, d_totals, 0, d_out);
    PLK_LAUNCHED();
    return;
  }
  perm_ratio_terms_kernel<P><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(a, d_num, d_den);
  PLK_LAUNCHED();
  launch_batch_inverse<P>(d_den, d_inv, m, d_flag, st);
  const unsigned tiles = (unsigned)((m + kScanTile - 1) / kScanTile);
  perm_tile_scan_kernel<P><<<tiles, kScanThreads, 0, st>>>(d_num, d_inv, m, d_den /* reuse: scanned */, d_totals);
  PLK_LAUNCHED();
  perm_totals_scan_kernel<P><<<1, 1024, 0, st>>>(d_totals, tiles);
  PLK_LAUNCHED();
  perm_finish_kernel<P><<<(unsigned)((m + 127) / 128), 128, 0, st>>>(d_den, d_totals, m, d_out);
  PLK_LAUNCHED();
  (void)sizeof(F);
}

}  // namespace plk

using namespace plk;

template <class P>
static void field_modulus64(const uint64_t** out) {
  static uint64_t m[P::LIMBS / 2];
  for (int i = 0; i < P::LIMBS / 2; ++i) m[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
  *out = m;
}
#define PLK_FIELD_DISPATCH(field, FN, ...)                                         \
  switch (field) {                                                                 \
    case PLK_FIELD_TWEEDLEDEE_BASE: FN<TweedledeeBaseParams>(__VA_ARGS__); break;  \
    case PLK_FIELD_TWEEDLEDUM_BASE: FN<TweedledumBaseParams>(__VA_ARGS__); break;  \
    case PLK_FIELD_BLS12_377_SCALAR: FN<Bls12377ScalarParams>(__VA_ARGS__); break; \
    case PLK_FIELD_BLS12_377_BASE: FN<Bls12377BaseParams>(__VA_ARGS__); break;     \
    default: fail(PLK_EINVAL, "unknown field id");                                 \
  }

void plk_launch_field_mul(int field, const void* d_a, const void* d_b, void* d_out, size_t n, cudaStream_t st) {
  PLK_FIELD_DISPATCH(field, launch_field_op, 2, d_a, d_b, d_out, n, st);
}
void plk_launch_field_inverse(int field, const void* d_in, void* d_out, size_t n, cudaStream_t st) {
  PLK_FIELD_DISPATCH(field, launch_field_op, 5, d_in, nullptr, d_out, n, st);
}

extern "C" {

const char* plk_status_string(int status) {
  switch (status) {
    case PLK_OK: return "ok";
    case PLK_EINVAL: return "invalid argument";
    case PLK_ELENGTH: return "precomputation / scalars length mismatch";
    case PLK_ENOTPOW2: return "Not a power of two";
    case PLK_ESIZE: return "length does not match the precomputation size";
    case PLK_EZERO: return "No inverse";
    case PLK_ECUDA: return "CUDA error";
    case PLK_ENOMEM: return "out of memory";
    case PLK_ETOOBIG: return "size exceeds the field's two-adicity";
  }
  return "unknown status";
}
const char* plk_last_error_message(void) { return t_last_error.c_str(); }
int plk_abi_version(void) { return PLK_ABI_VERSION; }
int plk_device_count(int* count) {
  return guarded([&] {
    if (!count) fail(PLK_EINVAL, "NULL count");
    PLK_CUDA(cudaGetDeviceCount(count));
  });
}
int plk_set_device(int device) {
  return guarded([&] { PLK_CUDA(cudaSetDevice(device)); });
}
int plk_field_limbs(int field) {
  switch (field) {
    case PLK_FIELD_TWEEDLEDEE_BASE:
    case PLK_FIELD_TWEEDLEDUM_BASE:
    case PLK_FIELD_BLS12_377_SCALAR: return 4;
    case PLK_FIELD_BLS12_377_BASE: return 6;
  }
  return 0;
}
int plk_curve_base_field(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return PLK_FIELD_TWEEDLEDEE_BASE;
    case PLK_CURVE_TWEEDLEDUM: return PLK_FIELD_TWEEDLEDUM_BASE;
    case PLK_CURVE_BLS12_377: return PLK_FIELD_BLS12_377_BASE;
  }
  return -1;
}
int plk_curve_scalar_field(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return PLK_FIELD_TWEEDLEDUM_BASE;
    case PLK_CURVE_TWEEDLEDUM: return PLK_FIELD_TWEEDLEDEE_BASE;
    case PLK_CURVE_BLS12_377: return PLK_FIELD_BLS12_377_SCALAR;
  }
  return -1;
}
uint64_t plk_kernel_launch_count(void) { return g_launches.load(); }
int plk_set_profiling(int enabled) { g_profiling.store(enabled ? 1 : 0); return PLK_OK; }

int plk_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  return guarded([&] {
    const int L = plk_field_limbs(field);
    if (!L) fail(PLK_EINVAL, "unknown field id");
    if (op < 0 || op > 9) fail(PLK_EINVAL, "unknown op");
    if (n == 0) return;
    if (!a || !out || ((op <= 2) && !b)) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t bytes = n * L * 8;
    void* da = thread_scratch(0, bytes);
    void* db = thread_scratch(1, bytes);
    void* dout = thread_scratch(2, bytes);
    PLK_CUDA(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, st));
    if (b) PLK_CUDA(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, st));
    if (op == 5 || op == 9) {
      // inverse of zero is an error in the reference (field.rs:159-165 returns None, Div panics)
      for (size_t i = 0; i < n; ++i) {
        bool z = true;
        for (int j = 0; j < L; ++j) z = z && a[i * L + j] == 0;
        if (z) fail(PLK_EZERO, "No inverse");
      }
    }
    PLK_FIELD_DISPATCH(field, launch_field_op, op, da, b ? db : nullptr, dout, n, st);
    PLK_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

// Field ToBytes / FromBytes (src/serialization.rs:17-30): to_canonical_u8_vec = the canonical value as 8*L little-endian
// bytes; from_canonical_u8_vec fails with "Out of range" for values >= the modulus (src/field/field.rs, the
// `from_canonical_u64_vec` check of every field).  The Montgomery <-> canonical conversions run on the device.
int plk_field_to_bytes(int field, const uint64_t* in, size_t n, uint8_t* out) {
  // little-endian host: the canonical limbs ARE the bytes
  return plk_field_op(field, 6, in, nullptr, reinterpret_cast<uint64_t*>(out), n);
}
int plk_field_from_bytes(int field, const uint8_t* in, size_t n, uint64_t* out) {
  return guarded([&] {
    const int L = plk_field_limbs(field);
    if (!L) fail(PLK_EINVAL, "unknown field id");
    if (n == 0) return;
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    // p - 1 in canonical form = to_canonical(NEG_ONE); the modulus limbs are host constants of the tables
    const uint64_t* mod = nullptr;
    PLK_FIELD_DISPATCH(field, field_modulus64, &mod);
    std::vector<uint64_t> tmp((size_t)n * L);
    memcpy(tmp.data(), in, (size_t)n * L * 8);
    for (size_t i = 0; i < n; ++i) {
      bool less = false;
      for (int j = 0; j < L; ++j) {                       // most significant limb decides last
        const uint64_t a = tmp[i * L + j], m = mod[j];
        less = a < m || (a == m && less);
      }
      if (!less) fail(PLK_EINVAL, "Out of range");
    }
    const int rc = plk_field_op(field, 7, tmp.data(), nullptr, out, n);
    if (rc != PLK_OK) fail(rc, plk_last_error_message());
  });
}

int plk_permutation_polynomial(int field, size_t degree, unsigned num_routed, const uint64_t* subgroup, const uint64_t* wires,
                               size_t wire_stride, const uint64_t* sigma, size_t sigma_row_len, size_t sigma_stride, const uint64_t* k_is,
                               const uint64_t* beta, const uint64_t* gamma, uint64_t* out) {
  return guarded([&] {
    const int L = plk_field_limbs(field);
    if (!L) fail(PLK_EINVAL, "unknown field id");
    if (degree == 0) return;
    if (!subgroup || !wires || !sigma || !k_is || !beta || !gamma || !out) fail(PLK_EINVAL, "NULL buffer");
    if (num_routed == 0 || num_routed > wire_stride) fail(PLK_EINVAL, "num_routed must be in 1..wire_stride");
    if (sigma_stride == 0 || (degree - 1) * sigma_stride >= sigma_row_len + (degree == 1 ? 1 : 0)) fail(PLK_EINVAL, "sigma rows too short for the stride");
    if (degree > ((size_t)1 << 22)) fail(PLK_EINVAL, "degree above 2^22 is not supported");
    cudaStream_t st = thread_stream();
    const size_t eb = (size_t)L * 8, n = degree;
    DevBuf d_sub(n * eb, st), d_w(n * wire_stride * eb, st), d_sig((size_t)num_routed * sigma_row_len * eb, st), d_k(num_routed * eb, st), d_bg(2 * eb, st);
    DevBuf d_num(n * eb, st), d_den(n * eb, st), d_inv(n * eb, st), d_tot((n / kScanTile + 2) * eb, st), d_out(n * eb, st), d_flag(16, st);
    PLK_CUDA(cudaMemcpyAsync(d_sub.p, subgroup, n * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_w.p, wires, n * wire_stride * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_sig.p, sigma, (size_t)num_routed * sigma_row_len * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_k.p, k_is, num_routed * eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_bg.p, beta, eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync((char*)d_bg.p + eb, gamma, eb, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
    PermArgs a;
    a.subgroup = d_sub.p; a.wires = d_w.p; a.sigma = d_sig.p; a.kis = d_k.p; a.beta_gamma = d_bg.p;
    a.n = n; a.wire_stride = wire_stride; a.sigma_row = sigma_row_len; a.sigma_stride = sigma_stride; a.routed = num_routed;
    PLK_FIELD_DISPATCH(field, launch_permutation, a, d_num.p, d_den.p, d_inv.p, d_tot.p, d_flag.as<int>(), d_out.p, st);
    int flag = 0;
    PLK_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out, d_out.p, n * eb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    if (flag) fail(PLK_EZERO, "No inverse");          // a zero denominator: the reference's `/` panics (field.rs Div)
  });
}

int plk_measure_mul_throughput(int field, double* products_per_s) {
  return guarded([&] {
    if (!products_per_s) fail(PLK_EINVAL, "NULL out");
    cudaStream_t st = thread_stream();
    PLK_FIELD_DISPATCH(field, measure_mul_throughput, products_per_s, st);
  });
}

int plk_batch_inverse(int field, const uint64_t* in, uint64_t* out, size_t n) {
  return guarded([&] {
    const int L = plk_field_limbs(field);
    if (!L) fail(PLK_EINVAL, "unknown field id");
    if (n == 0) return;
    if (!in || !out) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t bytes = n * L * 8;
    void* din = thread_scratch(0, bytes);
    void* dout = thread_scratch(1, bytes);
    int* dflag = reinterpret_cast<int*>(thread_scratch(2, 16));
    PLK_CUDA(cudaMemcpyAsync(din, in, bytes, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemsetAsync(dflag, 0, sizeof(int), st));
    PLK_FIELD_DISPATCH(field, launch_batch_inverse, din, dout, n, dflag, st);
    int flag = 0;
    PLK_CUDA(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    if (flag) fail(PLK_EZERO, "No inverse");      // field.rs:267
  });
}

}  // extern "C"
