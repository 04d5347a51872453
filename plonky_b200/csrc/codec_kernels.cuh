// Generator derivation and the point wire format on device.
//
//   blake_hash_base_field_to_curve / blake_hash_usize_to_curve   src/hash_to_curve.rs:13-76
//       (pedersen_g, pedersen_h, U: src/circuit_builder.rs:1127-1129, src/verifier.rs:151-174)
//   AffinePoint ToBytes / FromBytes                                src/serialization.rs:32-72
//   Field::square_root, is_quadratic_residue                      src/field/field.rs:377-391, 440-473
//
// The reference hashes with the blake3 crate (0.3.3, not vendored).  The inputs here are BYTES + 2 <= 64 bytes and
// the extended output BYTES + 1 <= 64 bytes, i.e. ONE compression of the published BLAKE3 function with the flags
// CHUNK_START | CHUNK_END | ROOT and output-block counter 0; that is what blake3_one_block implements.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace plk {

template <int N>
struct CodecConsts {
  uint32_t qr_exp[N];     // (p - 1) / 2        Euler's criterion, field.rs:382 (kept for reference; the kernels test a^T instead)
  uint32_t w_exp[N];      // (T - 1) / 2        field.rs:446
  uint32_t z[N];          // GENERATOR^T = primitive_root_of_unity(TWO_ADICITY), Montgomery form (field.rs:445)
};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
__device__ __forceinline__ void blake3_g(uint32_t (&v)[16], int a, int b, int c, int d, uint32_t mx, uint32_t my) {
  v[a] = v[a] + v[b] + mx;
  v[d] = rotr32(v[d] ^ v[a], 16);
  v[c] = v[c] + v[d];
  v[b] = rotr32(v[b] ^ v[c], 12);
  v[a] = v[a] + v[b] + my;
  v[d] = rotr32(v[d] ^ v[a], 8);
  v[c] = v[c] + v[d];
  v[b] = rotr32(v[b] ^ v[c], 7);
}
// hash of one block of `len` <= 64 bytes (zero padded in m), 64 bytes of output in `out`
__device__ __forceinline__ void blake3_one_block(const uint32_t (&msg)[16], uint32_t len, uint32_t (&out)[16]) {
  const uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  uint32_t v[16], m[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = iv[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[8 + i] = iv[i];
  v[12] = 0;            // output block counter, low / high
  v[13] = 0;
  v[14] = len;
  v[15] = 1u | 2u | 8u;   // CHUNK_START | CHUNK_END | ROOT
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = msg[i];
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    blake3_g(v, 0, 4, 8, 12, m[0], m[1]);
    blake3_g(v, 1, 5, 9, 13, m[2], m[3]);
    blake3_g(v, 2, 6, 10, 14, m[4], m[5]);
    blake3_g(v, 3, 7, 11, 15, m[6], m[7]);
    blake3_g(v, 0, 5, 10, 15, m[8], m[9]);
    blake3_g(v, 1, 6, 11, 12, m[10], m[11]);
    blake3_g(v, 2, 7, 8, 13, m[12], m[13]);
    blake3_g(v, 3, 4, 9, 14, m[14], m[15]);
    if (r < 6) {
      const uint32_t t[16] = {m[2], m[6], m[3], m[10], m[7], m[0], m[4], m[13], m[1], m[11], m[12], m[5], m[9], m[14], m[15], m[8]};
#pragma unroll
      for (int i = 0; i < 16; ++i) m[i] = t[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    out[i] = v[i] ^ v[i + 8];
    out[i + 8] = v[i + 8] ^ iv[i];
  }
}

template <class F>
__device__ __forceinline__ bool below_modulus(const F& c) {       // is_valid_canonical_u64 (cmp(v, ORDER) == Less)
  bool lt = false;
#pragma unroll
  for (int i = 0; i < F::N; ++i) lt = (c.l[i] < F::Params::mod(i)) || (c.l[i] == F::Params::mod(i) && lt);
  return lt;
}

// Field::square_root (field.rs:440-473), the reference's own Tonelli-Shanks loop so that the SAME root comes out.
// Returns false for a non-residue.
template <class F>
__device__ __noinline__ bool field_sqrt(const F& a, const CodecConsts<F::N>& k, F& root) {
  if (a.is_zero()) { root = a; return true; }
  const F one = F::one();
  // w, x, b exactly as field.rs:446-448.  Euler's criterion (is_quadratic_residue, field.rs:377-391) is evaluated on
  // b = a^T instead of with a second full-length exponentiation: a^((p-1)/2) = (a^T)^(2^(s-1)), s = TWO_ADICITY.
  F w = F::pow(a, k.w_exp, F::Params::BITS);
  F x = F::mul(w, a);
  F b = F::mul(x, w);
  F e = b;
  for (int i = 0; i < F::Params::TWO_ADICITY - 1; ++i) e = F::sqr(e);
  if (e != one) return false;
  F z;
#pragma unroll
  for (int i = 0; i < F::N; ++i) z.l[i] = k.z[i];
  int v = F::Params::TWO_ADICITY;
  while (b != one) {
    int kk = 0;
    F b2k = b;
    while (b2k != one) { b2k = F::sqr(b2k); ++kk; }
    const int j = v - kk - 1;
    w = z;
    for (int i = 0; i < j; ++i) w = F::sqr(w);
    z = F::sqr(w);
    b = F::mul(b, z);
    x = F::mul(x, w);
    v = kk;
  }
  root = x;
  return true;
}
template <class C>
__device__ __forceinline__ Fp<typename C::Base> curve_rhs(const Fp<typename C::Base>& x) {     // x^3 + a x + b with a = 0
  typedef Fp<typename C::Base> F;
  static_assert(C::A_SMALL == 0, "the supported curves have a = 0");
  F b = F::zero();
  b.l[0] = (uint32_t)C::B_SMALL;
  return F::add(F::mul(F::sqr(x), x), F::from_canonical(b));
}

// one thread per seed.  seeds == nullptr: seed_i = from_canonical_usize(seed_start + i) (blake_hash_usize_to_curve);
// else seeds[i] is a base-field element in Montgomery form (blake_hash_base_field_to_curve).
template <class C>
__global__ void __launch_bounds__(128) blake_hash_to_curve_kernel(const void* __restrict__ seeds, unsigned long long seed_start,
                                                                  unsigned long long n, CodecConsts<Fp<typename C::Base>::N> k,
                                                                  void* __restrict__ out_xy) {
  typedef Fp<typename C::Base> F;
  constexpr int N = F::N;                      // 32-bit words of a field element; BYTES = 4 N
  constexpr int TOP_SHIFT = 32 * N - C::Base::BITS;      // 8 * BYTES - BITS (hash_to_curve.rs:39)
  static_assert(N + 1 <= 16 && TOP_SHIFT < 8, "one BLAKE3 block in, one out");
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  F seed = F::zero();
  if (seeds) seed = F::to_canonical(load_fp<F>(seeds, i));
  else { seed.l[0] = (uint32_t)(seed_start + i); seed.l[1] = (uint32_t)((seed_start + i) >> 32); }
  uint32_t msg[16], h[16];
#pragma unroll
  for (int w = 0; w < 16; ++w) msg[w] = w < N ? seed.l[w] : 0u;
  for (unsigned it = 0;; ++it) {                       // blake_hash_base_field_to_curve's loop (MapToGroup)
    F x;
    bool y_neg;
    for (unsigned j = 0;; ++j) {                       // blake_field's retry loop
      msg[N] = (it & 0xffu) | ((j & 0xffu) << 8);      // bytes[BYTES] = iter, bytes[BYTES + 1] = j
      blake3_one_block(msg, 4 * N + 2, h);
#pragma unroll
      for (int w = 0; w < N; ++w) x.l[w] = h[w];
      x.l[N - 1] = (h[N - 1] & 0x00ffffffu) | (((h[N - 1] >> 24) >> TOP_SHIFT) << 24);     // hash_container[BYTES - 1] >>= 8 BYTES - BITS
      y_neg = (h[N] & 1u) != 0;                        // the extra byte
      if (below_modulus(x)) break;                     // from_canonical_u8_vec: Ok
    }
    const F xm = F::from_canonical(x);
    F y;
    if (field_sqrt<F>(curve_rhs<C>(xm), k, y)) {
      if (y_neg) y = F::neg(y);
      store_fp<F>(out_xy, 2 * i, xm);
      store_fp<F>(out_xy, 2 * i + 1, y);
      return;
    }
  }
}

// AffinePoint::write (serialization.rs:32-44): mask byte (bit 0 zero, bit 1 y odd) then x as BYTES canonical
// little-endian bytes; stride 1 + BYTES per point.
template <class C>
__global__ void points_compress_kernel(const void* __restrict__ xy, const unsigned char* __restrict__ zero, unsigned long long n,
                                       unsigned char* __restrict__ out) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F x = F::to_canonical(load_fp<F>(xy, 2 * i)), y = F::to_canonical(load_fp<F>(xy, 2 * i + 1));
  unsigned char* o = out + i * (1 + 4 * F::N);
  o[0] = (unsigned char)(((zero && zero[i]) ? 1u : 0u) | ((y.l[0] & 1u) << 1));
  for (int w = 0; w < F::N; ++w)
    for (int b = 0; b < 4; ++b) o[1 + 4 * w + b] = (unsigned char)(x.l[w] >> (8 * b));
}
// AffinePoint::read (serialization.rs:46-72).  status[i]: 0 ok, 1 "Out of range", 2 "Invalid x coordinate".
template <class C>
__global__ void points_decompress_kernel(const unsigned char* __restrict__ in, unsigned long long n, CodecConsts<Fp<typename C::Base>::N> k,
                                         void* __restrict__ out_xy, unsigned char* __restrict__ out_zero, unsigned char* __restrict__ status) {
  typedef Fp<typename C::Base> F;
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = in + i * (1 + 4 * F::N);
  const unsigned mask = p[0];
  F x = F::zero(), y = F::zero();
  unsigned char st = 0, z = 0;
  if (mask & 1u) {
    z = 1;                                             // AffinePoint { x: ZERO, y: ZERO, zero: true }
  } else {
    F c;
    for (int w = 0; w < F::N; ++w) c.l[w] = (uint32_t)p[1 + 4 * w] | ((uint32_t)p[2 + 4 * w] << 8) | ((uint32_t)p[3 + 4 * w] << 16) | ((uint32_t)p[4 + 4 * w] << 24);
    if (!below_modulus(c)) st = 1;
    else {
      x = F::from_canonical(c);
      if (!field_sqrt<F>(curve_rhs<C>(x), k, y)) { st = 2; x = F::zero(); y = F::zero(); }
      else if ((F::to_canonical(y).l[0] & 1u) != ((mask >> 1) & 1u)) y = F::neg(y);
    }
  }
  store_fp<F>(out_xy, 2 * i, x);
  store_fp<F>(out_xy, 2 * i + 1, y);
  out_zero[i] = z;
  status[i] = st;
}

struct CodecOps {
  void (*hash_to_curve)(const void* d_seeds, unsigned long long seed_start, size_t n, void* d_out_xy, cudaStream_t st);
  void (*compress)(const void* d_xy, const unsigned char* d_zero, size_t n, unsigned char* d_out, cudaStream_t st);
  void (*decompress)(const unsigned char* d_in, size_t n, void* d_out_xy, unsigned char* d_out_zero, unsigned char* d_status, cudaStream_t st);
};

template <class C>
CodecConsts<Fp<typename C::Base>::N> make_codec_consts() {
  typedef typename C::Base P;
  constexpr int N = P::LIMBS;
  CodecConsts<N> k;
  uint32_t pm1[N];
  for (int i = 0; i < N; ++i) pm1[i] = P::mod(i);
  pm1[0] -= 1;                                          // p is odd
  auto shr = [](const uint32_t (&in)[N], int bits, uint32_t (&out)[N]) {
    for (int i = 0; i < N; ++i) {
      const int src = i + bits / 32, off = bits % 32;
      uint64_t lo = src < N ? in[src] : 0, hi = src + 1 < N ? in[src + 1] : 0;
      out[i] = (uint32_t)(((hi << 32) | lo) >> off);
    }
  };
  shr(pm1, 1, k.qr_exp);
  uint32_t t[N];
  shr(pm1, P::TWO_ADICITY, t);                          // T, odd
  t[0] -= 1;
  shr(t, 1, k.w_exp);
  for (int i = 0; i < N; ++i) k.z[i] = FieldTables<P>::root(P::TWO_ADICITY)[i];
  return k;
}
template <class C>
void codec_hash_to_curve(const void* d_seeds, unsigned long long seed_start, size_t n, void* d_out_xy, cudaStream_t st) {
  if (n == 0) return;
  blake_hash_to_curve_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_seeds, seed_start, n, make_codec_consts<C>(), d_out_xy);
  PLK_LAUNCHED();
}
template <class C>
void codec_compress(const void* d_xy, const unsigned char* d_zero, size_t n, unsigned char* d_out, cudaStream_t st) {
  if (n == 0) return;
  points_compress_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_xy, d_zero, n, d_out);
  PLK_LAUNCHED();
}
template <class C>
void codec_decompress(const unsigned char* d_in, size_t n, void* d_out_xy, unsigned char* d_out_zero, unsigned char* d_status, cudaStream_t st) {
  if (n == 0) return;
  points_decompress_kernel<C><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_in, n, make_codec_consts<C>(), d_out_xy, d_out_zero, d_status);
  PLK_LAUNCHED();
}
template <class C>
const CodecOps* make_codec_ops() {
  static const CodecOps ops = {&codec_hash_to_curve<C>, &codec_compress<C>, &codec_decompress<C>};
  return &ops;
}
}  // namespace plk
