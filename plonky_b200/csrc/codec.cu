// C ABI of the generator derivation and the point wire format (include/plonky_b200.h).
#include "codec_kernels.cuh"

using namespace plk;

namespace plk {
const CodecOps* codec_ops_tweedledee();
const CodecOps* codec_ops_tweedledum();
const CodecOps* codec_ops_bls12_377();
}
namespace {
const CodecOps* codec_ops_for(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return codec_ops_tweedledee();
    case PLK_CURVE_TWEEDLEDUM: return codec_ops_tweedledum();
    case PLK_CURVE_BLS12_377: return codec_ops_bls12_377();
  }
  fail(PLK_EINVAL, "unknown curve id");
}
size_t base_limbs64(int curve) {
  const int bf = plk_curve_base_field(curve);
  if (bf < 0) fail(PLK_EINVAL, "unknown curve id");
  return (size_t)plk_field_limbs(bf);
}
}  // namespace

extern "C" {

int plk_blake_hash_usize_to_curve_dev(int curve, uint64_t seed_start, size_t n, void* d_points_xy, void* stream) {
  return guarded([&] {
    if (n && !d_points_xy) fail(PLK_EINVAL, "NULL buffer");
    codec_ops_for(curve)->hash_to_curve(nullptr, seed_start, n, d_points_xy, reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_blake_hash_usize_to_curve(int curve, uint64_t seed_start, size_t n, uint64_t* points_xy) {
  return guarded([&] {
    const size_t L = base_limbs64(curve);
    if (n == 0) return;
    if (!points_xy) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    DevBuf d(n * 2 * L * 8, st);
    codec_ops_for(curve)->hash_to_curve(nullptr, seed_start, n, d.p, st);
    PLK_CUDA(cudaMemcpyAsync(points_xy, d.p, n * 2 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_blake_hash_base_field_to_curve(int curve, const uint64_t* seeds, size_t n, uint64_t* points_xy) {
  return guarded([&] {
    const size_t L = base_limbs64(curve);
    if (n == 0) return;
    if (!seeds || !points_xy) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    DevBuf ds(n * L * 8, st), d(n * 2 * L * 8, st);
    PLK_CUDA(cudaMemcpyAsync(ds.p, seeds, n * L * 8, cudaMemcpyHostToDevice, st));
    codec_ops_for(curve)->hash_to_curve(ds.p, 0, n, d.p, st);
    PLK_CUDA(cudaMemcpyAsync(points_xy, d.p, n * 2 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
size_t plk_point_compressed_bytes(int curve) {
  const int bf = plk_curve_base_field(curve);
  return bf < 0 ? 0 : 1 + 8 * (size_t)plk_field_limbs(bf);
}
int plk_points_compress(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n, uint8_t* out) {
  return guarded([&] {
    const size_t L = base_limbs64(curve), cb = 1 + 8 * L;
    if (n == 0) return;
    if (!points_xy || !out) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    DevBuf d_xy(n * 2 * L * 8, st), d_z(n, st), d_out(n * cb, st);
    PLK_CUDA(cudaMemcpyAsync(d_xy.p, points_xy, n * 2 * L * 8, cudaMemcpyHostToDevice, st));
    if (zero) PLK_CUDA(cudaMemcpyAsync(d_z.p, zero, n, cudaMemcpyHostToDevice, st));
    codec_ops_for(curve)->compress(d_xy.p, zero ? d_z.as<unsigned char>() : nullptr, n, d_out.as<unsigned char>(), st);
    PLK_CUDA(cudaMemcpyAsync(out, d_out.p, n * cb, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_points_decompress(int curve, const uint8_t* in, size_t n, uint64_t* out_xy, uint8_t* out_zero, uint8_t* out_status) {
  return guarded([&] {
    const size_t L = base_limbs64(curve), cb = 1 + 8 * L;
    if (n == 0) return;
    if (!in || !out_xy || !out_zero) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    DevBuf d_in(n * cb, st), d_xy(n * 2 * L * 8, st), d_z(n, st), d_st(n, st);
    PLK_CUDA(cudaMemcpyAsync(d_in.p, in, n * cb, cudaMemcpyHostToDevice, st));
    codec_ops_for(curve)->decompress(d_in.as<unsigned char>(), n, d_xy.p, d_z.as<unsigned char>(), d_st.as<unsigned char>(), st);
    std::vector<uint8_t> status(n);
    PLK_CUDA(cudaMemcpyAsync(out_xy, d_xy.p, n * 2 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out_zero, d_z.p, n, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(status.data(), d_st.p, n, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
    bool bad = false;
    for (size_t i = 0; i < n; ++i) { if (out_status) out_status[i] = status[i]; bad = bad || status[i]; }
    // the reference returns io::Error for these ("Out of range", "Invalid x coordinate", serialization.rs:58-64)
    if (bad) fail(PLK_EINVAL, "invalid point encoding (out of range or x not on the curve)");
  });
}

}  // extern "C"
