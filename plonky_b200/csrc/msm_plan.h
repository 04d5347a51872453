// Host-side structures of the MSM (shared by the per-curve kernel translation units and the C ABI).
#pragma once
#include <mutex>
#include "common.cuh"

namespace plk {
constexpr int kTaskSize = 64;        // S: additions per accumulate task
constexpr int kRangeSize = 8;        // buckets per running-sum range
constexpr int kAccThreads = 128;

struct MsmGeom {
  unsigned long long n;     // terms
  int c;                    // window bits
  int nwin;                 // windows = ceil((BITS + 1) / c)
  unsigned nb;              // buckets = 2^(c-1)
};

}  // namespace plk

struct plk_msm_table {
  int curve = 0;
  size_t n = 0;
  unsigned w = 0;          // the caller's window (interface fidelity only)
  plk::MsmGeom g;
  size_t point_bytes = 64; // affine point
  plk::DevBuf table;            // nwin * n affine points, window-major
  // scratch (one execute at a time per table)
  std::mutex mu;
  plk::DevBuf counts, offsets, task_off, cursors, sorted, partials, buckets, ranges, result;
  size_t max_tasks = 0;
  plk::PhaseTimer timer;   // count | scan | scatter | accumulate | bucket_sum | range | final
};


namespace plk {
// per-curve entry points (one translation unit per curve keeps ptxas time parallel)
struct MsmOps {
  void (*table_build)(plk_msm_table* t, const void* d_points, cudaStream_t st);
  void (*import_points)(const void* d_raw, const unsigned char* d_zero, size_t n, int projective, void* d_out, cudaStream_t st);
  void (*execute_one)(plk_msm_table* t, const void* d_scalars, void* d_out_xyz, void* d_out_zero, void* d_partial, cudaStream_t st);
  void (*combine_partials)(const void* d_partials, size_t count, void* d_out_xyz, void* d_out_zero, cudaStream_t st);
  void (*generate_points)(uint64_t seed, size_t n, void* d_out, cudaStream_t st);
  void (*to_affine_batch)(const void* d_in, const unsigned char* d_zero, size_t n, void* d_out, unsigned char* d_out_zero, cudaStream_t st);
  void (*multisum)(const void* d_points, const unsigned char* d_zero, const unsigned long long* d_offsets, size_t lists, void* d_out_xyz,
                   unsigned char* d_out_zero, cudaStream_t st);
  void (*curve_mul)(const void* d_points_xy, const void* d_scalars, size_t n, void* d_out_xyz, unsigned char* d_out_zero, cudaStream_t st);
};
}  // namespace plk
