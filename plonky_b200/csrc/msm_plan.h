// Host-side structures of the MSM (shared by the per-curve kernel translation units and the C ABI).
#pragma once
#include <map>
#include <mutex>
#include "common.cuh"

namespace plk {
constexpr int kTaskSizeMax = 32;     // S: additions per accumulate task (smaller for small MSMs: more threads)
constexpr int kBigBucket = 32;       // buckets with more task partials than this are summed by a whole CTA
constexpr int kRangeSize = 8;        // buckets per running-sum range
constexpr int kAccThreads = 128;
constexpr int kMaddCompactDefault = 3;   // accumulate kernel: 0 fully inlined madd, 1 six products as calls, 2 all ten, 3 four as calls (see ec.cuh)
constexpr int kQuadThreads = 64;     // CTA size of the quad-cooperative tail kernels (16 quads)
constexpr int kFinalQuadsMax = 64;   // quads of the single-CTA final reduction
constexpr int kAffineRoundsHostMax = 3;  // = kAffineRoundsMax of msm_affine.cuh
constexpr int kPartsMax = 8;         // bucket ranges of the overlapped pipeline (execute_one)
constexpr unsigned kSortMaxBins = 32768;   // shared-memory histogram sort: nb * 4 B <= 128 KiB

struct MsmGeom {
  unsigned long long n;     // terms
  int c;                    // window bits
  int nwin;                 // windows = ceil((BITS + 1) / c)
  unsigned nb;              // buckets in total: nbw (fixed-base table: all windows share them) or nwin * nbw (variable base)
  unsigned nbw;             // buckets per window = 2^(c-1)
  int variable;             // 1: no table of powers, buckets per window, windows combined with doublings (msm_parallel)
  unsigned task;            // S: entries per accumulate task (power of two, <= kTaskSizeMax)
  int affine_rounds;        // batched-affine tree rounds before the XYZZ task kernel (0 = none)
  // merged batch (run_batch): `batch` scalar vectors of n_pts terms each against the same fixed-base table run as ONE
  // pipeline -- n = batch * n_pts terms, vector v owns the bucket set [v * nbw, (v + 1) * nbw), nb = batch * nbw.
  // batch == 1: n_pts == n.
  unsigned batch;
  unsigned long long n_pts;
};

}  // namespace plk

// Scratch of ONE in-flight execute.  A table owns a small pool of these, one per CUDA stream that has
// executed against it, so that executes issued on different streams (the batch entry points fork onto
// internal streams; independent host threads use their own) never share buffers.
struct plk_msm_scratch {
  plk::DevBuf counts, offsets, task_off, cursors, sorted, partials, buckets, ranges, big_list;   // big_list[0] = count
  plk::DevBuf cta_hist;    // sort_rows x nb per-CTA histograms / column prefixes (shared-memory sort)
  // batched-affine rounds (msm_affine.cuh): ping-pong point lists, the parked running products, per-round bucket offsets
  plk::DevBuf aff[2], aff_prefix, aff_off[3];
  int affine_rounds = 0;   // 0 = XYZZ accumulation straight from the table
  unsigned sort_rows = 0;  // CTAs of the shared-memory sort; 0 = global-atomic counting sort
  plk::PhaseTimer timer;   // count | scan | scatter | accumulate | bucket_sum | range | final
  plk::MsmGeom g;          // geometry this scratch was sized for: the table's, or a merged batch of it
  size_t max_tasks = 0;    // upper bound of accumulate tasks under g
  // overlapped pipeline (execute_one): streams of the accumulate parts 1.., the high-priority stream of the reduction
  // tails, their events, and two small ping-pong buffers for the per-part chunk sums
  plk::DevBuf chunks[2];
  cudaStream_t part_st[plk::kPartsMax] = {}, tail_st = nullptr;
  cudaEvent_t fork_ev = nullptr, tail_ev = nullptr, acc_ev[plk::kPartsMax] = {};
  int streams_made = 0;

  // Number of bucket ranges this execute is cut into; 1 = the serial pipeline, which is the default: measured on B200
  // the overlapped pipeline LOSES (2^20 terms: 3.19 ms serial, 3.27 / 3.59 / 4.33 ms with 2 / 4 / 8 parts) --
  // a tail running under the next part's accumulation shares each scheduler with twelve accumulate warps and its
  // dependent chain stretches from 0.3 to 1.3 ms.  PLK_MSM_OVERLAP_PARTS = 2 | 4 | 8 switches it on (tests, profiles).
  // Always serial while phase timings are taken, and for variable-base and batched-affine geometries.
  int parts_for(const plk::MsmGeom& g, bool temporary) const {
    if (g.variable || g.affine_rounds > 0 || g.batch > 1 || temporary || plk::g_profiling.load(std::memory_order_relaxed)) return 1;
    static const int want = getenv("PLK_MSM_OVERLAP_PARTS") ? atoi(getenv("PLK_MSM_OVERLAP_PARTS")) : 1;
    int p = 1;
    while (2 * p <= want && 2 * p <= plk::kPartsMax && g.nb / (2 * p) >= 8u * plk::kRangeSize) p *= 2;
    return p;
  }
  void ensure_streams(int parts) {
    if (!fork_ev) {
      int least = 0, greatest = 0;
      PLK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      PLK_CUDA(cudaStreamCreateWithPriority(&tail_st, cudaStreamNonBlocking, greatest));
      PLK_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
      PLK_CUDA(cudaEventCreateWithFlags(&tail_ev, cudaEventDisableTiming));
      for (auto& e : acc_ev) PLK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // the parts must be accumulated IN ORDER for their tails to overlap anything: equal-priority streams are drained
    // round-robin (all parts finish together, measured), so part k gets the k-th priority below the tail stream
    for (; streams_made < parts; ++streams_made) {
      int least = 0, greatest = 0;
      PLK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      int prio = greatest + 1 + streams_made;
      if (prio > least) prio = least;
      PLK_CUDA(cudaStreamCreateWithPriority(&part_st[streams_made], cudaStreamNonBlocking, prio));
    }
  }
  ~plk_msm_scratch() {
    for (auto q : part_st) if (q) cudaStreamDestroy(q);
    if (tail_st) cudaStreamDestroy(tail_st);
    if (fork_ev) cudaEventDestroy(fork_ev);
    if (tail_ev) cudaEventDestroy(tail_ev);
    for (auto e : acc_ev) if (e) cudaEventDestroy(e);
  }
};

struct plk_msm_table {
  int curve = 0;
  size_t n = 0;
  unsigned w = 0;          // the caller's window (interface fidelity only)
  plk::MsmGeom g;
  size_t point_bytes = 64; // affine point
  plk::DevBuf table;            // nwin * n affine points, window-major
  size_t max_tasks = 0;
  bool temporary = false;       // created and destroyed inside one call (msm_parallel): stream-ordered allocations
  cudaStream_t temp_stream = nullptr;
  std::mutex mu;                // guards the pools below (not the execution)
  std::mutex batch_mu;          // serialises the fork/join enqueue of the batch entry points
  std::map<std::pair<cudaStream_t, unsigned>, plk_msm_scratch*> scratch;   // keyed by the executing stream and the merged-batch size
  plk_msm_scratch* last = nullptr;                     // scratch of the most recent execute (phase timings)
  static constexpr int kSideStreamsMax = 8;
  int side_streams = 8;                                       // PLK_MSM_SIDE_STREAMS (1..8), read when the pool is created; 8 vs 4: prover mix 7.6 -> 7.2 ms
  cudaStream_t side[kSideStreamsMax] = {};                    // fork/join streams of the batch entry points
  cudaEvent_t fork_ev = nullptr, join_ev[kSideStreamsMax] = {};
  ~plk_msm_table() {
    for (auto& kv : scratch) delete kv.second;
    for (auto s : side) if (s) cudaStreamDestroy(s);
    if (fork_ev) cudaEventDestroy(fork_ev);
    for (auto e : join_ev) if (e) cudaEventDestroy(e);
  }
};

namespace plk {
// per-curve entry points (one translation unit per curve keeps ptxas time parallel)
struct MsmOps {
  void (*table_build)(plk_msm_table* t, const void* d_points, cudaStream_t st);
  void (*import_points)(const void* d_raw, const unsigned char* d_zero, size_t n, int projective, void* d_out, cudaStream_t st);
  void (*execute_one)(plk_msm_table* t, plk_msm_scratch* s, const void* d_scalars, void* d_out_xyz, void* d_out_zero, void* d_partial,
                      cudaStream_t st);
  void (*combine_partials)(const void* d_partials, size_t count, void* d_out_xyz, void* d_out_zero, cudaStream_t st);
  void (*generate_points)(uint64_t seed, size_t n, void* d_out, cudaStream_t st);
  void (*to_affine_batch)(const void* d_in, const unsigned char* d_zero, size_t n, void* d_out, unsigned char* d_out_zero, cudaStream_t st);
  void (*multisum)(const void* d_points, const unsigned char* d_zero, const unsigned long long* d_offsets, size_t lists, void* d_out_xyz,
                   unsigned char* d_out_zero, cudaStream_t st);
  void (*curve_mul)(const void* d_points_xy, const void* d_scalars, size_t n, void* d_out_xyz, unsigned char* d_out_zero, cudaStream_t st);
  void (*commit_blind)(const void* d_msm_xyz, const unsigned char* d_msm_zero, const void* d_blinding, const uint64_t* h_xy, bool h_zero, size_t k,
                       void* d_out_xy, unsigned char* d_out_zero, cudaStream_t st);
};
// msm_parallel on device buffers (variable base, no table of powers; src/curve/curve_msm.rs:54-61): n affine points
// (identity = (0, 0)) and n Montgomery scalars -> one normalised point (3*L u64) + zero flag, asynchronous on `st`.
void msm_variable_dev(int curve, const void* d_points_xy, const void* d_scalars, size_t n, void* d_out_xyz, void* d_out_zero,
                      cudaStream_t st);
// k executes against one fixed-base table on device buffers (k*n scalars, row-major; d_out_xyz k*3*L u64, d_out_zero k
// bytes), forked over the table's side streams and joined on `st` -- the body of plk_msm_execute_batch_dev.
void msm_execute_batch_on(plk_msm_table* t, const void* d_scalars, size_t k, void* d_out_xyz, void* d_out_zero, cudaStream_t st);
}  // namespace plk
