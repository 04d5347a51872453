// Host-side plan structures of the NTT (shared by the per-field kernel translation units and the C ABI).
#pragma once
#include <map>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace plk {
constexpr int kSubLog = 8;          // largest sub-transform: 2^8 rows
constexpr int kMaxPeers = 8;        // GPUs of one NVSwitch domain the domain-split transform stores into directly
constexpr int kMaxDigits = 6;       // 6 * 8 = 48 >= largest TWO_ADICITY (47)
constexpr int kTileColsLog = 3;     // 8 columns per tile
constexpr int kNttThreads = 256;
constexpr int kDirectLog = 24;      // passes with N_d <= 2^24 read their inter-pass twiddles from a full table (<= 512 MiB each;
                                    // PLK_NTT_DIRECT_LOG lowers it, larger passes fall back to two small tables + one product)
}  // namespace plk

struct CosetTables {
  plk::DevBuf fwd_lo, fwd_hi;    // s^j
  plk::DevBuf inv_lo, inv_hi;    // s^-j
};

struct plk_fft_plan {
  int field = 0;
  int log_n = 0;
  size_t n = 0;
  int device = 0;
  int m = 0;                 // passes
  int dig[plk::kMaxDigits];       // r_1 .. r_m
  int lo_bits = 0;
  int direct_log = plk::kDirectLog;   // passes with N_d <= 2^direct_log use a full twiddle table (0: always on the fly)
  size_t elem_bytes = 32;
  plk::DevBuf wsub[2];            // [0] forward, [1] inverse
  plk::DevBuf tw_lo[2], tw_hi[2];
  plk::DevBuf tw_hi_inv_scaled;   // inverse hi table with n^-1 folded in
  // full twiddle tables of the in-place passes whose N_d = R_d M_d is small (<= 2^kDirectLog entries):
  // [0] forward, [1] inverse, [2] inverse with n^-1 folded (last pass); indexed by digit
  plk::DevBuf direct[3][plk::kMaxDigits];
  plk::DevBuf n_inv;              // one element: n^-1
  plk::DevBuf pow2_inv;           // 2^-k, k <= TWO_ADICITY
  plk::DevBuf subgroup;           // w^k, k < n, natural order (built on first use: the transform of the unit vector e_1)
  std::mutex mu;
  std::map<std::vector<uint32_t>, CosetTables*> cosets;   // keyed by the shift's limbs
  std::map<size_t, plk::DevBuf*> zh_tables;                     // keyed by n_gates
  plk::PhaseTimer timer;          // one phase per pass of the last transform (written only while profiling, under timer_mu)
  std::mutex timer_mu;
  std::mutex sub_mu;              // guards the lazy construction of `subgroup`
  ~plk_fft_plan() {
    for (auto& kv : cosets) delete kv.second;
    for (auto& kv : zh_tables) delete kv.second;
  }
};


namespace plk {
struct FusedOps {
  const void* pre_lo = nullptr;
  const void* pre_hi = nullptr;
  const void* post_lo = nullptr;
  const void* post_hi = nullptr;
  const void* post_periodic = nullptr;
  unsigned long long post_mask = 0;
  // domain-split (multi-GPU) transform, phase A: twiddle w_N^(j_1 k') on the final values and packed store
  int post_rowmul = 0;
  unsigned long long post_row_base = 0;
  int post_lo_bits = -1;              // lo_bits of the post tables when they belong to another (larger) plan
  void* final_out = nullptr;          // last pass writes here (out of place) instead of d_out
  int remap = 0, remap_cl_log = 0;    // remap = 2: store into the peers' receive buffers (peer[], remap_row_base)
  unsigned long long remap_rows = 0;
  void* peer[kMaxPeers] = {};
  unsigned long long remap_row_base = 0;
};


// per-field entry points (one translation unit per field keeps ptxas time parallel)
struct NttOps {
  void (*plan_build)(plk_fft_plan*);
  void (*run)(const plk_fft_plan*, const void* d_in, size_t n_in, size_t in_stride, void* d_out, size_t k, bool inverse,
              const FusedOps* ops, cudaStream_t st);
  void (*coset)(plk_fft_plan*, const uint64_t* shift, bool inverse, FusedOps* ops, cudaStream_t st);
  void (*zh_table)(plk_fft_plan*, size_t n_gates, FusedOps* ops, cudaStream_t st);
  // one in-place pass of 2^r rows x 2^log_cols columns without twiddles (phase B of the domain-split transform)
  void (*final_pass)(const plk_fft_plan*, void* d_buf, int r, int log_cols, bool inverse, const void* d_scale, cudaStream_t st);
};
}  // namespace plk
