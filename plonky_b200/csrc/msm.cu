// C ABI of the MSM (include/plonky_b200.h); kernels live in msm_kernels.cuh, one translation unit per curve.
#include <stdlib.h>
#include "msm_plan.h"
#include "field_constants.cuh"

using namespace plk;

namespace plk {
const MsmOps* msm_ops_tweedledee();
const MsmOps* msm_ops_tweedledum();
const MsmOps* msm_ops_bls12_377();
}

namespace {

const MsmOps* ops_for(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return msm_ops_tweedledee();
    case PLK_CURVE_TWEEDLEDUM: return msm_ops_tweedledum();
    case PLK_CURVE_BLS12_377: return msm_ops_bls12_377();
  }
  fail(PLK_EINVAL, "unknown curve id");
}

int curve_scalar_bits(int curve) {
  switch (curve) {
    case PLK_CURVE_TWEEDLEDEE: return TweedledeeParams::Scalar::BITS;
    case PLK_CURVE_TWEEDLEDUM: return TweedledumParams::Scalar::BITS;
    case PLK_CURVE_BLS12_377: return Bls12377Params::Scalar::BITS;
  }
  return -1;
}
int curve_base_limbs64(int curve) { return curve == PLK_CURVE_BLS12_377 ? 6 : 4; }

int pick_window(size_t n) {
  // all windows share one bucket set (fixed-base table), so the bucket reduction costs ~2^c additions
  // once and the accumulation n * ceil(256 / c): large windows win early.  PLK_MSM_WINDOW overrides (tuning).
  if (const char* e = getenv("PLK_MSM_WINDOW")) {
    int c = atoi(e);
    if (c >= 2 && c <= 24) return c;
  }
  // Measured on B200 (execute, ms; c = 13 / 14 / 15 / 16):  2^10: .34 .59 .36 .39   2^12: .41 .44 .39 .45
  // 2^14: .59 .53 .45 -   2^16: .80 .74 .61 .63   2^18: 4.3 1.43 1.16 1.19   2^20: 16 wins (fewer windows).
  // Small windows leave hundreds of entries per bucket, i.e. many task partials per bucket for the bucket-sum
  // kernel; the running-sum tails cost ~0.27 ms whatever the bucket count, so up to 2^18 terms 15 bits are best.
  int lg = log2_ceil(n ? n : 1);
  if (lg >= 19) return 16;
  if (lg >= 11) return 15;
  return 13;
}

// variable base (msm_parallel): every window has its own buckets, so the reduction work is nwin * 2^c
int pick_window_variable(size_t n) {
  if (const char* e = getenv("PLK_MSM_WINDOW_VAR")) {
    int c = atoi(e);
    if (c >= 4 && c <= 20) return c;
  }
  int lg = log2_ceil(n ? n : 1);
  int c = lg - 5;
  if (c < 4) c = 4;
  if (c > 14) c = 14;
  return c;
}

// accumulate task size for `entries` sorted entries spread over `nb` buckets: aim at >= ~75 K accumulate threads (4 CTAs of
// 128 on each of the 148 SMs) so that small MSMs still fill the machine; large ones use the full 32-entry tasks
unsigned pick_task(unsigned long long entries, unsigned nb) {
  unsigned s = kTaskSizeMax;
  while (s > 8 && entries / s < 75000) s >>= 1;
  // very long buckets (2^21+ terms on one GPU): keep the task partials per bucket near 16, otherwise every bucket
  // overflows into the one-CTA-per-bucket reduction (measured: BLS12-377 2^22 on one GPU, bucket_sum 50 ms)
  while ((entries / (nb ? nb : 1)) / s > 16 && s < 1024) s <<= 1;
  if (const char* e = getenv("PLK_MSM_TASK")) { int v = atoi(e); if (v >= 1 && v <= 1024) s = (unsigned)v; }
  return s;
}

// scratch of the stream `st` for executes of `batch` merged vectors (created on first use)
plk_msm_scratch* scratch_for(plk_msm_table* t, cudaStream_t st, unsigned batch = 1) {
  std::lock_guard<std::mutex> lk(t->mu);
  auto it = t->scratch.find(std::make_pair(st, batch));
  if (it == t->scratch.end()) {
    MsmGeom g = t->g;
    if (batch > 1) {
      g.batch = batch;
      g.n_pts = t->g.n;
      g.n = t->g.n * batch;
      g.nb = t->g.nbw * batch;
      g.task = pick_task(g.n * (unsigned long long)g.nwin, g.nb);
    }
    const size_t entries = (size_t)g.n * g.nwin;
    const size_t max_tasks = ((entries >> g.affine_rounds) + g.nb) / g.task + g.nb + 1;
    const size_t xyzz = 2 * t->point_bytes;
    auto* s = new plk_msm_scratch();
    s->affine_rounds = g.affine_rounds;
    s->g = g;
    s->max_tasks = max_tasks;
    if (t->temporary)
      for (plk::DevBuf* b : {&s->counts, &s->offsets, &s->task_off, &s->cursors, &s->sorted, &s->partials, &s->buckets, &s->ranges, &s->big_list,
                             &s->cta_hist, &s->chunks[0], &s->chunks[1]})
        b->set_async(st);
    try {
      s->counts.alloc((size_t)g.nb * 4);
      s->offsets.alloc(((size_t)g.nb + 1) * 4);
      s->task_off.alloc(((size_t)g.nb + 1) * 4);
      s->cursors.alloc((size_t)g.nb * 4);
      s->sorted.alloc((entries ? entries : 1) * 4);
      s->partials.alloc(max_tasks * xyzz);
      s->buckets.alloc((size_t)g.nb * xyzz);
      s->ranges.alloc(((size_t)g.nb / kRangeSize + 1) * xyzz);
      s->big_list.alloc(((size_t)g.nb + 1 + plk::kPartsMax) * 4);        // one list (count + entries) per part of the overlapped pipeline
      if (!g.variable && !t->temporary)
        for (auto& cb : s->chunks) cb.alloc(((size_t)g.nb / kRangeSize / 2 + plk::kPartsMax + 16) * xyzz);
      if (g.affine_rounds > 0) {
        // round r leaves at most entries / 2^r + nb points (every bucket rounds up)
        const size_t t1 = (entries >> 1) + g.nb + 1, t2 = (entries >> 2) + g.nb + 1;
        s->aff[0].alloc(t1 * t->point_bytes);
        if (g.affine_rounds > 1) s->aff[1].alloc(t2 * t->point_bytes);
        s->aff_prefix.alloc(t1 * (t->point_bytes / 2));
        for (int r = 0; r < g.affine_rounds; ++r) s->aff_off[r].alloc(((size_t)g.nb + 1) * 4);
      }
      // fixed-base geometry whose histogram fits in shared memory: per-CTA histograms instead of global atomics
      static const bool smem_sort = !(getenv("PLK_MSM_SORT") && atoi(getenv("PLK_MSM_SORT")) == 0);
      if (smem_sort && !g.variable && g.nb <= kSortMaxBins && g.n >= 4096) {
        int dev = 0, sms = 0;
        PLK_CUDA(cudaGetDevice(&dev));
        PLK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        size_t rows = g.n / 2048;
        if (rows > (size_t)sms) rows = sms;
        if (rows < 1) rows = 1;
        s->sort_rows = (unsigned)rows;
        s->cta_hist.alloc(rows * g.nb * 4);
      }
    } catch (...) {
      delete s;
      throw;
    }
    it = t->scratch.emplace(std::make_pair(st, batch), s).first;
  }
  t->last = it->second;
  return it->second;
}
void alloc_scratch(plk_msm_table* t) {
  const MsmGeom& g = t->g;
  t->max_tasks = ((((size_t)g.n * g.nwin) >> g.affine_rounds) + g.nb) / g.task + g.nb + 1;
}
// Merged batches (opt-in): up to kMergeMax vectors of a short MSM run as one pipeline with one bucket set per
// vector -- the sort, the accumulation and the reduction tails are each ONE launch over all vectors.  Measured on B200,
// 9 vectors of 2^16 terms: 2.74 ms merged against 2.61 ms for the default fork/join over 8 side streams (16 x 2^12:
// 1.34 vs 1.33 ms); the accumulation floor is 1.5 ms either way, and the merged bucket set (9 x 2^14 buckets) no longer
// fits the shared-memory sort.  PLK_MSM_BATCH_MERGE = m (2..16) enables it with at most m vectors per merged execute.
constexpr size_t kMergeMax = 16;
size_t merge_width(const plk_msm_table* t, size_t k) {
  static const int env = getenv("PLK_MSM_BATCH_MERGE") ? atoi(getenv("PLK_MSM_BATCH_MERGE")) : 0;
  const MsmGeom& g = t->g;
  if (env <= 1 || k < 2 || g.variable || g.affine_rounds > 0 || t->temporary || g.n == 0) return 1;
  if (g.n * (unsigned long long)g.nwin > (1ull << 22)) return 1;          // long MSMs already fill the machine on their own
  size_t m = k < (size_t)env ? k : (size_t)env;
  if (m > kMergeMax) m = kMergeMax;
  while (m > 1 && ((unsigned long long)g.nbw * m > (1u << 22) || g.n * (unsigned long long)g.nwin * m >= (1ull << 31))) --m;
  return m;
}
void run_one(plk_msm_table* t, const void* d_scalars, void* d_out_xyz, void* d_out_zero, void* d_partial, cudaStream_t st) {
  ops_for(t->curve)->execute_one(t, scratch_for(t, st), d_scalars, d_out_xyz, d_out_zero, d_partial, st);
}
// k executes on device buffers, forked round-robin onto the table's side streams and joined back on `st`: the
// low-occupancy reduction tails of one MSM overlap with the sort / accumulation of the next ones.
// h_scalars (optional): host copy of the k scalar vectors; vector j is then copied to d_scalars + j * sbytes on the side
// stream that executes it, so the copy of one vector overlaps the accumulation of the previous ones.
void run_batch(plk_msm_table* t, const char* d_scalars, size_t k, char* d_out_xyz, char* d_out_zero, cudaStream_t st,
               const char* h_scalars = nullptr) {
  const size_t L = curve_base_limbs64(t->curve), sbytes = t->n * 32;
  if (k == 1) {
    if (h_scalars && sbytes) PLK_CUDA(cudaMemcpyAsync(const_cast<char*>(d_scalars), h_scalars, sbytes, cudaMemcpyHostToDevice, st));
    run_one(t, d_scalars, d_out_xyz, d_out_zero, nullptr, st);
    return;
  }
  std::lock_guard<std::mutex> batch_lock(t->batch_mu);
  const size_t width = merge_width(t, k);
  if (width > 1) {
    // merged executes of `width` vectors each, back to back on the caller's stream (each one fills the machine)
    if (h_scalars && sbytes) PLK_CUDA(cudaMemcpyAsync(const_cast<char*>(d_scalars), h_scalars, k * sbytes, cudaMemcpyHostToDevice, st));
    for (size_t j = 0; j < k; j += width) {
      const size_t m = k - j < width ? k - j : width;
      if (m == 1) { run_one(t, d_scalars + j * sbytes, d_out_xyz + j * 3 * L * 8, d_out_zero + j, nullptr, st); continue; }
      ops_for(t->curve)->execute_one(t, scratch_for(t, st, (unsigned)m), d_scalars + j * sbytes, d_out_xyz + j * 3 * L * 8, d_out_zero + j,
                                     nullptr, st);
    }
    return;
  }
  {
    std::lock_guard<std::mutex> lk(t->mu);
    if (!t->fork_ev) {
      if (const char* e = getenv("PLK_MSM_SIDE_STREAMS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= plk_msm_table::kSideStreamsMax) t->side_streams = v;
      }
      PLK_CUDA(cudaEventCreateWithFlags(&t->fork_ev, cudaEventDisableTiming));
      for (int i = 0; i < t->side_streams; ++i) {
        PLK_CUDA(cudaStreamCreateWithFlags(&t->side[i], cudaStreamNonBlocking));
        PLK_CUDA(cudaEventCreateWithFlags(&t->join_ev[i], cudaEventDisableTiming));
      }
    }
  }
  const int S = t->side_streams;
  PLK_CUDA(cudaEventRecord(t->fork_ev, st));
  for (int i = 0; i < S && (size_t)i < k; ++i) PLK_CUDA(cudaStreamWaitEvent(t->side[i], t->fork_ev, 0));
  for (size_t j = 0; j < k; ++j) {
    if (h_scalars && sbytes)
      PLK_CUDA(cudaMemcpyAsync(const_cast<char*>(d_scalars) + j * sbytes, h_scalars + j * sbytes, sbytes, cudaMemcpyHostToDevice, t->side[j % S]));
    run_one(t, d_scalars + j * sbytes, d_out_xyz + j * 3 * L * 8, d_out_zero + j, nullptr, t->side[j % S]);
  }
  for (int i = 0; i < S && (size_t)i < k; ++i) {
    PLK_CUDA(cudaEventRecord(t->join_ev[i], t->side[i]));
    PLK_CUDA(cudaStreamWaitEvent(st, t->join_ev[i], 0));
  }
}

plk_msm_table* new_table(int curve, size_t n, unsigned w, bool variable = false) {
  const int bits = curve_scalar_bits(curve);
  if (bits < 0) fail(PLK_EINVAL, "unknown curve id");
  if (w < 1 || w > 32) fail(PLK_EINVAL, "window size out of range");
  auto* t = new plk_msm_table();
  t->curve = curve;
  t->n = n;
  t->w = w;
  t->g.n = n;
  t->g.variable = variable ? 1 : 0;
  t->g.c = variable ? pick_window_variable(n) : pick_window(n);
  t->g.nwin = (bits + 1 + t->g.c - 1) / t->g.c;
  t->g.nbw = 1u << (t->g.c - 1);
  t->g.nb = variable ? t->g.nbw * (unsigned)t->g.nwin : t->g.nbw;
  // Batched-affine tree rounds in front of the XYZZ task kernel (msm_affine.cuh).  OFF by default: measured on B200 at
  // 2^20 terms (profiles/r2_msm_affine_round_ncu.txt) the first round is bound by the random 64-byte gathers of the table
  // walk, which it has to do twice (denominators, then the additions): 4.7 GB of DRAM reads at the ~2.6 TB/s random-access
  // ceiling (tools/gather_probe.cu) = 1.54 ms for 8.4 M additions, against 1.26 ms for the same additions as mixed XYZZ
  // additions that read every point once.  PLK_MSM_AFFINE_ROUNDS=1..3 enables it (parity-tested for all three).
  // task size: aim at >= ~75 K accumulate threads (4 CTAs of 128 on each of the 148 SMs) so that small MSMs
  // still fill the machine; large ones use the full 32-entry tasks
  {
    const unsigned long long entries = (unsigned long long)n * t->g.nwin;
    int rounds = 0;
    if (const char* e = getenv("PLK_MSM_AFFINE_ROUNDS")) { int v = atoi(e); if (v >= 0 && v <= kAffineRoundsHostMax && !variable) rounds = v; }
    t->g.affine_rounds = rounds;
    t->g.task = pick_task(entries >> rounds, t->g.nb);           // entries the XYZZ stage still sees
    t->g.batch = 1;
    t->g.n_pts = n;
  }
  if ((unsigned long long)n * t->g.nwin >= (1ull << 31)) { delete t; fail(PLK_EINVAL, "too many terms for 31-bit table slots"); }
  return t;
}

// build from raw host points (projective or affine layout)
void precompute_host(int curve, const uint64_t* pts, const uint8_t* zero, size_t n, unsigned w, int projective, plk_msm_table** out,
                     bool variable = false) {
  if (!out) fail(PLK_EINVAL, "NULL out");
  *out = nullptr;
  if (n && !pts) fail(PLK_EINVAL, "NULL points");
  plk_msm_table* t = new_table(curve, n, w, variable);
  try {
    cudaStream_t st = thread_stream();
    if (variable) {               // lives for one msm_parallel call on this thread's stream
      t->temporary = true;
      t->temp_stream = st;
      t->table.set_async(st);
    }
    const size_t L = curve_base_limbs64(curve);
    t->point_bytes = 2 * L * 8;
    const size_t raw_bytes = n * (projective ? 3 : 2) * L * 8;
    DevBuf d_raw(raw_bytes, st), d_zero(n, st), d_aff(n * t->point_bytes, st);
    if (n) PLK_CUDA(cudaMemcpyAsync(d_raw.p, pts, raw_bytes, cudaMemcpyHostToDevice, st));
    if (n && zero) PLK_CUDA(cudaMemcpyAsync(d_zero.p, zero, n, cudaMemcpyHostToDevice, st));
    ops_for(curve)->import_points(d_raw.p, zero ? d_zero.as<unsigned char>() : nullptr, n, projective, d_aff.p, st);
    ops_for(curve)->table_build(t, d_aff.p, st);
    alloc_scratch(t);
    PLK_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    delete t;
    throw;
  }
  *out = t;
}

void execute_host(plk_msm_table* t, const uint64_t* scalars, size_t n, size_t k, uint64_t* out_xyz, uint8_t* out_zero) {
  if (!t) fail(PLK_EINVAL, "NULL table");
  if (n != t->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");   // curve_msm.rs:67,106
  if ((n && !scalars) || !out_xyz || !out_zero) fail(PLK_EINVAL, "NULL buffer");
  cudaStream_t st = thread_stream();
  const size_t L = curve_base_limbs64(t->curve);
  const size_t sbytes = n * 32;
  char* d_s = reinterpret_cast<char*>(thread_scratch(0, sbytes * k + 16));
  char* d_o = reinterpret_cast<char*>(thread_scratch(1, k * (3 * L * 8 + 8)));
  run_batch(t, d_s, k, d_o, d_o + k * 3 * L * 8, st, reinterpret_cast<const char*>(scalars));
  PLK_CUDA(cudaMemcpyAsync(out_xyz, d_o, k * 3 * L * 8, cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaMemcpyAsync(out_zero, d_o + k * 3 * L * 8, k, cudaMemcpyDeviceToHost, st));
  PLK_CUDA(cudaStreamSynchronize(st));
}

}  // namespace

namespace plk {
void msm_variable_dev(int curve, const void* d_points_xy, const void* d_scalars, size_t n, void* d_out_xyz, void* d_out_zero,
                      cudaStream_t st) {
  plk_msm_table* t = new_table(curve, n, 8, true);
  try {
    t->temporary = true;
    t->temp_stream = st;
    t->table.set_async(st);
    t->point_bytes = 2 * curve_base_limbs64(curve) * 8;
    ops_for(curve)->table_build(t, d_points_xy, st);
    alloc_scratch(t);
    run_one(t, d_scalars, d_out_xyz, d_out_zero, nullptr, st);
  } catch (...) {
    delete t;
    throw;
  }
  delete t;      // stream-ordered frees: the buffers outlive the kernels queued above
}
void msm_execute_batch_on(plk_msm_table* t, const void* d_scalars, size_t k, void* d_out_xyz, void* d_out_zero, cudaStream_t st) {
  run_batch(t, reinterpret_cast<const char*>(d_scalars), k, reinterpret_cast<char*>(d_out_xyz), reinterpret_cast<char*>(d_out_zero), st);
}
}  // namespace plk

extern "C" {

int plk_msm_precompute(int curve, const uint64_t* points_xyz, const uint8_t* zero, size_t n, unsigned w, plk_msm_table** out) {
  return guarded([&] { precompute_host(curve, points_xyz, zero, n, w, 1, out); });
}
int plk_msm_precompute_affine(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n, unsigned w, plk_msm_table** out) {
  return guarded([&] { precompute_host(curve, points_xy, zero, n, w, 0, out); });
}
int plk_msm_precompute_affine_dev(int curve, const void* d_points_xy, size_t n, unsigned w, plk_msm_table** out) {
  return guarded([&] {
    if (!out) fail(PLK_EINVAL, "NULL out");
    *out = nullptr;
    if (n && !d_points_xy) fail(PLK_EINVAL, "NULL points");
    plk_msm_table* t = new_table(curve, n, w);
    try {
      cudaStream_t st = thread_stream();
      t->point_bytes = 2 * curve_base_limbs64(curve) * 8;
      ops_for(curve)->table_build(t, d_points_xy, st);
      alloc_scratch(t);
      PLK_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
      delete t;
      throw;
    }
    *out = t;
  });
}
size_t plk_msm_table_len(const plk_msm_table* t) { return t ? t->n : 0; }
unsigned plk_msm_table_window(const plk_msm_table* t) { return t ? t->w : 0; }
void plk_msm_free(plk_msm_table* t) { delete t; }
int plk_msm_table_info(const plk_msm_table* t, unsigned* out, int cap) {
  if (!t || !out) return 0;
  const unsigned v[5] = {(unsigned)t->g.c, (unsigned)t->g.nwin, t->g.affine_rounds > 0 ? 1u : 0u, 10u, (unsigned)t->g.affine_rounds};
  int k = 0;
  for (; k < cap && k < 5; ++k) out[k] = v[k];
  return k;
}

int plk_msm_execute(const plk_msm_table* t, const uint64_t* scalars, size_t n, uint64_t* out_xyz, uint8_t* out_zero) {
  return guarded([&] { execute_host(const_cast<plk_msm_table*>(t), scalars, n, 1, out_xyz, out_zero); });
}
int plk_msm_execute_batch(const plk_msm_table* t, const uint64_t* scalars, size_t n, size_t k, uint64_t* out_xyz, uint8_t* out_zero) {
  return guarded([&] {
    if (k == 0) return;
    execute_host(const_cast<plk_msm_table*>(t), scalars, n, k, out_xyz, out_zero);
  });
}
int plk_msm_parallel(int curve, const uint64_t* scalars, const uint64_t* points_xyz, const uint8_t* zero, size_t n, unsigned w,
                     uint64_t* out_xyz, uint8_t* out_zero) {
  return guarded([&] {
    // variable base: no table of powers is built (the reference builds and discards one, curve_msm.rs:54-61)
    plk_msm_table* t = nullptr;
    precompute_host(curve, points_xyz, zero, n, w, 1, &t, true);
    try {
      execute_host(t, scalars, n, 1, out_xyz, out_zero);
    } catch (...) {
      delete t;
      throw;
    }
    delete t;
  });
}
int plk_msm_execute_dev(const plk_msm_table* tc, const void* d_scalars, size_t n, void* d_out_xyz, void* d_out_zero, void* stream) {
  return guarded([&] {
    auto* t = const_cast<plk_msm_table*>(tc);
    if (!t) fail(PLK_EINVAL, "NULL table");
    if (n != t->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");
    if ((n && !d_scalars) || !d_out_xyz || !d_out_zero) fail(PLK_EINVAL, "NULL buffer");
    run_one(t, d_scalars, d_out_xyz, d_out_zero, nullptr, reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_msm_execute_partial_dev(const plk_msm_table* tc, const void* d_scalars, size_t n, void* d_partial, void* stream) {
  return guarded([&] {
    auto* t = const_cast<plk_msm_table*>(tc);
    if (!t) fail(PLK_EINVAL, "NULL table");
    if (n != t->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");
    if ((n && !d_scalars) || !d_partial) fail(PLK_EINVAL, "NULL buffer");
    run_one(t, d_scalars, nullptr, nullptr, d_partial, reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_msm_combine_partials_dev(int curve, const void* d_partials, size_t count, void* d_out_xyz, void* d_out_zero, void* stream) {
  return guarded([&] {
    if (!d_partials || !d_out_xyz || !d_out_zero || count == 0) fail(PLK_EINVAL, "bad arguments");
    ops_for(curve)->combine_partials(d_partials, count, d_out_xyz, d_out_zero, reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_msm_last_phase_ms(const plk_msm_table* tc, float* out_ms, int cap) {
  int n = 0;
  int rc = guarded([&] {
    auto* t = const_cast<plk_msm_table*>(tc);
    if (!t || !out_ms) fail(PLK_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> lk(t->mu);
    if (!t->last) fail(PLK_EINVAL, "no execute has run against this table");
    n = t->last->timer.read(out_ms, cap);
  });
  return rc == PLK_OK ? n : -rc;
}
int plk_msm_execute_batch_dev(const plk_msm_table* tc, const void* d_scalars, size_t n, size_t k, void* d_out_xyz, void* d_out_zero,
                              void* stream) {
  return guarded([&] {
    auto* t = const_cast<plk_msm_table*>(tc);
    if (!t) fail(PLK_EINVAL, "NULL table");
    if (n != t->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");
    if (k == 0) return;
    if ((n && !d_scalars) || !d_out_xyz || !d_out_zero) fail(PLK_EINVAL, "NULL buffer");
    run_batch(t, reinterpret_cast<const char*>(d_scalars), k, reinterpret_cast<char*>(d_out_xyz), reinterpret_cast<char*>(d_out_zero),
              reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_commit_batch(const plk_msm_table* tc, const uint64_t* scalars, size_t n, size_t k, const uint64_t* blinding, const uint64_t* h_xy,
                     uint8_t h_zero, uint64_t* out_xy, uint8_t* out_zero) {
  return guarded([&] {
    auto* t = const_cast<plk_msm_table*>(tc);
    if (!t) fail(PLK_EINVAL, "NULL table");
    if (n != t->n) fail(PLK_ELENGTH, "precomputation / scalars length mismatch");   // curve_msm.rs:106
    if (k == 0) return;
    if ((n && !scalars) || !out_xy || !out_zero || (blinding && !h_xy && !h_zero)) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t L = curve_base_limbs64(t->curve), sbytes = n * 32;
    char* d_s = reinterpret_cast<char*>(thread_scratch(0, sbytes * k + 16));
    // msm results (k * 3L u64 + k flags), blinding factors, outputs (k * 2L u64 + k flags)
    const size_t o_msm = 0, o_mz = k * 3 * L * 8, o_bl = o_mz + ((k + 15) / 16) * 16, o_out = o_bl + k * 32, o_oz = o_out + k * 2 * L * 8;
    char* d_o = reinterpret_cast<char*>(thread_scratch(1, o_oz + k + 16));
    if (blinding) PLK_CUDA(cudaMemcpyAsync(d_o + o_bl, blinding, k * 32, cudaMemcpyHostToDevice, st));
    run_batch(t, d_s, k, d_o + o_msm, d_o + o_mz, st, reinterpret_cast<const char*>(scalars));
    ops_for(t->curve)->commit_blind(d_o + o_msm, reinterpret_cast<unsigned char*>(d_o + o_mz), blinding ? d_o + o_bl : nullptr, h_xy, h_zero != 0, k,
                                    d_o + o_out, reinterpret_cast<unsigned char*>(d_o + o_oz), st);
    PLK_CUDA(cudaMemcpyAsync(out_xy, d_o + o_out, k * 2 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out_zero, d_o + o_oz, k, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_msm_parallel_dev(int curve, const void* d_scalars, const void* d_points_xy, size_t n, void* d_out_xyz, void* d_out_zero,
                         void* stream) {
  return guarded([&] {
    if (curve_scalar_bits(curve) < 0) fail(PLK_EINVAL, "unknown curve id");
    if ((n && (!d_scalars || !d_points_xy)) || !d_out_xyz || !d_out_zero) fail(PLK_EINVAL, "NULL buffer");
    msm_variable_dev(curve, d_points_xy, d_scalars, n, d_out_xyz, d_out_zero, reinterpret_cast<cudaStream_t>(stream));
  });
}
size_t plk_msm_partial_limbs(int curve) { return 4 * (size_t)curve_base_limbs64(curve); }

int plk_points_generate_dev(int curve, uint64_t seed, size_t n, void* d_points_xy, void* stream) {
  return guarded([&] {
    if (n && !d_points_xy) fail(PLK_EINVAL, "NULL buffer");
    ops_for(curve)->generate_points(seed, n, d_points_xy, reinterpret_cast<cudaStream_t>(stream));
  });
}
int plk_points_generate(int curve, uint64_t seed, size_t n, uint64_t* points_xy) {
  return guarded([&] {
    if (n && !points_xy) fail(PLK_EINVAL, "NULL buffer");
    if (curve_scalar_bits(curve) < 0) fail(PLK_EINVAL, "unknown curve id");
    cudaStream_t st = thread_stream();
    const size_t bytes = n * 2 * curve_base_limbs64(curve) * 8;
    DevBuf d(bytes, st);
    ops_for(curve)->generate_points(seed, n, d.p, st);
    PLK_CUDA(cudaMemcpyAsync(points_xy, d.p, bytes, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_affine_multisummation(int curve, const uint64_t* points_xy, const uint8_t* zero, const uint64_t* offsets, size_t lists,
                               uint64_t* out_xyz, uint8_t* out_zero) {
  return guarded([&] {
    if (curve_scalar_bits(curve) < 0) fail(PLK_EINVAL, "unknown curve id");
    if (lists == 0) return;
    if (!offsets || !out_xyz || !out_zero) fail(PLK_EINVAL, "NULL buffer");
    for (size_t i = 0; i < lists; ++i) if (offsets[i] > offsets[i + 1]) fail(PLK_EINVAL, "offsets must be non-decreasing");
    const size_t n = offsets[lists];
    if (n && !points_xy) fail(PLK_EINVAL, "NULL points");
    cudaStream_t st = thread_stream();
    const size_t L = curve_base_limbs64(curve);
    DevBuf d_pts(n * 2 * L * 8, st), d_z(n, st), d_off((lists + 1) * 8, st), d_out(lists * 3 * L * 8, st), d_oz(lists, st);
    if (n) PLK_CUDA(cudaMemcpyAsync(d_pts.p, points_xy, n * 2 * L * 8, cudaMemcpyHostToDevice, st));
    if (n && zero) PLK_CUDA(cudaMemcpyAsync(d_z.p, zero, n, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_off.p, offsets, (lists + 1) * 8, cudaMemcpyHostToDevice, st));
    ops_for(curve)->multisum(d_pts.p, zero ? d_z.as<unsigned char>() : nullptr, d_off.as<unsigned long long>(), lists, d_out.p,
                             d_oz.as<unsigned char>(), st);
    PLK_CUDA(cudaMemcpyAsync(out_xyz, d_out.p, lists * 3 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out_zero, d_oz.p, lists, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_affine_summation(int curve, const uint64_t* points_xy, const uint8_t* zero, size_t n, uint64_t* out_xyz, uint8_t* out_zero) {
  const uint64_t offsets[2] = {0, n};
  return plk_affine_multisummation(curve, points_xy, zero, offsets, 1, out_xyz, out_zero);
}
int plk_curve_mul(int curve, const uint64_t* points_xyz, const uint8_t* zero, const uint64_t* scalars, size_t n, uint64_t* out_xyz,
                  uint8_t* out_zero) {
  return guarded([&] {
    if (curve_scalar_bits(curve) < 0) fail(PLK_EINVAL, "unknown curve id");
    if (n == 0) return;
    if (!points_xyz || !scalars || !out_xyz || !out_zero) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t L = curve_base_limbs64(curve);
    DevBuf d_raw(n * 3 * L * 8, st), d_z(n, st), d_aff(n * 2 * L * 8, st), d_s(n * 32, st), d_out(n * 3 * L * 8, st), d_oz(n, st);
    PLK_CUDA(cudaMemcpyAsync(d_raw.p, points_xyz, n * 3 * L * 8, cudaMemcpyHostToDevice, st));
    if (zero) PLK_CUDA(cudaMemcpyAsync(d_z.p, zero, n, cudaMemcpyHostToDevice, st));
    PLK_CUDA(cudaMemcpyAsync(d_s.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    ops_for(curve)->import_points(d_raw.p, zero ? d_z.as<unsigned char>() : nullptr, n, 1, d_aff.p, st);
    ops_for(curve)->curve_mul(d_aff.p, d_s.p, n, d_out.p, d_oz.as<unsigned char>(), st);
    PLK_CUDA(cudaMemcpyAsync(out_xyz, d_out.p, n * 3 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out_zero, d_oz.p, n, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}
int plk_batch_to_affine(int curve, const uint64_t* points_xyz, const uint8_t* zero, size_t n, uint64_t* out_xy, uint8_t* out_zero) {
  return guarded([&] {
    if (curve_scalar_bits(curve) < 0) fail(PLK_EINVAL, "unknown curve id");
    if (n == 0) return;
    if (!points_xyz || !out_xy || !out_zero) fail(PLK_EINVAL, "NULL buffer");
    cudaStream_t st = thread_stream();
    const size_t L = curve_base_limbs64(curve);
    DevBuf d_in(n * 3 * L * 8, st), d_z(n, st), d_out(n * 2 * L * 8, st), d_oz(n, st);
    PLK_CUDA(cudaMemcpyAsync(d_in.p, points_xyz, n * 3 * L * 8, cudaMemcpyHostToDevice, st));
    if (zero) PLK_CUDA(cudaMemcpyAsync(d_z.p, zero, n, cudaMemcpyHostToDevice, st));
    ops_for(curve)->to_affine_batch(d_in.p, zero ? d_z.as<unsigned char>() : nullptr, n, d_out.p, d_oz.as<unsigned char>(), st);
    PLK_CUDA(cudaMemcpyAsync(out_xy, d_out.p, n * 2 * L * 8, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaMemcpyAsync(out_zero, d_oz.p, n, cudaMemcpyDeviceToHost, st));
    PLK_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
