// TMA-staged Stockham NTT pass for sm_100a (32-byte field elements).
//
// Same pass decomposition and the same fused factors as ntt_pass_kernel (ntt_kernels.cuh); what changes is how a tile
// moves and how the sub-transform is indexed:
//   * the tile -- 2^r rows x 4 columns, 128 contiguous bytes per row at a constant row stride -- is one box of a 3-D
//     tensor map and is brought in by ONE cp.async.bulk.tensor (TMA, SASS: UTMALDG) issued by one thread and signalled on
//     an mbarrier; the in-place passes write their tile back with ONE cp.async.bulk.tensor store (UTMASTG): whole 128-byte
//     rows, no per-thread address arithmetic, no half-written sectors.  Rows of a zero-padded input (LDE) lie outside the
//     tensor's bounds and arrive as zeros without touching memory;
//   * shared memory holds the tile in the TMA engine's own SWIZZLE_128B layout (16-byte chunk c of row q sits at chunk
//     c ^ (q & 7)), which makes the 16-byte accesses of a quarter warp (4 columns x 2 rows of different parity) hit eight
//     different bank groups;
//   * TMA cannot permute rows, so the decimation-in-time butterflies with their bit-reversed load become a Stockham
//     autosort transform: natural order in, natural order out, radix-4 rounds (two layers per round trip) that read one
//     buffer and write the other -- one barrier per round instead of a bit reversal.
//     Round with sub-transform size Ns -> 4 Ns, thread (j, col), k = j mod Ns, W = w_(4 Ns):
//        a_t = in[j + t R/4],  b0 = a0 + W^2k a2, b1 = a0 - W^2k a2, c0 = a1 + W^2k a3, c1 = a1 - W^2k a3
//        out[4 (j - k) + k + {0, 2, 1, 3} Ns] = b0 + W^k c0, b0 - W^k c0, b1 + W^(k + Ns) c1, b1 - W^(k + Ns) c1
//     (four products per four elements like two radix-2 layers; all twiddles come from the 128-entry table w_256^e).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "ntt_kernels.cuh"

namespace plk {

constexpr int kTmaColsLog = 2;                 // 4 columns x 32 B = 128 B per row = the SWIZZLE_128B span
constexpr int kTmaThreads = 256;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
// 1-D bulk copy global -> shared, signalled on the same mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// element (row, col) of a SWIZZLE_128B tile: two 16-byte chunks
template <class F>
__device__ __forceinline__ F tile_ld(const uint4* buf, int row, int col) {
  const uint4* r = buf + row * 8;
  const int s = row & 7;
  const uint4 a = r[(2 * col) ^ s], b = r[(2 * col + 1) ^ s];
  F x;
  x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
  x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
  return x;
}
template <class F>
__device__ __forceinline__ void tile_st(uint4* buf, int row, int col, const F& x) {
  uint4* r = buf + row * 8;
  const int s = row & 7;
  r[(2 * col) ^ s] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
  r[(2 * col + 1) ^ s] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
}

struct TmaTile {
  unsigned long long k0;       // in-place: k'' of column 0; first pass: jlow of column 0
  unsigned long long gbase;    // global element index of (row 0, col 0) inside its batch row
  int rowshift;
};

// Input-side factors applied in place on the freshly loaded tile (rows in natural order): coset shift s^j (first pass),
// two-level inter-pass twiddle, constant scale -- the cases that are not the fused direct-table load of the first round.
template <class F>
__device__ __noinline__ void tma_pre_pass(const NttPassParams& p, const TmaTile& g, uint4* buf) {
  const int elems = 1 << (p.r + kTmaColsLog);
#pragma unroll 1
  for (int idx = threadIdx.x; idx < elems; idx += kTmaThreads) {
    const int col = idx & 3, row = idx >> kTmaColsLog;
    F fac;
    bool have = false;
    if (p.first) {
      const unsigned long long j = g.gbase + col + ((unsigned long long)row << g.rowshift);
      if (p.pre_lo && j < p.n_in && j != 0) { fac = two_level<F>(p.pre_lo, p.pre_hi, p.lo_bits, j); have = true; }
    } else if (!p.tw_none) {
      const unsigned long long kk = g.k0 + col;
      const unsigned long long ex = ((unsigned long long)row * kk) << (p.log_n - p.r - p.log_m);
      if (p.tw_direct) {
        if (p.tw_all || ex != 0) { fac = load_fp<F>(p.tw_direct, ((size_t)row << p.log_m) + (size_t)kk); have = true; }
      } else if (p.tw_all) { fac = two_level_always<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex); have = true; }
      else if (ex != 0) { fac = two_level<F>(p.tw_lo, p.tw_hi, p.lo_bits, ex); have = true; }
    }
    if (p.scale) {
      const F s = load_fp<F>(p.scale, 0);
      fac = have ? F::mul(fac, s) : s;
      have = true;
    }
    if (have) tile_st<F>(buf, row, col, F::mul(tile_ld<F>(buf, row, col), fac));
  }
  __syncthreads();
}
// Output-side factors on the final natural-order values (see ntt_post_factors)
template <class F>
__device__ __noinline__ void tma_post_pass(const NttPassParams& p, const TmaTile& g, uint4* buf) {
  const int elems = 1 << (p.r + kTmaColsLog);
#pragma unroll 1
  for (int idx = threadIdx.x; idx < elems; idx += kTmaThreads) {
    const int col = idx & 3, row = idx >> kTmaColsLog;
    const unsigned long long k = p.first ? (unsigned long long)row : (g.k0 + col + ((unsigned long long)row << p.log_m));
    const unsigned long long pe = p.post_rowmul ? (p.post_row_base + blockIdx.y) * k : k;
    F x = tile_ld<F>(buf, row, col);
    if (p.post_lo && pe != 0) x = F::mul(x, two_level<F>(p.post_lo, p.post_hi, p.post_lo_bits, pe));
    if (p.post_periodic) x = F::mul(x, load_fp<F>(p.post_periodic, (size_t)(k & p.post_mask)));
    tile_st<F>(buf, row, col, x);
  }
}

// One Stockham round of radix 2^Q from `src` to `dst`: sub-transforms of size 2^ls grow to 2^(ls + Q).
// FUSE_TW: the inter-pass twiddle of the first round comes from the full table (one load, one product per element).
template <class F, int Q, bool FUSE_TW>
__device__ __forceinline__ void stockham_round(const NttPassParams& p, const TmaTile& g, const uint4* src, uint4* dst, int ls, const void* wsub) {
  const int R = 1 << p.r;
  const int groups = (R >> Q) << kTmaColsLog;
  const int Ns = 1 << ls;
  for (int gi = threadIdx.x; gi < groups; gi += kTmaThreads) {
    const int col = gi & 3, j = gi >> kTmaColsLog;
    const int k = j & (Ns - 1);
    const int in_stride = R >> Q;
    const int o = ((j - k) << Q) + k;
    F a[1 << Q];
#pragma unroll
    for (int t = 0; t < (1 << Q); ++t) {
      const int row = j + t * in_stride;
      a[t] = tile_ld<F>(src, row, col);
      if (FUSE_TW) {
        const size_t kk = (size_t)(g.k0 + col);
        if (p.tw_all || (row != 0 && kk != 0)) a[t] = F::mul(a[t], load_fp<F>(p.tw_direct, ((size_t)row << p.log_m) + kk));
      }
    }
    if (Q == 2) {
      // W = w_(4 Ns) = w_256^(64 / Ns): W^2k at index k * 128 / Ns, W^k at k * 64 / Ns, W^(k + Ns) at k * 64 / Ns + 64
      const int e1 = k << (6 - ls);
      F b0, b1, c0, c1;
      if (k == 0) {                                // first layer trivial (every group of the first round; one in Ns of the others)
        b0 = F::add(a[0], a[2]); b1 = F::sub(a[0], a[2]);
        c0 = F::add(a[1], a[3]); c1 = F::sub(a[1], a[3]);
      } else {
        const F w2 = load_fp<F>(wsub, (size_t)(2 * e1));
        const F t2 = F::mul(a[2], w2), t3 = F::mul(a[3], w2);
        b0 = F::add(a[0], t2); b1 = F::sub(a[0], t2);
        c0 = F::add(a[1], t3); c1 = F::sub(a[1], t3);
      }
      const F u0 = (k == 0) ? c0 : F::mul(c0, load_fp<F>(wsub, (size_t)e1));
      const F u1 = F::mul(c1, load_fp<F>(wsub, (size_t)(e1 + 64)));
      a[0] = F::add(b0, u0); a[2] = F::sub(b0, u0);
      a[1] = F::add(b1, u1); a[3] = F::sub(b1, u1);
    } else {
      const F t1 = (k == 0) ? a[1] : F::mul(a[1], load_fp<F>(wsub, (size_t)(k << (7 - ls))));
      const F s = F::add(a[0], t1);
      a[1] = F::sub(a[0], t1);
      a[0] = s;
    }
#pragma unroll
    for (int u = 0; u < (1 << Q); ++u) tile_st<F>(dst, o + u * Ns, col, a[u]);
  }
}

template <class F>
__global__ void __launch_bounds__(kTmaThreads, 3) ntt_tma_pass_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                                                                     NttPassParams p, int tma_store) {
  static_assert(F::N == 8, "32-byte elements");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int R = 1 << p.r;
  const int buf_vec = (R < 8 ? 8 : R) * 8;          // each buffer starts on a 1024-byte boundary (the swizzle pattern's period)
  uint4* buf0 = reinterpret_cast<uint4*>(smem_raw);
  uint4* buf1 = buf0 + buf_vec;
  uint4* wsub_s = buf1 + buf_vec;                    // the 128 sub-transform twiddles w_256^e (4 KiB), brought in by a bulk copy
  uint64_t* bar = reinterpret_cast<uint64_t*>(wsub_s + (1 << (kSubLog - 1)) * 2);
  const unsigned long long tile = blockIdx.x;
  TmaTile g;
  int c0, c2;
  if (p.first) {
    g.k0 = tile << kTmaColsLog;
    g.gbase = g.k0;
    g.rowshift = p.log_n - p.r;
    c0 = (int)(g.k0 * 4);
    c2 = (int)blockIdx.y;
  } else {
    const int tph = p.log_m - kTmaColsLog;
    const unsigned long long hi = tile >> tph;
    g.k0 = (tile & ((1ull << tph) - 1)) << kTmaColsLog;
    g.gbase = g.k0 + (hi << (p.log_m + p.r));
    g.rowshift = p.log_m;
    c0 = (int)(g.k0 * 4);
    c2 = (int)(((unsigned long long)blockIdx.y << (p.log_n - p.log_m - p.r)) + hi);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (unsigned)(R * 128 + (128 << (kSubLog - 1 - 2))));
    tma_load_3d(buf0, &map_in, c0, 0, c2, bar);
    bulk_load_1d(wsub_s, p.wsub, 128u << (kSubLog - 1 - 2), bar);
  }
  mbar_wait(bar, 0);

  // rounds: radix-4 while two layers remain, one radix-2 round for an odd digit
  const bool need_pre = (p.first && p.pre_lo) || (!p.first && !p.tw_none) || p.scale;
  const bool fuse_tw = !p.first && !p.tw_none && p.tw_direct && !p.scale;
  if (need_pre && !fuse_tw) tma_pre_pass<F>(p, g, buf0);
  const uint4* src = buf0;
  uint4* dst = buf1;
  int ls = 0;
  const int nrounds = (p.r + 1) >> 1;
  for (int rd = 0; rd < nrounds; ++rd) {
    const int q = (p.r - ls >= 2) ? 2 : 1;
    if (rd == 0 && fuse_tw) {
      if (q == 2) stockham_round<F, 2, true>(p, g, src, dst, ls, wsub_s);
      else stockham_round<F, 1, true>(p, g, src, dst, ls, wsub_s);
    } else {
      if (q == 2) stockham_round<F, 2, false>(p, g, src, dst, ls, wsub_s);
      else stockham_round<F, 1, false>(p, g, src, dst, ls, wsub_s);
    }
    ls += q;
    const uint4* t = src;
    src = dst;
    dst = const_cast<uint4*>(t);
    if (rd + 1 < nrounds) __syncthreads();
  }
  if (p.last && (p.post_lo || p.post_periodic)) {
    __syncthreads();
    tma_post_pass<F>(p, g, const_cast<uint4*>(src));
  }
  // `src` now holds the natural-order result
  if (tma_store) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA engine
    __syncthreads();
    if (threadIdx.x == 0) tma_store_3d(&map_out, src, c0, 0, c2);
    return;
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(p.out);
  const int elems = R << kTmaColsLog;
  if (!p.first) {
    for (int idx = threadIdx.x; idx < elems * 2; idx += kTmaThreads) {
      const int piece = idx & 1, e = idx >> 1;
      const int col = e & 3, row = e >> kTmaColsLog;
      const unsigned long long gidx = g.gbase + col + ((unsigned long long)row << g.rowshift);
      uint4* dstp = p.last ? ntt_out_ptr(p, out, gidx, blockIdx.y, 2) : out + ((unsigned long long)blockIdx.y * p.out_stride + gidx) * 2;
      dstp[piece] = src[row * 8 + ((2 * col + piece) ^ (row & 7))];
    }
  } else {
    // column c of the tile becomes a run of R contiguous outputs at rev_digits(jlow) * R
    for (int idx = threadIdx.x; idx < elems * 2; idx += kTmaThreads) {
      const int piece = idx & 1, e = idx >> 1;
      const int row = e & (R - 1), col = e >> p.r;
      unsigned long long x = g.k0 + col, pos = 0;
      for (int i = 0; i < p.ndig; ++i) {
        pos = (pos << p.digs[i]) | (x & ((1ull << p.digs[i]) - 1));
        x >>= p.digs[i];
      }
      const unsigned long long gidx = (pos << p.r) + row;
      uint4* dstp = p.last ? ntt_out_ptr(p, out, gidx, blockIdx.y, 2) : out + ((unsigned long long)blockIdx.y * p.out_stride + gidx) * 2;
      dstp[piece] = src[row * 8 + ((2 * col + piece) ^ (row & 7))];
    }
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tensorMapEncodeTiled tensor_map_encoder() {
  static PFN_tensorMapEncodeTiled fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<PFN_tensorMapEncodeTiled>(f);
  }();
  return fn;
}
// 3-D view (u64 words): dim0 = inner words (contiguous), dim1 = rows at stride1 bytes, dim2 = slabs at stride2 bytes; box = 16 words x box_rows x 1
inline bool make_tile_map(CUtensorMap* map, const void* base, unsigned long long inner_words, unsigned long long rows, unsigned long long stride1,
                          unsigned long long slabs, unsigned long long stride2, unsigned box_rows) {
  PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
  if (!enc) return false;
  if (inner_words == 0 || rows == 0 || slabs == 0) return false;
  if (inner_words >= (1ull << 32) || rows >= (1ull << 32) || slabs >= (1ull << 32)) return false;
  if (stride1 >= (1ull << 40) || stride2 >= (1ull << 40) || (stride1 & 15) || (stride2 & 15)) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  cuuint64_t dims[3] = {inner_words, rows, slabs};
  cuuint64_t strides[2] = {stride1, stride2};
  cuuint32_t box[3] = {16, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch one pass through the TMA kernel if its geometry qualifies; returns false to let the caller use ntt_pass_kernel.
template <class F>
bool launch_tma_pass(const NttPassParams& p, size_t n, size_t k, cudaStream_t st) {
  if constexpr (F::N != 8) {
    return false;
  } else {
  // Policy (measured on B200, 2^24 TweedledeeBase, profiles/r2_ntt_tma_ncu.txt): on full-length passes this kernel runs
  // 4 % behind the cp.async radix-4 kernel with 8-column tiles (3.71 vs 3.55 ms per transform: half-size tiles, the tile load
  // is not overlapped with the CTA's own butterflies); on the first pass of a zero-padded input (LDE, n_in <= n / 2) it is
  // ahead (coset LDE 2^21 -> 2^24: 3.88 vs 3.97 ms) because the padding rows are out-of-bounds box rows the TMA engine
  // zero-fills without memory traffic.  Default: that case only.  PLK_NTT_TMA=1 takes every qualifying pass, 0 none.
  static const int mode = getenv("PLK_NTT_TMA") ? atoi(getenv("PLK_NTT_TMA")) : -1;
  if (mode == 0) return false;
  if (mode < 0 && !(p.first && p.n_in * 2 <= n)) return false;
  const int cols_log = p.first ? (p.log_n - p.r) : p.log_m;
  if (cols_log < kTmaColsLog || p.r < 1) return false;
  const unsigned R = 1u << p.r;
  CUtensorMap map_in, map_out;
  int tma_store = 0;
  if (p.first) {
    const unsigned long long row_len = 1ull << (p.log_n - p.r);            // elements between consecutive rows j_m
    if (p.n_in == 0) return false;
    unsigned long long rows_present = p.n_in >= n ? R : p.n_in / row_len;
    if (p.n_in < n && (p.n_in % row_len) != 0) return false;               // a partial last row is not a rectangle
    if (rows_present == 0) return false;
    // rows beyond rows_present are out of bounds: the TMA engine fills them with zeros (the LDE's zero padding)
    if (!make_tile_map(&map_in, p.in, row_len * 4, rows_present, row_len * 32, k, p.in_stride * 32, R)) return false;
    map_out = map_in;
  } else {
    const unsigned long long M = 1ull << p.log_m, slabs = (unsigned long long)k << (p.log_n - p.log_m - p.r);
    if (p.in_stride != n || p.out_stride != n) return false;
    if (!make_tile_map(&map_in, p.in, M * 4, R, M * 32, slabs, (M * 32) << p.r, R)) return false;
    if (!(p.last && p.remap)) {
      if (!make_tile_map(&map_out, p.out, M * 4, R, M * 32, slabs, (M * 32) << p.r, R)) return false;
      tma_store = 1;
    } else {
      map_out = map_in;
    }
  }
  const size_t tiles = n >> (p.r + kTmaColsLog);
  if (tiles == 0 || tiles > 0x7fffffffull || k > 65535) return false;
  const size_t smem = (size_t)(R < 8 ? 8 : R) * 128 * 2 + (32 << (kSubLog - 1)) + 16;
  static std::mutex attr_mu;
  static bool attr_done[64] = {};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(attr_mu);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      PLK_CUDA(cudaFuncSetAttribute(ntt_tma_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 * 2 + (32 << (kSubLog - 1)) + 16));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  ntt_tma_pass_kernel<F><<<dim3((unsigned)tiles, (unsigned)k), kTmaThreads, smem, st>>>(map_in, map_out, p, tma_store);
  PLK_LAUNCHED();
  return true;
  }
}

}  // namespace plk
