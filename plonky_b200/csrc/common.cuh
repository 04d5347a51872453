// Shared host-side plumbing for the C ABI: status codes, error capture, launch counting,
// per-thread streams, RAII device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>
#include "../../include/plonky_b200.h"

namespace plk {

extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_profiling;     // plk_set_profiling: record CUDA events between pipeline phases
void set_last_error(const std::string& s);

struct CudaError {
  cudaError_t e;
  const char* what;
  const char* file;
  int line;
};

#define PLK_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) throw ::plk::CudaError{e_, #call, __FILE__, __LINE__}; \
  } while (0)

// count + check a kernel launch
#define PLK_LAUNCHED()                                  \
  do {                                                  \
    ::plk::g_launches.fetch_add(1, std::memory_order_relaxed); \
    PLK_CUDA(cudaGetLastError());                       \
  } while (0)

struct StatusError {
  int status;
  std::string msg;
};
[[noreturn]] inline void fail(int status, const std::string& msg) { throw StatusError{status, msg}; }

// Run `body`, translating exceptions into plk_status (the ABI never aborts or throws).
template <class Fn>
int guarded(Fn&& body) {
  try {
    body();
    return PLK_OK;
  } catch (const StatusError& s) {
    set_last_error(s.msg);
    return s.status;
  } catch (const CudaError& c) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", c.what, cudaGetErrorString(c.e), c.file, c.line);
    set_last_error(buf);
    cudaGetLastError();   // clear sticky-less errors
    return c.e == cudaErrorMemoryAllocation ? PLK_ENOMEM : PLK_ECUDA;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return PLK_ENOMEM;
  } catch (...) {
    set_last_error("unknown exception");
    return PLK_ECUDA;
  }
}

// One non-blocking stream per calling host thread (the reference is called from rayon workers,
// src/plonk_util.rs:173-189, src/halo.rs:119-123).
cudaStream_t thread_stream();
// Grow-only device scratch owned by the calling host thread (slots 0..7), for the host-pointer entry
// points: avoids a cudaMalloc/cudaFree pair (and its implicit device synchronisation) per call.
void* thread_scratch(int slot, size_t bytes);

// Device buffer.  Long-lived buffers (tables, plans) use cudaMalloc; temporaries of one call use the
// stream-ordered allocator (cudaMallocFromPoolAsync / cudaFreeAsync on the call's stream, a private pool per device), which
// costs microseconds instead of the ~0.1-0.5 ms and implicit device synchronisation of cudaMalloc / cudaFree.
cudaMemPool_t async_pool();      // the library's private pool on the current device
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool async = false;
  cudaStream_t astream = nullptr;
  DevBuf() {}
  explicit DevBuf(size_t b) { alloc(b); }
  DevBuf(size_t b, cudaStream_t st) { set_async(st); alloc(b); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes), async(o.async), astream(o.astream) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; async = o.async; astream = o.astream; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void set_async(cudaStream_t st) { async = true; astream = st; }
  void alloc(size_t b) {
    release();
    if (b == 0) b = 16;
    if (async) PLK_CUDA(cudaMallocFromPoolAsync(&p, b, async_pool(), astream));
    else PLK_CUDA(cudaMalloc(&p, b));
    bytes = b;
  }
  void ensure(size_t b) { if (b > bytes) alloc(b); }
  void release() {
    if (p) {
      if (async) cudaFreeAsync(p, astream); else cudaFree(p);
      p = nullptr;
      bytes = 0;
    }
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Phase timer: CUDA events recorded on the launching stream between the kernels of one call (only when
// profiling is enabled; bench.py uses it for the live per-kernel roofline numbers).
struct PhaseTimer {
  static constexpr int kMax = 12;
  cudaEvent_t ev[kMax + 1];
  int n = 0;        // events recorded in the last call
  bool created = false;
  void begin(cudaStream_t st) {
    if (!g_profiling.load(std::memory_order_relaxed)) return;          // not profiling: the timer is never written (shared plans stay read-only)
    if (!created) { for (auto& e : ev) PLK_CUDA(cudaEventCreate(&e)); created = true; }
    n = 0;
    mark(st);
  }
  void mark(cudaStream_t st) {
    if (!created || !g_profiling.load(std::memory_order_relaxed) || n > kMax) return;
    PLK_CUDA(cudaEventRecord(ev[n++], st));
  }
  // elapsed ms of each phase of the last call; returns the number of phases
  int read(float* out, int cap) {
    if (n < 2) return 0;
    PLK_CUDA(cudaEventSynchronize(ev[n - 1]));
    int k = 0;
    for (int i = 0; i + 1 < n && k < cap; ++i, ++k) PLK_CUDA(cudaEventElapsedTime(&out[k], ev[i], ev[i + 1]));
    return k;
  }
  ~PhaseTimer() { if (created) for (auto& e : ev) cudaEventDestroy(e); }
};

inline bool is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }
inline int log2_floor(size_t n) { int k = 0; while ((n >> k) > 1) ++k; return k; }
inline int log2_ceil(size_t n) { int k = 0; while (((size_t)1 << k) < n) ++k; return k; }   // util.rs:11-13

}  // namespace plk
