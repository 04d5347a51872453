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

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  explicit DevBuf(size_t b) { alloc(b); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t b) {
    release();
    if (b == 0) b = 16;
    PLK_CUDA(cudaMalloc(&p, b));
    bytes = b;
  }
  void ensure(size_t b) { if (b > bytes) alloc(b); }
  void release() { if (p) { cudaFree(p); p = nullptr; bytes = 0; } }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline bool is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }
inline int log2_floor(size_t n) { int k = 0; while ((n >> k) > 1) ++k; return k; }
inline int log2_ceil(size_t n) { int k = 0; while (((size_t)1 << k) < n) ++k; return k; }   // util.rs:11-13

}  // namespace plk
