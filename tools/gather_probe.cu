// Probe: DRAM bytes fetched per random 64-byte gather from a 1 GiB table (the MSM table walk), for the plain
// ld.global.v4, the .L2::64B / .L2::128B prefetch-size qualifiers and cudaLimitMaxL2FetchGranularity.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe tools/gather_probe.cu
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather_probe <granularity limit or 0>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 v;
  if (MODE == 0) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long mix(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// BYTES = 64: both halves of a 64-byte point; 32: only the first half (x)
template <int MODE, int BYTES>
__global__ void gather(const uint4* __restrict__ table, unsigned long long slots, unsigned per_thread, uint4* out) {
  const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (unsigned k = 0; k < per_thread; ++k) {
    const unsigned long long s = mix(t * per_thread + k) % slots;
    const uint4 a = ld<MODE>(table + 4 * s);
    acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
    const uint4 b = ld<MODE>(table + 4 * s + 1);
    acc.x ^= b.x; acc.y ^= b.y; acc.z ^= b.z; acc.w ^= b.w;
    if (BYTES == 64) {
      const uint4 c = ld<MODE>(table + 4 * s + 2), d = ld<MODE>(table + 4 * s + 3);
      acc.x ^= c.x ^ d.x; acc.y ^= c.y ^ d.y; acc.z ^= c.z ^ d.z; acc.w ^= c.w ^ d.w;
    }
  }
  if (acc.x == 0x12345678u) out[t] = acc;
}
int main(int argc, char** argv) {
  const int limit = argc > 1 ? atoi(argv[1]) : 0;
  if (limit) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)limit);
    printf("cudaDeviceSetLimit(MaxL2FetchGranularity, %d): %s\n", limit, cudaGetErrorString(e));
  }
  size_t got = 0;
  cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
  printf("MaxL2FetchGranularity = %zu\n", got);
  const unsigned long long slots = 1ull << 24;          // 16 M points x 64 B = 1 GiB
  uint4 *table, *out;
  cudaMalloc(&table, slots * 64);
  cudaMemset(table, 1, slots * 64);
  cudaMalloc(&out, (size_t)148 * 8 * 256 * 16);
  const unsigned per = 64;
  const dim3 grid(148 * 8), block(256);                 // 303 K threads x 64 gathers = 19.4 M gathers = 1.24 GB of points
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(MODE, BYTES, NAME)                                                           \
  gather<MODE, BYTES><<<grid, block>>>(table, slots, per, out);                          \
  cudaDeviceSynchronize();                                                               \
  cudaEventRecord(e0);                                                                   \
  gather<MODE, BYTES><<<grid, block>>>(table, slots, per, out);                          \
  cudaEventRecord(e1); cudaEventSynchronize(e1);                                         \
  { float ms; cudaEventElapsedTime(&ms, e0, e1);                                         \
    printf("%-28s %8.3f ms  %7.1f GB/s useful\n", NAME, ms, 148.0 * 8 * 256 * per * BYTES / ms / 1e6); }
  RUN(0, 64, "64B plain")
  RUN(1, 64, "64B .L2::64B")
  RUN(2, 64, "64B .L2::128B")
  RUN(3, 64, "64B .nc")
  RUN(4, 64, "64B .cg")
  RUN(0, 32, "32B plain")
  RUN(1, 32, "32B .L2::64B")
  RUN(4, 32, "32B .cg")
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
