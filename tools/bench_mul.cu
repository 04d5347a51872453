// Throughput micro-benchmark of the Montgomery product (tuning tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iplonky_b200/csrc [-DPLK_SQR_DEDICATED=0|1] tools/bench_mul.cu -o /tmp/bench_mul
// Every thread runs 4 independent chains of dependent products (like the 4-deep ILP of a mixed addition); prints
// products/s over the whole GPU and the checksum (identical across variants = same values).
#include <cstdio>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace plk;
template <class P, bool SQR>
__global__ void __launch_bounds__(128) mul_loop(const uint32_t* seed, int iters, uint32_t* out) {
  typedef Fp<P> F;
  F a[4], b;
  for (int k = 0; k < F::N; ++k) {
    b.l[k] = seed[k] ^ (threadIdx.x * 2654435761u);
    for (int c = 0; c < 4; ++c) a[c].l[k] = seed[F::N + k] + c + blockIdx.x;
  }
  b.l[F::N - 1] &= 0x0fffffffu;
  for (int c = 0; c < 4; ++c) a[c].l[F::N - 1] &= 0x0fffffffu;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = SQR ? F::sqr(a[c]) : F::mul(a[c], b);
  }
  uint32_t acc = 0;
  for (int c = 0; c < 4; ++c) for (int k = 0; k < F::N; ++k) acc ^= a[c].l[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <class P, bool SQR>
void run(const char* name) {
  const int blocks = 148 * 4, threads = 128, iters = 2000;
  uint32_t h_seed[32];
  for (int i = 0; i < 32; ++i) h_seed[i] = 0x9e3779b9u * (i + 1);
  uint32_t *d_seed, *d_out;
  cudaMalloc(&d_seed, sizeof(h_seed));
  cudaMalloc(&d_out, blocks * threads * 4);
  cudaMemcpy(d_seed, h_seed, sizeof(h_seed), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mul_loop<P, SQR><<<blocks, threads>>>(d_seed, 10, d_out);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    mul_loop<P, SQR><<<blocks, threads>>>(d_seed, iters, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  static uint32_t h_out[148 * 4 * 128];
  cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
  uint32_t sum = 0;
  for (auto v : h_out) sum = sum * 31 + v;
  const double prods = (double)blocks * threads * iters * 4;
  printf("%-22s: %.3f ms, %.3e products/s, checksum %08x (%s)\n", name, best, prods / (best * 1e-3), sum,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<TweedledeeBaseParams, false>("TweedledeeBase mul");
  run<TweedledeeBaseParams, true>("TweedledeeBase sqr");
  run<Bls12377BaseParams, false>("Bls12377Base mul");
  run<Bls12377BaseParams, true>("Bls12377Base sqr");
  return 0;
}
