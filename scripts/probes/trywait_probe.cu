// How long does a FAILING mbarrier.try_wait take (the hardware suspends the thread for a bounded time)?  Sets the
// real duration of the spin-count watchdog in ptx.cuh (2^26 attempts).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fbk-fairseq-st_b200/csrc -o scripts/probes/trywait_probe scripts/probes/trywait_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace fbkst;

__global__ void probe(unsigned long long* out, int n) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  int ok = 0;
  for (int i = 0; i < n; ++i) ok += mbar_try_wait(&bar, 0) ? 1 : 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = ok;
  }
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  for (int threads : {32, 128}) {
    const int n = 1 << 14;
    probe<<<1, threads>>>(d, n);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%3d threads: %d failing try_wait attempts in %.3f ms = %.2f us each -> 2^26 attempts = %.1f s\n", threads, n,
           h[0] * 1e-6, h[0] * 1e-3 / n, h[0] * 1e-9 / n * (double)(1 << 26));
  }
  return 0;
}
