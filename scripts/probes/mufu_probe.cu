// Throughput of the exponential candidates for the attention exp pass (elements per clock per SM):
//   ex2.approx.ftz.f32 | ex2.approx.f16x2 | ex2.approx.ftz.bf16x2 | degree-3 polynomial 2^x on the FMA pipe
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = seed * (threadIdx.x + j) * 1e-3f - 3.0f;
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = 0xB800B800u + j;  // packed halves / bfloat16s near -0.5
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
        a[j] -= 1.5f;
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[j]));
        h[j] ^= 0x80008000u;
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[j]));
        h[j] ^= 0x80008000u;
      } else {  // 2^x = 2^floor * p(frac): magic-number floor, degree-3 minimax, exponent add
        float x = fmaxf(a[j], -126.f);
        const float r = x + 12582912.f;  // round to nearest integer in the low mantissa bits
        const float fl = r - 12582912.f;
        const float f = x - fl;  // [-0.5, 0.5]
        float p = fmaf(f, 0.0555041f, 0.2402265f);
        p = fmaf(p, f, 0.6931472f);
        p = fmaf(p, f, 1.0f);
        a[j] = __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23)) - 1.5f;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j] + __uint_as_float(h[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int per_instr) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 4 * 512);
  const int iters = 20000;
  probe<MODE><<<sms * 4, 512>>>(out, 100, 1.f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<sms * 4, 512>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double elems = (double)sms * 4 * 512 * iters * 8 * per_instr;
  printf("%-28s %8.3f ms  %7.2f G elem/s  %6.2f elem/clk/SM (at %d MHz nominal)  err=%s\n", name, ms,
         elems / ms * 1e-6, elems / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("poly3 on the FMA pipe", 1);
  return 0;
}
