// Issue / pipe cost of the instructions of the attention exp pass, per warp instruction, with W warps per SM
// (W = 8: two per scheduler, as in the wide kernel; W = 16: four, as in the decoupled kernel):
//   MUFU.EX2 | F2FP.BF16.F32.PACK_AB | FADD2 | FFMA2 | FMNMX3 | the mix of one 8-score chunk
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void fadd2(float& x, float& y, float a, float b) {
  uint64_t r, p, q;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x), "f"(y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(a), "f"(b));
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p), "l"(q));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r));
}
__device__ __forceinline__ void ffma2(float& x, float& y, float a, float b) {
  uint64_t r, p, q;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x), "f"(y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(a), "f"(b));
  asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(r) : "l"(p), "l"(q));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r));
}

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
  float a[16];
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = seed * (threadIdx.x + j) * 1e-3f - 3.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = j;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 5) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
    }
    if (MODE == 1) {  // 8 chains cvt -> shl -> cvt (the shift keeps the input loop-variant; MODE 6 times it alone)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(a[j]), "f"(a[j + 8]));
        a[j] = __uint_as_float(h[j] << 16);
      }
    }
    if (MODE == 6) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("shl.b32 %0, %1, 16;" : "=r"(h[j]) : "r"(__float_as_uint(a[j]) + h[j]));
        a[j] = __uint_as_float(h[j]);
      }
    }
    if (MODE == 5) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(a[2 * j]), "f"(a[2 * j + 1]));
        a[2 * j + 8] += __uint_as_float(h[j] << 16);
      }
    }
    if (MODE == 2 || MODE == 5) {
#pragma unroll
      for (int j = 0; j < 4; ++j) fadd2(a[2 * j + 8], a[2 * j + 9], 1.0f, 0.5f);
      if (MODE == 5) {
#pragma unroll
        for (int j = 0; j < 4; ++j) fadd2(a[2 * j + 8], a[2 * j + 9], 0.25f, 0.125f);
      }
    }
    if (MODE == 3 || MODE == 5) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ffma2(a[2 * j + 8], a[2 * j + 9], 0.999f, 1e-3f);
    }
    if (MODE == 4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(a[j + 8]), "f"(seed));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) s += __uint_as_float(h[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
static void run(const char* what, int warps, int per_iter) {
  float* o;
  long long* c;
  cudaMalloc(&o, 148 * 1024 * 4);
  cudaMalloc(&c, 8);
  const int iters = 2000;
  probe<MODE><<<148, warps * 32>>>(o, c, iters, 1.0f);
  probe<MODE><<<148, warps * 32>>>(o, c, iters, 1.0f);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  // cycles per warp instruction on one scheduler: the warps/4 warps of a scheduler each issue per_iter per iteration
  printf("%-44s %2d warps/SM: %6.2f clk per warp instruction and scheduler (%5.1f clk per iteration and warp)\n", what,
         warps, (double)h / iters / per_iter / (warps / 4), (double)h / iters);
  cudaFree(o);
  cudaFree(c);
}

int main() {
  for (int w = 4; w <= 16; w *= 2) {
    run<0>("MUFU.EX2 x8", w, 8);
    run<1>("F2FP.BF16.F32.PACK_AB x8 (+ 8 SHL)", w, 8);
    run<6>("IADD + SHL x8", w, 8);
    run<2>("FADD2 x4", w, 4);
    run<3>("FFMA2 x4", w, 4);
    run<4>("FMNMX3 x8", w, 8);
    run<5>("chunk mix: 8 MUFU + 4 F2FP + 8 FADD2 + 4 FFMA2 (+ 4 SHL, 4 FADD)", w, 24);
  }
  return 0;
}
