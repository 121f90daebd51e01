// Latency of ONE TMA box load as the attention kernel issues them: a {64 bf16, 1, ROWS} box of the [L, B, 3D] qkv
// tensor (rows of 128 B, 196 608 B apart at cfg2), from L2-resident data, issued by one thread of one CTA per SM
// (all 148 at once) or by CTA 0 alone.  Also a dense 2-D box of the same bytes for comparison.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fbk-fairseq-st_b200/csrc -lcuda \
//        -o fbk-fairseq-st_b200/build/tma_probe scripts/probes/tma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace fbkst;

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, long long* out, int reps, int rows_bytes,
                                                int only_cta0, int L, int B) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && (!only_cta0 || blockIdx.x == 0)) {
    long long tot = 0, mx = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      const int b = (blockIdx.x * 7 + r) % B, t0 = ((blockIdx.x + r * 5) % 4) * 64, c0 = ((blockIdx.x + r) % 24) * 64;
      const long long t = clock64();
      mbar_arrive_expect_tx(bar, rows_bytes);
      tma_load_3d(smem, &tm, bar, c0, b, t0);
      mbar_wait(bar, ph);
      ph ^= 1;
      const long long d = clock64() - t;
      tot += d;
      mx = d > mx ? d : mx;
    }
    if (blockIdx.x == 0) {
      out[0] = tot / reps;
      out[1] = mx;
    }
  }
}

int main() {
  const int L = 375, B = 64, D3 = 1536;
  __nv_bfloat16* q;
  cudaMalloc(&q, (size_t)L * B * D3 * 2);
  cudaMemset(q, 0, (size_t)L * B * D3 * 2);
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  for (int rows : {16, 32, 64, 128}) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)D3, (cuuint64_t)B, (cuuint64_t)L};
    cuuint64_t strides[2] = {(cuuint64_t)D3 * 2, (cuuint64_t)B * D3 * 2};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)rows}, es[3] = {1, 1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box, es,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
    for (int only0 : {1, 0}) {
      for (int pass = 0; pass < 2; ++pass) {  // second pass: L2-warm
        probe<<<148, 128, 70 * 1024>>>(tm, d, 200, rows * 128, only0, L, B);
        cudaDeviceSynchronize();
      }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("box {64 x bf16, 1, %3d rows} (%5d B), %s: mean %6lld cycles, max %6lld   %s\n", rows, rows * 128,
             only0 ? "CTA 0 alone        " : "148 CTAs at once   ", h[0], h[1], cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
