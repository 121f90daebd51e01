// Issue / execution cost of one tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, K = 16) as a function of N and of
// where the A operand lives (shared memory descriptor vs tensor memory), back to back from ONE thread per CTA and
// from two warps at once.  Motivation: the attention kernel's 128 x 64 x 16 UMMAs keep the tensor pipe 15 % busy
// while every PV product completes thousands of cycles after it was requested (profiles/r02w_attention_timeline.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fbk-fairseq-st_b200/csrc \
//        -o fbk-fairseq-st_b200/build/umma_probe scripts/probes/umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace fbkst;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode: 0 SS back to back | 1 TS back to back | 2 SS with a commit after every 4 | 3 SS from two warps (own D each)
// 4 / 5 / 6: SS from one thread, round-robin over 2 / 4 / 3 independent accumulators (D columns i * (512 / n))
template <int N, int MM = 128>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;              // 128 rows x 128 B
  uint8_t* sB = smem + 16384;      // N rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = idesc_bf16_f32(MM, N, 0, 0);
  const uint64_t adesc = desc_kmajor_sw128(smem_u32(sA)), bdesc = desc_kmajor_sw128(smem_u32(sB));
  const int issuers = (mode == 3) ? 2 : (mode == 7 ? 4 : 1);
  if (warp < issuers) {
    const uint32_t d = tmem + warp * (mode == 7 ? 128 : 256);  // D: N <= 256 (128 with four issuers) columns per issuer
    const uint32_t a_t = tmem + 448;               // TS: A (128 x 16 bf16 = 8 columns per k-step) in the last columns
    // Warp-uniform control flow with ONE elected lane issuing, as the production kernels do: `if (lane == 0)` makes
    // ptxas wrap every UTCHMMA in an ELECT / BRA.U.ANY waterfall that costs ~190 cycles per instruction by itself
    // (build with -DPROBE_LANE0 to see that).
    long long t0 = 0, t1 = 0, t2 = 0;
#ifdef PROBE_LANE0
    if (lane == 0) {
#else
    {
#endif
      t0 = clock64();
      uint32_t ph = 0;
      for (int r = 0; r < reps; ++r) {
#ifndef PROBE_LANE0
        if (elect_one()) {
#else
        {
#endif
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (mode == 1) {
              umma_bf16_ts(d, a_t + 8 * k, bdesc + 2 * k, idesc, 1);
            } else if (mode >= 4 && mode <= 6) {
              const int nacc = mode == 4 ? 2 : (mode == 5 ? 4 : 3);
              for (int a = 0; a < nacc; ++a)
                umma_bf16_ss(tmem + a * (mode == 6 ? 128 : 512 / nacc), adesc + 2 * k, bdesc + 2 * k, idesc, 1);
            } else {
              umma_bf16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
            }
          }
          if (mode == 2) umma_commit(&bars[warp]);
        }
        __syncwarp();
        if (mode == 2) {
          mbar_wait(&bars[warp], ph);
          ph ^= 1;
        }
      }
      t1 = clock64();
      if (mode != 2) {
        if (elect_one()) umma_commit(&bars[warp]);
        __syncwarp();
        mbar_wait(&bars[warp], 0);
      }
      t2 = clock64();
      if (blockIdx.x == 0 && lane == 0) {
        out[warp * 2 + 0] = t1 - t0;  // issue loop
        out[warp * 2 + 1] = t2 - t0;  // until everything has completed
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, int MM = 128>
static void run(const char* what, int mode, int reps) {
  long long* d;
  cudaMalloc(&d, 128);
  cudaMemset(d, 0, 128);
  const int smem = 16384 + 32768 + 1024 + 256;
  cudaFuncSetAttribute(probe<N, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N, MM><<<148, 128, smem>>>(d, reps, mode);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const double n = 4.0 * reps * (mode == 4 ? 2 : mode == 5 ? 4 : mode == 6 ? 3 : 1);
  printf("%-34s M=%3d N=%3d  issue %7.1f clk/UMMA  complete %7.1f clk/UMMA  (math at full rate %5.1f)%s", what, MM, N,
         h[0] / n, h[1] / n, (double)MM * N * 16 * 2 / 8192.0, mode == 3 ? "" : "\n");
  if (mode == 3) printf("  | warp 1: issue %7.1f complete %7.1f\n", h[2] / n, h[3] / n);
  if (mode == 7) printf("four issuing warps N=%d: complete %7.1f %7.1f %7.1f %7.1f clk/UMMA each\n", N, h[1] / n, h[3] / n, h[5] / n, h[7] / n);
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int reps = 2000;
  run<64>("SS back to back", 0, reps);
  run<128>("SS back to back", 0, reps);
  run<256>("SS back to back", 0, reps);
  run<64, 64>("SS back to back, M = 64", 0, reps);
  run<128, 64>("SS back to back, M = 64", 0, reps);
  run<256, 64>("SS back to back, M = 64", 0, reps);
  run<128, 64>("SS, two issuing warps, M = 64", 3, reps);
  run<256>("SS, two issuing warps", 3, reps);
  run<64>("SS, commit + wait after every 4", 2, reps);
  run<128>("SS, commit + wait after every 4", 2, reps);
  run<64>("SS, two issuing warps", 3, reps);
  run<128>("SS, two issuing warps", 3, reps);
  run<64>("SS, four issuing warps", 7, reps);
  run<128>("SS, four issuing warps", 7, reps);
  run<64>("SS, 2 accumulators round-robin", 4, reps);
  run<128>("SS, 2 accumulators round-robin", 4, reps);
  run<256>("SS, 2 accumulators round-robin", 4, reps);
  run<64>("SS, 4 accumulators round-robin", 5, reps);
  run<128>("SS, 4 accumulators round-robin", 5, reps);
  run<128>("SS, 3 accumulators round-robin", 6, reps);
  run<64>("TS (A in TMEM) back to back", 1, reps);
  run<128>("TS (A in TMEM) back to back", 1, reps);
  run<256>("TS (A in TMEM) back to back", 1, reps);
  return 0;
}
