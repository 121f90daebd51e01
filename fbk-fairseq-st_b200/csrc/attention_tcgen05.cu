// Fused self-attention core for the ST encoder (reference: local_attention.py:115-139 with
// LogPenalty conv_transformer_layer.py:22-27): flash-style streaming softmax with
//   S = Q K^T      tcgen05.mma, both operands K-major in smem (TMA, 128B swizzle), S in TMEM
//   P = softmax    128 threads, one query row each: key-padding mask from lengths, the
//                  log-distance penalty as a 255-entry per-tile LUT in log2 domain, fp32 exp2
//   O += P V       tcgen05.mma, P written to smem as bf16 (K-major), V consumed MN-major
// One CTA = one (128-query tile, utterance, head); 2 CTAs per SM interleave MMA and softmax.
//
//   warp 0     TMA producer (Q once, K double-buffered, V single-buffered)
//   warp 1     TMEM allocator + MMA issuer
//   warps 2-5  softmax / correction / output (TMEM lane quarter = warp % 4)
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per tile
constexpr int AT_HD = 64;    // head dim (all reference archs: embed_dim / heads = 64)
constexpr int AT_TILE = AT_BM * AT_HD * 2;  // 16 KB
constexpr int AT_SMEM = AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + AT_TILE /*V*/ + 2 * AT_TILE /*P*/ +
                        2 * 256 * 4 /*LUT*/ + 128 /*barriers*/ + 1024 /*align*/;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int LOGPEN>
__global__ void __launch_bounds__(192, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out,
                         const int* __restrict__ lengths, int L, int B, int H) {
  const int D = H * AT_HD;
  const int q0 = blockIdx.x * AT_BM;
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int len = min(__ldg(lengths + b), L);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (q0 >= len) {  // tile of padded queries: defined output, no work (CTA-uniform exit)
    const int tid = threadIdx.x;
    if (tid < 128) {
      const int i = q0 + tid;
      if (i < L) {
        uint4* op = reinterpret_cast<uint4*>(out + ((size_t)i * B + b) * D + h * AT_HD);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;      // 2 stages
  uint8_t* sV = sK + 2 * AT_TILE;  // 1 stage
  uint8_t* sP = sV + AT_TILE;      // 128 x 128 bf16 as two K-major 64-column halves
  float* sLut = reinterpret_cast<float*>(sP + 2 * AT_TILE);  // 2 x 256
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 512);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* pv_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int n_kv = (len + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1);
    mbar_init(&k_full[1], 1);
    mbar_init(&k_empty[0], 1);
    mbar_init(&k_empty[1], 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  if (warp == 0) {
    if (lane == 0) {
      const int cq = h * AT_HD, ck = D + h * AT_HD, cv = 2 * D + h * AT_HD;
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_3d(sQ, &tmQKV, q_full, cq, b, q0);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j & 1;
        mbar_wait(&k_empty[ks], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], AT_TILE);
        tma_load_3d(sK + ks * AT_TILE, &tmQKV, &k_full[ks], ck, b, j * AT_BN);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, AT_TILE);
        tma_load_3d(sV, &tmQKV, v_full, cv, b, j * AT_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = idesc_bf16_f32(AT_BM, AT_BN, 0, 0);
      constexpr uint32_t IDESC_PV = idesc_bf16_f32(AT_BM, AT_HD, 0, 1);
      const uint64_t qdesc = desc_kmajor_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int j) {
        const int ks = j & 1;
        mbar_wait(&k_full[ks], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t kdesc = desc_kmajor_sw128(smem_u32(sK + ks * AT_TILE));
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16_ss(tmem_S, qdesc + 2 * k, kdesc + 2 * k, IDESC_QK, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[ks]);
      };
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(p_full, j & 1);  // S_j consumed, P_j in smem, O rescaled
        mbar_wait(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {
          const uint64_t pdesc =
              desc_kmajor_sw128(smem_u32(sP + (kk >> 2) * AT_TILE)) + 2 * (kk & 3);
          const uint64_t vdesc = desc_mnmajor_sw128(smem_u32(sV + kk * 2048), AT_TILE);
          umma_bf16_ss(tmem_O, pdesc, vdesc, IDESC_PV, (j | kk) != 0);
        }
        umma_commit(pv_done);
        umma_commit(v_empty);
        if (j + 1 < n_kv) issue_qk(j + 1);
      }
    }
  } else {
    // ---- softmax / correction / output: thread <-> query row
    const int q = (warp & 3) * 32 + lane;  // row in the tile == TMEM lane
    const int st = threadIdx.x - 64;       // 0..127, LUT writer index
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float m = -INFINITY, l = 0.0f;
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * AT_BN;
      const int nvalid = min(AT_BN, len - k0);
      float* lut = sLut + (j & 1) * 256;
      if (LOGPEN) {
        const int delta = k0 - q0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int idx = st + r * 128;
          const int d = abs(delta + idx - 127);
          lut[idx] = (d > 1) ? __log2f((float)d) : 0.0f;
        }
      }
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max of the raw scores over the valid keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (c * 32 + jj < nvalid) mx = fmaxf(mx, __uint_as_float(v[jj]));
      }
      const float m_new = fmaxf(m, mx * kLog2e);
      const float alpha = exp2f(m - m_new);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // O_{j-1} final, P smem free
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_O + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) * alpha);
          tmem_st32(tmem_O + lane_addr + c * 32, v);
        }
        tmem_st_wait();
      }
      if (LOGPEN) named_bar_sync(1, 128);  // LUT_j complete
      // pass 2: p = exp2(s*log2e - pen2 - m_new), P -> smem (bf16, swizzled K-major)
      float sum = 0.0f;
      const float* lrow = lut + (127 - q);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          float t0 = fmaf(__uint_as_float(v[jj]), kLog2e, -m_new);
          float t1 = fmaf(__uint_as_float(v[jj + 1]), kLog2e, -m_new);
          if (LOGPEN) {
            t0 -= lrow[c * 32 + jj];
            t1 -= lrow[c * 32 + jj + 1];
          }
          const float p0 = (c * 32 + jj < nvalid) ? exp2f(t0) : 0.0f;
          const float p1 = (c * 32 + jj + 1 < nvalid) ? exp2f(t1) : 0.0f;
          sum += p0 + p1;
          pk[jj >> 1] = pack_bf16x2(p0, p1);
        }
        uint8_t* prow = sP + (c >> 1) * AT_TILE + q * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = ((c & 1) * 4 + g) ^ (q & 7);
          *reinterpret_cast<uint4*>(prow + ch * 16) =
              make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      l = l * alpha + sum;
      m = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const int i = q0 + q;
    const float inv = 1.0f / l;
    uint4* op = reinterpret_cast<uint4*>(out + ((size_t)i * B + b) * D + h * AT_HD);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (i < L) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          op[c * 4 + g] = make_uint4(
              pack_bf16x2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_attention_fwd(const void* qkv, void* out, const int32_t* lengths, int L, int B,
                                   int H, int log_penalty, fbkst_stream_t stream) {
  FBKST_REQUIRE(qkv && out && lengths, "fbkst_attention_fwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && H > 0, "fbkst_attention_fwd: bad shape L=%d B=%d H=%d", L, B, H);
  FBKST_REQUIRE((long long)B * H <= 65535, "fbkst_attention_fwd: B*H=%d exceeds the grid limit", B * H);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = H * AT_HD;
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)B, (uint64_t)L};
  uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)B * 3 * D * 2};
  uint32_t box[3] = {AT_HD, 1, AT_BM};
  int rc = make_tensor_map(&tm, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box,
                           nullptr);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    configured = true;
  }
  dim3 grid((L + AT_BM - 1) / AT_BM, B * H);
  if (log_penalty)
    attention_fwd_kernel<1><<<grid, 192, AT_SMEM, st>>>(tm, (__nv_bfloat16*)out, lengths, L, B, H);
  else
    attention_fwd_kernel<0><<<grid, 192, AT_SMEM, st>>>(tm, (__nv_bfloat16*)out, lengths, L, B, H);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
