// Fused self-attention core for the ST encoder (reference: local_attention.py:115-139 with
// LogPenalty conv_transformer_layer.py:22-27): flash-style streaming softmax on tcgen05.
//
//   S_j = Q K_j^T   tcgen05.mma (M128 x N64 x K64), operands K-major in smem (TMA, 128B swizzle),
//                   S double-buffered in TMEM so QK of tile j+1/j+2 runs under the softmax of tile j
//   P_j = softmax   128 threads, one query row each, the 64 scores of the row held in registers:
//                   key-padding mask (last tile only), log-distance penalty from a per-CTA LUT in
//                   log2 domain indexed by (key - query), lazy max (O is rescaled only when the
//                   running max grows by more than 2^8), exp2 in fp32
//   O  += P_j V_j   tcgen05.mma (M128 x N64 x K64), P as bf16 in swizzled smem (double-buffered),
//                   V consumed MN-major straight from its TMA tile
// One CTA = one (128-query tile, utterance, head); 2 CTAs per SM.
//
//   warp 0     TMA producer (Q once; K 3-stage, V 2-stage rings)
//   warp 1     TMEM allocator + MMA issuer
//   warps 2-5  softmax / correction / output (TMEM lane quarter = warp % 4)
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

constexpr int AT_BM = 128;  // queries per CTA
constexpr int AT_BN = 64;   // keys per tile
constexpr int AT_HD = 64;   // head dim (all reference archs: embed_dim / heads = 64)
constexpr int AT_QB = AT_BM * AT_HD * 2;  // 16 KB
constexpr int AT_KB = AT_BN * AT_HD * 2;  // 8 KB
constexpr int AT_KST = 3, AT_VST = 2;
// shared memory without the penalty LUT (its size depends on L: see attention_smem_bytes)
constexpr int AT_SMEM_FIXED = AT_QB + AT_KST * AT_KB + AT_VST * AT_KB + 2 * AT_QB /*P x2*/ +
                              256 /*barriers*/ + 1024 /*align*/;
static inline int attention_lut_floats(int L) { return ((L + AT_BN - 1) / AT_BN) * AT_BN + 128; }
static inline int attention_smem_bytes(int L) { return AT_SMEM_FIXED + 4 * attention_lut_floats(L); }
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Optional per-CTA timeline (debug only; null in production): 16 x int64 per CTA.
__device__ long long* g_attn_trace = nullptr;
#define AT_TRACE(slot)                                                                  \
  do {                                                                                  \
    if (trace != nullptr && threadIdx.x == 64) trace[slot] = clock64();                 \
  } while (0)

template <int LOGPEN>
__global__ void __launch_bounds__(192, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         __nv_bfloat16* __restrict__ out, const int* __restrict__ lengths, int L, int B,
                         int H) {
  const int D = H * AT_HD;
  const int q0 = blockIdx.x * AT_BM;
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int len = min(__ldg(lengths + b), L);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  long long* trace = g_attn_trace;
  if (trace != nullptr) {
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    trace = (cta < 4096) ? trace + (size_t)cta * 16 : nullptr;
    if (trace != nullptr && threadIdx.x == 64) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trace[15] = smid;
    }
  }
  AT_TRACE(0);

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
  // 1024-byte alignment by pointer arithmetic (keeps the shared address space visible to ptxas)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_QB;            // AT_KST stages
  uint8_t* sV = sK + AT_KST * AT_KB;   // AT_VST stages
  uint8_t* sP = sV + AT_VST * AT_KB;   // 2 x [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_QB);
  float* sLut = reinterpret_cast<float*>(bars + 32);  // [n_kv*64 + 128]
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [2]
  uint64_t* v_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* pv_done = bars + 15;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int n_kv = (len + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
    }
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
  AT_TRACE(1);
  const uint32_t tmem_O = tmem_base + 128;  // S[0] = +0, S[1] = +64, O = +128 (64 columns each)

  if (warp == 0) {
    if (lane == 0) {
      const int cq = h * AT_HD, ck = D + h * AT_HD, cv = 2 * D + h * AT_HD;
      mbar_arrive_expect_tx(q_full, AT_QB);
      tma_load_3d(sQ, &tmQ, q_full, cq, b, q0);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j % AT_KST, vs = j & 1;
        mbar_wait(&k_empty[ks], ((j / AT_KST) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], AT_KB);
        tma_load_3d(sK + ks * AT_KB, &tmKV, &k_full[ks], ck, b, j * AT_BN);
        mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], AT_KB);
        tma_load_3d(sV + vs * AT_KB, &tmKV, &v_full[vs], cv, b, j * AT_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = idesc_bf16_f32(AT_BM, AT_BN, 0, 0);
      constexpr uint32_t IDESC_PV = idesc_bf16_f32(AT_BM, AT_HD, 0, 1);
      const uint64_t qdesc = desc_kmajor_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int j) {
        const int ks = j % AT_KST;
        mbar_wait(&k_full[ks], (j / AT_KST) & 1);
        tc_fence_after();
        const uint64_t kdesc = desc_kmajor_sw128(smem_u32(sK + ks * AT_KB));
        const uint32_t d_tmem = tmem_base + (j & 1) * AT_BN;
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16_ss(d_tmem, qdesc + 2 * k, kdesc + 2 * k, IDESC_QK, k != 0);
        umma_commit(&s_full[j & 1]);
        umma_commit(&k_empty[ks]);
      };
      issue_qk(0);
      if (n_kv > 1) issue_qk(1);
      for (int j = 0; j < n_kv; ++j) {
        const int pb = j & 1;
        mbar_wait(&p_full[pb], (j >> 1) & 1);  // S_j consumed, P_j in smem, O rescaled if needed
        mbar_wait(&v_full[pb], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + pb * AT_QB), va = smem_u32(sV + pb * AT_KB);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk)
          umma_bf16_ss(tmem_O, desc_kmajor_sw128(pa) + 2 * kk, desc_mnmajor_sw128(va + kk * 2048, AT_KB),
                       IDESC_PV, (j | kk) != 0);
        umma_commit(&pv_done[pb]);
        umma_commit(&v_empty[pb]);
        if (j + 2 < n_kv) issue_qk(j + 2);
      }
    }
  } else {
    // ---- softmax / correction / output: thread <-> query row
    const int q = (warp & 3) * 32 + lane;  // row in the tile == TMEM lane
    const int st = threadIdx.x - 64;       // 0..127
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int swz = q & 7;
    // Penalty LUT, once per CTA: entry o <-> (key - query) = o - 127 - q0, i.e. key k and tile row r
    // read entry k - r + 127.  pen2 = log2(max(1, |key - query|)).
    if (LOGPEN) {
      const int n_lut = n_kv * AT_BN + 128;
      for (int o = st; o < n_lut; o += 128) {
        const int d = abs(o - 127 - q0);
        sLut[o] = (d > 1) ? __log2f((float)d) : 0.0f;
      }
      named_bar_sync(1, 128);
    }
    AT_TRACE(2);
    float m_used = -INFINITY, l = 0.0f;
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const int k0 = j * AT_BN;
      const int nvalid = min(AT_BN, len - k0);
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      if (j == 0) AT_TRACE(3);
      if (j == 2) AT_TRACE(12);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tmem_base + lane_addr + sb * AT_BN, s0);
      tmem_ld32(tmem_base + lane_addr + sb * AT_BN + 32, s1);
      tmem_ld_wait();
      if (j == 2) AT_TRACE(13);
      float mx = -INFINITY;
      if (nvalid == AT_BN) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // 4 independent chains
#pragma unroll
        for (int c = 0; c < 32; ++c)
          m4[c & 3] = fmaxf(m4[c & 3], fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < nvalid) mx = fmaxf(mx, __uint_as_float(s0[c]));
          if (c + 32 < nvalid) mx = fmaxf(mx, __uint_as_float(s1[c]));
        }
      }
      const float m_new = fmaxf(m_used, mx * kLog2e);
      // lazy rescale: only when some row of the warp grew by more than 2^8 (always at j == 0)
      const bool grow = m_new > m_used + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_next = grow ? m_new : m_used;
        if (j > 0) {
          const float alpha = ex2(m_used - m_next);
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // O_{j-1} final
          tc_fence_after();
          uint32_t o0[32], o1[32];
          tmem_ld32(tmem_O + lane_addr, o0);
          tmem_ld32(tmem_O + lane_addr + 32, o1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            o0[c] = __float_as_uint(__uint_as_float(o0[c]) * alpha);
            o1[c] = __float_as_uint(__uint_as_float(o1[c]) * alpha);
          }
          tmem_st32(tmem_O + lane_addr, o0);
          tmem_st32(tmem_O + lane_addr + 32, o1);
          tmem_st_wait();
          l *= alpha;
        }
        m_used = m_next;
      }
      if (j >= 2) mbar_wait(&pv_done[sb], ((j >> 1) & 1) ^ 1);  // P buffer sb free (PV_{j-2} done)
      if (j == 2) AT_TRACE(14);
      const float* lrow = sLut + (127 - q) + k0;  // penalty of key k0+c for this row: lrow[c]
      uint4* prow = reinterpret_cast<uint4*>(sP + sb * AT_QB + q * 128);
      const float negm = -m_used;
      // stage A (in place): t = s*log2e - m - pen2   (all 64 LDS independent -> full ILP)
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float t0 = fmaf(__uint_as_float(s0[c]), kLog2e, negm);
        float t1 = fmaf(__uint_as_float(s1[c]), kLog2e, negm);
        if (LOGPEN) {
          t0 -= lrow[c];
          t1 -= lrow[c + 32];
        }
        s0[c] = __float_as_uint(t0);
        s1[c] = __float_as_uint(t1);
      }
      // stage B (in place): p = 2^t, masked keys -> 0
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        s0[c] = __float_as_uint(ex2(__uint_as_float(s0[c])));
        s1[c] = __float_as_uint(ex2(__uint_as_float(s1[c])));
      }
      if (nvalid != AT_BN) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c >= nvalid) s0[c] = 0u;
          if (c + 32 >= nvalid) s1[c] = 0u;
        }
      }
      if (j == 2) AT_TRACE(8);
      // stage C: row sum (4 chains) + bf16 pack -> swizzled K-major P row
      float sm4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 32; ++c) sm4[c & 3] += __uint_as_float(s0[c]) + __uint_as_float(s1[c]);
      const float sum = (sm4[0] + sm4[1]) + (sm4[2] + sm4[3]);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* sv = (g < 4) ? s0 : s1;
        const int o = (g & 3) * 8;
        prow[g ^ swz] = make_uint4(
            pack_bf16x2(__uint_as_float(sv[o]), __uint_as_float(sv[o + 1])),
            pack_bf16x2(__uint_as_float(sv[o + 2]), __uint_as_float(sv[o + 3])),
            pack_bf16x2(__uint_as_float(sv[o + 4]), __uint_as_float(sv[o + 5])),
            pack_bf16x2(__uint_as_float(sv[o + 6]), __uint_as_float(sv[o + 7])));
      }
      l += sum;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[sb]);
      if (j < 4) AT_TRACE(4 + j);
      if (j == 5) AT_TRACE(9);
    }
    mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    AT_TRACE(10);
    tc_fence_after();
    const int i = q0 + q;
    const float inv = 1.0f / l;
    uint32_t o0[32], o1[32];
    tmem_ld32(tmem_O + lane_addr, o0);
    tmem_ld32(tmem_O + lane_addr + 32, o1);
    tmem_ld_wait();
    if (i < L) {
      uint4* op = reinterpret_cast<uint4*>(out + ((size_t)i * B + b) * D + h * AT_HD);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t* v = (g < 4) ? o0 : o1;
        const int o = (g & 3) * 8;
        op[g] = make_uint4(
            pack_bf16x2(__uint_as_float(v[o]) * inv, __uint_as_float(v[o + 1]) * inv),
            pack_bf16x2(__uint_as_float(v[o + 2]) * inv, __uint_as_float(v[o + 3]) * inv),
            pack_bf16x2(__uint_as_float(v[o + 4]) * inv, __uint_as_float(v[o + 5]) * inv),
            pack_bf16x2(__uint_as_float(v[o + 6]) * inv, __uint_as_float(v[o + 7]) * inv));
      }
    }
  }
  AT_TRACE(11);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace fbkst

// debug hook (not part of the public ABI): buffer of 4096*16 int64, or NULL to disable
extern "C" int fbkst_debug_set_attention_trace(long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(fbkst::g_attn_trace, &buf, sizeof(buf));
  return e == cudaSuccess ? 0 : -2;
}

using namespace fbkst;

extern "C" int fbkst_attention_fwd(const void* qkv, void* out, const int32_t* lengths, int L, int B,
                                   int H, int log_penalty, fbkst_stream_t stream) {
  FBKST_REQUIRE(qkv && out && lengths, "fbkst_attention_fwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && H > 0, "fbkst_attention_fwd: bad shape L=%d B=%d H=%d", L, B, H);
  FBKST_REQUIRE((long long)B * H <= 65535, "fbkst_attention_fwd: B*H=%d exceeds the grid limit", B * H);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = H * AT_HD;
  CUtensorMap tmQ, tmKV;
  uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)B, (uint64_t)L};
  uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)B * 3 * D * 2};
  uint32_t boxq[3] = {AT_HD, 1, AT_BM};
  uint32_t boxk[3] = {AT_HD, 1, AT_BN};
  int rc = make_tensor_map(&tmQ, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, boxq,
                           nullptr);
  if (rc) return rc;
  rc = make_tensor_map(&tmKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, boxk, nullptr);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int smem = attention_smem_bytes(L);
  FBKST_REQUIRE(smem <= 227 * 1024, "fbkst_attention_fwd: L=%d needs %d B of shared memory", L, smem);
  dim3 grid((L + AT_BM - 1) / AT_BM, B * H);
  if (log_penalty)
    attention_fwd_kernel<1><<<grid, 192, smem, st>>>(tmQ, tmKV, (__nv_bfloat16*)out, lengths, L, B, H);
  else
    attention_fwd_kernel<0><<<grid, 192, smem, st>>>(tmQ, tmKV, (__nv_bfloat16*)out, lengths, L, B, H);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
