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
// PERSISTENT: the grid is 2 CTAs per SM; every CTA walks a static, strided list of work items
// (128-query tile, utterance, head) and treats the key tiles of all its items as ONE flattened
// stream, so barrier init, TMEM allocation, the penalty LUT and -- above all -- the TMA latency
// of the first Q/K/V tiles of an item are paid once per CTA instead of once per item (the
// one-CTA-per-item version spent ~45 % of its life in those prologues: profiles/r01a).  The K/V
// rings, the S/P double buffers and all mbarrier phases simply keep counting across items.
//
//   warp 0     TMA producer (Q single-buffered; K 3-stage, V 2-stage rings)
//   warp 1     TMEM allocator + MMA issuer (QK runs two key tiles ahead of PV, across items)
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
// penalty LUT: entry o <-> (key - query) = o - lut_off, lut_off = nq*128; keys < nkv*64
static inline int attention_lut_floats(int L) {
  return ((L + AT_BM - 1) / AT_BM) * AT_BM + ((L + AT_BN - 1) / AT_BN) * AT_BN;
}
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

// Static work list of a CTA: items w = blockIdx.x, blockIdx.x + gridDim.x, ...; item index
// w = (b*H + h)*nq + q_tile, so CTAs running side by side share K/V of one (b, h) through L2.
struct AttnItem {
  int w, q0, b, h, len, n_kv;  // n_kv == 0: tile of padded queries (zero fill, no pipeline work)
};
__device__ __forceinline__ void attn_decode(AttnItem& it, const int* __restrict__ lengths, int L,
                                            int H, int nq, int n_items) {
  if (it.w >= n_items) return;
  const int qt = it.w % nq, bh = it.w / nq;
  it.b = bh / H;
  it.h = bh - it.b * H;
  it.q0 = qt * AT_BM;
  it.len = min(__ldg(lengths + it.b), L);
  it.n_kv = (it.q0 < it.len) ? (it.len + AT_BN - 1) / AT_BN : 0;
}
// advance to the next item that has pipeline work (n_kv > 0)
__device__ __forceinline__ void attn_next_work(AttnItem& it, const int* __restrict__ lengths, int L,
                                               int H, int nq, int n_items, bool first) {
  if (!first) it.w += gridDim.x;
  for (;;) {
    attn_decode(it, lengths, L, H, nq, n_items);
    if (it.w >= n_items || it.n_kv > 0) return;
    it.w += gridDim.x;
  }
}

template <int LOGPEN>
__global__ void __launch_bounds__(192, 2)
    attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         __nv_bfloat16* __restrict__ out, const int* __restrict__ lengths, int L, int B,
                         int H) {
  const int D = H * AT_HD;
  const int nq = (L + AT_BM - 1) / AT_BM;
  const int n_items = nq * B * H;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (keeps the shared address space visible to ptxas)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_QB;            // AT_KST stages
  uint8_t* sV = sK + AT_KST * AT_KB;   // AT_VST stages
  uint8_t* sP = sV + AT_VST * AT_KB;   // 2 x [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_QB);
  float* sLut = reinterpret_cast<float*>(bars + 32);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [2]
  uint64_t* v_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* pv_done = bars + 15;  // [2]
  uint64_t* q_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
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
  const uint32_t tmem_O = tmem_base + 128;  // S[0] = +0, S[1] = +64, O = +128 (64 columns each)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      AttnItem it;
      it.w = blockIdx.x;
      uint32_t g = 0, n = 0;  // g: key tiles loaded so far, n: items loaded so far
      for (attn_next_work(it, lengths, L, H, nq, n_items, true); it.w < n_items;
           attn_next_work(it, lengths, L, H, nq, n_items, false), ++n) {
        const int cq = it.h * AT_HD, ck = D + cq, cv = 2 * D + cq;
        mbar_wait(q_empty, (n & 1) ^ 1);  // every QK of the previous item has completed
        mbar_arrive_expect_tx(q_full, AT_QB);
        tma_load_3d(sQ, &tmQ, q_full, cq, it.b, it.q0);
        for (int j = 0; j < it.n_kv; ++j, ++g) {
          const uint32_t ks = g % AT_KST, vs = g & 1;
          mbar_wait(&k_empty[ks], ((g / AT_KST) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], AT_KB);
          tma_load_3d(sK + ks * AT_KB, &tmKV, &k_full[ks], ck, it.b, j * AT_BN);
          mbar_wait(&v_empty[vs], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], AT_KB);
          tma_load_3d(sV + vs * AT_KB, &tmKV, &v_full[vs], cv, it.b, j * AT_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = idesc_bf16_f32(AT_BM, AT_BN, 0, 0);
      constexpr uint32_t IDESC_PV = idesc_bf16_f32(AT_BM, AT_HD, 0, 1);
      const uint64_t qdesc = desc_kmajor_sw128(smem_u32(sQ));
      AttnItem qk, pv;  // two cursors over the same item list: QK runs two key tiles ahead of PV
      qk.w = pv.w = blockIdx.x;
      attn_next_work(qk, lengths, L, H, nq, n_items, true);
      attn_next_work(pv, lengths, L, H, nq, n_items, true);
      int qk_j = 0, pv_j = 0;
      uint32_t gq = 0, gp = 0, nqk = 0;  // global key-tile counters, items started by QK
      auto issue_qk = [&]() {
        if (qk_j == 0) mbar_wait(q_full, nqk & 1);
        const uint32_t ks = gq % AT_KST;
        mbar_wait(&k_full[ks], (gq / AT_KST) & 1);
        tc_fence_after();
        const uint64_t kdesc = desc_kmajor_sw128(smem_u32(sK + ks * AT_KB));
        const uint32_t d_tmem = tmem_base + (gq & 1) * AT_BN;
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k)
          umma_bf16_ss(d_tmem, qdesc + 2 * k, kdesc + 2 * k, IDESC_QK, k != 0);
        umma_commit(&s_full[gq & 1]);
        umma_commit(&k_empty[ks]);
        ++gq;
        if (++qk_j == qk.n_kv) {
          umma_commit(q_empty);  // Q may be overwritten once these MMAs have completed
          qk_j = 0;
          ++nqk;
          attn_next_work(qk, lengths, L, H, nq, n_items, false);
        }
      };
      if (qk.w < n_items) issue_qk();
      if (qk.w < n_items) issue_qk();
      while (pv.w < n_items) {
        const uint32_t pb = gp & 1, ph = (gp >> 1) & 1;
        mbar_wait(&p_full[pb], ph);  // S consumed, P in smem, O rescaled / read out if needed
        mbar_wait(&v_full[pb], ph);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + pb * AT_QB), va = smem_u32(sV + pb * AT_KB);
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk)
          umma_bf16_ss(tmem_O, desc_kmajor_sw128(pa) + 2 * kk, desc_mnmajor_sw128(va + kk * 2048, AT_KB),
                       IDESC_PV, (pv_j | kk) != 0);
        umma_commit(&pv_done[pb]);
        umma_commit(&v_empty[pb]);
        ++gp;
        if (++pv_j == pv.n_kv) {
          pv_j = 0;
          attn_next_work(pv, lengths, L, H, nq, n_items, false);
        }
        if (qk.w < n_items) issue_qk();
      }
    }
  } else {
    // ---- softmax / correction / output: thread <-> query row
    const int q = (warp & 3) * 32 + lane;  // row in the tile == TMEM lane
    const int st = threadIdx.x - 64;       // 0..127
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int swz = q & 7;
    // Penalty LUT, once per CTA: entry o <-> (key - query) = o - lut_off;
    // pen2 = log2(max(1, |key - query|)).
    const int lut_off = nq * AT_BM;
    if (LOGPEN) {
      const int n_lut = lut_off + ((L + AT_BN - 1) / AT_BN) * AT_BN;
      for (int o = st; o < n_lut; o += 128) {
        const int d = abs(o - lut_off);
        sLut[o] = (d > 1) ? __log2f((float)d) : 0.0f;
      }
      named_bar_sync(1, 128);
    }
    uint32_t g = 0;  // key tiles consumed so far (all items)
    AttnItem it;
    for (it.w = blockIdx.x; it.w < n_items; it.w += gridDim.x) {
      attn_decode(it, lengths, L, H, nq, n_items);
      const int i = it.q0 + q;
      __nv_bfloat16* orow = out + ((size_t)i * B + it.b) * D + it.h * AT_HD;
      if (it.n_kv == 0) {  // tile of padded queries: defined (finite) output, no pipeline work
        if (i < L) {
          uint4* op = reinterpret_cast<uint4*>(orow);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      float m_used = -INFINITY, l = 0.0f;
      for (int j = 0; j < it.n_kv; ++j, ++g) {
        const uint32_t sb = g & 1, ph = (g >> 1) & 1;
        const int k0 = j * AT_BN;
        const int nvalid = min(AT_BN, it.len - k0);
        mbar_wait(&s_full[sb], ph);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld32(tmem_base + lane_addr + sb * AT_BN, s0);
        tmem_ld32(tmem_base + lane_addr + sb * AT_BN + 32, s1);
        tmem_ld_wait();
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
            mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // O of the previous tile final
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
        if (g >= 2) mbar_wait(&pv_done[sb], ph ^ 1);  // P buffer sb free (PV of tile g-2 done)
        const float* lrow = sLut + (lut_off - i) + k0;  // penalty of key k0+c for this row: lrow[c]
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
        // stage C: row sum (4 chains) + bf16 pack -> swizzled K-major P row
        float sm4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 32; ++c) sm4[c & 3] += __uint_as_float(s0[c]) + __uint_as_float(s1[c]);
        const float sum = (sm4[0] + sm4[1]) + (sm4[2] + sm4[3]);
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const uint32_t* sv = (gq < 4) ? s0 : s1;
          const int o = (gq & 3) * 8;
          prow[gq ^ swz] = make_uint4(
              pack_bf16x2(__uint_as_float(sv[o]), __uint_as_float(sv[o + 1])),
              pack_bf16x2(__uint_as_float(sv[o + 2]), __uint_as_float(sv[o + 3])),
              pack_bf16x2(__uint_as_float(sv[o + 4]), __uint_as_float(sv[o + 5])),
              pack_bf16x2(__uint_as_float(sv[o + 6]), __uint_as_float(sv[o + 7])));
        }
        l += sum;
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&p_full[sb]);
      }
      // item epilogue: O / l -> bf16 -> global.  The next item's first PV (accumulate = 0) is issued
      // only after all 128 threads arrive on its p_full, i.e. after every thread has read O here.
      mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_O + lane_addr, o0);
      tmem_ld32(tmem_O + lane_addr + 32, o1);
      tmem_ld_wait();
      if (i < L) {
        uint4* op = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          const uint32_t* v = (gq < 4) ? o0 : o1;
          const int o = (gq & 3) * 8;
          op[gq] = make_uint4(
              pack_bf16x2(__uint_as_float(v[o]) * inv, __uint_as_float(v[o + 1]) * inv),
              pack_bf16x2(__uint_as_float(v[o + 2]) * inv, __uint_as_float(v[o + 3]) * inv),
              pack_bf16x2(__uint_as_float(v[o + 4]) * inv, __uint_as_float(v[o + 5]) * inv),
              pack_bf16x2(__uint_as_float(v[o + 6]) * inv, __uint_as_float(v[o + 7]) * inv));
        }
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
  const int smem = attention_smem_bytes(log_penalty ? L : 0);
  FBKST_REQUIRE(smem <= 227 * 1024, "fbkst_attention_fwd: L=%d needs %d B of shared memory", L, smem);
  // persistent grid: as many CTAs as fit (2 per SM up to L ~ 3000), never more than work items
  const int per_sm = (2 * (smem + 1024) <= 228 * 1024) ? 2 : 1;
  const long long n_items = (long long)((L + AT_BM - 1) / AT_BM) * B * H;
  FBKST_REQUIRE(n_items < (1ll << 31), "fbkst_attention_fwd: too many work items");
  int grid = num_sms() * per_sm;
  if (grid > n_items) grid = (int)n_items;
  if (log_penalty)
    attention_fwd_kernel<1><<<grid, 192, smem, st>>>(tmQ, tmKV, (__nv_bfloat16*)out, lengths, L, B, H);
  else
    attention_fwd_kernel<0><<<grid, 192, smem, st>>>(tmQ, tmKV, (__nv_bfloat16*)out, lengths, L, B, H);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
