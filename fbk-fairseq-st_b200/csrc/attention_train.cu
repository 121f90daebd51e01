// Training-side self-attention of the ST encoder on tcgen05 (row T of the scope table):
// reference local_attention.py:115-139 (scores - log-distance penalty, key-padding mask, fp32 softmax,
// attention dropout, P V) and its autograd.  Three kernels from ONE skeleton:
//
//   FWD   own rows = 128 queries.  pass A: S_j = Q K_j^T -> exact row max / sum -> LSE;  pass B: S_j again,
//         P_j = 2^(s - LSE) (normalised), dropout, O += P_j V_j.  Writes O (bf16) and LSE (the backward needs it).
//   DQ    own rows = 128 queries.  S_j = Q K_j^T, dP_j = dO V_j^T, dS = P o (dP o keep - delta) * scale,
//         dQ += dS K_j.
//   DKV   own rows = 128 keys.     S^T_i = K Q_i^T, dP^T_i = V dO_i^T, dV += (P o keep)^T dO_i, dK += dS^T Q_i.
//
// The skeleton: a CTA of 128 threads owns a 128-row block of one (utterance, head) and walks the OTHER side
// in tiles of 64 rows.  Per tile: [TMA of the next tile] -> one or two S-type UMMAs (M128 x N64 x K64, operands
// K-major from TMA tiles) into TMEM -> thread <-> own row: TMEM -> registers -> exp2 / mask / dropout ->
// bf16 tile(s) in swizzled shared memory (the K-major A operand) -> accumulate UMMA(s) with the other tile
// consumed MN-major straight from its TMA bytes (the same tile serves K-major for S and MN-major for the
// accumulation).  All MMA operand patterns are the ones of the inference kernel (attention_tcgen05.cu).
// Recomputing S in the backward (instead of storing P) keeps HBM traffic at the q/k/v/dO level: no L^2 buffer.
//
// This is the correctness-first training path: one tile in flight per CTA, two CTAs per SM cover each other's
// round trips.  256 threads per CTA: TWO threads per own row (warps w and w + 4 share a TMEM lane quarter),
// each handling 32 of a tile's 64 columns -- the first version (one thread per row, 8 warps per SM) was bound by
// the latency of its dependent instruction chains (253 us per cfg2 layer for FWD, tensor pipe 5 %: r02k ncu).  Query rows at or beyond the utterance's length get zero output / zero gradient (the reference
// computes finite garbage there that nothing reads: the loss masks padded positions).
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

constexpr int TR_HD = 64;
constexpr int TR_OWN = 128;   // own rows per CTA
constexpr int TR_OTH = 64;    // other rows per tile
constexpr int TR_OWN_BYTES = TR_OWN * TR_HD * 2;  // 16 KB
constexpr int TR_OTH_BYTES = TR_OTH * TR_HD * 2;  // 8 KB
enum { TR_FWD = 0, TR_DQ = 1, TR_DKV = 2 };

struct AttnTrainParams {
  __nv_bfloat16* out;    // FWD: O [L*B, D];  DQ / DKV: dqkv [L*B, 3D]
  float* lse;            // [B*H, L]  log2-domain log-sum-exp of the penalised, scaled scores
  const float* delta;    // [L*B, H]  rowsum(dO o O)
  const int* lengths;    // [B]
  int L, B, H;
  float scale;           // head_dim^-0.5 (local_attention.py:98)
  float scale_log2e;     // scale * log2(e)
  int log_penalty;
  float drop_scale;        // 1 / (1 - p); 1 when off
  uint32_t drop_threshold; // keep iff hash >= threshold; 0 when off
  uint32_t drop_key;       // mix of (seed, site)
};

__device__ __forceinline__ float tr_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// keep-scale of attention probability (q, k) of head-batch bh: stateless per-element hash (lowbias32), the
// same value from either side (the DQ kernel walks keys per query row, the DKV kernel queries per key row)
__device__ __forceinline__ float tr_keep(const AttnTrainParams& p, uint32_t bh, uint32_t q, uint32_t k) {
  if (p.drop_threshold == 0u) return 1.0f;
  uint32_t x = (q * 0x9E3779B1u) ^ (k * 0x85EBCA77u) ^ (bh * 0xC2B2AE3Du) ^ p.drop_key;
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x >= p.drop_threshold ? p.drop_scale : 0.0f;
}

// one row of a [128 x 64] bf16 K-major SW128 tile: 8 chunks of 16 B, chunk c stored at (c ^ (row & 7))
__device__ __forceinline__ void tr_store_chunk(uint8_t* tile, int row, int chunk, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                 pack_bf16x2(v[6], v[7]));
}

template <int MODE>
__global__ void __launch_bounds__(256, 2)
    attn_train_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                      const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                      const AttnTrainParams p) {
  const int L = p.L, B = p.B, H = p.H, D = H * TR_HD;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int r0 = blockIdx.x * TR_OWN;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rl = tid & 127;   // own row inside the block == TMEM lane
  const int ch = tid >> 7;    // which 32 of a tile's 64 columns this thread handles
  const int len = min(__ldg(p.lengths + b), L);
  const int row = r0 + rl;  // this thread's own row (query for FWD / DQ, key for DKV)

  // ---- blocks with nothing to compute: defined zeros, no pipeline (CTA-uniform branch)
  if (r0 >= len) {
    if (row < L && ch == 0) {
      if (MODE == TR_FWD) {
        uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)row * B + b) * D + h * TR_HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) op[c] = make_uint4(0, 0, 0, 0);
        p.lse[(size_t)bh * L + row] = 0.0f;
      } else if (MODE == TR_DQ) {
        uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)row * B + b) * 3 * D + h * TR_HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) op[c] = make_uint4(0, 0, 0, 0);
      } else {
        uint4* ok = reinterpret_cast<uint4*>(p.out + ((size_t)row * B + b) * 3 * D + D + h * TR_HD);
        uint4* ov = reinterpret_cast<uint4*>(p.out + ((size_t)row * B + b) * 3 * D + 2 * D + h * TR_HD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          ok[c] = make_uint4(0, 0, 0, 0);
          ov[c] = make_uint4(0, 0, 0, 0);
        }
      }
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOwnA = smem;                          // FWD/DQ: Q;  DKV: K
  uint8_t* sOwnB = sOwnA + TR_OWN_BYTES;          // DQ: dO;     DKV: V      (unused by FWD)
  uint8_t* sOth = sOwnB + TR_OWN_BYTES;           // [stage][which] x 8 KB: FWD/DQ: (K_j, V_j); DKV: (Q_i, dO_i)
  uint8_t* sT0 = sOth + 4 * TR_OTH_BYTES;         // written tile 0: FWD P, DQ dS, DKV (P o keep)^T
  uint8_t* sT1 = sT0 + TR_OWN_BYTES;              // written tile 1: DKV dS^T
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT1 + TR_OWN_BYTES);
  uint64_t* bar_own = bars + 0;
  uint64_t* bar_full = bars + 1;  // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_acc = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sLse = reinterpret_cast<float*>(bars + 8);   // [2][64]   (DKV)
  float* sDelta = sLse + 128;                         // [2][64]   (DKV)
  float* sXch = sDelta + 128;                         // [2][2][128]  (FWD: (m | l) of the two column halves)
  float* sLut = sXch + 512;                           // [2 * lut_off]
  const int lut_off = ((L + 127) / 128) * 128;        // entry o <-> (k - q) = o - lut_off

  if (tid == 0) {
    tma_prefetch_desc(&tmQKV128);
    tma_prefetch_desc(&tmQKV64);
    if (MODE != TR_FWD) {
      tma_prefetch_desc(&tmDO128);
      tma_prefetch_desc(&tmDO64);
    }
    mbar_init(bar_own, 1);
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // penalty LUT in log2 domain, negated: sLut[o] = -log2(max(1, |o - lut_off|))  (conv_transformer_layer.py:26-27)
  for (int o = tid; o < 2 * lut_off; o += 256) {
    const int d = abs(o - lut_off);
    sLut[o] = (p.log_penalty && d > 1) ? -__log2f((float)d) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t tS = tmem_base + lane_addr + ch * 32;        // columns   0.. 63: S (or S^T); this thread's half
  const uint32_t tP = tmem_base + lane_addr + 64 + ch * 32;   // columns  64..127: dP (or dP^T)
  const uint32_t tA0 = tmem_base + lane_addr + 128 + ch * 32; // columns 128..191: accumulator 0 (O / dQ / dV)
  const uint32_t tA1 = tmem_base + lane_addr + 192 + ch * 32; // columns 192..255: accumulator 1 (dK)

  constexpr uint32_t IDESC_S = idesc_bf16_f32(TR_OWN, TR_OTH, 0, 0);
  constexpr uint32_t IDESC_ACC = idesc_bf16_f32(TR_OWN, TR_HD, 0, 1);
  const int cq = h * TR_HD, ck = D + h * TR_HD, cv = 2 * D + h * TR_HD;  // column offsets inside qkv rows

  // ---- own tiles
  if (tid == 0) {
    if (MODE == TR_FWD) {
      mbar_arrive_expect_tx(bar_own, TR_OWN_BYTES);
      tma_load_3d(sOwnA, &tmQKV128, bar_own, cq, b, r0);
    } else if (MODE == TR_DQ) {
      mbar_arrive_expect_tx(bar_own, 2 * TR_OWN_BYTES);
      tma_load_3d(sOwnA, &tmQKV128, bar_own, cq, b, r0);
      tma_load_3d(sOwnB, &tmDO128, bar_own, cq, b, r0);
    } else {
      mbar_arrive_expect_tx(bar_own, 2 * TR_OWN_BYTES);
      tma_load_3d(sOwnA, &tmQKV128, bar_own, ck, b, r0);
      tma_load_3d(sOwnB, &tmQKV128, bar_own, cv, b, r0);
    }
  }
  // number of other-side tiles: keys < len (FWD / DQ), queries < len (DKV: padded queries carry no gradient)
  const int n_oth = (len + TR_OTH - 1) / TR_OTH;

  // issue the TMA of other-side tile `j` into stage `st` (one thread)
#define TR_LOAD_OTHER(j, st, with_second)                                                              \
  do {                                                                                                 \
    uint8_t* dst_ = sOth + (st) * 2 * TR_OTH_BYTES;                                                    \
    mbar_arrive_expect_tx(&bar_full[(st)], (with_second) ? 2 * TR_OTH_BYTES : TR_OTH_BYTES);           \
    if (MODE == TR_DKV) {                                                                              \
      tma_load_3d(dst_, &tmQKV64, &bar_full[(st)], cq, b, (j) * TR_OTH);                               \
      tma_load_3d(dst_ + TR_OTH_BYTES, &tmDO64, &bar_full[(st)], cq, b, (j) * TR_OTH);                 \
    } else {                                                                                           \
      tma_load_3d(dst_, &tmQKV64, &bar_full[(st)], ck, b, (j) * TR_OTH);                               \
      if (with_second) tma_load_3d(dst_ + TR_OTH_BYTES, &tmQKV64, &bar_full[(st)], cv, b, (j) * TR_OTH); \
    }                                                                                                  \
  } while (0)

  uint32_t n_full[2] = {0, 0};  // loads completed-and-consumed per stage (parity bookkeeping, CTA-uniform)
  uint32_t n_s = 0, n_acc = 0;  // S-type / accumulate commits consumed so far
  uint32_t n_ld = 0;            // other-side tiles issued so far (stage = n_ld & 1)

  // per-row constants
  const bool own_valid = row < len;  // FWD/DQ: a real query;  DKV: a real (unmasked) key
  float lse2 = 0.0f, dlt = 0.0f;

  mbar_wait(bar_own, 0);

  if (MODE == TR_FWD) {
    // ================================================================ pass A: exact LSE of every own row
    float m_run = -INFINITY, l_run = 0.0f;
    if (tid == 0) TR_LOAD_OTHER(0, 0, false);
    n_ld = 1;
    for (int j = 0; j < n_oth; ++j) {
      const int st = j & 1;
      if (tid == 0 && j + 1 < n_oth) TR_LOAD_OTHER(j + 1, (j + 1) & 1, false);  // stage free: S(j-1) was waited for
      if (j + 1 < n_oth) ++n_ld;
      mbar_wait(&bar_full[st], n_full[st] & 1);
      ++n_full[st];
      tc_fence_after();
      if (tid == 0) {
        const uint64_t adesc = desc_kmajor_sw128(smem_u32(sOwnA));
        const uint64_t bdesc = desc_kmajor_sw128(smem_u32(sOth + st * 2 * TR_OTH_BYTES));
#pragma unroll
        for (int k = 0; k < TR_HD / 16; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC_S, k != 0);
        umma_commit(bar_s);
      }
      mbar_wait(bar_s, n_s & 1);
      ++n_s;
      tc_fence_after();
      float s2[32];
      {
        uint32_t v[32];
        tmem_ld32(tS, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int k = j * TR_OTH + ch * 32 + c;
          const float x = fmaf(__uint_as_float(v[c]), p.scale_log2e, sLut[k - row + lut_off]);
          s2[c] = (k < len) ? x : -INFINITY;
        }
      }
      float tm = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) tm = fmaxf(tm, s2[c]);
      const float m_new = fmaxf(m_run, tm);
      if (m_new != -INFINITY) {  // this half of the tile holds at least one valid key (or an earlier one did)
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc += tr_ex2(s2[c] - m_new);
        l_run = l_run * tr_ex2(m_run - m_new) + acc;
        m_run = m_new;
      }
      tc_fence_before();
      __syncthreads();  // every thread has read S(j) before the next S-type MMA overwrites it
    }
    // combine the two column halves of the row (the other thread of the row is tid ^ 128)
    sXch[(ch * 2 + 0) * 128 + rl] = m_run;
    sXch[(ch * 2 + 1) * 128 + rl] = l_run;
    __syncthreads();
    {
      const float m_o = sXch[((ch ^ 1) * 2 + 0) * 128 + rl], l_o = sXch[((ch ^ 1) * 2 + 1) * 128 + rl];
      const float mm = fmaxf(m_run, m_o);  // finite: key 0 is always valid
      const float a_me = (m_run == -INFINITY) ? 0.0f : l_run * tr_ex2(m_run - mm);
      const float a_ot = (m_o == -INFINITY) ? 0.0f : l_o * tr_ex2(m_o - mm);
      lse2 = mm + __log2f(a_me + a_ot);
    }
    if (row < L && ch == 0) p.lse[(size_t)bh * L + row] = lse2;
  } else if (MODE == TR_DQ) {
    if (row < L) {
      lse2 = p.lse[(size_t)bh * L + row];
      dlt = p.delta[((size_t)row * B + b) * H + h];
    }
  }

  // ==================================================================== main pass
  if (tid == 0) TR_LOAD_OTHER(0, n_ld & 1, true);
  const uint32_t ld_base = n_ld;  // tile j of the main pass lives in stage (ld_base + j) & 1
  for (int j = 0; j < n_oth; ++j) {
    const int st = (ld_base + j) & 1;
    uint8_t* oth = sOth + st * 2 * TR_OTH_BYTES;
    if (MODE == TR_DKV) {  // per-query statistics of the tile (64 queries): lse = +inf -> p = 0 for padded queries
      if (tid < 64) {
        const int q = j * TR_OTH + tid;
        sLse[st * 64 + tid] = (q < len) ? p.lse[(size_t)bh * L + q] : INFINITY;
        sDelta[st * 64 + tid] = (q < len) ? p.delta[((size_t)q * B + b) * H + h] : 0.0f;
      }
    }
    mbar_wait(&bar_full[st], n_full[st] & 1);
    ++n_full[st];
    tc_fence_after();
    if (tid == 0) {
      const uint64_t bdesc = desc_kmajor_sw128(smem_u32(oth));
      const uint64_t adesc = desc_kmajor_sw128(smem_u32(sOwnA));
#pragma unroll
      for (int k = 0; k < TR_HD / 16; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC_S, k != 0);
      if (MODE != TR_FWD) {  // dP = dO V_j^T  (DQ)   /   dP^T = V dO_i^T  (DKV)
        const uint64_t a2 = desc_kmajor_sw128(smem_u32(sOwnB));
        const uint64_t b2 = desc_kmajor_sw128(smem_u32(oth + TR_OTH_BYTES));
#pragma unroll
        for (int k = 0; k < TR_HD / 16; ++k)
          umma_bf16_ss(tmem_base + 64, a2 + 2 * k, b2 + 2 * k, IDESC_S, k != 0);
      }
      umma_commit(bar_s);
    }
    __syncthreads();  // sLse / sDelta of this stage visible (DKV); harmless otherwise
    mbar_wait(bar_s, n_s & 1);
    ++n_s;
    tc_fence_after();

    // ---- TMEM -> registers -> probabilities / gradients (8 columns at a time)
    float t0[32], t1[32];
    {
      uint32_t v[32], w[32];
      tmem_ld32(tS, v);
      if (MODE != TR_FWD) tmem_ld32(tP, w);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int cc = ch * 32 + c;
        const int o = j * TR_OTH + cc;  // other-side index: key (FWD / DQ) or query (DKV)
        if (MODE == TR_DKV) {
          const float x = fmaf(__uint_as_float(v[c]), p.scale_log2e, sLut[row - o + lut_off]);
          const float pr = own_valid ? tr_ex2(x - sLse[st * 64 + cc]) : 0.0f;
          const float keep = tr_keep(p, bh, o, row);
          t0[c] = pr * keep;
          t1[c] = pr * (__uint_as_float(w[c]) * keep - sDelta[st * 64 + cc]) * p.scale;
        } else {
          const float x = fmaf(__uint_as_float(v[c]), p.scale_log2e, sLut[o - row + lut_off]);
          const float pr = (o < len) ? tr_ex2(x - lse2) : 0.0f;
          const float keep = tr_keep(p, bh, row, o);
          if (MODE == TR_FWD)
            t0[c] = pr * keep;
          else
            t0[c] = own_valid ? pr * (__uint_as_float(w[c]) * keep - dlt) * p.scale : 0.0f;
        }
      }
    }
    // the accumulate MMAs of the previous tile must have finished reading sT0 / sT1 and its stage
    if (j > 0) {
      mbar_wait(bar_acc, n_acc & 1);
      ++n_acc;
      tc_fence_after();
    }
    if (tid == 0 && j + 1 < n_oth) TR_LOAD_OTHER(j + 1, (ld_base + j + 1) & 1, true);
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // this thread's 4 of the row's 8 sixteen-byte chunks
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = t0[k * 8 + e];
      tr_store_chunk(sT0, rl, ch * 4 + k, a);
      if (MODE == TR_DKV) {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = t1[k * 8 + e];
        tr_store_chunk(sT1, rl, ch * 4 + k, a);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ta = smem_u32(sT0);
      // FWD: O += P V_j;  DQ: dQ += dS K_j;  DKV: dV += (P o keep)^T dO_i
      const uint32_t bsrc = smem_u32(oth + ((MODE == TR_DQ) ? 0 : TR_OTH_BYTES));
#pragma unroll
      for (int kk = 0; kk < TR_OTH / 16; ++kk)
        umma_bf16_ss(tmem_base + 128, desc_kmajor_sw128(ta) + 2 * kk,
                     desc_mnmajor_sw128(bsrc + kk * 2048, TR_OTH_BYTES), IDESC_ACC, (j > 0) || kk != 0);
      if (MODE == TR_DKV) {  // dK += dS^T Q_i
        const uint32_t tb = smem_u32(sT1), qsrc = smem_u32(oth);
#pragma unroll
        for (int kk = 0; kk < TR_OTH / 16; ++kk)
          umma_bf16_ss(tmem_base + 192, desc_kmajor_sw128(tb) + 2 * kk,
                       desc_mnmajor_sw128(qsrc + kk * 2048, TR_OTH_BYTES), IDESC_ACC, (j > 0) || kk != 0);
      }
      umma_commit(bar_acc);
    }
  }
  mbar_wait(bar_acc, n_acc & 1);
  tc_fence_after();

  // ---- epilogue: accumulators -> global (two threads per own row, 32 columns = 64 contiguous bytes each)
  {
    uint32_t v[32];
    const size_t base = (MODE == TR_FWD) ? ((size_t)row * B + b) * D + h * TR_HD
                                         : ((size_t)row * B + b) * 3 * D + h * TR_HD;
#pragma unroll
    for (int acc = 0; acc < (MODE == TR_DKV ? 2 : 1); ++acc) {
      // DKV: accumulator 0 = dV -> column block 2D, accumulator 1 = dK -> column block D
      const size_t off = (MODE == TR_DKV) ? (acc == 0 ? 2 * (size_t)D : (size_t)D) : 0;
      tmem_ld32(acc == 0 ? tA0 : tA1, v);
      tmem_ld_wait();
      if (row < L) {
        const bool zero = (MODE == TR_DQ) && !own_valid;  // padded query rows: zero gradient
        uint4* op = reinterpret_cast<uint4*>(p.out + base + off + ch * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int o = c * 8;
          op[c] = zero ? make_uint4(0, 0, 0, 0)
                       : make_uint4(pack_bf16x2(__uint_as_float(v[o]), __uint_as_float(v[o + 1])),
                                    pack_bf16x2(__uint_as_float(v[o + 2]), __uint_as_float(v[o + 3])),
                                    pack_bf16x2(__uint_as_float(v[o + 4]), __uint_as_float(v[o + 5])),
                                    pack_bf16x2(__uint_as_float(v[o + 6]), __uint_as_float(v[o + 7])));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static inline int attn_train_smem(int L) {
  const int lut_off = ((L + 127) / 128) * 128;
  return 2 * TR_OWN_BYTES + 4 * TR_OTH_BYTES + 2 * TR_OWN_BYTES + 64 /*barriers*/ + 4 * 128 * 2 /*stats*/ +
         4 * 512 /*(m, l) exchange*/ + 4 * 2 * lut_off + 1024 /*align*/;
}

static uint32_t mix_key(uint64_t seed, int site) {
  uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(site + 1));
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  return (uint32_t)x ^ (uint32_t)(x >> 32);
}

static int attn_train_launch(int mode, const void* qkv, const void* dO, void* out, float* lse, const float* delta,
                             const int32_t* lengths, int L, int B, int H, int log_penalty, float drop_p,
                             uint64_t seed, int site, cudaStream_t st) {
  const int D = H * TR_HD;
  CUtensorMap tm128, tm64, tmdo128, tmdo64;
  {
    uint64_t dims[3] = {(uint64_t)3 * D, (uint64_t)B, (uint64_t)L};
    uint64_t strides[2] = {(uint64_t)3 * D * 2, (uint64_t)B * 3 * D * 2};
    uint32_t box128[3] = {TR_HD, 1, TR_OWN}, box64[3] = {TR_HD, 1, TR_OTH};
    int rc = make_tensor_map(&tm128, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box128, nullptr);
    if (rc) return rc;
    rc = make_tensor_map(&tm64, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box64, nullptr);
    if (rc) return rc;
  }
  if (dO != nullptr) {
    uint64_t dims[3] = {(uint64_t)D, (uint64_t)B, (uint64_t)L};
    uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)B * D * 2};
    uint32_t box128[3] = {TR_HD, 1, TR_OWN}, box64[3] = {TR_HD, 1, TR_OTH};
    int rc = make_tensor_map(&tmdo128, dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box128, nullptr);
    if (rc) return rc;
    rc = make_tensor_map(&tmdo64, dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3, dims, strides, box64, nullptr);
    if (rc) return rc;
  } else {
    tmdo128 = tm128;
    tmdo64 = tm64;
  }
  AttnTrainParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.delta = delta;
  p.lengths = lengths;
  p.L = L;
  p.B = B;
  p.H = H;
  p.scale = 0.125f;  // head_dim 64
  p.scale_log2e = 0.125f * 1.4426950408889634f;
  p.log_penalty = log_penalty;
  p.drop_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const double t = (double)drop_p * 4294967296.0;
  p.drop_threshold = drop_p > 0.f ? (t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t) : 0u;
  if (drop_p > 0.f && p.drop_threshold == 0u) p.drop_threshold = 1u;
  p.drop_key = mix_key(seed, site);
  const int smem = attn_train_smem(L);
  FBKST_REQUIRE(smem <= 227 * 1024, "attention (training): L=%d needs %d B of shared memory", L, smem);
  static PerDeviceFlag configured;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attn_train_kernel<TR_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attn_train_kernel<TR_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(attn_train_kernel<TR_DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
    configured = true;
  }
  dim3 grid((L + TR_OWN - 1) / TR_OWN, B * H);
  if (mode == TR_FWD)
    attn_train_kernel<TR_FWD><<<grid, 256, smem, st>>>(tm128, tm64, tmdo128, tmdo64, p);
  else if (mode == TR_DQ)
    attn_train_kernel<TR_DQ><<<grid, 256, smem, st>>>(tm128, tm64, tmdo128, tmdo64, p);
  else
    attn_train_kernel<TR_DKV><<<grid, 256, smem, st>>>(tm128, tm64, tmdo128, tmdo64, p);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_attention_train_fwd(const void* qkv, void* out, float* lse, const int32_t* lengths, int L,
                                         int B, int H, int log_penalty, float dropout_p, uint64_t seed, int site,
                                         fbkst_stream_t stream) {
  FBKST_REQUIRE(qkv && out && lse && lengths, "fbkst_attention_train_fwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && H > 0 && B * H <= 65535, "fbkst_attention_train_fwd: bad shape L=%d B=%d H=%d", L,
                B, H);
  FBKST_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "fbkst_attention_train_fwd: p must be in [0, 1)");
  return attn_train_launch(TR_FWD, qkv, nullptr, out, lse, nullptr, lengths, L, B, H, log_penalty, dropout_p, seed,
                           site, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int fbkst_attention_train_bwd(const void* qkv, const void* dO, const float* lse, const float* delta,
                                         void* dqkv, const int32_t* lengths, int L, int B, int H,
                                         int log_penalty, float dropout_p, uint64_t seed, int site,
                                         fbkst_stream_t stream) {
  FBKST_REQUIRE(qkv && dO && lse && delta && dqkv && lengths, "fbkst_attention_train_bwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && H > 0 && B * H <= 65535, "fbkst_attention_train_bwd: bad shape L=%d B=%d H=%d", L,
                B, H);
  FBKST_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "fbkst_attention_train_bwd: p must be in [0, 1)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = attn_train_launch(TR_DQ, qkv, dO, dqkv, const_cast<float*>(lse), delta, lengths, L, B, H, log_penalty,
                             dropout_p, seed, site, st);
  if (rc) return rc;
  return attn_train_launch(TR_DKV, qkv, dO, dqkv, const_cast<float*>(lse), delta, lengths, L, B, H, log_penalty,
                           dropout_p, seed, site, st);
}
