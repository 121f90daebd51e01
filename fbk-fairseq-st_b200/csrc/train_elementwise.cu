// Training-side byte work of the encoder (row T of the scope table): everything between the tensor-core
// contractions of the backward pass.  All kernels are HBM-bound streaming / tiled-transpose kernels.
//
//   dropout_add_ln     x1 = res + dropout(y);  LN(x1) -> bf16            (transformer_layer.py:120-133)
//   ln_bwd             dx (+)= LayerNorm backward(dy, x, gamma), per-block partial dgamma / dbeta
//   grad_prep          g -> [ReLU mask] -> [dropout mask] -> bf16 copy, TRANSPOSED bf16 copy (the wgrad
//                      GEMM's operand), per-row-tile column sums (the bias gradient), optional row remap
//   reduce_sum         sum of G partial vectors / of the split-K slices of a wgrad GEMM
//   attn_delta         delta[m, h] = <dO[m, h, :], O[m, h, :]>           (flash-attention backward)
//   ctc_compress_bwd   dx[t] = W[t, seg(t)] * dout[seg(t)]               (conv_transformer.py:290)
//   dropout_inplace    activation / embedding dropout
//
// Dropout masks come from a stateless counter-based generator (Philox4x32-7) keyed by (seed, site) and
// indexed by the element's position, so the backward pass regenerates the forward mask instead of
// storing it.  One Philox call yields the four 32-bit words of FOUR CONSECUTIVE COLUMNS
// (counter = row * (N/4) + col/4): every kernel that touches a mask processes column quads.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

struct DropoutParams {
  float p;             // drop probability; 0 = off
  float scale;         // 1 / (1 - p)
  uint32_t threshold;  // keep iff word >= threshold  (threshold = p * 2^32)
  uint32_t seed_lo, seed_hi;
  uint32_t site;       // call-site id: independent streams per dropout site
};

__device__ __forceinline__ uint4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-scale of the four columns of quad `quad` (a 64-bit index)
__device__ __forceinline__ float4 dropout_scale4(const DropoutParams& d, unsigned long long quad) {
  const uint4 r = philox4x32_7((uint32_t)quad, (uint32_t)(quad >> 32), d.site, 0x5eedu, d.seed_lo, d.seed_hi);
  return make_float4(r.x >= d.threshold ? d.scale : 0.f, r.y >= d.threshold ? d.scale : 0.f,
                     r.z >= d.threshold ? d.scale : 0.f, r.w >= d.threshold ? d.scale : 0.f);
}

// ------------------------------------------------------------------ x1 = res + dropout(y); LN(x1)
template <int NV>
__global__ void __launch_bounds__(256)
    dropout_add_ln_kernel(const float* __restrict__ y, const float* __restrict__ res, float* __restrict__ x1,
                          __nv_bfloat16* __restrict__ ln_out, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, int M, DropoutParams dp) {
  constexpr int D = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float4* yp = reinterpret_cast<const float4*>(y + (size_t)row * D);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = yp[i * 32 + lane];
  if (dp.p > 0.f) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 k = dropout_scale4(dp, (unsigned long long)row * (D / 4) + i * 32 + lane);
      v[i].x *= k.x; v[i].y *= k.y; v[i].z *= k.z; v[i].w *= k.w;
    }
  }
  if (res != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(res + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 r = rp[i * 32 + lane];
      v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
    }
  }
  if (x1 != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(x1 + (size_t)row * D)[i * 32 + lane] = v[i];
  }
  if (ln_out == nullptr) return;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* gp = reinterpret_cast<const float4*>(gamma);
  const float4* bp = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(gp + i * 32 + lane), b = __ldg(bp + i * 32 + lane);
    reinterpret_cast<uint2*>(ln_out + (size_t)row * D)[i * 32 + lane] =
        make_uint2(pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y),
                   pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// xhat = (x - mean) rstd;  g = dy gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat));
// dgamma += dy xhat;  dbeta += dy.   Warp per row, grid-stride over rows; each block leaves ONE partial
// (dgamma | dbeta) vector, summed in a fixed order (warp 0..7) -> run-to-run identical.
template <int NV>
__global__ void __launch_bounds__(256)
    ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                  float* __restrict__ dx, int accumulate, float* __restrict__ partial, float eps, int M) {
  constexpr int D = NV * 128;
  __shared__ float sred[2 * D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float4 dg[NV], db[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
  }
  for (int row = blockIdx.x * nw + warp; row < M; row += gridDim.x * nw) {
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)row * D);
    const float4* dp = reinterpret_cast<const float4*>(dy + (size_t)row * D);
    float4 v[NV], d[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xp[i * 32 + lane];
      d[i] = dp[i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
      dg[i].x = fmaf(d[i].x, v[i].x, dg[i].x); dg[i].y = fmaf(d[i].y, v[i].y, dg[i].y);
      dg[i].z = fmaf(d[i].z, v[i].z, dg[i].z); dg[i].w = fmaf(d[i].w, v[i].w, dg[i].w);
      db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;  // g = dy gamma
      c1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      c2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    c1 = warp_sum(c1) * (1.0f / D);
    c2 = warp_sum(c2) * (1.0f / D);
    float4* op = reinterpret_cast<float4*>(dx + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rstd * (d[i].x - c1 - v[i].x * c2);
      o.y = rstd * (d[i].y - c1 - v[i].y * c2);
      o.z = rstd * (d[i].z - c1 - v[i].z * c2);
      o.w = rstd * (d[i].w - c1 - v[i].w * c2);
      if (accumulate) {
        const float4 a = op[i * 32 + lane];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      op[i * 32 + lane] = o;
    }
  }
  for (int w = 0; w < nw; ++w) {  // fixed order: deterministic
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4* a = reinterpret_cast<float4*>(sred) + i * 32 + lane;
        float4* b = reinterpret_cast<float4*>(sred + D) + i * 32 + lane;
        if (w == 0) {
          *a = dg[i];
          *b = db[i];
        } else {
          float4 t = *a;
          t.x += dg[i].x; t.y += dg[i].y; t.z += dg[i].z; t.w += dg[i].w;
          *a = t;
          t = *b;
          t.x += db[i].x; t.y += db[i].y; t.z += db[i].z; t.w += db[i].w;
          *b = t;
        }
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * D + i] = sred[i];
}

// ------------------------------------------------------------------ gradient preparation (tiled transpose)
// One 64 x 64 tile per block (256 threads: 16 column quads x 16 rows, 4 row passes).
//   v = g[in_row(m), n]                       (fp32 or bf16; in_row applies the optional (inner, outer) remap)
//   v = (act[m, n] > 0) ? v : 0               if act != nullptr   (ReLU backward; act is the saved output)
//   v *= act_scale                            (activation dropout: the saved output is already masked)
//   v *= dropout keep-scale(m, n)             if dp.p > 0         (regenerated residual-dropout mask)
//   gb[m, n]  = bf16(v)  for n < n_pad (0 beyond N)      optional
//   gT[n, m]  = bf16(v)                                  optional   (pitch ldt)
//   colsum[blockIdx.y, n] = sum over the tile's rows     optional   (fp32, reduced later: bias gradient)
struct GradPrepParams {
  const void* g;
  long long ldg;
  const __nv_bfloat16* act;
  long long lda;
  float act_scale;
  int remap_inner, remap_outer;  // in_row(m) = (m % inner) * outer + m / inner   (0: identity)
  __nv_bfloat16* gb;
  long long ldb;
  int n_pad;
  __nv_bfloat16* gT;
  long long ldt;
  float* colsum;  // [row tiles, n_pad_cs]
  int ld_cs;
  int M, N;
  int vec_ok;  // rows of g (and act) are 16-/8-byte aligned: vector loads
  DropoutParams dp;
  int dp_cols;  // columns per row of the tensor the dropout mask was drawn for (counter pitch = dp_cols / 4)
};

// one 64 x 64 tile (tile column bx, tile row by) of the grad_prep pass
template <int IN_T>  // 0: bf16, 1: fp32, 2: IEEE fp16 (the conv front end's activations)
__device__ __forceinline__ void grad_prep_tile(const GradPrepParams& p, const int bx, const int by,
                                               __nv_bfloat16 (*tileT)[72], float (*cs)[65]) {
  constexpr int IN_F32 = IN_T == 1;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = by * 64, n0 = bx * 64;
  const int n = n0 + tx * 4;
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = ty + pass * 16, m = m0 + r;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < p.M && n < p.N) {
      const long long mi = p.remap_inner ? (long long)(m % p.remap_inner) * p.remap_outer + m / p.remap_inner : m;
      if (IN_F32) {
        const float* gp = reinterpret_cast<const float*>(p.g) + mi * p.ldg + n;
        if (p.vec_ok && n + 3 < p.N) {
          const float4 t = *reinterpret_cast<const float4*>(gp);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) v[j] = gp[j];
        }
      } else if (IN_T == 2) {
        const __half* gp = reinterpret_cast<const __half*>(p.g) + mi * p.ldg + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) v[j] = __half2float(gp[j]);
      } else {
        const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(p.g) + mi * p.ldg + n;
        if (p.vec_ok && n + 3 < p.N) {
          const uint2 t = *reinterpret_cast<const uint2*>(gp);
          v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
          v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) v[j] = __bfloat162float(gp[j]);
        }
      }
      if (p.act != nullptr) {
        const __nv_bfloat16* ap = p.act + (long long)m * p.lda + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) v[j] = (__bfloat162float(ap[j]) > 0.f) ? v[j] * p.act_scale : 0.f;
      }
      if (p.dp.p > 0.f) {
        // the mask lives in the layout of the tensor g was drawn for: its row is the (remapped) input row
        const float4 k = dropout_scale4(p.dp, (unsigned long long)mi * (p.dp_cols / 4) + (n >> 2));
        v[0] *= k.x; v[1] *= k.y; v[2] *= k.z; v[3] *= k.w;
      }
    }
    // round once: the bf16 copy, the transposed copy and the column sums all see the same values
    __nv_bfloat16 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(v[j]);
      csum[j] += __bfloat162float(h[j]);
      tileT[tx * 4 + j][r] = h[j];
    }
    if (p.gb != nullptr && m < p.M && n < p.n_pad) {
      __nv_bfloat16* bp = p.gb + (long long)m * p.ldb + n;
      if (n + 3 < p.n_pad && (p.ldb & 3) == 0) {
        *reinterpret_cast<uint2*>(bp) =
            make_uint2((uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
                       (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.n_pad) bp[j] = h[j];
      }
    }
  }
  if (p.colsum != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) cs[ty][tx * 4 + j] = csum[j];
  }
  __syncthreads();
  if (p.colsum != nullptr && threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += cs[k][threadIdx.x];
    if (n0 + (int)threadIdx.x < p.ld_cs) p.colsum[(size_t)by * p.ld_cs + n0 + threadIdx.x] = s;
  }
  if (p.gT != nullptr) {
    const int nr = threadIdx.x >> 2, seg = (threadIdx.x & 3) * 16;  // output row n0+nr, columns m0+seg..+15
    if (n0 + nr < p.N) {
      __nv_bfloat16* op = p.gT + (long long)(n0 + nr) * p.ldt + m0 + seg;
      const uint4* sp = reinterpret_cast<const uint4*>(&tileT[nr][seg]);
      if (m0 + seg + 15 < p.M && (p.ldt & 7) == 0) {
        reinterpret_cast<uint4*>(op)[0] = sp[0];
        reinterpret_cast<uint4*>(op)[1] = sp[1];
      } else {
        for (int j = 0; j < 16; ++j)
          if (m0 + seg + j < p.M) op[j] = tileT[nr][seg + j];
      }
    }
  }
}

template <int IN_T>
__global__ void __launch_bounds__(256) grad_prep_kernel(const GradPrepParams p) {
  __shared__ __align__(16) __nv_bfloat16 tileT[64][72];  // [n][m], 144-byte rows
  __shared__ float cs[16][65];
  grad_prep_tile<IN_T>(p, blockIdx.x, blockIdx.y, tileT, cs);
}

// ---- many (cast, transpose) jobs in ONE launch.  A training step derives ~70 bf16 operand copies (+ their
// transposes) from the fp32 master weights, and ~45 token-contiguous activation copies for the weight-gradient
// GEMMs; one launch per matrix made those stretches of the step bound by the HOST's launch rate.  The job
// table travels in the kernel parameters (no staging copy); a block finds its job by a scan of the tile-offset
// prefix.
constexpr int PREP_BATCH_MAX = 48;
struct PrepBatch {
  fbkst_prep_desc_t d[PREP_BATCH_MAX];
  int tile0[PREP_BATCH_MAX + 1];
  int n;
};
__global__ void __launch_bounds__(256) prep_batch_kernel(const __grid_constant__ PrepBatch tb) {
  __shared__ __align__(16) __nv_bfloat16 tileT[64][72];
  __shared__ float cs[16][65];
  int di = 0;
  while (di + 1 < tb.n && (int)blockIdx.x >= tb.tile0[di + 1]) ++di;
  const fbkst_prep_desc_t& d = tb.d[di];
  const int t = blockIdx.x - tb.tile0[di];
  const int tiles_x = (d.cols + 63) / 64;
  GradPrepParams q;
  q.g = d.src;
  q.ldg = d.ld_src;
  q.act = nullptr;
  q.lda = 0;
  q.act_scale = 1.f;
  q.remap_inner = q.remap_outer = 0;
  q.gb = reinterpret_cast<__nv_bfloat16*>(d.copy);
  q.ldb = d.ld_copy;
  q.n_pad = d.cols;
  q.gT = reinterpret_cast<__nv_bfloat16*>(d.transposed);
  q.ldt = d.ld_transposed;
  q.colsum = nullptr;
  q.ld_cs = 0;
  q.M = d.rows;
  q.N = d.cols;
  q.vec_ok = d.reserved;  // (set by the host: source rows aligned for vector loads)
  q.dp.p = 0.f;
  q.dp.scale = 1.f;
  q.dp.threshold = q.dp.seed_lo = q.dp.seed_hi = q.dp.site = 0u;
  q.dp_cols = 4;
  const int bx = t % tiles_x, by = t / tiles_x;
  if (d.src_type == 1)
    grad_prep_tile<1>(q, bx, by, tileT, cs);
  else if (d.src_type == 2)
    grad_prep_tile<2>(q, bx, by, tileT, cs);
  else
    grad_prep_tile<0>(q, bx, by, tileT, cs);
}

// ---- many fixed-order reductions in ONE launch (bias gradients, split-K slices of the weight-gradient GEMMs,
// LayerNorm parameter gradients: ~120 per training step, each a few microseconds of a nearly empty GPU)
constexpr int REDUCE_BATCH_MAX = 56;
struct ReduceBatch {
  fbkst_reduce_desc_t d[REDUCE_BATCH_MAX];
  int blk0[REDUCE_BATCH_MAX + 1];
  int n;
};
__global__ void __launch_bounds__(256) reduce_sum_batch_kernel(const __grid_constant__ ReduceBatch tb) {
  __shared__ float red[32][9];
  int di = 0;
  while (di + 1 < tb.n && (int)blockIdx.x >= tb.blk0[di + 1]) ++di;
  const fbkst_reduce_desc_t& d = tb.d[di];
  const int blk = blockIdx.x - tb.blk0[di], nblk = tb.blk0[di + 1] - tb.blk0[di];
  const long long total = (long long)d.rows * d.cols;
  if (d.G <= 16) {  // few slices, many outputs: thread per output element
    for (long long i = (long long)blk * 256 + threadIdx.x; i < total; i += (long long)nblk * 256) {
      const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
      const float* ip = d.in + (long long)r * d.ldi + c;
      float s = 0.f;
      for (int g = 0; g < d.G; ++g) s += ip[(long long)g * d.g_stride];
      d.out[(long long)r * d.ldo + c] = s * d.scale;
    }
    return;
  }
  const int e = threadIdx.x & 7, gl = threadIdx.x >> 3;  // many slices: 32 lanes split the slices of 8 outputs
  for (long long i0 = (long long)blk * 8; i0 < total; i0 += (long long)nblk * 8) {
    const long long i = i0 + e;
    float s = 0.f;
    if (i < total) {
      const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
      const float* ip = d.in + (long long)r * d.ldi + c;
      float s0 = 0.f, s1 = 0.f;
      int g = gl;
      for (; g + 32 < d.G; g += 64) {
        s0 += ip[(long long)g * d.g_stride];
        s1 += ip[(long long)(g + 32) * d.g_stride];
      }
      if (g < d.G) s0 += ip[(long long)g * d.g_stride];
      s = s0 + s1;
    }
    red[gl][e] = s;
    __syncthreads();
    if (gl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) t += red[l][e];
      const int r = (int)(i / d.cols), c = (int)(i - (long long)r * d.cols);
      d.out[(long long)r * d.ldo + c] = t * d.scale;
    }
    __syncthreads();
  }
}

// out[r, c] = scale * sum_{g < G} in[g * g_stride + r * ldi + c], fixed summation order.
// Block = 8 consecutive output elements x 32 g-lanes: lane l sums g = l, l + 32, ... (independent loads in
// flight), then the 32 lane sums are added in lane order.  (The first version walked all G partials in one
// thread: 375 dependent loads for a bias gradient -- 36 us per call, 4.3 ms of a training step.)
__global__ void __launch_bounds__(256)
    reduce_sum_kernel(const float* __restrict__ in, int G, long long g_stride, int rows, int cols,
                      long long ldi, float* __restrict__ out, long long ldo, float scale) {
  __shared__ float red[32][9];
  const int e = threadIdx.x & 7, gl = threadIdx.x >> 3;
  const long long total = (long long)rows * cols;
  for (long long i0 = (long long)blockIdx.x * 8; i0 < total; i0 += (long long)gridDim.x * 8) {
    const long long i = i0 + e;
    float s = 0.f;
    if (i < total) {
      const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
      const float* ip = in + (long long)r * ldi + c;
      float s0 = 0.f, s1 = 0.f;
      int g = gl;
      for (; g + 32 < G; g += 64) {
        s0 += ip[(long long)g * g_stride];
        s1 += ip[(long long)(g + 32) * g_stride];
      }
      if (g < G) s0 += ip[(long long)g * g_stride];
      s = s0 + s1;
    }
    red[gl][e] = s;
    __syncthreads();
    if (gl == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int l = 0; l < 32; ++l) t += red[l][e];
      const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
      out[(long long)r * ldo + c] = t * scale;
    }
    __syncthreads();
  }
}

// Few partials, many outputs (the split-K slices of a wgrad GEMM): one thread per output element.
__global__ void __launch_bounds__(256)
    reduce_sum_small_kernel(const float* __restrict__ in, int G, long long g_stride, int rows, int cols,
                            long long ldi, float* __restrict__ out, long long ldo, float scale) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float* ip = in + (long long)r * ldi + c;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += ip[(long long)g * g_stride];
    out[(long long)r * ldo + c] = s * scale;
  }
}

// delta[m, h] = sum_d dO[m, h*64 + d] * O[m, h*64 + d]   (warp per row m; bf16 inputs)
__global__ void __launch_bounds__(256)
    attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O,
                      float* __restrict__ delta, int M, int H) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(dO + (size_t)row * H * 64);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(O + (size_t)row * H * 64);
  for (int h = 0; h < H; ++h) {
    const uint32_t u = a[h * 32 + lane], w = b[h * 32 + lane];
    float s = __uint_as_float(u << 16) * __uint_as_float(w << 16) +
              __uint_as_float(u & 0xffff0000u) * __uint_as_float(w & 0xffff0000u);
    s = warp_sum(s);
    if (lane == 0) delta[(size_t)row * H + h] = s;
  }
}

// dx[t*B+b, :] = weight[t, b] * dout[seg_id[t, b]*B + b, :]  (0 for padded frames)
__global__ void __launch_bounds__(256)
    ctc_compress_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ seg_id,
                            const float* __restrict__ weight, float* __restrict__ dx, int rows, int B, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = row % B;
  const int sid = __ldg(seg_id + row);
  const float w = sid >= 0 ? __ldg(weight + row) : 0.f;
  float4* op = reinterpret_cast<float4*>(dx + (size_t)row * D);
  const float4* ip = reinterpret_cast<const float4*>(dout + ((size_t)(sid >= 0 ? sid : 0) * B + b) * D);
  for (int i = lane; i < D / 4; i += 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sid >= 0) {
      v = __ldg(ip + i);
      v.x *= w; v.y *= w; v.z *= w; v.w *= w;
    }
    op[i] = v;
  }
}

// in-place dropout of a [M, N] tensor (N % 4 == 0), bf16 or fp32
template <int IS_F32>
__global__ void __launch_bounds__(256) dropout_inplace_kernel(void* x, long long quads, DropoutParams dp) {
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    const float4 k = dropout_scale4(dp, (unsigned long long)q);
    if (IS_F32) {
      float4* p = reinterpret_cast<float4*>(x) + q;
      float4 v = *p;
      v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
      *p = v;
    } else {
      uint2* p = reinterpret_cast<uint2*>(x) + q;
      const uint2 t = *p;
      *p = make_uint2(pack_bf16x2(__uint_as_float(t.x << 16) * k.x, __uint_as_float(t.x & 0xffff0000u) * k.y),
                      pack_bf16x2(__uint_as_float(t.y << 16) * k.z, __uint_as_float(t.y & 0xffff0000u) * k.w));
    }
  }
}

static DropoutParams make_dropout(float p, uint64_t seed, int site) {
  DropoutParams d;
  d.p = p;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const double t = (double)p * 4294967296.0;
  d.threshold = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.site = (uint32_t)site;
  return d;
}

}  // namespace fbkst

using namespace fbkst;

#define FBKST_NV_DISPATCH(D, CALL)                                   \
  switch ((D) / 128) {                                               \
    case 1: { constexpr int NV = 1; CALL; } break;                   \
    case 2: { constexpr int NV = 2; CALL; } break;                   \
    case 3: { constexpr int NV = 3; CALL; } break;                   \
    case 4: { constexpr int NV = 4; CALL; } break;                   \
    case 6: { constexpr int NV = 6; CALL; } break;                   \
    case 8: { constexpr int NV = 8; CALL; } break;                   \
    default: return set_error(FBKST_ERR_ARG, "width %d not supported (128, 256, 384, 512, 768, 1024)", (D)); \
  }

extern "C" int fbkst_dropout_add_ln(const float* y, const float* res, float* x1, void* ln_out_bf16,
                                    const float* gamma, const float* beta, float eps, int M, int D, float p,
                                    uint64_t seed, int site, fbkst_stream_t stream) {
  FBKST_REQUIRE(y && M > 0 && D > 0 && D % 128 == 0, "fbkst_dropout_add_ln: bad arguments");
  FBKST_REQUIRE(ln_out_bf16 == nullptr || (gamma && beta), "fbkst_dropout_add_ln: LayerNorm needs gamma and beta");
  FBKST_REQUIRE(p >= 0.f && p < 1.f, "fbkst_dropout_add_ln: p must be in [0, 1)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DropoutParams dp = make_dropout(p, seed, site);
  FBKST_NV_DISPATCH(D, (dropout_add_ln_kernel<NV><<<(M + 7) / 8, 256, 0, st>>>(
                           y, res, x1, reinterpret_cast<__nv_bfloat16*>(ln_out_bf16), gamma, beta, eps, M, dp)));
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ln_bwd_blocks(int M) {
  const int cap = num_sms() * 4;
  const int need = (M + 7) / 8;
  return need < cap ? need : cap;
}

extern "C" int fbkst_ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, int accumulate,
                            float* partial, float eps, int M, int D, fbkst_stream_t stream) {
  FBKST_REQUIRE(dy && x && gamma && dx && partial && M > 0 && D > 0 && D % 128 == 0,
                "fbkst_ln_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = fbkst_ln_bwd_blocks(M);
  FBKST_NV_DISPATCH(D, (ln_bwd_kernel<NV><<<grid, 256, 0, st>>>(dy, x, gamma, dx, accumulate, partial, eps, M)));
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_grad_prep(const void* g, int g_is_f32, int64_t ldg, const void* act_bf16, int64_t lda,
                               float act_scale, int remap_inner, int remap_outer, void* gb, int64_t ldb,
                               int n_pad, void* gT, int64_t ldt, float* colsum, int ld_cs, int M, int N,
                               float p, uint64_t seed, int site, int dp_cols, fbkst_stream_t stream) {
  FBKST_REQUIRE(g && M > 0 && N > 0, "fbkst_grad_prep: bad arguments");
  FBKST_REQUIRE(gb == nullptr || (n_pad >= N && ldb >= n_pad), "fbkst_grad_prep: bad gb pitch / padding");
  FBKST_REQUIRE(gT == nullptr || ldt >= M, "fbkst_grad_prep: bad transposed pitch");
  FBKST_REQUIRE(p >= 0.f && p < 1.f, "fbkst_grad_prep: p must be in [0, 1)");
  FBKST_REQUIRE(p == 0.f || (dp_cols > 0 && dp_cols % 4 == 0), "fbkst_grad_prep: dropout needs dp_cols %% 4 == 0");
  FBKST_REQUIRE((remap_inner == 0) == (remap_outer == 0), "fbkst_grad_prep: remap dims come together");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GradPrepParams q;
  q.g = g;
  q.ldg = ldg;
  q.act = reinterpret_cast<const __nv_bfloat16*>(act_bf16);
  q.lda = lda;
  q.act_scale = act_scale;
  q.remap_inner = remap_inner;
  q.remap_outer = remap_outer;
  q.gb = reinterpret_cast<__nv_bfloat16*>(gb);
  q.ldb = ldb;
  q.n_pad = gb ? n_pad : N;
  q.gT = reinterpret_cast<__nv_bfloat16*>(gT);
  q.ldt = ldt;
  q.colsum = colsum;
  q.ld_cs = ld_cs;
  q.M = M;
  q.N = N;
  const size_t esz = g_is_f32 == 1 ? 4 : 2;
  q.vec_ok = ((reinterpret_cast<uintptr_t>(g) % (4 * esz)) == 0 && (ldg * esz) % (4 * esz) == 0) ? 1 : 0;
  q.dp = make_dropout(p, seed, site);
  q.dp_cols = dp_cols > 0 ? dp_cols : 4;
  const int ncols = q.n_pad > N ? q.n_pad : N;
  dim3 grid((ncols + 63) / 64, (M + 63) / 64);
  if (g_is_f32 == 1)
    grad_prep_kernel<1><<<grid, 256, 0, st>>>(q);
  else if (g_is_f32 == 2)
    grad_prep_kernel<2><<<grid, 256, 0, st>>>(q);
  else
    grad_prep_kernel<0><<<grid, 256, 0, st>>>(q);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_reduce_sum(const float* in, int G, int64_t g_stride, int rows, int cols, int64_t ldi,
                                float* out, int64_t ldo, float scale, fbkst_stream_t stream) {
  FBKST_REQUIRE(in && out && G > 0 && rows > 0 && cols > 0, "fbkst_reduce_sum: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)rows * cols;
  if (G <= 16) {
    long long grid = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (grid > cap) grid = cap;
    reduce_sum_small_kernel<<<(int)grid, 256, 0, st>>>(in, G, g_stride, rows, cols, ldi, out, ldo, scale);
  } else {
    long long grid = (total + 7) / 8;
    const long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    reduce_sum_kernel<<<(int)grid, 256, 0, st>>>(in, G, g_stride, rows, cols, ldi, out, ldo, scale);
  }
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_prep_batch(const fbkst_prep_desc_t* descs, int n, fbkst_stream_t stream) {
  using namespace fbkst;
  FBKST_REQUIRE(descs && n > 0, "fbkst_prep_batch: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i0 = 0; i0 < n; i0 += PREP_BATCH_MAX) {
    PrepBatch tb;
    tb.n = n - i0 < PREP_BATCH_MAX ? n - i0 : PREP_BATCH_MAX;
    long long tiles = 0;
    for (int i = 0; i < tb.n; ++i) {
      fbkst_prep_desc_t d = descs[i0 + i];
      FBKST_REQUIRE(d.src && (d.copy || d.transposed) && d.rows > 0 && d.cols > 0 && d.src_type >= 0 &&
                        d.src_type <= 2 && d.ld_src >= d.cols,
                    "fbkst_prep_batch: bad job %d", i0 + i);
      FBKST_REQUIRE(d.copy == nullptr || d.ld_copy >= d.cols, "fbkst_prep_batch: job %d: bad copy pitch", i0 + i);
      FBKST_REQUIRE(d.transposed == nullptr || d.ld_transposed >= d.rows,
                    "fbkst_prep_batch: job %d: bad transposed pitch", i0 + i);
      const size_t esz = d.src_type == 1 ? 4 : 2;
      d.reserved = ((reinterpret_cast<uintptr_t>(d.src) % (4 * esz)) == 0 && d.ld_src % 4 == 0) ? 1 : 0;
      tb.d[i] = d;
      tb.tile0[i] = (int)tiles;
      tiles += (long long)((d.cols + 63) / 64) * ((d.rows + 63) / 64);
      FBKST_REQUIRE(tiles < (1ll << 30), "fbkst_prep_batch: too many tiles");
    }
    tb.tile0[tb.n] = (int)tiles;
    prep_batch_kernel<<<(int)tiles, 256, 0, st>>>(tb);
    FBKST_CHECK_CUDA(cudaGetLastError());
  }
  return FBKST_OK;
}

extern "C" int fbkst_reduce_sum_batch(const fbkst_reduce_desc_t* descs, int n, fbkst_stream_t stream) {
  using namespace fbkst;
  FBKST_REQUIRE(descs && n > 0, "fbkst_reduce_sum_batch: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i0 = 0; i0 < n; i0 += REDUCE_BATCH_MAX) {
    ReduceBatch tb;
    tb.n = n - i0 < REDUCE_BATCH_MAX ? n - i0 : REDUCE_BATCH_MAX;
    int blocks = 0;
    for (int i = 0; i < tb.n; ++i) {
      const fbkst_reduce_desc_t& d = descs[i0 + i];
      FBKST_REQUIRE(d.in && d.out && d.G > 0 && d.rows > 0 && d.cols > 0, "fbkst_reduce_sum_batch: bad job %d",
                    i0 + i);
      tb.d[i] = d;
      tb.blk0[i] = blocks;
      const long long total = (long long)d.rows * d.cols;
      long long nb = d.G <= 16 ? (total + 255) / 256 : (total + 7) / 8;
      const long long cap = d.G <= 16 ? 2 * num_sms() : 4 * num_sms();
      if (nb > cap) nb = cap;
      blocks += (int)nb;
    }
    tb.blk0[tb.n] = blocks;
    reduce_sum_batch_kernel<<<blocks, 256, 0, st>>>(tb);
    FBKST_CHECK_CUDA(cudaGetLastError());
  }
  return FBKST_OK;
}

extern "C" int fbkst_attn_delta(const void* dO, const void* O, float* delta, int M, int H,
                                fbkst_stream_t stream) {
  FBKST_REQUIRE(dO && O && delta && M > 0 && H > 0, "fbkst_attn_delta: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  attn_delta_kernel<<<(M + 7) / 8, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dO),
                                                 reinterpret_cast<const __nv_bfloat16*>(O), delta, M, H);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ctc_compress_bwd(const float* dout, const int32_t* seg_id, const float* weight, float* dx,
                                      int L, int B, int D, fbkst_stream_t stream) {
  FBKST_REQUIRE(dout && seg_id && weight && dx && L > 0 && B > 0 && D > 0 && D % 4 == 0,
                "fbkst_ctc_compress_bwd: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = L * B;
  ctc_compress_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(dout, seg_id, weight, dx, rows, B, D);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_dropout_inplace(void* x, int is_f32, int64_t numel, float p, uint64_t seed, int site,
                                     fbkst_stream_t stream) {
  FBKST_REQUIRE(x && numel > 0 && numel % 4 == 0, "fbkst_dropout_inplace: numel must be a positive multiple of 4");
  FBKST_REQUIRE(p >= 0.f && p < 1.f, "fbkst_dropout_inplace: p must be in [0, 1)");
  if (p == 0.f) return FBKST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long quads = numel / 4;
  long long grid = (quads + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (grid > cap) grid = cap;
  const DropoutParams dp = make_dropout(p, seed, site);
  if (is_f32)
    dropout_inplace_kernel<1><<<(int)grid, 256, 0, st>>>(x, quads, dp);
  else
    dropout_inplace_kernel<0><<<(int)grid, 256, 0, st>>>(x, quads, dp);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
