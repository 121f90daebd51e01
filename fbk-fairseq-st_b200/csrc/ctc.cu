// CTC-argmax compression (reference: conv_transformer.py:278-291 + CTCCompressStrategy
// :385-426), entirely on device:
//   1. fbkst_ctc_argmax   warp per frame: vectorised streaming arg-max (+ softmax prob of it)
//   2. fbkst_ctc_segment  CTA per utterance: run-boundary flags -> block scan -> segment ids,
//                         starts, new lengths, per-frame pooling weight (avg|weighted|softmax)
//   3. fbkst_ctc_compress segmented weighted reduction of x into the ragged output
// All three are HBM-bound integer/float streaming kernels; no tensor cores.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

struct ArgMax {
  float v;
  int i;
  float s;  // sum of exp(x - v) over the elements seen
};

__device__ __forceinline__ void argmax_push(ArgMax& a, float x, int i, bool want_sum) {
  if (x > a.v) {
    if (want_sum) a.s = a.s * __expf(a.v - x) + 1.0f;
    a.v = x;
    a.i = i;
  } else if (want_sum) {
    a.s += __expf(x - a.v);
  }
}

__device__ __forceinline__ void argmax_merge(ArgMax& a, float v, int i, float s, bool want_sum) {
  const bool take = (v > a.v) || (v == a.v && i < a.i);
  if (want_sum) {
    const float m = fmaxf(a.v, v);
    const float sa = (a.v == -INFINITY) ? 0.0f : a.s * __expf(a.v - m);
    const float sb = (v == -INFINITY) ? 0.0f : s * __expf(v - m);
    a.s = sa + sb;
  }
  if (take) {
    a.v = v;
    a.i = i;
  }
}

constexpr float kLog2eCtc = 1.4426950408889634f;
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// One warp per (t, b) row.  BF16 = 1: 8 logits per 16-byte load; else fp32, 4 per load.
template <int IS_BF16>
__global__ void __launch_bounds__(256)
    ctc_argmax_kernel(const void* __restrict__ logits, long long ldv,
                      const int* __restrict__ lengths, int* __restrict__ labels,
                      float* __restrict__ top_prob, float* __restrict__ lse, int rows, int B,
                      int V) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const bool want_sum = top_prob != nullptr || lse != nullptr;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_per_grid) {
    const int t = row / B, b = row - t * B;
    if (t >= __ldg(lengths + b)) {
      if (lane == 0) {
        labels[row] = -1;
        if (top_prob) top_prob[row] = 0.0f;
        if (lse) lse[row] = 0.0f;
      }
      continue;
    }
    ArgMax a{-INFINITY, 0x7fffffff, 0.0f};
    // Rows need not be 16-byte aligned (odd V / column-narrowed views): scalar head up to the
    // first aligned element, vector body, scalar tail.  Each lane still visits increasing indices.
    if (IS_BF16) {
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(logits) + (size_t)row * ldv;
      const int head = min(V, (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(rp) & 15u)) & 15u) >> 1));
      if (lane < head) argmax_push(a, __bfloat162float(rp[lane]), lane, want_sum);
      const uint4* vp = reinterpret_cast<const uint4*>(rp + head);
      const int nvec = (V - head) >> 3;
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) {
        const uint4 u = __ldg(vp + i);
        const int c = head + (i << 3);
        argmax_push(a, bf16_lo(u.x), c + 0, want_sum);
        argmax_push(a, bf16_hi(u.x), c + 1, want_sum);
        argmax_push(a, bf16_lo(u.y), c + 2, want_sum);
        argmax_push(a, bf16_hi(u.y), c + 3, want_sum);
        argmax_push(a, bf16_lo(u.z), c + 4, want_sum);
        argmax_push(a, bf16_hi(u.z), c + 5, want_sum);
        argmax_push(a, bf16_lo(u.w), c + 6, want_sum);
        argmax_push(a, bf16_hi(u.w), c + 7, want_sum);
      }
      for (int c = head + (nvec << 3) + lane; c < V; c += 32)
        argmax_push(a, __bfloat162float(rp[c]), c, want_sum);
    } else {
      const float* rp = reinterpret_cast<const float*>(logits) + (size_t)row * ldv;
      const int head = min(V, (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(rp) & 15u)) & 15u) >> 2));
      if (lane < head) argmax_push(a, rp[lane], lane, want_sum);
      const float4* vp = reinterpret_cast<const float4*>(rp + head);
      const int nvec = (V - head) >> 2;
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) {
        const float4 u = __ldg(vp + i);
        const int c = head + (i << 2);
        argmax_push(a, u.x, c + 0, want_sum);
        argmax_push(a, u.y, c + 1, want_sum);
        argmax_push(a, u.z, c + 2, want_sum);
        argmax_push(a, u.w, c + 3, want_sum);
      }
      for (int c = head + (nvec << 2) + lane; c < V; c += 32) argmax_push(a, rp[c], c, want_sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v = __shfl_xor_sync(0xffffffffu, a.v, o);
      const int i = __shfl_xor_sync(0xffffffffu, a.i, o);
      const float s = __shfl_xor_sync(0xffffffffu, a.s, o);
      argmax_merge(a, v, i, s, want_sum);
    }
    if (lane == 0) {
      labels[row] = a.i;
      if (top_prob) top_prob[row] = 1.0f / a.s;
      if (lse) lse[row] = a.v + logf(a.s);
    }
  }
}

// Fast path (rows 16-byte aligned: base and pitch multiples of 16 B -- the layout fbkst_linear_bf16
// produces).  One warp per row; every lane issues UNROLL independent 16-byte loads before touching
// any of them (the generic kernel above interleaves load and use: one load in flight per lane).
// The streaming loop only tracks, per lane, the running maximum and the index of the 16-byte
// VECTOR that holds it (bf16: packed HMNMX2, ~1 instruction per element instead of 4); the element
// index is recovered once per row by re-reading that one vector.  The optional sum of exponentials
// is rescaled once per vector.  Ties resolve to the lowest index at every level (strict '>' in
// index order inside a lane, (value, index) merge across lanes).
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int IS_BF16, int WANT_SUM>
__global__ void __launch_bounds__(256)
    ctc_argmax_vec_kernel(const uint4* __restrict__ logits, long long row_vecs,
                          const int* __restrict__ lengths, int* __restrict__ labels,
                          float* __restrict__ top_prob, float* __restrict__ lse, int rows, int B,
                          int V) {
  constexpr int EPV = IS_BF16 ? 8 : 4;  // elements per 16-byte vector
  constexpr int UNROLL = 8;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int t = row / B, b = row - t * B;
  if (t >= __ldg(lengths + b)) {
    if (lane == 0) {
      labels[row] = -1;
      if (WANT_SUM && top_prob) top_prob[row] = 0.0f;
      if (WANT_SUM && lse) lse[row] = 0.0f;
    }
    return;
  }
  const uint4* vp = logits + (size_t)row * row_vecs;
  const int nvec = (V + EPV - 1) / EPV;  // the last vector may hold columns >= V (masked below)
  const int tail = V - (nvec - 1) * EPV;  // valid elements of the last vector (1..EPV)
  float bv = -INFINITY, bs = 0.0f;  // running max / sum of exp(x - bv) of this lane
  int bvec = -1;                    // vector that holds the running max
  // columns >= V of the last vector -> -inf
  auto mask_tail = [&](uint4& u) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&u);
    if (IS_BF16) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k >= tail) w[k >> 1] = (k & 1) ? ((w[k >> 1] & 0x0000ffffu) | 0xff800000u)
                                           : ((w[k >> 1] & 0xffff0000u) | 0x0000ff80u);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k >= tail) w[k] = 0xff800000u;
    }
  };
  int bi0 = 0x7fffffff;  // WANT_SUM == 0: element index tracked directly (memory-bound already)
  auto consume = [&](uint4 u, int vi) {
    if (vi == nvec - 1 && tail != EPV) mask_tail(u);
    if (!WANT_SUM) {
      float x[EPV];
      if (IS_BF16) {
        x[0] = bf16_lo(u.x); x[1] = bf16_hi(u.x); x[2] = bf16_lo(u.y); x[3] = bf16_hi(u.y);
        x[4 % EPV] = bf16_lo(u.z); x[5 % EPV] = bf16_hi(u.z);
        x[6 % EPV] = bf16_lo(u.w); x[7 % EPV] = bf16_hi(u.w);
      } else {
        x[0] = __uint_as_float(u.x); x[1] = __uint_as_float(u.y);
        x[2] = __uint_as_float(u.z); x[3] = __uint_as_float(u.w);
      }
#pragma unroll
      for (int k = 0; k < EPV; ++k) {  // strict '>' keeps the lowest index of equal values
        const bool g2 = x[k] > bv;
        bv = g2 ? x[k] : bv;
        bi0 = g2 ? vi * EPV + k : bi0;
      }
      return;
    }
    float m;
    if (IS_BF16) {
      const uint32_t p = bf16x2_max(bf16x2_max(u.x, u.y), bf16x2_max(u.z, u.w));
      m = fmaxf(bf16_lo(p), bf16_hi(p));
    } else {
      m = fmaxf(fmaxf(__uint_as_float(u.x), __uint_as_float(u.y)),
                fmaxf(__uint_as_float(u.z), __uint_as_float(u.w)));
    }
    const bool gt = m > bv;  // strict: an earlier vector keeps equal values
    const float mn = gt ? m : bv;
    if (WANT_SUM) {
      if (mn != -INFINITY) {
        const float nb = -mn * kLog2eCtc;
        float x[EPV];
        if (IS_BF16) {
          x[0] = bf16_lo(u.x); x[1] = bf16_hi(u.x); x[2] = bf16_lo(u.y); x[3] = bf16_hi(u.y);
          x[4 % EPV] = bf16_lo(u.z); x[5 % EPV] = bf16_hi(u.z);
          x[6 % EPV] = bf16_lo(u.w); x[7 % EPV] = bf16_hi(u.w);
        } else {
          x[0] = __uint_as_float(u.x); x[1] = __uint_as_float(u.y);
          x[2] = __uint_as_float(u.z); x[3] = __uint_as_float(u.w);
        }
        float add0 = 0.0f, add1 = 0.0f;
#pragma unroll
        for (int k = 0; k < EPV; k += 2) {
          add0 += exp2f_approx(fmaf(x[k], kLog2eCtc, nb));
          add1 += exp2f_approx(fmaf(x[k + 1], kLog2eCtc, nb));
        }
        bs = fmaf(bs, exp2f_approx(fmaf(bv, kLog2eCtc, nb)), add0 + add1);
      }
    }
    bv = mn;
    bvec = gt ? vi : bvec;
  };
  int i = lane;
  for (; i + 32 * (UNROLL - 1) < nvec; i += 32 * UNROLL) {
    uint4 u[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) u[k] = ld_nc_na(vp + i + 32 * k);
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) consume(u[k], i + 32 * k);
  }
  {
    uint4 u[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
      if (i + 32 * k < nvec) u[k] = ld_nc_na(vp + i + 32 * k);
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
      if (i + 32 * k < nvec) consume(u[k], i + 32 * k);
  }
  // element index of the lane's maximum: first element of vector bvec equal to bv
  int bi = bi0;
  if (WANT_SUM && bvec >= 0) {
    uint4 u = __ldg(vp + bvec);
    if (bvec == nvec - 1 && tail != EPV) mask_tail(u);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
    for (int k = EPV - 1; k >= 0; --k) {
      const float x = IS_BF16 ? ((k & 1) ? bf16_hi(w[k >> 1]) : bf16_lo(w[k >> 1])) : __uint_as_float(w[k]);
      if (x == bv) bi = bvec * EPV + k;
    }
  }
  ArgMax a{bv, bi, bs};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(0xffffffffu, a.v, o);
    const int ii = __shfl_xor_sync(0xffffffffu, a.i, o);
    const float ss = __shfl_xor_sync(0xffffffffu, a.s, o);
    argmax_merge(a, v, ii, ss, WANT_SUM);
  }
  if (lane == 0) {
    labels[row] = a.i;
    if (WANT_SUM && top_prob) top_prob[row] = 1.0f / a.s;
    if (WANT_SUM && lse) lse[row] = a.v + logf(a.s);  // log-sum-exp of the row (natural log)
  }
}

// Merge of the per-chunk arg-max partials the ctc_fc GEMM epilogue leaves (gemm2_tcgen05.cu ArgmaxEpi): one
// WARP per row: lane l folds chunks l, l + 32, ... (coalesced 16-byte loads, increasing column order: strict
// '>' keeps the lowest index), then a (value, index) butterfly.  (One thread per row walked 63 strided float4
// loads: 41 us at cfg2 in the ncu launch list, r02k.)
__global__ void __launch_bounds__(256)
    ctc_argmax_merge_kernel(const float4* __restrict__ partial, int chunks, const int* __restrict__ lengths,
                            int* __restrict__ labels, float* __restrict__ top_prob, float* __restrict__ lse,
                            int rows, int B, const float* __restrict__ logits, long long ldv, int V) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int t = row / B, b = row - t * B;
  if (t >= __ldg(lengths + b)) {
    if (lane == 0) {
      labels[row] = -1;
      if (top_prob) top_prob[row] = 0.0f;
      if (lse) lse[row] = 0.0f;
    }
    return;
  }
  const bool want_sum = top_prob != nullptr || lse != nullptr;
  ArgMax a{-INFINITY, 0x7fffffff, 0.0f};
  const float4* p = partial + (size_t)row * chunks;
  bool rescan = false;
  for (int c = lane; c < chunks; c += 32) {
    const float4 v = __ldg(p + c);
    rescan = v.w != 0.0f;  // (uniform over the row: set by the lean epilogue)
    argmax_merge(a, v.x, __float_as_int(v.y), v.z, want_sum);
  }
  rescan = __any_sync(0xffffffffu, rescan);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(0xffffffffu, a.v, o);
    const int i = __shfl_xor_sync(0xffffffffu, a.i, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, a.s, o);
    argmax_merge(a, v, i, s2, want_sum);
  }
  if (rescan) {
    // lean partials: a.i is the first column of the winning 128-column chunk (the lowest one among equal maxima);
    // the arg-max is the first stored logit of that chunk equal to the maximum (lane <-> 4 consecutive columns;
    // rows are 16-byte aligned and the chunk starts at a multiple of 128 columns)
    const int c0 = a.i + 4 * lane;
    int first = 0x7fffffff;
    if (c0 < V) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(logits + (size_t)row * ldv + c0));
      if (c0 + 3 < V && x.w == a.v) first = c0 + 3;
      if (c0 + 2 < V && x.z == a.v) first = c0 + 2;
      if (c0 + 1 < V && x.y == a.v) first = c0 + 1;
      if (x.x == a.v) first = c0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    a.i = first;
  }
  if (lane == 0) {
    labels[row] = a.i;
    if (top_prob) top_prob[row] = 1.0f / a.s;
    if (lse) lse[row] = a.v + logf(a.s);
  }
}

// One CTA per utterance.  smem: lab[L] | start[L+1] ints.
__global__ void __launch_bounds__(512)
    ctc_segment_kernel(const int* __restrict__ labels, const float* __restrict__ top_prob,
                       const int* __restrict__ lengths, int strategy, int* __restrict__ seg_id,
                       int* __restrict__ seg_start, float* __restrict__ weight,
                       int* __restrict__ new_lengths, int* __restrict__ max_new_len, int L, int B) {
  extern __shared__ int sm[];
  int* lab = sm;
  int* start = sm + L;
  __shared__ int warp_tot[16];
  __shared__ int carry_s;
  const int b = blockIdx.x;
  const int len = min(lengths[b], L);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int t = tid; t < L; t += blockDim.x) lab[t] = (t < len) ? labels[(size_t)t * B + b] : -1;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < L; t0 += blockDim.x) {
    const int t = t0 + tid;
    const int flag = (t < len) && (t == 0 || lab[t] != lab[t - 1]);
    int incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int base = carry_s;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    const int sid = base + incl - 1;
    if (t < L) seg_id[(size_t)t * B + b] = (t < len) ? sid : -1;
    if (flag) {
      start[sid] = t;
      seg_start[(size_t)sid * B + b] = t;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = carry_s;
      for (int w = 0; w < nw; ++w) tot += warp_tot[w];
      carry_s = tot;
    }
    __syncthreads();
  }
  const int nseg = carry_s;
  if (tid == 0) {
    start[nseg] = len;
    new_lengths[b] = nseg;
    atomicMax(max_new_len, nseg);
  }
  for (int t = len + tid; t < L; t += blockDim.x) weight[(size_t)t * B + b] = 0.0f;
  __syncthreads();
  for (int s = tid; s < nseg; s += blockDim.x) {
    const int a = start[s], e = start[s + 1];
    if (strategy == FBKST_CTC_AVG) {
      const float w = 1.0f / (float)(e - a);
      for (int t = a; t < e; ++t) weight[(size_t)t * B + b] = w;
    } else if (strategy == FBKST_CTC_WEIGHTED) {
      float sum = 0.0f;
      for (int t = a; t < e; ++t) sum += top_prob[(size_t)t * B + b];
      for (int t = a; t < e; ++t) weight[(size_t)t * B + b] = top_prob[(size_t)t * B + b] / sum;
    } else {  // softmax over the run of the PROBABILITIES (conv_transformer.py:422)
      float mx = -INFINITY;
      for (int t = a; t < e; ++t) mx = fmaxf(mx, top_prob[(size_t)t * B + b]);
      float sum = 0.0f;
      for (int t = a; t < e; ++t) sum += expf(top_prob[(size_t)t * B + b] - mx);
      for (int t = a; t < e; ++t)
        weight[(size_t)t * B + b] = expf(top_prob[(size_t)t * B + b] - mx) / sum;
    }
  }
}

// Generic D (any multiple of 4; small test shapes): one CTA per output row, threads over D.
__global__ void __launch_bounds__(128)
    ctc_compress_generic_kernel(const float* __restrict__ x, const int* __restrict__ seg_start,
                                const float* __restrict__ weight, const int* __restrict__ lengths,
                                const int* __restrict__ new_lengths,
                                const int* __restrict__ max_new_len, float* __restrict__ out, int L,
                                int B, int D, int guard) {
  const int rows = min(__ldg(max_new_len) + guard, L) * B;  // (guard rows: see fbkst_ctc_compress)
  const int nvec = D >> 2;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int s = row / B, b = row - s * B;
    const int nl = __ldg(new_lengths + b);
    float4* op = reinterpret_cast<float4*>(out + (size_t)row * D);
    const bool pad = s >= nl;
    const int a = pad ? 0 : __ldg(seg_start + (size_t)s * B + b);
    const int e = pad ? 0 : ((s + 1 < nl) ? __ldg(seg_start + (size_t)(s + 1) * B + b) : min(__ldg(lengths + b), L));
    for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = a; t < e; ++t) {
        const float w = __ldg(weight + (size_t)t * B + b);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)t * B + b) * D) + j);
        acc.x = fmaf(w, v.x, acc.x);
        acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z);
        acc.w = fmaf(w, v.w, acc.w);
      }
      op[j] = acc;
    }
  }
}

// out[s*B+b, :] = sum_{t in segment s of b} weight[t,b] * x[t*B+b, :].
// INPUT-stationary and PERSISTENT.  A warp task is (chunk of CH = 16 consecutive frames, utterance b,
// 128-column group); the task OWNS the runs that START inside its chunk and reads exactly their
// frames [begin, end) -- `end` may lie beyond the chunk when the last run continues -- so every frame
// of x is read by exactly one task per column group (once from DRAM, never again from L2), every
// output row has exactly one writer and a fixed summation order (ascending t): no atomics,
// run-to-run identical results.
//   * (begin, end) come from a 32-frame window of run ids [t0-1, t0+31) that the lanes fetch ONE TASK
//     AHEAD (the previous iteration issues the loads), so the only memory round trip on a task's
//     critical path is the batch of <= 16 frame loads (512 B per warp each, all issued before any is
//     used) with the per-frame run id / pooling weight fetched by lanes 0..15 in the same flight.
//   * runs longer than the window (merged blank runs reach 60-80 frames) take their end from
//     seg_start[run + 1] and simply need more 16-frame batches; no per-batch metadata dependency.
//   * the grid is sized to the resident-warp capacity and each warp walks tasks in a grid-stride
//     loop: warps drift out of phase, so frame loads stream continuously instead of in waves
//     (the one-task-per-warp version ran 2.6 waves of issue -> wait -> fold: 46 % of HBM peak,
//     profiles/r01e_ncu_ctc.txt; one task per OUTPUT row was worse still: profiles/r01b).
// The task of chunk c also zero-fills the padding rows new_len[b] <= s < max_new_len + guard, s in chunk c.
constexpr int CTC_CH = 16;

__device__ __forceinline__ int ctc_window_sid(const int* __restrict__ seg_id, int task, int ncg, int B,
                                              int L, int lane) {
  const int b = (task / ncg) % B, c = task / (ncg * B);
  const int t = c * CTC_CH - 1 + lane;
  return (t >= 0 && t < L) ? __ldg(seg_id + (size_t)t * B + b) : (t < 0 ? -2 : -1);
}

__global__ void __launch_bounds__(256)
    ctc_compress_kernel(const float* __restrict__ x, const int* __restrict__ seg_id,
                        const int* __restrict__ seg_start, const float* __restrict__ weight,
                        const int* __restrict__ lengths, const int* __restrict__ new_lengths,
                        const int* __restrict__ max_new_len, float* __restrict__ out, int L, int B, int D,
                        int tasks, int guard) {
  constexpr int FR = CTC_CH;
  const int ncg = D >> 7;  // 128-column groups per row
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= tasks) return;
  const size_t fstride = (size_t)B * (D >> 2);  // float4 elements between consecutive frames
  const int mx = min(__ldg(max_new_len) + guard, L);  // zero fill incl. the guard rows (fbkst_ctc_compress)
  int sidw = ctc_window_sid(seg_id, task, ncg, B, L, lane);
  for (; task < tasks; task += nwarps) {
    const int cg = task % ncg, b = (task / ncg) % B, c = task / (ncg * B);
    const int t0 = c * FR;
    const int sid = sidw;
    if (task + nwarps < tasks) sidw = ctc_window_sid(seg_id, task + nwarps, ncg, B, L, lane);  // one task ahead
    const int nl = __ldg(new_lengths + b);
    const float4* xcol = reinterpret_cast<const float4*>(x + (size_t)b * D) + cg * 32 + lane;
    float4* ocol = reinterpret_cast<float4*>(out + (size_t)b * D) + cg * 32 + lane;
    // lane j <-> frame t0 - 1 + j.  Runs that start in the chunk: boundary at some lane 1..FR.
    const int prv = __shfl_up_sync(0xffffffffu, sid, 1);
    const unsigned bmask = __ballot_sync(0xffffffffu, lane >= 1 && lane <= FR && sid >= 0 && sid != prv);
    if (bmask != 0) {  // warp-uniform
      const int jb = __ffs(bmask) - 1;
      const int s_last = __shfl_sync(0xffffffffu, sid, FR);  // run of the chunk's last frame (-1: padding)
      int begin = t0 - 1 + jb, end;
      if (s_last < 0) {
        const unsigned em = __ballot_sync(0xffffffffu, lane > jb && sid < 0);
        end = t0 - 1 + (__ffs(em) - 1);
      } else {
        const unsigned em = __ballot_sync(0xffffffffu, lane > FR && sid != s_last);
        if (em != 0)
          end = t0 - 1 + (__ffs(em) - 1);
        else  // the run leaves the window: its end is the next run's start (or the utterance's end)
          end = (s_last + 1 < nl) ? __ldg(seg_start + (size_t)(s_last + 1) * B + b) : min(__ldg(lengths + b), L);
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur = -1;  // run being accumulated (warp-uniform)
      for (int fb = begin; fb < end; fb += FR) {
        float4 v[FR];
#pragma unroll
        for (int k = 0; k < FR; ++k)
          v[k] = (fb + k < end) ? ld_nc_na(xcol + (size_t)(fb + k) * fstride) : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool has = lane < FR && fb + lane < end;
        const int sl = has ? __ldg(seg_id + (size_t)(fb + lane) * B + b) : -1;
        const float wl = has ? __ldg(weight + (size_t)(fb + lane) * B + b) : 0.0f;
#pragma unroll
        for (int k = 0; k < FR; ++k) {
          const int sk = __shfl_sync(0xffffffffu, sl, k);
          const float wk = __shfl_sync(0xffffffffu, wl, k);
          if (sk != cur) {  // warp-uniform
            if (cur >= 0) ocol[(size_t)cur * fstride] = acc;
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            cur = sk;
          }
          if (sk >= 0) {
            acc.x = fmaf(wk, v[k].x, acc.x);
            acc.y = fmaf(wk, v[k].y, acc.y);
            acc.z = fmaf(wk, v[k].z, acc.z);
            acc.w = fmaf(wk, v[k].w, acc.w);
          }
        }
      }
      if (cur >= 0) ocol[(size_t)cur * fstride] = acc;
    }
    // padding rows of the compressed output that fall into this chunk's index range
    for (int s = max(t0, nl); s < min(t0 + FR, mx); ++s) ocol[(size_t)s * fstride] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_ctc_argmax_lse(const void* logits, int logits_dtype, int64_t ldv,
                                    const int32_t* lengths, int32_t* labels, float* top_prob,
                                    float* lse, int L, int B, int V, fbkst_stream_t stream) {
  FBKST_REQUIRE(logits && lengths && labels, "fbkst_ctc_argmax: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && V > 0 && ldv >= V, "fbkst_ctc_argmax: bad shape");
  FBKST_REQUIRE(logits_dtype == FBKST_BF16 || logits_dtype == FBKST_F32,
                "fbkst_ctc_argmax: dtype must be bf16 or fp32");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = L * B;
  int grid = (rows + 7) / 8;
  const size_t esz = logits_dtype == FBKST_BF16 ? 2 : 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15u) == 0 && ((size_t)ldv * esz) % 16 == 0 &&
                       (size_t)((V + 16 / esz - 1) / (16 / esz)) * 16 <= (size_t)ldv * esz;
  if (aligned) {
    const uint4* lp = reinterpret_cast<const uint4*>(logits);
    const long long rv = (long long)((size_t)ldv * esz / 16);
    const bool sum = top_prob || lse;
    if (logits_dtype == FBKST_BF16 && sum)
      ctc_argmax_vec_kernel<1, 1><<<grid, 256, 0, st>>>(lp, rv, lengths, labels, top_prob, lse, rows, B, V);
    else if (logits_dtype == FBKST_BF16)
      ctc_argmax_vec_kernel<1, 0><<<grid, 256, 0, st>>>(lp, rv, lengths, labels, top_prob, lse, rows, B, V);
    else if (sum)
      ctc_argmax_vec_kernel<0, 1><<<grid, 256, 0, st>>>(lp, rv, lengths, labels, top_prob, lse, rows, B, V);
    else
      ctc_argmax_vec_kernel<0, 0><<<grid, 256, 0, st>>>(lp, rv, lengths, labels, top_prob, lse, rows, B, V);
    FBKST_CHECK_CUDA(cudaGetLastError());
    return FBKST_OK;
  }
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  if (logits_dtype == FBKST_BF16)
    ctc_argmax_kernel<1><<<grid, 256, 0, st>>>(logits, ldv, lengths, labels, top_prob, lse, rows, B, V);
  else
    ctc_argmax_kernel<0><<<grid, 256, 0, st>>>(logits, ldv, lengths, labels, top_prob, lse, rows, B, V);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ctc_argmax(const void* logits, int logits_dtype, int64_t ldv,
                                const int32_t* lengths, int32_t* labels, float* top_prob, int L,
                                int B, int V, fbkst_stream_t stream) {
  return fbkst_ctc_argmax_lse(logits, logits_dtype, ldv, lengths, labels, top_prob, nullptr, L, B, V,
                              stream);
}

extern "C" int fbkst_ctc_segment(const int32_t* labels, const float* top_prob,
                                 const int32_t* lengths, int strategy, int32_t* seg_id,
                                 int32_t* seg_start, float* weight, int32_t* new_lengths,
                                 int32_t* max_new_len, int L, int B, fbkst_stream_t stream) {
  FBKST_REQUIRE(labels && lengths && seg_id && seg_start && weight && new_lengths && max_new_len,
                "fbkst_ctc_segment: null pointer");
  FBKST_REQUIRE(strategy == FBKST_CTC_AVG || top_prob != nullptr,
                "fbkst_ctc_segment: weighted/softmax need top_prob");
  FBKST_REQUIRE(strategy >= 0 && strategy <= 2, "fbkst_ctc_segment: unknown strategy %d", strategy);
  FBKST_REQUIRE(L > 0 && B > 0 && L <= 24000, "fbkst_ctc_segment: L=%d out of range (1..24000)", L);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(int) * (2 * (size_t)L + 1);
  static PerDeviceFlag configured;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_segment_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  FBKST_CHECK_CUDA(cudaMemsetAsync(max_new_len, 0, sizeof(int32_t), st));
  ctc_segment_kernel<<<B, 512, smem, st>>>(labels, top_prob, lengths, strategy, seg_id, seg_start,
                                           weight, new_lengths, max_new_len, L, B);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ctc_compress(const float* x, const int32_t* seg_id, const int32_t* seg_start,
                                  const float* weight, const int32_t* lengths,
                                  const int32_t* new_lengths, const int32_t* max_new_len, float* out,
                                  int L, int B, int D, fbkst_stream_t stream) {
  FBKST_REQUIRE(x && seg_id && seg_start && weight && lengths && new_lengths && max_new_len && out,
                "fbkst_ctc_compress: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && D > 0 && D % 4 == 0, "fbkst_ctc_compress: bad shape (D %% 4)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Guard rows: the layers after compression run on worst-case grids with a device-side row limit
  // (max_new_len * B rows) that their GEMM tiles round up to 256 rows, so up to 255 rows past the limit are
  // read-modify-written (x += f(x)) by every residual epilogue.  Left untouched here they carry their values
  // from one forward to the next in a persistent workspace, grow geometrically (x1.4 per step measured) and
  // reach inf after ~70 forwards -- and attention then reads them as V rows of a partially valid key tile
  // (0 * inf = NaN in every row of the utterance).  Zeroing them with the other padding rows restarts them
  // at every forward.
  const int guard = (512 + B - 1) / B + 1;
  if (D % 128 != 0 || D > 1024) {
    int g = L * B;
    const int gcap = num_sms() * 16;
    if (g > gcap) g = gcap;
    ctc_compress_generic_kernel<<<g, 128, 0, st>>>(x, seg_start, weight, lengths, new_lengths,
                                                   max_new_len, out, L, B, D, guard);
    FBKST_CHECK_CUDA(cudaGetLastError());
    return FBKST_OK;
  }
  const long long tasks = (long long)((L + CTC_CH - 1) / CTC_CH) * B * (D / 128);
  FBKST_REQUIRE(tasks < (1ll << 31), "fbkst_ctc_compress: too many tasks");
  // persistent: as many CTAs as can be resident (registers allow 2 x 256 threads per SM)
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    FBKST_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, ctc_compress_kernel, 256, 0));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  long long grid = (tasks + 7) / 8;
  if (grid > (long long)num_sms() * ctas_per_sm) grid = (long long)num_sms() * ctas_per_sm;
  ctc_compress_kernel<<<(int)grid, 256, 0, st>>>(x, seg_id, seg_start, weight, lengths, new_lengths,
                                                 max_new_len, out, L, B, D, (int)tasks, guard);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ctc_argmax_merge(const float* partial, int chunks, const float* logits, int64_t ldv, int V,
                                      const int32_t* lengths, int32_t* labels, float* top_prob, float* lse, int L,
                                      int B, fbkst_stream_t stream) {
  FBKST_REQUIRE(partial && lengths && labels && chunks > 0 && L > 0 && B > 0, "fbkst_ctc_argmax_merge: bad arguments");
  FBKST_REQUIRE(logits && V > 0 && ldv >= V && ldv % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0,
                "fbkst_ctc_argmax_merge: the logits of fbkst_linear_argmax_f32 (16-byte aligned rows) are required");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = L * B;
  ctc_argmax_merge_kernel<<<(rows + 7) / 8, 256, 0, st>>>(reinterpret_cast<const float4*>(partial), chunks,
                                                             lengths, labels, top_prob, lse, rows, B, logits, ldv, V);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
