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

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// One warp per (t, b) row.  BF16 = 1: 8 logits per 16-byte load; else fp32, 4 per load.
template <int IS_BF16>
__global__ void __launch_bounds__(256)
    ctc_argmax_kernel(const void* __restrict__ logits, long long ldv,
                      const int* __restrict__ lengths, int* __restrict__ labels,
                      float* __restrict__ top_prob, int rows, int B, int V) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const bool want_sum = top_prob != nullptr;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += warps_per_grid) {
    const int t = row / B, b = row - t * B;
    if (t >= __ldg(lengths + b)) {
      if (lane == 0) {
        labels[row] = -1;
        if (want_sum) top_prob[row] = 0.0f;
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
      if (want_sum) top_prob[row] = 1.0f / a.s;
    }
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

// out[s*B+b, :] = sum_{t in segment s of b} weight[t,b] * x[t*B+b, :]; one CTA row-loop,
// threads over D as float4.  Grid covers the worst case (L*B rows); rows >= max_new_len exit.
__global__ void __launch_bounds__(128)
    ctc_compress_kernel(const float* __restrict__ x, const int* __restrict__ seg_start,
                        const float* __restrict__ weight, const int* __restrict__ lengths,
                        const int* __restrict__ new_lengths, const int* __restrict__ max_new_len,
                        float* __restrict__ out, int L, int B, int D) {
  const int rows = min(__ldg(max_new_len), L) * B;
  const int nvec = D >> 2;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int s = row / B, b = row - s * B;
    const int nl = __ldg(new_lengths + b);
    float4* op = reinterpret_cast<float4*>(out + (size_t)row * D);
    if (s >= nl) {
      for (int j = threadIdx.x; j < nvec; j += blockDim.x) op[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int a = __ldg(seg_start + (size_t)s * B + b);
    const int e = (s + 1 < nl) ? __ldg(seg_start + (size_t)(s + 1) * B + b) : min(__ldg(lengths + b), L);
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

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_ctc_argmax(const void* logits, int logits_dtype, int64_t ldv,
                                const int32_t* lengths, int32_t* labels, float* top_prob, int L,
                                int B, int V, fbkst_stream_t stream) {
  FBKST_REQUIRE(logits && lengths && labels, "fbkst_ctc_argmax: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && V > 0 && ldv >= V, "fbkst_ctc_argmax: bad shape");
  FBKST_REQUIRE(logits_dtype == FBKST_BF16 || logits_dtype == FBKST_F32,
                "fbkst_ctc_argmax: dtype must be bf16 or fp32");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = L * B;
  int grid = (rows + 7) / 8;
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  if (logits_dtype == FBKST_BF16)
    ctc_argmax_kernel<1><<<grid, 256, 0, st>>>(logits, ldv, lengths, labels, top_prob, rows, B, V);
  else
    ctc_argmax_kernel<0><<<grid, 256, 0, st>>>(logits, ldv, lengths, labels, top_prob, rows, B, V);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
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
  static bool configured = false;
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

extern "C" int fbkst_ctc_compress(const float* x, const int32_t* seg_start, const float* weight,
                                  const int32_t* lengths, const int32_t* new_lengths,
                                  const int32_t* max_new_len, float* out, int L, int B, int D,
                                  fbkst_stream_t stream) {
  FBKST_REQUIRE(x && seg_start && weight && lengths && new_lengths && max_new_len && out,
                "fbkst_ctc_compress: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && D > 0 && D % 4 == 0, "fbkst_ctc_compress: bad shape (D %% 4)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int grid = L * B;
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  ctc_compress_kernel<<<grid, 128, 0, st>>>(x, seg_start, weight, lengths, new_lengths, max_new_len,
                                            out, L, B, D);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
