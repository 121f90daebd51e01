// Next row N1: the CTC criterion that consumes the encoder's `ctc_out`, on device.
//   reference: examples/speech_recognition/criterions/CTC_loss.py:31-74 (compute_ctc_uer: per-utterance
//   argmax().tolist() -> python groupby -> blank removal -> EditDistance.align with costs 0/3/3/4,
//   examples/speech_recognition/utils/wer_utils.py:80-96,141-202) and :143-151 (F.ctc_loss, sum
//   reduction, zero_infinity=True, on log_softmax(ctc_out.float())).
//
//   fbkst_ctc_uer       CTA per utterance: collapse the frame labels that fbkst_ctc_argmax already
//                       produced (run boundaries -> block scan -> compaction in shared memory), then
//                       the alignment DP as an anti-diagonal wavefront that carries, with each cell's
//                       cost, the number of non-match steps of the path the reference's backtrace
//                       would follow (same strict-'<' tie-breaks: diagonal, then column-1, then row-1).
//                       Integer work: errors / totals are bit-exact.
//   fbkst_ctc_loss_fwd  CTA per utterance: log-space alpha recursion over the blank-extended target;
//                       the per-frame log-probabilities of the U+1 distinct columns (labels + blank)
//                       are gathered in chunks of frames (independent loads, no load->use chain per
//                       time step) and normalised with the log-sum-exp fbkst_ctc_argmax_lse wrote.
// Both are latency-bound (B CTAs, O(L) dependent steps); nothing here touches tensor cores.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace fbkst {

// ------------------------------------------------------------------------------------------------
// UER.  smem ints: pred[L] | sc[3][U+1] | ne[3][U+1] | tgt[U]
__global__ void __launch_bounds__(256)
    ctc_uer_kernel(const int* __restrict__ labels, const int* __restrict__ in_lengths,
                   const long long* __restrict__ targets, long long ldt,
                   const int* __restrict__ target_lengths, int blank, int* __restrict__ errors,
                   int* __restrict__ pred_lengths, int L, int B, int Umax) {
  extern __shared__ int sm[];
  int* pred = sm;
  int* sc = pred + L;
  int* ne = sc + 3 * (Umax + 1);
  int* tgt = ne + 3 * (Umax + 1);
  __shared__ int warp_tot[8];
  __shared__ int carry_s;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int len = max(0, min(in_lengths[b], L));
  const int U = max(0, min(target_lengths[b], Umax));
  for (int j = tid; j < U; j += blockDim.x) tgt[j] = (int)targets[(size_t)b * ldt + j];
  if (tid == 0) carry_s = 0;
  __syncthreads();
  // 1. dedup consecutive frames (groupby), drop blanks, compact (CTC_loss.py:50-58)
  for (int t0 = 0; t0 < len; t0 += blockDim.x) {
    const int t = t0 + tid;
    int cur = -1, keep = 0;
    if (t < len) {
      cur = __ldg(labels + (size_t)t * B + b);
      const int prv = t > 0 ? __ldg(labels + (size_t)(t - 1) * B + b) : -2;
      keep = (cur != prv) && (cur != blank);
    }
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int base = carry_s;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    if (keep) pred[base + incl - 1] = cur;
    __syncthreads();
    if (tid == 0) {
      int tot = carry_s;
      for (int w = 0; w < nw; ++w) tot += warp_tot[w];
      carry_s = tot;
    }
    __syncthreads();
  }
  const int P = carry_s;  // rows of the DP = predicted tokens (the reference passes them as `refs`)
  // 2. alignment DP over anti-diagonals d = i + j; cell (i, j) lives at column j of diagonal d.
  //    cost: match 0, row/column step 3, mismatch 4 (wer_utils.py:90-96); a cell takes the
  //    diagonal first, then (i, j-1) if strictly cheaper, then (i-1, j) if strictly cheaper
  //    (wer_utils.py:170-192).  ne = non-match codes on the backtrace path = the reference's count.
  const int W = Umax + 1;
  for (int d = 0; d <= P + U; ++d) {
    int* s0 = sc + (d % 3) * W;            // diagonal d
    int* n0 = ne + (d % 3) * W;
    const int* s1 = sc + ((d + 2) % 3) * W;  // diagonal d-1
    const int* n1 = ne + ((d + 2) % 3) * W;
    const int* s2 = sc + ((d + 1) % 3) * W;  // diagonal d-2
    const int* n2 = ne + ((d + 1) % 3) * W;
    const int jlo = max(0, d - P), jhi = min(d, U);
    for (int j = jlo + tid; j <= jhi; j += blockDim.x) {
      const int i = d - j;
      int best, nerr;
      if (i == 0 && j == 0) {
        best = 0; nerr = 0;
      } else if (i == 0) {
        best = s1[j - 1] + 3; nerr = n1[j - 1] + 1;
      } else if (j == 0) {
        best = s1[0] + 3; nerr = n1[0] + 1;
      } else {
        const bool same = pred[i - 1] == tgt[j - 1];
        best = s2[j - 1] + (same ? 0 : 4);
        nerr = n2[j - 1] + (same ? 0 : 1);
        const int ins = s1[j - 1] + 3;
        if (ins < best) { best = ins; nerr = n1[j - 1] + 1; }
        const int del = s1[j] + 3;
        if (del < best) { best = del; nerr = n1[j] + 1; }
      }
      s0[j] = best;
      n0[j] = nerr;
    }
    __syncthreads();
  }
  if (tid == 0) {
    // both sequences empty: the reference's align() returns NaN and the criterion then fails on
    // `.codes`; the device path defines that case as 0 errors
    errors[b] = ne[((P + U) % 3) * W + U];
    pred_lengths[b] = P;
  }
}

// totals[0] = sum errors, totals[1] = sum target lengths (fixed order: integer, exact either way)
__global__ void ctc_uer_totals_kernel(const int* __restrict__ errors,
                                      const int* __restrict__ target_lengths, int Umax,
                                      long long* __restrict__ totals, int B) {
  long long e = 0, n = 0;
  for (int b = threadIdx.x; b < B; b += 32) {
    e += errors[b];
    n += max(0, min(target_lengths[b], Umax));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if (threadIdx.x == 0) {
    totals[0] = e;
    totals[1] = n;
  }
}

// ------------------------------------------------------------------------------------------------
// CTC negative log-likelihood.
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

template <int IS_BF16>
__device__ __forceinline__ float load_logit(const void* p, size_t idx) {
  if (IS_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  return reinterpret_cast<const float*>(p)[idx];
}

// smem: alpha[2][S] floats | lp[TCH][U+1] floats (column U = blank) | tgt[U] ints
template <int IS_BF16>
__global__ void __launch_bounds__(256)
    ctc_loss_fwd_kernel(const void* __restrict__ logits, long long ldv, const float* __restrict__ lse,
                        const int* __restrict__ in_lengths, const long long* __restrict__ targets,
                        long long ldt, const int* __restrict__ target_lengths, int blank,
                        float* __restrict__ nll, int L, int B, int Umax, int TCH) {
  extern __shared__ float smf[];
  const int Smax = 2 * Umax + 1;
  float* alpha = smf;
  float* lp = alpha + 2 * Smax;
  int* tgt = reinterpret_cast<int*>(lp + (size_t)TCH * (Umax + 1));
  const int b = blockIdx.x, tid = threadIdx.x;
  const int len = max(0, min(in_lengths[b], L));
  const int U = max(0, min(target_lengths[b], Umax));
  const int S = 2 * U + 1;
  for (int j = tid; j < U; j += blockDim.x) tgt[j] = (int)targets[(size_t)b * ldt + j];
  __syncthreads();
  if (len == 0) {  // F.ctc_loss: empty input aligns only with an empty target
    if (tid == 0) nll[b] = 0.0f;  // U == 0: -log 1 = 0;  U > 0: inf -> 0 under zero_infinity
    return;
  }
  const int W = U + 1;
  int cur = 0;
  for (int t0 = 0; t0 < len; t0 += TCH) {
    const int nt = min(TCH, len - t0);
    // gather log p_t(c) for the U labels and the blank, nt frames at once (independent loads)
    for (int e = tid; e < nt * W; e += blockDim.x) {
      const int tt = e / W, u = e - tt * W;
      const size_t row = (size_t)(t0 + tt) * B + b;
      const int col = (u == U) ? blank : tgt[u];
      lp[tt * W + u] = load_logit<IS_BF16>(logits, row * (size_t)ldv + col) - __ldg(lse + row);
    }
    __syncthreads();
    for (int tt = 0; tt < nt; ++tt) {
      const float* a0 = alpha + cur * Smax;
      float* a1 = alpha + (cur ^ 1) * Smax;
      const float* lpt = lp + tt * W;
      for (int s = tid; s < S; s += blockDim.x) {
        const int u = s >> 1;
        const bool is_label = s & 1;
        const float p = is_label ? lpt[u] : lpt[U];
        float v;
        if (t0 + tt == 0) {
          v = (s <= 1) ? p : -INFINITY;
        } else {
          const float x0 = a0[s];
          const float x1 = s >= 1 ? a0[s - 1] : -INFINITY;
          const float x2 = (is_label && s >= 3 && tgt[u] != tgt[u - 1]) ? a0[s - 2] : -INFINITY;
          v = lse3(x0, x1, x2) + p;
        }
        a1[s] = v;
      }
      cur ^= 1;
      __syncthreads();
    }
  }
  if (tid == 0) {
    const float* a = alpha + cur * Smax;
    const float ll = (S > 1) ? lse3(a[S - 1], a[S - 2], -INFINITY) : a[0];
    const float v = -ll;
    nll[b] = (isinf(v) || isnan(v)) ? 0.0f : v;  // zero_infinity=True (CTC_loss.py:150)
  }
}


// ------------------------------------------------------------------------------------------------
// CTC loss backward w.r.t. the LOGITS (autograd of F.ctc_loss(log_softmax(z)), CTC_loss.py:143-151):
//   d nll / d z[t, c] = softmax(z[t])[c] - sum_{s : l'_s = c} gamma_t(s),
//   gamma_t(s) = exp(alpha_t(s) + beta_t(s) - lp_t(l'_s) - ll)        (alpha, beta both include frame t)
// (1) ctc_softmax_grad_kernel: the dense term g * softmax for the valid frames (zeros beyond each length):
//     HBM streaming, one warp per row;
// (2) ctc_loss_bwd_kernel: one CTA per utterance: alpha pass (kept in a global workspace), beta pass, and the
//     sparse subtraction of g * gamma at the target / blank columns.  Infeasible alignments (ll = -inf,
//     zero_infinity=True) get an all-zero gradient.  Duplicate labels are folded through a "next position
//     with the same label" chain so that every (t, column) has exactly one writer (no atomics).
template <int IS_BF16>
__global__ void __launch_bounds__(256)
    ctc_softmax_grad_kernel(const void* __restrict__ logits, long long ldv, const float* __restrict__ lse,
                            const int* __restrict__ in_lengths, const float* __restrict__ grad_loss,
                            float* __restrict__ dz, long long ldd, int rows, int B, int V) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int t = row / B, b = row - t * B;
  float* op = dz + (size_t)row * ldd;
  if (t >= __ldg(in_lengths + b)) {
    for (int c = lane; c < V; c += 32) op[c] = 0.0f;
    return;
  }
  const float g = __ldg(grad_loss), l = __ldg(lse + row);
  for (int c = lane; c < V; c += 32)
    op[c] = g * expf(load_logit<IS_BF16>(logits, (size_t)row * ldv + c) - l);
}

// The recursions are latency chains (one CTA per utterance, one __syncthreads per frame), so nothing on the chain
// may wait for global memory: alpha_{t-1} / beta_{t+1} live in shared memory (alpha_t is ALSO streamed to the
// global workspace for the beta pass, fire-and-forget), a chunk of TCH frames has its log-probabilities and --
// in the beta pass -- its alpha rows staged in shared memory with all loads in flight at once, the gammas of a
// chunk are collected in shared memory and subtracted from dz after the chunk by independent read-modify-writes.
// (First version: alpha_{t-1} re-read from global every frame, two barriers and a dependent global RMW per
// frame of the beta pass: 1.09 ms per cfg4 step for 64 utterances of 375 frames.)
// smem floats: ab[2][S] | lp[TCH][W] | al[TCH][S] | gm[TCH][W] | bp[TCH][8];  ints: tgt[U] | nxt[U] | hd[U]
// (W = U + 1, column U = blank; S = 2U + 1)
template <int IS_BF16>
__global__ void __launch_bounds__(256)
    ctc_loss_bwd_kernel(const void* __restrict__ logits, long long ldv, const float* __restrict__ lse,
                        const int* __restrict__ in_lengths, const long long* __restrict__ targets,
                        long long ldt, const int* __restrict__ target_lengths, int blank,
                        const float* __restrict__ grad_loss, float* __restrict__ alpha_ws,
                        float* __restrict__ dz, long long ldd, int L, int B, int V, int Umax, int TCH) {
  extern __shared__ float smf[];
  const int Smax = 2 * Umax + 1, Wmax = Umax + 1;
  float* ab = smf;                            // [2][Smax]  alpha / beta double buffer
  float* lp = ab + 2 * Smax;                  // [TCH][Wmax]
  float* al = lp + (size_t)TCH * Wmax;        // [TCH][Smax]
  float* gm = al + (size_t)TCH * Smax;        // [TCH][Wmax] label gammas (column U unused)
  float* bp = gm + (size_t)TCH * Wmax;        // [TCH][8]    per-warp partial sums of the blank gammas
  int* tgt = reinterpret_cast<int*>(bp + (size_t)TCH * 8);
  int* nxt = tgt + Umax;
  int* hd = nxt + Umax;
  __shared__ float s_ll;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int len = max(0, min(in_lengths[b], L));
  const int U = max(0, min(target_lengths[b], Umax));
  const int S = 2 * U + 1, W = U + 1;
  for (int j = tid; j < U; j += blockDim.x) tgt[j] = (int)targets[(size_t)b * ldt + j];
  __syncthreads();
  // nxt[u]: next position with the same label (-1: none); a position is a chain HEAD if no earlier one matches
  for (int j = tid; j < U; j += blockDim.x) {
    int n = -1;
    for (int k = j + 1; k < U; ++k)
      if (tgt[k] == tgt[j]) {
        n = k;
        break;
      }
    bool head = true;
    for (int k = 0; k < j; ++k)
      if (tgt[k] == tgt[j]) {
        head = false;
        break;
      }
    nxt[j] = n;
    hd[j] = head ? 1 : 0;
  }
  if (len == 0) return;  // no frames: nothing to subtract (the dense kernel wrote zeros)
  float* aw = alpha_ws + (size_t)b * L * Smax;
  const float g = __ldg(grad_loss);
  // ---- alpha pass (as ctc_loss_fwd_kernel), every alpha_t streamed to the workspace
  int cur = 0;
  for (int t0 = 0; t0 < len; t0 += TCH) {
    const int nt = min(TCH, len - t0);
    __syncthreads();
    for (int e = tid; e < nt * W; e += blockDim.x) {
      const int tt = e / W, u = e - tt * W;
      const size_t row = (size_t)(t0 + tt) * B + b;
      const int col = (u == U) ? blank : tgt[u];
      lp[tt * W + u] = load_logit<IS_BF16>(logits, row * (size_t)ldv + col) - __ldg(lse + row);
    }
    __syncthreads();
    for (int tt = 0; tt < nt; ++tt) {
      const int t = t0 + tt;
      const float* a0 = ab + cur * Smax;
      float* a1 = ab + (cur ^ 1) * Smax;
      float* aout = aw + (size_t)t * Smax;
      const float* lpt = lp + tt * W;
      for (int s = tid; s < S; s += blockDim.x) {
        const int u = s >> 1;
        const bool is_label = s & 1;
        const float p = is_label ? lpt[u] : lpt[U];
        float v;
        if (t == 0) {
          v = (s <= 1) ? p : -INFINITY;
        } else {
          const float x0 = a0[s];
          const float x1 = s >= 1 ? a0[s - 1] : -INFINITY;
          const float x2 = (is_label && s >= 3 && tgt[u] != tgt[u - 1]) ? a0[s - 2] : -INFINITY;
          v = lse3(x0, x1, x2) + p;
        }
        a1[s] = v;
        aout[s] = v;
      }
      cur ^= 1;
      __syncthreads();  // alpha_t visible to the whole CTA before step t + 1
    }
  }
  if (tid == 0) {
    const float* a = ab + cur * Smax;  // alpha_{len-1}
    s_ll = (S > 1) ? lse3(a[S - 1], a[S - 2], -INFINITY) : a[0];
  }
  __syncthreads();
  const float ll = s_ll;
  if (isinf(ll) || isnan(ll)) {  // zero_infinity: no gradient for this utterance at all
    for (int t = 0; t < len; ++t) {
      float* op = dz + ((size_t)t * B + b) * ldd;
      for (int c = tid; c < V; c += blockDim.x) op[c] = 0.0f;
    }
    return;
  }
  // ---- beta pass + sparse subtraction, chunk by chunk from the end
  cur = 0;
  for (int t1 = len; t1 > 0; t1 -= TCH) {
    const int t0 = max(0, t1 - TCH), nt = t1 - t0;
    __syncthreads();  // the previous chunk's gm / bp / lp / al have been consumed
    for (int e = tid; e < nt * W; e += blockDim.x) {
      const int tt = e / W, u = e - tt * W;
      const size_t row = (size_t)(t0 + tt) * B + b;
      const int col = (u == U) ? blank : tgt[u];
      lp[tt * W + u] = load_logit<IS_BF16>(logits, row * (size_t)ldv + col) - __ldg(lse + row);
    }
    for (int e = tid; e < nt * S; e += blockDim.x) {  // (written by this CTA's own threads before a barrier)
      const int tt = e / S, s2 = e - tt * S;
      al[tt * S + s2] = aw[(size_t)(t0 + tt) * Smax + s2];
    }
    __syncthreads();
    for (int tt = nt - 1; tt >= 0; --tt) {
      const int t = t0 + tt;
      const float* b0 = ab + cur * Smax;        // beta_{t+1}
      float* b1 = ab + (cur ^ 1) * Smax;        // beta_t
      const float* at = al + tt * S;
      const float* lpt = lp + tt * W;
      float blank_part = 0.0f;
      for (int s = tid; s < S; s += blockDim.x) {
        const int u = s >> 1;
        const bool is_label = s & 1;
        const float p = is_label ? lpt[u] : lpt[U];
        float v;
        if (t == len - 1) {
          v = (s >= S - 2) ? p : -INFINITY;
        } else {
          const float x0 = b0[s];
          const float x1 = s + 1 < S ? b0[s + 1] : -INFINITY;
          const float x2 = (is_label && s + 2 < S && tgt[u + 1] != tgt[u]) ? b0[s + 2] : -INFINITY;
          v = lse3(x0, x1, x2) + p;
        }
        b1[s] = v;
        const float lg = at[s] + v - p - ll;
        const float gv = (lg == -INFINITY || isnan(lg)) ? 0.0f : expf(lg);
        if (is_label)
          gm[tt * W + u] = gv;
        else
          blank_part += gv;
      }
      blank_part = warp_sum(blank_part);  // blank column: sum over the even states, per warp here ...
      if (lane == 0) bp[tt * 8 + wid] = blank_part;
      cur ^= 1;
      __syncthreads();  // beta_t visible before step t - 1
    }
    // ... and across warps (fixed order) at the subtraction: one writer per (frame, column), no atomics
    for (int e = tid; e < nt * W; e += blockDim.x) {
      const int tt = e / W, u = e - tt * W;
      float* op = dz + ((size_t)(t0 + tt) * B + b) * ldd;
      if (u == U) {
        float tot = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += bp[tt * 8 + w];
        op[blank] -= g * tot;
      } else if (hd[u] && tgt[u] != blank) {  // the first position of each label owns the column
        float tot = gm[tt * W + u];
        for (int k = nxt[u]; k >= 0; k = nxt[k]) tot += gm[tt * W + k];
        op[tgt[u]] -= g * tot;
      }
    }
  }
}

__global__ void ctc_loss_sum_kernel(const float* __restrict__ nll, float* __restrict__ loss, int B) {
  // fixed order: lane-strided partial sums, then a butterfly -> run-to-run identical
  float s = 0.0f;
  for (int b = threadIdx.x; b < B; b += 32) s += nll[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) loss[0] = s;
}

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_ctc_uer(const int32_t* labels, const int32_t* in_lengths, const int64_t* targets,
                             int64_t ldt, const int32_t* target_lengths, int blank, int32_t* errors,
                             int32_t* pred_lengths, int64_t* totals, int L, int B, int Umax,
                             fbkst_stream_t stream) {
  FBKST_REQUIRE(labels && in_lengths && target_lengths && errors && pred_lengths && totals,
                "fbkst_ctc_uer: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && Umax >= 0 && (Umax == 0 || (targets && ldt >= Umax)),
                "fbkst_ctc_uer: bad shape");
  const size_t smem = sizeof(int) * ((size_t)L + 6 * ((size_t)Umax + 1) + (size_t)Umax);
  FBKST_REQUIRE(smem <= 200 * 1024, "fbkst_ctc_uer: L=%d / Umax=%d exceed shared memory", L, Umax);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDeviceFlag configured;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_uer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
    configured = true;
  }
  ctc_uer_kernel<<<B, 256, smem, st>>>(labels, in_lengths, reinterpret_cast<const long long*>(targets),
                                       (long long)ldt, target_lengths, blank, errors, pred_lengths, L,
                                       B, Umax);
  FBKST_CHECK_CUDA(cudaGetLastError());
  ctc_uer_totals_kernel<<<1, 32, 0, st>>>(errors, target_lengths, Umax,
                                          reinterpret_cast<long long*>(totals), B);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_ctc_loss_fwd(const void* logits, int logits_dtype, int64_t ldv, const float* lse,
                                  const int32_t* in_lengths, const int64_t* targets, int64_t ldt,
                                  const int32_t* target_lengths, int blank, float* nll, float* loss,
                                  int L, int B, int V, int Umax, fbkst_stream_t stream) {
  FBKST_REQUIRE(logits && lse && in_lengths && target_lengths && nll && loss,
                "fbkst_ctc_loss_fwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && V > 0 && ldv >= V && blank >= 0 && blank < V && Umax >= 0 &&
                    (Umax == 0 || (targets && ldt >= Umax)),
                "fbkst_ctc_loss_fwd: bad shape");
  FBKST_REQUIRE(logits_dtype == FBKST_BF16 || logits_dtype == FBKST_F32,
                "fbkst_ctc_loss_fwd: dtype must be bf16 or fp32");
  const size_t fixed = sizeof(float) * 2 * (2 * (size_t)Umax + 1) + sizeof(int) * (size_t)Umax;
  const size_t per_frame = sizeof(float) * ((size_t)Umax + 1);
  FBKST_REQUIRE(fixed + per_frame <= 200 * 1024, "fbkst_ctc_loss_fwd: Umax=%d exceeds shared memory",
                Umax);
  int tch = (int)((96 * 1024 - (fixed < 96 * 1024 ? fixed : 96 * 1024)) / per_frame);
  if (tch > 32) tch = 32;
  if (tch < 1) tch = 1;
  const size_t smem = fixed + per_frame * tch;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDeviceFlag configured;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_fwd_kernel<0>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_fwd_kernel<1>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (logits_dtype == FBKST_BF16)
    ctc_loss_fwd_kernel<1><<<B, 256, smem, st>>>(logits, (long long)ldv, lse, in_lengths,
                                                 reinterpret_cast<const long long*>(targets),
                                                 (long long)ldt, target_lengths, blank, nll, L, B, Umax,
                                                 tch);
  else
    ctc_loss_fwd_kernel<0><<<B, 256, smem, st>>>(logits, (long long)ldv, lse, in_lengths,
                                                 reinterpret_cast<const long long*>(targets),
                                                 (long long)ldt, target_lengths, blank, nll, L, B, Umax,
                                                 tch);
  FBKST_CHECK_CUDA(cudaGetLastError());
  ctc_loss_sum_kernel<<<1, 32, 0, st>>>(nll, loss, B);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" long long fbkst_ctc_loss_bwd_workspace(int L, int B, int Umax) {
  return (long long)B * L * (2 * (long long)Umax + 1);
}

/* dlogits [L*B, ldd] fp32 = grad_loss[0] * d(sum_b nll_b)/d logits  (rows t*B+b; zero for frames beyond each
 * input length and for utterances whose alignment is infeasible).  grad_loss: DEVICE scalar (the upstream
 * gradient of the summed loss).  alpha_ws: fbkst_ctc_loss_bwd_workspace(L, B, Umax) floats. */
extern "C" int fbkst_ctc_loss_bwd(const void* logits, int logits_dtype, int64_t ldv, const float* lse,
                                  const int32_t* in_lengths, const int64_t* targets, int64_t ldt,
                                  const int32_t* target_lengths, int blank, const float* grad_loss,
                                  float* alpha_ws, float* dlogits, int64_t ldd, int L, int B, int V, int Umax,
                                  fbkst_stream_t stream) {
  FBKST_REQUIRE(logits && lse && in_lengths && target_lengths && grad_loss && alpha_ws && dlogits,
                "fbkst_ctc_loss_bwd: null pointer");
  FBKST_REQUIRE(L > 0 && B > 0 && V > 0 && ldv >= V && ldd >= V && blank >= 0 && blank < V && Umax >= 0 &&
                    (Umax == 0 || (targets && ldt >= Umax)),
                "fbkst_ctc_loss_bwd: bad shape");
  FBKST_REQUIRE(logits_dtype == FBKST_BF16 || logits_dtype == FBKST_F32,
                "fbkst_ctc_loss_bwd: dtype must be bf16 or fp32");
  const size_t fixed = sizeof(float) * 2 * (2 * (size_t)Umax + 1) + sizeof(int) * 3 * (size_t)Umax;
  // per staged frame: lp [W] + alpha [S] + gamma [W] + 8 blank partials
  const size_t per_frame = sizeof(float) * (2 * ((size_t)Umax + 1) + (2 * (size_t)Umax + 1) + 8);
  FBKST_REQUIRE(fixed + per_frame <= 200 * 1024, "fbkst_ctc_loss_bwd: Umax=%d exceeds shared memory", Umax);
  int tch = (int)((160 * 1024 - (fixed < 160 * 1024 ? fixed : 160 * 1024)) / per_frame);
  if (tch > 64) tch = 64;
  if (tch < 1) tch = 1;
  const size_t smem = fixed + per_frame * tch;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDeviceFlag configured;
  if (!configured) {
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
    FBKST_CHECK_CUDA(cudaFuncSetAttribute(ctc_loss_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
    configured = true;
  }
  const int rows = L * B;
  const long long* tg = reinterpret_cast<const long long*>(targets);
  if (logits_dtype == FBKST_BF16) {
    ctc_softmax_grad_kernel<1><<<(rows + 7) / 8, 256, 0, st>>>(logits, (long long)ldv, lse, in_lengths, grad_loss,
                                                             dlogits, (long long)ldd, rows, B, V);
    FBKST_CHECK_CUDA(cudaGetLastError());
    ctc_loss_bwd_kernel<1><<<B, 256, smem, st>>>(logits, (long long)ldv, lse, in_lengths, tg, (long long)ldt,
                                                 target_lengths, blank, grad_loss, alpha_ws, dlogits,
                                                 (long long)ldd, L, B, V, Umax, tch);
  } else {
    ctc_softmax_grad_kernel<0><<<(rows + 7) / 8, 256, 0, st>>>(logits, (long long)ldv, lse, in_lengths, grad_loss,
                                                             dlogits, (long long)ldd, rows, B, V);
    FBKST_CHECK_CUDA(cudaGetLastError());
    ctc_loss_bwd_kernel<0><<<B, 256, smem, st>>>(logits, (long long)ldv, lse, in_lengths, tg, (long long)ldt,
                                                 target_lengths, blank, grad_loss, alpha_ws, dlogits,
                                                 (long long)ldd, L, B, V, Umax, tch);
  }
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
