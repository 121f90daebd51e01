// Next row N3: the training-time batch augmentations that sit between the collater and the
// encoder -- SpecAugment (examples/speech_recognition/modules/specaugment.py:55-112) and TimeStretch
// (modules/time_stretch.py:18-57), called at tasks/speech_recognition.py:254-258 -- applied to the
// batch where it already lives (HBM) instead of per-sample Python loops, a deep copy and a second H2D.
// The random draws stay on the host (they are a handful of numbers and must consume Python's and
// numpy's RNG streams exactly like the reference); everything that touches frames is here.
// All of it is HBM-bound byte work: coalesced float4 accesses, no tensor cores.
#include "host_common.h"

namespace fbkst {

// x[b, t, f] = 0 for f in a frequency band or t in a time band of utterance b.  Write-only: nothing
// is read from x, and only the masked elements are written.
// bands [B, n_freq + n_time, 2] int32 = (start, width); width 0 = no band.
__global__ void __launch_bounds__(256)
    specaugment_kernel(float* __restrict__ x, const int* __restrict__ bands, int T, int F, int n_freq,
                       int n_time, int rows_per_block) {
  extern __shared__ int sb[];  // (start, end) pairs
  const int b = blockIdx.y, nb = n_freq + n_time;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int s = __ldg(bands + ((size_t)b * nb + i) * 2), w = __ldg(bands + ((size_t)b * nb + i) * 2 + 1);
    sb[2 * i] = s;
    sb[2 * i + 1] = s + w;
  }
  __syncthreads();
  const int t0 = blockIdx.x * rows_per_block, t1 = min(t0 + rows_per_block, T);
  bool any = false;
  for (int i = 0; i < n_freq; ++i) any |= sb[2 * i + 1] > sb[2 * i];
  for (int i = n_freq; i < nb; ++i) any |= sb[2 * i] < t1 && sb[2 * i + 1] > t0;
  if (!any) return;  // block-uniform
  float* xp = x + ((size_t)b * T + t0) * F;
  const int n = (t1 - t0) * F;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = t0 + i / F, f = i % F;
    bool m = false;
    for (int k = 0; k < n_freq; ++k) m |= f >= sb[2 * k] && f < sb[2 * k + 1];
    for (int k = n_freq; k < nb; ++k) m |= t >= sb[2 * k] && t < sb[2 * k + 1];
    if (m) xp[i] = 0.0f;
  }
}

// One thread per stretch window: ids[out_off + r] = round(linspace(first, last, count))[r], the
// fp32 arithmetic of torch.linspace (step = (last - first) / (count - 1); the first half counts up
// from `first`, the second half down from `last`) followed by round-half-to-even
// (time_stretch.py:50-51).  Separate multiply and add (no FMA contraction): the index is an integer
// contract, and a fused multiply-add rounds differently half an ulp away from x.5.
// windows [n, 4] int32 = (first, last, count, out_off); out_off indexes the flat [B * T_out] ids.
__global__ void __launch_bounds__(256)
    stretch_ids_kernel(const int4* __restrict__ windows, int* __restrict__ ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 w = __ldg(windows + i);
  const int count = w.z;
  if (count <= 0) return;
  int* o = ids + w.w;
  const float first = (float)w.x, last = (float)w.y;
  if (count == 1) {
    o[0] = w.x;
    return;
  }
  const float step = __fdiv_rn(__fsub_rn(last, first), (float)(count - 1));
  const int half = count / 2;
  for (int r = 0; r < count; ++r) {
    const float v = r < half ? __fadd_rn(first, __fmul_rn(step, (float)r))
                             : __fsub_rn(last, __fmul_rn(step, (float)(count - r - 1)));
    o[r] = (int)rintf(v);
  }
}

// out[b, j, :] = ids[b, j] >= 0 ? x[b, ids[b, j], :] : 0   (ids = -1 beyond the new length).
template <typename V>
__global__ void __launch_bounds__(256)
    gather_frames_kernel(const V* __restrict__ x, const int* __restrict__ ids, V* __restrict__ out, int T,
                         int T_out, int FV, long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / FV;
    const int c = (int)(e - row * FV);
    const int b = (int)(row / T_out);
    const int id = __ldg(ids + row);
    V v;
    if (id >= 0 && id < T)
      v = __ldg(x + ((size_t)b * T + id) * FV + c);
    else
      memset(&v, 0, sizeof(V));
    out[e] = v;
  }
}

}  // namespace fbkst

using namespace fbkst;

extern "C" int fbkst_specaugment_f32(float* x, const int32_t* bands, int B, int T, int F, int n_freq,
                                     int n_time, fbkst_stream_t stream) {
  FBKST_REQUIRE(x && (bands || n_freq + n_time == 0), "fbkst_specaugment_f32: null pointer");
  FBKST_REQUIRE(B > 0 && T > 0 && F > 0 && n_freq >= 0 && n_time >= 0 && n_freq + n_time <= 1024,
                "fbkst_specaugment_f32: bad shape B=%d T=%d F=%d bands=%d+%d", B, T, F, n_freq, n_time);
  if (n_freq + n_time == 0) return FBKST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = 64;
  dim3 grid((T + rows - 1) / rows, B);
  specaugment_kernel<<<grid, 256, sizeof(int) * 2 * (n_freq + n_time), st>>>(x, bands, T, F, n_freq, n_time,
                                                                          rows);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}

extern "C" int fbkst_time_stretch_f32(const float* x, const int32_t* windows, int n_windows, int32_t* ids,
                                      float* out, int B, int T, int T_out, int F, fbkst_stream_t stream) {
  FBKST_REQUIRE(x && ids && out && (windows || n_windows == 0), "fbkst_time_stretch_f32: null pointer");
  FBKST_REQUIRE(B > 0 && T > 0 && T_out > 0 && F > 0 && n_windows >= 0,
                "fbkst_time_stretch_f32: bad shape B=%d T=%d T_out=%d F=%d", B, T, T_out, F);
  FBKST_REQUIRE((reinterpret_cast<uintptr_t>(windows) & 15u) == 0, "fbkst_time_stretch_f32: windows must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FBKST_CHECK_CUDA(cudaMemsetAsync(ids, 0xFF, sizeof(int32_t) * (size_t)B * T_out, st));
  if (n_windows > 0)
    stretch_ids_kernel<<<(n_windows + 255) / 256, 256, 0, st>>>(reinterpret_cast<const int4*>(windows), ids,
                                                               n_windows);
  const bool vec = F % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  const int FV = vec ? F / 4 : F;
  const long long total = (long long)B * T_out * FV;
  long long g = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (vec)
    gather_frames_kernel<float4><<<(int)g, 256, 0, st>>>(reinterpret_cast<const float4*>(x), ids,
                                                         reinterpret_cast<float4*>(out), T, T_out, FV, total);
  else
    gather_frames_kernel<float><<<(int)g, 256, 0, st>>>(x, ids, out, T, T_out, FV, total);
  FBKST_CHECK_CUDA(cudaGetLastError());
  return FBKST_OK;
}
